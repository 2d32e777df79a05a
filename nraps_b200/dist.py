"""One process per GPU: histories shard across ranks, tallies all-reduce.

Mirrors the reference's thread fork / ordered join (src/mc_code.rs:302-338) at
GPU granularity: rank r transports the contiguous history range
``shard_range(H, r, world)`` of every generation, the integer tally buffer is
summed across ranks with one ``all_reduce`` (NCCL over NVLink on GPUs), and
every rank then runs the same finalize, so all ranks hold identical k / flux.
Because history streams depend only on the global history index and the tally
is an integer sum, the result is bit-identical for any world size.
"""
from __future__ import annotations

from typing import Protocol


def shard_range(histories: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, begin+count) of rank `rank`; the ranges tile [0, histories)."""
    begin = histories * rank // world
    end = histories * (rank + 1) // world
    return begin, end - begin


class GenerationEngine(Protocol):
    generations: int
    histories: int

    def transport(self, gen: int, hist_begin: int, hist_count: int, stream=None) -> None: ...
    def finalize_generation(self, gen: int, stream=None) -> None: ...


def run_generations(engine: GenerationEngine, tally, rank: int, world: int, *, all_reduce=None, stream=None,
                    first_gen: int = 0, n_gens: int | None = None) -> None:
    """Drive `n_gens` generations of `engine` for this rank.

    `tally` is the buffer the engine accumulates into (an int64 tensor);
    `all_reduce(tally)` sums it in place across ranks (skipped when world == 1).
    """
    begin, count = shard_range(engine.histories, rank, world)
    last = engine.generations if n_gens is None else first_gen + n_gens
    for gen in range(first_gen, last):
        engine.transport(gen, begin, count, stream)
        if world > 1:
            all_reduce(tally)
        engine.finalize_generation(gen, stream)


def monte_carlo_distributed(variables, xsdata, delta_x, meshid, fuel_indices, k_new: float = 1.0, *, generations=None,
                            histories=None, skip=None, **options):
    """`monte_carlo` across the ranks of the default torch.distributed group (NCCL, one GPU per rank)."""
    import torch
    import torch.distributed as dist

    from .api import MonteCarloContext

    rank, world = dist.get_rank(), dist.get_world_size()
    device = options.pop("device", torch.cuda.current_device())
    with torch.cuda.device(device):
        ctx = MonteCarloContext(variables, xsdata, delta_x, meshid, fuel_indices, k_new, generations=generations,
                                histories=histories, skip=skip, device=device, **options)
        try:
            tally = torch.zeros(ctx.n_words, dtype=torch.int64, device=f"cuda:{device}")
            ctx.use_tally_tensor(tally)
            stream = torch.cuda.current_stream().cuda_stream
            run_generations(ctx, tally, rank, world, all_reduce=lambda t: dist.all_reduce(t), stream=stream)
            return ctx.fetch(stream)
        finally:
            ctx.close()
