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


def gather_bank(local_sites, world: int, all_gather_counts, all_gather_padded):
    """All-gather variable-length site lists into one bank in rank order (= canonical history order).

    `local_sites`: 1-D int64 tensor of this rank's sites; `all_gather_counts(n) -> list[int]`;
    `all_gather_padded(padded_tensor, max_n) -> [world, max_n] tensor`.  NCCL all-gather needs equal
    sizes, so every rank pads to the largest count and the valid prefixes are concatenated afterwards.
    """
    import torch

    counts = all_gather_counts(int(local_sites.numel()))
    max_n = max(counts)
    if max_n == 0:
        return local_sites[:0], counts
    padded = torch.zeros(max_n, dtype=local_sites.dtype, device=local_sites.device)
    padded[: local_sites.numel()] = local_sites
    gathered = all_gather_padded(padded, max_n)
    return torch.cat([gathered[r, : counts[r]] for r in range(world)]), counts


def run_generations(engine: GenerationEngine, tally, rank: int, world: int, *, all_reduce=None, stream=None,
                    first_gen: int = 0, n_gens: int | None = None, bank=None) -> None:
    """Drive `n_gens` generations of `engine` for this rank.

    `tally` is the buffer the engine accumulates into (an int64 tensor);
    `all_reduce(tally)` sums it in place across ranks (skipped when world == 1).
    `bank(gen)`, when given (fission_bank source mode), compacts / gathers the
    fission bank and installs it as the source of generation gen+1.
    """
    begin, count = shard_range(engine.histories, rank, world)
    last = engine.generations if n_gens is None else first_gen + n_gens
    for gen in range(first_gen, last):
        engine.transport(gen, begin, count, stream)
        if world > 1:
            all_reduce(tally)
        engine.finalize_generation(gen, stream)
        if bank is not None:
            bank(gen)


def make_bank_callback(ctx, world: int, device: int, stream):
    """bank(gen) for run_generations on GPUs: compact locally, all-gather over NCCL, install as next source."""
    import torch
    import torch.distributed as dist

    from .api import _DevArray

    dev = f"cuda:{device}"

    def counts_fn(n):
        t = torch.tensor([n], dtype=torch.int64, device=dev)
        out = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return [int(v) for v in out.tolist()]

    def padded_fn(padded, max_n):
        out = torch.empty(world * max_n, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, padded)
        return out.view(world, max_n)

    def bank(gen):
        ctx.bank_compact(gen, stream)
        if world == 1:
            ctx.bank_set_source(gen, None, stream)
            return
        ptr, n = ctx.bank_local(stream)
        local = torch.as_tensor(_DevArray(ptr, n), device=dev) if n else torch.zeros(0, dtype=torch.int64, device=dev)
        full, _ = gather_bank(local, world, counts_fn, padded_fn)
        ctx.bank_set_source(gen, full.contiguous() if full.numel() else None, stream)

    return bank


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
            bank = make_bank_callback(ctx, world, device, stream) if options.get("source_mode") == "fission_bank" else None
            run_generations(ctx, tally, rank, world, all_reduce=lambda t: dist.all_reduce(t), stream=stream, bank=bank)
            return ctx.fetch(stream)
        finally:
            ctx.close()
