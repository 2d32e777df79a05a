"""One process per GPU: histories shard across ranks, tallies all-reduce.

Mirrors the reference's thread fork / ordered join (src/mc_code.rs:302-338) at
GPU granularity: rank r transports the contiguous history range
``shard_range(H, r, world)`` of every generation, the integer tally buffer is
summed across ranks with one ``all_reduce`` (NCCL over NVLink on GPUs), and
every rank then runs the same finalize, so all ranks hold identical k / flux.
Because history streams depend only on the global history index and the tally
is an integer sum, the result is bit-identical for any world size.

Two refinements (round 2):

* uniform source: generations are independent (k cancels, src/mc_code.rs:346-351), so the all-reduce and the
  finalize of generation g run on a side stream while generation g+1 is already transporting into a second tally
  buffer (``OverlappedReducer``): the collective leaves the critical path; the transports alternate between two
  streams, so that the launch of g+1 fills the tail of g.
* fission_bank source: nothing is gathered.  Every rank keeps the bank it compacted in a peer-mapped buffer, the
  source kernel of generation g+1 turns a site index into (rank, offset) from the ranks' site counts and loads the
  site over NVLink (``setup_bank_peers`` shares the buffers once: cuMemCreate allocations, exported file descriptors).  The per-generation all-reduce --
  issued after the local compaction -- is the only collective and doubles as the barrier between "every rank has
  compacted bank g" and "any rank samples from it"; the bank's cell histogram rides in the same buffer.
"""
from __future__ import annotations

from typing import Protocol


def shard_range(histories: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, begin+count) of rank `rank`; the ranges tile [0, histories)."""
    begin = histories * rank // world
    end = histories * (rank + 1) // world
    return begin, end - begin


def resolve_site(index: int, counts: list[int]) -> tuple[int, int]:
    """(rank, offset) of global site `index` of a bank held as one dense list per rank, canonical order = rank order.

    Host mirror of the lookup in source_kernel (nraps_b200/csrc/mc_source.cu): the first rank r whose cumulative
    count exceeds `index`; ranks with an empty bank are skipped."""
    first = 0
    for rank, n in enumerate(counts):
        if index < first + n:
            return rank, index - first
        first += n
    raise IndexError(f"site {index} of a bank of {first}")


class GenerationEngine(Protocol):
    generations: int
    histories: int

    def transport(self, gen: int, hist_begin: int, hist_count: int, stream=None) -> None: ...
    def finalize_generation(self, gen: int, stream=None) -> None: ...


def run_generations(engine: GenerationEngine, tally, rank: int, world: int, *, all_reduce=None, stream=None,
                    first_gen: int = 0, n_gens: int | None = None, bank: bool = False) -> None:
    """Drive `n_gens` generations of `engine` for this rank, one after the other.

    `tally` is the buffer the engine accumulates into (an int64 tensor); `all_reduce(tally)` sums it in place across
    ranks (skipped when world == 1).  With ``bank`` (fission_bank source mode) the local bank is compacted *before*
    the all-reduce -- the collective is then also the point after which every rank's bank of this generation is
    complete -- and becomes the source of generation gen+1 after the finalize.
    """
    begin, count = shard_range(engine.histories, rank, world)
    last = engine.generations if n_gens is None else first_gen + n_gens
    for gen in range(first_gen, last):
        engine.transport(gen, begin, count, stream)
        if bank:
            engine.bank_compact(gen, stream)
        if world > 1:
            all_reduce(tally)
        engine.finalize_generation(gen, stream)
        if bank:
            engine.bank_advance(gen, stream)


class OverlappedReducer:
    """Uniform source on GPUs: generations are independent, so their launches are pipelined.

    * all-reduce + finalize of generation g run on a side stream while g+1 transports (two tally tensors alternate);
    * the transports themselves alternate between two streams and the context's two scratch lanes
      (``select_lane``): the launch of g+1 is already queued on the device while the tail of g runs -- the last
      neutrons of a persistent launch finish one by one, ~0.6 ms with most of the GPU idle -- and its blocks move in as
      blocks of g retire (8.72e8 -> 8.90e8 histories/s on config 3).  ``pipeline=False`` keeps
      every transport on the main stream.

    Order kept by events: transport(g) waits for the finalize of g-2 (same tally tensor, same lane and stream); the side
    stream takes the generations in order (k and the running flux sums are sequential f32 accumulations)."""

    def __init__(self, ctx, world: int, device: int, main_stream=None, pipeline: bool = True):
        import torch

        import os

        self.torch, self.ctx, self.world = torch, ctx, world
        dev = f"cuda:{device}"
        pipeline = pipeline and os.environ.get("NRAPS_PIPELINE", "1") != "0"  # 0: every transport on the main stream (A/B)
        self.pipelined = pipeline
        self.tallies = [torch.zeros(ctx.n_words, dtype=torch.int64, device=dev) for _ in range(2)]
        self.main = main_stream if main_stream is not None else torch.cuda.current_stream()
        self.streams = [self.main, torch.cuda.Stream(device=dev) if pipeline else self.main]
        self.side = torch.cuda.Stream(device=dev)
        self.done = [None, None]   # side-stream events: finalize of the generation that last used the tensor
        self.last = [None, None]   # transport-stream events: end of the last launch on each stream
        self.launched = 0
        start = torch.cuda.Event()
        start.record(self.main)    # whatever the main stream was doing (tables, earlier runs) comes first on both
        self.streams[1].wait_event(start)

    def step(self, gen: int, hist_begin: int, hist_count: int, before_transport=None, after_transport=None):
        torch = self.torch
        import torch.distributed as dist

        lane = gen & 1
        t, stream = self.tallies[lane], self.streams[lane]
        with torch.cuda.stream(stream):   # the callbacks' torch work (L2 flush, timing events) goes where the launch goes
            if self.done[lane] is not None:
                stream.wait_event(self.done[lane])
            self.ctx.select_lane(lane)
            self.ctx.use_tally_tensor(t)
            if before_transport is not None:
                before_transport()
            self.ctx.transport(gen, hist_begin, hist_count, stream.cuda_stream)
            if after_transport is not None:
                after_transport()
            ready = torch.cuda.Event()
            ready.record(stream)
        self.last[lane] = ready
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if self.world > 1:
                dist.all_reduce(t)
            # the context reads its current tally pointer at launch time: point it at this generation's tensor
            self.ctx.use_tally_tensor(t)
            self.ctx.finalize_generation(gen, self.side.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[lane] = ev
        self.launched += 1

    def drain(self):
        """Main stream waits for everything the other streams still owe."""
        for ev in self.done + self.last:
            if ev is not None:
                self.main.wait_event(ev)


def setup_bank_peers(ctx, rank: int, world: int, shard_max: int) -> None:
    """fission_bank mode on several GPUs: make every rank's two bank buffers readable by every other rank.

    Allocates the buffers shareable (cuMemCreate, 2 MB pages), exchanges their tickets (pid + exported file descriptor)
    through the default process group and maps the peers' buffers; afterwards no bank data moves except the 8-byte
    sites the source kernel asks for."""
    import torch.distributed as dist

    if world == 1:
        return
    ctx.bank_reserve(shard_max)
    mine = ctx.bank_export()
    handles: list = [None] * world
    dist.all_gather_object(handles, mine)
    ctx.bank_import(world, rank, b"".join(handles))
    dist.barrier()  # nobody starts generation 0 before every mapping exists


def monte_carlo_distributed(variables, xsdata, delta_x, meshid, fuel_indices, k_new: float = 1.0, *, generations=None,
                            histories=None, skip=None, **options):
    """`monte_carlo` across the ranks of the default torch.distributed group (NCCL, one GPU per rank)."""
    import torch
    import torch.distributed as dist

    from .api import MonteCarloContext

    rank, world = dist.get_rank(), dist.get_world_size()
    device = options.pop("device", torch.cuda.current_device())
    bank = options.get("source_mode") == "fission_bank"
    with torch.cuda.device(device):
        ctx = MonteCarloContext(variables, xsdata, delta_x, meshid, fuel_indices, k_new, generations=generations,
                                histories=histories, skip=skip, device=device, **options)
        try:
            stream = torch.cuda.current_stream()
            begin, count = shard_range(ctx.histories, rank, world)
            if bank:
                setup_bank_peers(ctx, rank, world, max(shard_range(ctx.histories, r, world)[1] for r in range(world)))
                tally = torch.zeros(ctx.n_words, dtype=torch.int64, device=f"cuda:{device}")
                ctx.use_tally_tensor(tally)
                run_generations(ctx, tally, rank, world, all_reduce=lambda t: dist.all_reduce(t), stream=stream.cuda_stream, bank=True)
            else:
                red = OverlappedReducer(ctx, world, device, stream)
                for gen in range(ctx.generations):
                    red.step(gen, begin, count)
                red.drain()
            res = ctx.fetch(stream.cuda_stream)
            if world > 1:
                dist.barrier()  # peers may still be reading this rank's bank buffers: close nothing before all are done
            return res
        finally:
            ctx.close()
