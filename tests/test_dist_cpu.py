"""N > 1 host logic on CPU: world_size-2 gloo.  The GPU engine is replaced by an
oracle-backed stand-in with the same transport / finalize interface, so what is
exercised is exactly nraps_b200.dist: history sharding, the per-generation
integer all-reduce of the tally buffer, and the claim that the result is
bit-identical for any world size."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from nraps_b200.dist import resolve_site, run_generations, shard_range  # noqa: E402


def test_shard_ranges_tile_the_generation():
    for H in (1, 7, 100_000, 10**7 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == H
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_site_lookup_is_the_rank_ordered_concatenation():
    """The source kernel never sees a gathered bank: it maps a global site index to (rank, offset) from the ranks'
    site counts.  That map must be exactly the rank-ordered concatenation of the local banks -- what makes generation
    g+1 independent of the number of GPUs -- including ragged and empty local banks."""
    for counts in ([5, 0, 3], [0, 4], [7, 0], [1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 9, 0]):
        flat = [(r, i) for r, n in enumerate(counts) for i in range(n)]
        assert [resolve_site(s, counts) for s in range(len(flat))] == flat
        with pytest.raises(IndexError):
            resolve_site(len(flat), counts)
    # the device's index draw: site = (u32 * B) >> 32 covers [0, B) and nothing else
    B = sum([5, 0, 3])
    assert {(u * B) >> 32 for u in (0, 1, 2**31, 2**32 - 1)} <= set(range(B)) and ((2**32 - 1) * B) >> 32 == B - 1


def test_bank_protocol_order():
    """fission_bank mode: the local compaction comes before the all-reduce (the collective is the barrier after which
    every rank's bank is complete) and the bank becomes the next source only after the finalize."""
    calls = []

    class Engine:
        generations, histories = 2, 10

        def transport(self, gen, b, n, stream=None): calls.append(("transport", gen, b, n))
        def bank_compact(self, gen, stream=None): calls.append(("compact", gen))
        def finalize_generation(self, gen, stream=None): calls.append(("finalize", gen))
        def bank_advance(self, gen, stream=None): calls.append(("advance", gen))

    run_generations(Engine(), None, 1, 2, all_reduce=lambda t: calls.append(("all_reduce",)), bank=True)
    assert calls == [("transport", 0, 5, 5), ("compact", 0), ("all_reduce",), ("finalize", 0), ("advance", 0),
                     ("transport", 1, 5, 5), ("compact", 1), ("all_reduce",), ("finalize", 1), ("advance", 1)]


class OracleEngine:
    """Stand-in for MonteCarloContext on CPU: transport = the oracle on a history sub-range."""

    def __init__(self, case, generations, histories, tally):
        from oracle import oracle as orc
        from tests.util import load_case, oracle_inputs

        self.orc = orc
        self.deck, self.mesh = oracle_inputs(*load_case(case))
        self.generations, self.histories = generations, histories
        self.tally = tally  # torch int64 [G*N]
        self.k = []
        self.per_gen = []

    def transport(self, gen, hist_begin, hist_count, stream=None):
        import torch

        # the oracle derives streams from (gen, y), so running generation `gen` alone needs gens = gen+1
        r = self.orc.monte_carlo(self.deck, self.mesh, generations=gen + 1, histories=self.histories, skip=0, threads=2,
                                 hist_begin=hist_begin, hist_count=hist_count, want_tally=True) if hist_count else None
        flat = r.tally_fixed[gen].reshape(-1).astype(np.int64) if r is not None else 0
        self.tally.copy_(torch.as_tensor(flat) if r is not None else torch.zeros_like(self.tally))

    def finalize_generation(self, gen, stream=None):
        self.per_gen.append(self.tally.numpy().copy())


def _free_port() -> int:
    """A TCP port nobody listens on right now (asked of the OS), so two suites on one host do not collide."""
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.util import load_case

    G, N = load_case("a")[0].energygroups, 408
    tally = torch.zeros(G * N, dtype=torch.int64)
    eng = OracleEngine("a", generations=2, histories=9001, tally=tally)
    run_generations(eng, tally, rank, world, all_reduce=lambda t: dist.all_reduce(t))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack(eng.per_gen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_equals_single_rank(tmp_path):
    import torch
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)  # every rank holds the same reduced tally

    from tests.util import load_case

    G, N = load_case("a")[0].energygroups, 408
    tally = torch.zeros(G * N, dtype=torch.int64)
    eng = OracleEngine("a", generations=2, histories=9001, tally=tally)
    run_generations(eng, tally, 0, 1)
    assert np.array_equal(np.stack(eng.per_gen), r0)  # and it is the single-rank tally, bit for bit
    assert r0.any()
