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

from nraps_b200.dist import gather_bank, run_generations, shard_range  # noqa: E402


def test_shard_ranges_tile_the_generation():
    for H in (1, 7, 100_000, 10**7 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == H
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_gather_bank_concatenates_in_rank_order():
    """Variable-length site lists -> one bank in rank (= history) order; simulated 3-rank all-gather."""
    import torch

    locals_ = [torch.arange(5, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), torch.arange(100, 103, dtype=torch.int64)]
    counts = [t.numel() for t in locals_]
    pads = {}

    def counts_fn(n):
        return counts

    def make_padded_fn(rank):
        def fn(padded, max_n):
            pads[rank] = padded
            rows = [torch.zeros(max_n, dtype=torch.int64) for _ in range(3)]
            for r, t in enumerate(locals_):
                rows[r][: t.numel()] = t
            return torch.stack(rows)
        return fn

    for rank in range(3):
        full, got_counts = gather_bank(locals_[rank], 3, counts_fn, make_padded_fn(rank))
        assert got_counts == counts and full.tolist() == [0, 1, 2, 3, 4, 100, 101, 102]
        assert pads[rank].numel() == 5 and pads[rank][: counts[rank]].tolist() == locals_[rank].tolist()


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


def _bank_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def counts_fn(n):  # same shape of exchange as make_bank_callback's NCCL version
        out = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(out, torch.tensor([n], dtype=torch.int64))
        return [int(t.item()) for t in out]

    def padded_fn(padded, max_n):
        out = [torch.zeros(max_n, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(out, padded)
        return torch.stack(out)

    banks = []
    for gen, sizes in enumerate([(5, 3), (0, 4), (7, 0), (0, 0)]):  # ragged, one side empty, both empty
        local = torch.arange(sizes[rank], dtype=torch.int64) + 1000 * rank + 100 * gen
        full, counts = gather_bank(local, world, counts_fn, padded_fn)
        assert counts == list(sizes)
        banks.append(full.numpy().copy())
    np.save(os.path.join(out_dir, f"bank{rank}.npy"), np.concatenate(banks))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_bank_gather_is_rank_ordered_and_identical_everywhere(tmp_path):
    """The fission-bank exchange of nraps_b200.dist over real collectives (gloo, world size 2): every rank ends with
    the same bank, rank 0's sites first -- what makes generation g+1 independent of the number of GPUs."""
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_bank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "bank0.npy"), np.load(tmp_path / "bank1.npy")
    assert np.array_equal(b0, b1)
    want = np.r_[np.arange(5), 1000 + np.arange(3), 1100 + np.arange(4), 200 + np.arange(7)]
    assert np.array_equal(b0, want)
