"""GPU parity of the experimental block-level event pipeline (kernel_variant = "block_event"): bit-exact against the
oracle and against the default lane kernel, through the C ABI.

Opt-in (NRAPS_TEST_BLOCK_EVENT=1): the variant was written after this round's GPU minutes were spent.  Its per-thread
body is verified on CPU threads (tests/test_block_event_emul.py); the CUDA-only context around it (warp-aggregated list
claims, launch geometry, the shared-memory carve-up) has not run on a device yet, so these tests stay out of the default
`-m gpu` run until they have."""
import os

import numpy as np
import pytest

import nraps_b200 as nb
from oracle import oracle as orc
from tests.util import bits, load_case, oracle_inputs, synthetic_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NRAPS_TEST_BLOCK_EVENT", "0") != "1", reason="opt-in: NRAPS_TEST_BLOCK_EVENT=1")]


def _identical(got, want):
    assert np.array_equal(got.tally_fixed, want.tally_fixed)
    for name in ("k", "k_fund", "flux", "assembly_average", "fission_source"):
        assert np.array_equal(bits(getattr(got, name)), bits(getattr(want, name))), name
    for c in ("histories", "collisions", "flights", "leaks", "truncated"):
        assert got.counters[c] == want.counters[c], c


@pytest.mark.parametrize("case,H,gens", [("a", 100_000, 4), ("b", 150_000, 3), ("c", 150_000, 3)])
def test_results_bit_exact(case, H, gens):
    args = load_case(case)
    got = nb.monte_carlo(*args, 1.0, generations=gens, histories=H, skip=1, want_tally=True, kernel_variant="block_event")
    deck, m = oracle_inputs(*args)
    want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=8, want_tally=True)
    _identical(got, want)


@pytest.mark.parametrize("kw", [dict(threads_per_block=64, blocks_per_sm=1, slots_per_thread=1), dict(threads_per_block=1024, blocks_per_sm=1, slots_per_thread=3),
                                dict(threads_per_block=256, blocks_per_sm=4, slots_per_thread=5, chunk=33), dict(walk_cap=3), dict(spawn_batch=5),
                                dict(spawn_batch=2, walk_cap=6)])
def test_geometry_does_not_change_a_bit(kw):
    args = load_case("c")
    ref = nb.monte_carlo(*args, 1.0, generations=2, histories=200_000, skip=1, want_tally=True)
    got = nb.monte_carlo(*args, 1.0, generations=2, histories=200_000, skip=1, want_tally=True, kernel_variant="block_event", **kw)
    _identical(got, ref)


@pytest.mark.parametrize("bl,br", [(0.0, 0.0), (0.5, 1.0), (1.0, 0.0)])
def test_walls_and_switches(bl, br):
    v, xs, dx, mesh, fuel = load_case("b")
    v.boundl, v.boundr = bl, br
    for kw in (dict(), dict(scatter_mode="rust_182", stale_xs=False, seed=7, stream=3, stride=1000)):
        ref = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=80_000, skip=1, want_tally=True, **kw)
        got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=80_000, skip=1, want_tally=True,
                             kernel_variant="block_event", **kw)
        _identical(got, ref)


@pytest.mark.parametrize("M,G,pins,mpfr,mpwr,bl,br", [
    (2, 3, [1, 0, 1], 3, 2, 1.0, 1.0), (3, 5, [2, 0, 2, 1, 2, 0, 2], 5, 4, 1.0, 0.0), (5, 8, [4, 0, 3, 1, 2, 0, 4], 4, 6, 0.7, 1.0),
    (2, 2, [0], 1, 0, 1.0, 1.0), (3, 4, [2, 1, 2], 64, 2, 0.0, 0.0),
])
def test_synthetic_shapes(M, G, pins, mpfr, mpwr, bl, br):
    args = synthetic_case(M, G, pins, mpfr, mpwr, seed=M * 10 + G, boundl=bl, boundr=br)
    ref = nb.monte_carlo(*args, 1.0, generations=3, histories=30_000, skip=1, want_tally=True)
    got = nb.monte_carlo(*args, 1.0, generations=3, histories=30_000, skip=1, want_tally=True, kernel_variant="block_event")
    _identical(got, ref)


def test_flight_cap_and_fine_mesh():
    args = load_case("a")
    ref = nb.monte_carlo(*args, 1.0, generations=2, histories=20_000, skip=1, want_tally=True, max_flights=5)
    got = nb.monte_carlo(*args, 1.0, generations=2, histories=20_000, skip=1, want_tally=True, max_flights=5, kernel_variant="block_event")
    _identical(got, ref)
    assert got.counters["truncated"] > 0
    args = load_case("c", mpfr=80, mpwr=40)
    ref = nb.monte_carlo(*args, 1.0, generations=2, histories=50_000, skip=1, want_tally=True)
    got = nb.monte_carlo(*args, 1.0, generations=2, histories=50_000, skip=1, want_tally=True, kernel_variant="block_event")
    _identical(got, ref)
