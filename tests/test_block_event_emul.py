"""The block-level event pipeline (nraps_b200/csrc/mc_block_event.cuh, kernel_variant = block_event, experimental)
without a GPU: its per-thread body -- the very code block_event_kernel runs -- executed by CPU threads (tests/emul,
one pthread per CUDA thread, real barriers, real atomics) and bit-compared with the oracle: every tally bin and the
event counters of one generation.  What this cannot cover is the CUDA-only context (warp-aggregated list claims and the
shared-space PTX helpers shared with the lane kernels); the GPU parity tests of the variant do that."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import load_case, oracle_inputs, synthetic_case

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "emul", "_build", "libbev_emul.so")
SCATTER = {"single_xi": 0, "rust_pre182": 1, "rust_182": 2}


@pytest.fixture(scope="module")
def emul():
    r = subprocess.run(["make", "-C", os.path.join(HERE, "emul")], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    if r.returncode != 0:
        pytest.fail("tests/emul does not build:\n" + r.stdout + r.stderr)
    L = C.CDLL(LIB)
    L.bev_emul_generation.argtypes = [C.POINTER(orc.Problem), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return L


def _problem(args, histories, generations):
    """oracle.Problem has the field order of nraps_problem (include/nraps_mc.h); the arrays are kept alive by the caller."""
    v, xs, dx, mesh, fuel = args
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    keep = dict(sigt=f(deck.sigt), sigs=f(deck.sigs), mu=f(deck.mu), siga=f(deck.siga), sigf=f(deck.sigf), nut=f(deck.nut),
                chit=f(deck.chit), inv_sigtr=f(deck.inv_sigtr), scat=f(deck.scat), matid=np.ascontiguousarray(m[0], dtype=np.uint8),
                dx=f(m[1]), left=f(m[2]), right=f(m[3]), fuel=np.ascontiguousarray(m[4], dtype=np.uint64))
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    p = orc.Problem(M=int(deck.mattypes), G=int(deck.energygroups), N=len(keep["matid"]), NF=len(keep["fuel"]), numass=int(deck.numass),
                    generations=generations, histories=histories, skip=1, boundl=float(deck.boundl), boundr=float(deck.boundr),
                    dx_fuel=float(deck.dx_fuel), dx_water=float(deck.dx_water), k0=1.0,
                    sigt=fp(keep["sigt"]), sigs=fp(keep["sigs"]), mu=fp(keep["mu"]), siga=fp(keep["siga"]), sigf=fp(keep["sigf"]),
                    nut=fp(keep["nut"]), chit=fp(keep["chit"]), inv_sigtr=fp(keep["inv_sigtr"]), scat=fp(keep["scat"]),
                    matid=keep["matid"].ctypes.data_as(C.POINTER(C.c_uint8)), dx=fp(keep["dx"]), left=fp(keep["left"]),
                    right=fp(keep["right"]), fuel_indices=keep["fuel"].ctypes.data_as(C.POINTER(C.c_uint64)))
    return p, keep, deck, m


def _run(emul, args, *, H, gen, blocks=3, threads=8, slots=40, chunk=16, walk_cap=8, max_flights=1 << 24, scatter_mode="single_xi",
         stale_xs=True, seed=42, seq=54, stride=152917):
    p, keep, deck, m = _problem(args, H, gen + 1)
    G, N = p.G, p.N
    tally = np.zeros(G * N, np.uint64)
    counters = np.zeros(8, np.uint64)
    rc = emul.bev_emul_generation(C.byref(p), gen, 0, H, seed, seq, stride, SCATTER[scatter_mode], int(stale_xs), walk_cap, max_flights,
                                  blocks, threads, slots, chunk, tally.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  counters.ctypes.data_as(C.POINTER(C.c_uint64)), None)
    assert rc == 0
    want = orc.monte_carlo(deck, m, generations=gen + 1, histories=H, skip=1, threads=4, want_tally=True, trace_gen=gen,
                           scatter_mode=scatter_mode, stale_xs=stale_xs, seed=seed, seq=seq, stride=stride,
                           max_flights=0 if max_flights == 1 << 24 else max_flights)
    tr = want.trace
    assert np.array_equal(tally.reshape(G, N), want.tally_fixed[gen]), "tally bins differ from the oracle"
    assert counters[0] == H
    assert counters[1] == tr[:, 0].sum() and counters[3] == tr[:, 2].sum()           # collisions, flights
    assert counters[5] == (tr[:, 8] == 2).sum() and counters[6] == (tr[:, 8] == 3).sum()  # leaks, truncated
    return counters


@pytest.mark.parametrize("case,gen", [("a", 0), ("b", 1), ("c", 0), ("c", 2)])
def test_shipped_decks_bit_exact(emul, case, gen):
    _run(emul, load_case(case), H=3000, gen=gen)


@pytest.mark.parametrize("kw", [
    dict(blocks=1, threads=1, slots=1, chunk=1),        # one neutron at a time: the degenerate schedule
    dict(blocks=2, threads=32, slots=32, chunk=7),      # one slot per thread, ragged chunks
    dict(blocks=5, threads=4, slots=64, chunk=1000),    # more slots than a chunk can feed; one block takes most of the work
    dict(blocks=1, threads=16, slots=200, chunk=64, walk_cap=3),  # walks suspended every three crossings
])
def test_block_geometry_does_not_change_a_bit(emul, kw):
    _run(emul, load_case("c"), H=2000, gen=1, **kw)


@pytest.mark.parametrize("mode", ["rust_pre182", "rust_182"])
def test_probe_orders_and_the_fixed_index(emul, mode):
    _run(emul, load_case("c"), H=1500, gen=0, scatter_mode=mode, stale_xs=False, seed=7, seq=3, stride=1000)


@pytest.mark.parametrize("bl,br", [(0.0, 0.0), (0.5, 1.0), (1.0, 0.0)])
def test_vacuum_and_albedo_walls(emul, bl, br):
    v, xs, dx, mesh, fuel = load_case("b")
    v.boundl, v.boundr = bl, br
    ct = _run(emul, (v, xs, dx, mesh, fuel), H=2500, gen=0)
    assert (ct[5] > 0) == (bl == 0.0 or br == 0.0)


@pytest.mark.parametrize("threshold", ["2", "5"])
def test_walk_classes_by_predicted_crossings_change_nothing(emul, monkeypatch, threshold):
    """nraps_options.spawn_batch = T for this variant: the two walk lists are split by a prediction of the crossing
    count instead of the run length.  The prediction only sorts, so every bin must stay the same."""
    monkeypatch.setenv("BEV_EMUL_CLASS_T", threshold)
    v, xs, dx, mesh, fuel = load_case("c")
    v.boundl = 0.5
    _run(emul, (v, xs, dx, mesh, fuel), H=2000, gen=0, blocks=2, threads=8, slots=48, chunk=32)


def test_flight_cap_truncates_like_the_oracle(emul):
    ct = _run(emul, load_case("a"), H=800, gen=0, max_flights=5)
    assert ct[6] > 0


@pytest.mark.parametrize("M,G,pins,mpfr,mpwr,bl,br", [
    (2, 3, [1, 0, 1], 3, 2, 1.0, 1.0), (3, 5, [2, 0, 2, 1, 2, 0, 2], 5, 4, 1.0, 0.0), (5, 8, [4, 0, 3, 1, 2, 0, 4], 4, 6, 0.7, 1.0),
    (2, 2, [0], 1, 0, 1.0, 1.0), (3, 4, [2, 1, 2], 64, 2, 0.0, 0.0),
])
def test_synthetic_shapes(emul, M, G, pins, mpfr, mpwr, bl, br):
    """Generic group counts, five materials, a single cell (both walls in it), one long run."""
    args = synthetic_case(M, G, pins, mpfr, mpwr, seed=M * 10 + G, boundl=bl, boundr=br)
    _run(emul, args, H=1500, gen=1, walk_cap=24)


def test_shards_of_a_generation_add_up(emul):
    """Two ranks' shards [0, H/3) and [H/3, H) of one generation (contiguous history ranges, as nraps_b200.dist cuts
    them): the integer tallies add up to the oracle's whole generation."""
    args = load_case("c")
    H, gen = 2400, 1
    p, keep, deck, m = _problem(args, H, gen + 1)
    total = np.zeros(p.G * p.N, np.uint64)
    hist = 0
    for begin, count in ((0, H // 3), (H // 3, H - H // 3)):
        tally = np.zeros(p.G * p.N, np.uint64)
        counters = np.zeros(8, np.uint64)
        rc = emul.bev_emul_generation(C.byref(p), gen, begin, count, 42, 54, 152917, 0, 1, 8, 1 << 24, 2, 8, 24, 16,
                                      tally.ctypes.data_as(C.POINTER(C.c_uint64)), counters.ctypes.data_as(C.POINTER(C.c_uint64)), None)
        assert rc == 0
        total += tally
        hist += int(counters[0])
    want = orc.monte_carlo(deck, m, generations=gen + 1, histories=H, skip=1, threads=4, want_tally=True)
    assert hist == H and np.array_equal(total.reshape(p.G, p.N), want.tally_fixed[gen])


def test_warp_claim_arithmetic(emul):
    """DevCtx::claim2 (mc_block_event.cu): one atomic on the packed length word of an arena for both of its lists.  Lane
    by lane, for random ballots: the pushing lanes of either list get consecutive positions starting at the list's
    old length, in lane order, and the word grows by the two counts."""
    emul.bev_emul_claim2.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    emul.bev_emul_claim2.restype = C.c_uint32
    rng = np.random.default_rng(3)
    pos = (C.c_uint32 * 32)()
    for _ in range(3000):
        up = int(rng.integers(0, 1 << 32)) & int(rng.integers(0, 1 << 32))
        down = int(rng.integers(0, 1 << 32)) & ~up & 0xFFFFFFFF  # a lane pushes to at most one list
        if rng.integers(0, 8) == 0:
            up, down = (0xFFFFFFFF, 0) if rng.integers(0, 2) else (0, 0xFFFFFFFF)
        n_up, n_down = int(rng.integers(0, 60000)), int(rng.integers(0, 5000))
        new = emul.bev_emul_claim2(up, down, n_up | (n_down << 16), pos)
        ups = [pos[i] for i in range(32) if (up >> i) & 1]
        downs = [pos[i] for i in range(32) if (down >> i) & 1]
        assert ups == list(range(n_up, n_up + len(ups))) and downs == list(range(n_down, n_down + len(downs)))
        assert new == (n_up + len(ups)) | ((n_down + len(downs)) << 16)
