"""GPU parity tests proper: the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.  Integer results (event counts, RNG
states, fixed-point tally bins) and everything derived from them with the
reference's binary32 arithmetic (k, flux, fission source, CSV bytes) must be
BIT-EXACT.  Statistical agreement between independent streams is stated at
3 sigma combined / per-bin chi-square."""
import os

import numpy as np
import pytest

import nraps_b200 as nb
from nraps_b200 import _lib
from oracle import host_oracle as ho
from oracle import oracle as orc
from tests.util import bits, load_case, oracle_inputs

pytestmark = pytest.mark.gpu
f32 = np.float32


def _both(case, *, generations, histories, skip=1, oracle_kw=None, gpu_kw=None, threads=8, **common):
    v, xs, dx, mesh, fuel = load_case(case)
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=generations, histories=histories, skip=skip,
                         want_tally=True, **common, **(gpu_kw or {}))
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    okw = dict(common)
    if "stream" in okw:
        okw["seq"] = okw.pop("stream")
    want = orc.monte_carlo(deck, m, generations=generations, histories=histories, skip=skip, threads=threads,
                           want_tally=True, **okw, **(oracle_kw or {}))
    return got, want


def _assert_identical(got, want):
    assert np.array_equal(got.tally_fixed, want.tally_fixed)
    for name in ("k", "k_fund", "flux", "assembly_average", "fission_source"):
        assert np.array_equal(bits(getattr(got, name)), bits(getattr(want, name))), name
    for c in ("histories", "collisions", "flights", "leaks", "truncated"):
        assert got.counters[c] == want.counters[c], c


def test_device_logf_bit_exact():
    rng = np.random.default_rng(0)
    k = np.r_[rng.integers(0, 1 << 23, 2_000_000), 0, 1, 2, (1 << 23) - 1, (1 << 22), (1 << 22) - 1, 5931641, 5931642, 5931643]
    x = ((k.astype(f32) + f32(0.5)) * f32(2.0 ** -23)).astype(f32)
    got = nb.dev_logf(x)
    L = orc.lib()
    want = np.array([L.oracle_logf_f(float(v)) for v in x[:200_000]], f32)
    assert np.array_equal(bits(got[:200_000]), bits(want))
    tail = np.array([L.oracle_logf_f(float(v)) for v in x[-9:]], f32)
    assert np.array_equal(bits(got[-9:]), bits(tail))
    assert np.max(np.abs(got.astype(np.float64) - np.log(x.astype(np.float64))) / np.spacing(np.abs(np.log(x.astype(np.float64))).astype(f32))) < 1.0


def test_walk_division_is_ieee_exact():
    """The hoisted-reciprocal division of the walk loop equals the device's (and numpy's) correctly
    rounded division on the operand domain of the path: |mu| in [2^-23, 1], t = 0 or 1e-14 <= |t| <= 64."""
    rng = np.random.default_rng(3)
    n = 4_000_000
    mu = (2.0 * ((rng.integers(0, 1 << 23, n).astype(f32) + f32(0.5)) * f32(2.0 ** -23)) - 1.0).astype(f32)
    t = np.concatenate([
        rng.uniform(-0.12, 0.12, n // 4), rng.uniform(-64, 64, n // 4),
        (10.0 ** rng.uniform(-14, 1.8, n // 4)) * rng.choice([-1, 1], n // 4), np.zeros(n // 8),
        rng.choice([0.1175, 0.0805, -0.1175, -0.0805, 0.11749995, 0.08049999], n // 8),
    ]).astype(f32)
    mu[:8] = f32([2.0 ** -23, -(2.0 ** -23), 1 - 2.0 ** -23, -(1 - 2.0 ** -23), 2.0 ** -22, 0.5, -0.5, 3 * 2.0 ** -23])
    fast, ieee = nb.dev_div(t, mu)
    # the path only ever scores |t / mu|, so the sign of a zero quotient is immaterial
    bad = np.flatnonzero(bits(np.abs(fast)) != bits(np.abs(ieee)))
    assert bad.size == 0, (bad.size, t[bad[:5]], mu[bad[:5]], fast[bad[:5]], ieee[bad[:5]])
    assert np.array_equal(bits(np.abs(ieee)), bits(np.abs((t / mu).astype(f32))))


def test_device_pcg32_streams():
    import ctypes as C
    u, xi = nb.dev_pcg32(42, 54, 152917, 0, 6)
    assert [hex(v) for v in u] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    for hid in (1, 12345, 10**7 * 200 - 1, 2**40 + 17):
        u, xi = nb.dev_pcg32(42, 54, 152917, hid, 64)
        st = (C.c_uint64 * 2)()
        orc.lib().oracle_pcg32_state(42, 54, (hid * 152917) % 2**64, st)
        state, inc, out = st[0], st[1], []
        for _ in range(64):
            old = state
            state = (old * 6364136223846793005 + inc) % 2**64
            xs_ = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
            rot = old >> 59
            out.append(((xs_ >> rot) | (xs_ << ((32 - rot) & 31))) & 0xFFFFFFFF)
        assert u.tolist() == out
        assert np.array_equal(xi, ((np.array(out, np.uint64) >> 9).astype(f32) + f32(0.5)) * f32(2.0 ** -23))
        assert np.all((xi > 0) & (xi < 1) & (xi != 0.5))


@pytest.mark.parametrize("case,gen", [("a", 0), ("b", 3), ("c", 0), ("c", 7)])
def test_replay_every_history_bit_exact(case, gen):
    """Per-history collisions / crossings / flights / reflections, final RNG state, cell, x bits, fate, group."""
    v, xs, dx, mesh, fuel = load_case(case)
    H = 100_000
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=gen + 1, histories=H, skip=0) as ctx:
        rec = ctx.trace(gen, 0, H)
        tally, counters = ctx.read_tally()
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=gen + 1, histories=H, skip=0, threads=8, trace_gen=gen, want_tally=True)
    assert np.array_equal(rec, want.trace)
    assert np.array_equal(tally, want.tally_fixed[gen])
    assert counters["crossings"] == int(want.trace[:, 1].sum()) and counters["collisions"] == int(want.trace[:, 0].sum())
    assert rec[:, 0].max() > 100 and rec[:, 8].min() == 1  # long tails exist; every history was absorbed


@pytest.mark.parametrize("case,tracking,H,gens", [("a", "surface", 501, 70), ("c", "surface", 3001, 9),
                                                  ("b", "woodcock", 2000, 11)])
def test_batched_generations_bit_exact(case, tracking, H, gens):
    """nraps_mc_run lets one launch carry several small uniform-source generations (each into its own tally rows);
    every per-generation tally, k and the folded results must equal the oracle's one-generation-at-a-time run, and
    the generation-level API (one generation per launch) must give the same tallies."""
    got, want = _both(case, generations=gens, histories=H, skip=2, tracking_mode=tracking)
    _assert_identical(got, want)
    v, xs, dx, mesh, fuel = load_case(case)
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=2, tracking_mode=tracking) as ctx:
        for gen in (0, gens // 2, gens - 1):
            ctx.transport(gen)
            tally, _ = ctx.read_tally()
            assert np.array_equal(tally, got.tally_fixed[gen])


def test_pipelined_generations_bit_exact():
    """Uniform source: the launches of consecutive generations alternate between two streams and the context's two
    scratch lanes (nraps_mc_select_lane), all-reduce + finalize on a side stream (OverlappedReducer).  k, flux and the
    fission source must equal nraps_mc_run's, whatever overlaps on the device; small generations so that launches
    really are in flight together, and a larger one."""
    import torch

    from nraps_b200.dist import OverlappedReducer

    for case, H, gens in (("c", 2_000, 12), ("a", 300_000, 7)):
        v, xs, dx, mesh, fuel = load_case(case)
        want = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True)  # one stream
        for pipeline in (True, False):
            with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1) as ctx:
                red = OverlappedReducer(ctx, 1, 0, torch.cuda.current_stream(), pipeline=pipeline)
                for gen in range(gens):
                    red.step(gen, 0, H)
                red.drain()
                got = ctx.fetch(torch.cuda.current_stream().cuda_stream)
            for name in ("k", "k_fund", "flux", "assembly_average", "fission_source"):
                assert np.array_equal(bits(getattr(got, name)), bits(getattr(want, name))), (case, pipeline, name)
            assert got.counters["collisions"] == want.counters["collisions"]


def test_large_generations_pipelined_or_batched_bit_exact(monkeypatch):
    """The tail of a persistent launch (the last neutrons finish one by one) is hidden two ways in nraps_mc_run, uniform
    source only: by default the launches of consecutive generations alternate between two streams and two scratch
    lanes; NRAPS_TAIL_BATCH=3 lets one launch carry three generations instead (up to 2^25 histories).  Both, and the
    plain one-stream run (NRAPS_PIPELINE=0), must give the oracle's k, flux and fission source bit for bit; the batched
    run is also compared tally by tally (the pipelined run does not read tallies back: that path is sequential)."""
    H, gens = 3_000_000, 5
    v, xs, dx, mesh, fuel = load_case("c")
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=8, want_tally=True)

    def same(got, tallies):
        for name in ("k", "k_fund", "flux", "assembly_average", "fission_source"):
            assert np.array_equal(bits(getattr(got, name)), bits(getattr(want, name))), name
        assert got.counters["collisions"] == want.counters["collisions"]
        if tallies:
            assert np.array_equal(got.tally_fixed, want.tally_fixed)

    same(nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1), False)          # pipelined
    same(nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, tracking_mode="surface",
                        want_tally=True), True)                                                               # sequential (tally read-back)
    monkeypatch.setenv("NRAPS_PIPELINE", "0")
    same(nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1), False)
    monkeypatch.delenv("NRAPS_PIPELINE")
    monkeypatch.setenv("NRAPS_TAIL_BATCH", "3")
    same(nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True), True)  # 3 + 2 per launch


@pytest.mark.parametrize("case,H,gens,skip", [("a", 100_000, 6, 4), ("b", 150_000, 3, 1), ("c", 150_000, 3, 1)])
def test_results_bit_exact(case, H, gens, skip):
    got, want = _both(case, generations=gens, histories=H, skip=skip)
    _assert_identical(got, want)


def test_csv_output_bytes_identical(tmp_path):
    got, want = _both("c", generations=4, histories=50_000, skip=1)
    v, xs, dx, mesh, fuel = load_case("c")
    nb.plot_solution(got, v.energygroups, 4, len(mesh), float(mesh.mesh_right[-1]), str(tmp_path))
    files = ho.csv_files(want.flux, want.assembly_average, want.fission_source, want.k, want.k_fund, mesh.mesh_right[-1], len(mesh), 4)
    for name, text in files.items():
        assert (tmp_path / name).read_text() == text, name


@pytest.mark.parametrize("mode", ["rust_pre182", "rust_182"])
def test_scatter_probe_orders_bit_exact(mode):
    got, want = _both("c", generations=2, histories=60_000, scatter_mode=mode)
    _assert_identical(got, want)


def test_fixed_stale_xs_switch_bit_exact():
    got, want = _both("c", generations=2, histories=60_000, stale_xs=False)
    _assert_identical(got, want)
    base, _ = _both("c", generations=2, histories=60_000)
    assert float(np.mean(got.k)) < float(np.mean(base.k)) - 0.03  # fixing the stale index lowers k_C by ~0.08


def test_other_seeds_and_strides_bit_exact():
    got, want = _both("b", generations=2, histories=40_000, seed=7, stream=11, stride=100_003)
    _assert_identical(got, want)


def _with_bounds(case, boundl, boundr):
    v, xs, dx, mesh, fuel = load_case(case)
    v.boundl, v.boundr = boundl, boundr
    return v, xs, dx, mesh, fuel


@pytest.mark.parametrize("bl,br", [(0.0, 0.0), (0.5, 1.0), (1.0, 0.0)])
def test_vacuum_and_albedo_boundaries_bit_exact(bl, br):
    args = _with_bounds("a", bl, br)
    got = nb.monte_carlo(*args, 1.0, generations=2, histories=50_000, skip=1, want_tally=True)
    deck, m = oracle_inputs(*args)
    want = orc.monte_carlo(deck, m, generations=2, histories=50_000, skip=1, threads=8, want_tally=True)
    _assert_identical(got, want)
    if bl == 0.0 or br == 0.0:
        assert got.counters["leaks"] > 0


def test_launch_geometry_and_sharding_do_not_change_a_bit():
    v, xs, dx, mesh, fuel = load_case("c")
    H = 70_001
    ref = None
    for kw in [dict(), dict(threads_per_block=256, blocks_per_sm=4, chunk=32), dict(threads_per_block=1024, blocks_per_sm=1, chunk=1000),
               dict(threads_per_block=64, blocks_per_sm=1, chunk=1), dict(walk_cap=3), dict(walk_cap=-1),
               dict(walk_cap=5, spawn_batch=8)]:
        with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=H, skip=1, **kw) as ctx:
            ctx.transport(1)
            tally, counters = ctx.read_tally()
            assert counters["histories"] == H
            if ref is None:
                ref = tally.copy()
                # two shards of the same generation sum to the whole (what multi-GPU relies on)
                ctx.transport(1, 0, 12_345)
                a, _ = ctx.read_tally()
                ctx.transport(1, 12_345, H - 12_345)
                b, _ = ctx.read_tally()
                assert np.array_equal(a + b, ref)
                ctx.transport(1, 500, 0)
                z, cz = ctx.read_tally()
                assert not z.any() and cz["histories"] == 0
            assert np.array_equal(tally, ref), kw


@pytest.mark.parametrize("kw", [dict(), dict(source_mode="fission_bank"), dict(tracking_mode="woodcock", source_mode="fission_bank")])
def test_shards_taken_in_sub_shards_bit_exact(monkeypatch, kw):
    """A shard larger than 2^27 histories is transported in sub-shards (births of a piece, its transport, the next piece)
    so that the birth-record buffer stays bounded.  NRAPS_SUBSHARD shrinks the piece to 7001 histories: tallies, k, bank
    and per-history replay records must not notice."""
    monkeypatch.setenv("NRAPS_SUBSHARD", "7001")
    got, want = _both("c", generations=3, histories=50_000, **kw)
    _assert_identical(got, want)
    assert np.array_equal(got.bank_sizes, want.bank_sizes)
    if not kw:
        v, xs, dx, mesh, fuel = load_case("c")
        with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=30_000, skip=0) as ctx:
            rec = ctx.trace(1, 100, 29_000)
        deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
        ref = orc.monte_carlo(deck, m, generations=2, histories=30_000, skip=0, threads=8, trace_gen=1, hist_begin=100, hist_count=29_000)
        assert np.array_equal(rec, ref.trace)


def test_single_history_and_flight_cap():
    got, want = _both("a", generations=2, histories=1, skip=1)
    _assert_identical(got, want)
    got, want = _both("c", generations=1, histories=5_000, skip=0, max_flights=20)
    _assert_identical(got, want)
    assert got.counters["truncated"] > 0


def test_fine_mesh_bit_exact():
    v, xs, dx, mesh, fuel = load_case("c", mpfr=80, mpwr=40)
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=20_000, skip=1, want_tally=True)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=2, histories=20_000, skip=1, threads=8, want_tally=True)
    _assert_identical(got, want)


@pytest.mark.parametrize("case,mpfr,bl,br,source", [("c", 80, 1.0, 1.0, "uniform_fuel"), ("a", 160, 0.0, 0.5, "uniform_fuel"),
                                                    ("b", 40, 1.0, 0.0, "fission_bank"), ("c", 320, 1.0, 1.0, "uniform_fuel")])
def test_closed_form_strides_do_not_change_a_bit(case, mpfr, bl, br, source):
    """Fine meshes: the surface kernel strides over the surely-crossed cells of a segment in closed form (DESIGN section
    5).  With the strides switched off (walk_cap = -2: cell-by-cell loop only) every tally bin, k and counter must be
    the same -- 2 and 4 groups, reflecting / vacuum / albedo walls, both source modes, tallies in shared memory (N = 2040,
    4080) and in global memory (N = 8160, 16320) -- and equal to the oracle's."""
    v, xs, dx, mesh, fuel = load_case(case, mpfr=mpfr, mpwr=mpfr // 2)
    v.boundl, v.boundr = bl, br
    kw = dict(generations=3, histories=40_000, skip=1, want_tally=True, source_mode=source)
    on = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, **kw)
    off = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, walk_cap=-2, **kw)
    _assert_identical(on, off)
    assert np.array_equal(on.bank_sizes, off.bank_sizes)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=3, histories=40_000, skip=1, threads=8, want_tally=True, source_mode=source)
    _assert_identical(on, want)


def test_statistical_parity_independent_streams():
    """k within 3 sigma combined and per-bin chi-square between the GPU (seed 1) and the oracle (seed 2)."""
    v, xs, dx, mesh, fuel = load_case("c")
    gens, H = 12, 100_000
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True, seed=1, stream=3, stride=152917)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=8, want_tally=True, seed=2, seq=5)
    kg, ko = got.k.astype(np.float64), want.k.astype(np.float64)
    sigma = np.hypot(kg.std(ddof=1), ko.std(ddof=1)) / np.sqrt(gens)
    assert abs(kg.mean() - ko.mean()) < 3 * sigma, (kg.mean(), ko.mean(), sigma)
    tg = got.tally_fixed.astype(np.float64).reshape(gens, -1)
    to = want.tally_fixed.astype(np.float64).reshape(gens, -1)
    var = (tg.var(axis=0, ddof=1) + to.var(axis=0, ddof=1)) / gens
    z2 = (tg.mean(axis=0) - to.mean(axis=0)) ** 2 / var
    chi2, dof = z2.sum(), z2.size
    assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof) * 1.2, (chi2, dof)  # ~t-distributed terms: slightly heavier than chi2


def test_config3_size_properties():
    """BASELINE config 3 at full size (10^7 histories/generation): size-independent properties."""
    v, xs, dx, mesh, fuel = load_case("c")
    H = 10_000_000
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=200, histories=H, skip=1) as ctx:
        ctx.transport(199)
        tally, counters = ctx.read_tally()
        ctx.finalize_generation(199)
        k = ctx.fetch().k[199]
        assert counters["histories"] == H and counters["truncated"] == 0 and counters["leaks"] == 0
        assert abs(counters["collisions"] / H - 29.55) < 0.05
        assert abs(float(k) - 1.8226) < 0.004  # survey anchor +- ~6 sigma of one 10^7 generation
        # determinism: a second pass of the same generation gives the same bins, whatever the scheduling
        ctx.transport(199)
        again, _ = ctx.read_tally()
        assert np.array_equal(tally, again)
        # k is the nu-fission-weighted tally sum / H (k cancels, src/mc_code.rs:346-351)
        nusigf = (xs.nut * xs.sigf).reshape(v.energygroups, v.mattypes)[:, mesh.matid]
        k64 = float((tally.astype(np.float64) * 2.0 ** -28 * nusigf).sum() / H)
        assert abs(k64 - float(k)) < 2e-5 * k64


def test_flux_moments_extension_matches_the_per_generation_tallies():
    """nraps_results.flux_moments (extension): sum and sum of squares over generations >= skip of flux * conversion, checked
    in f64 against the per-generation tallies (which are bit-identical to the oracle's); tolerance = f32 rounding of
    the per-generation term, 2e-6 relative."""
    v, xs, dx, mesh, fuel = load_case("c")
    gens, skip, H = 12, 3, 40_000
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=skip, want_tally=True)
    M = v.mattypes
    L = float(mesh.mesh_right[-1])
    scale = 3565e6 * 36.2 / (200e6 * 1.602176634e-19 * float(xs.nut[0 + M * 1]) * L)  # conversion / k, src/mc_code.rs:353-357
    term = got.tally_fixed[skip:].astype(np.float64) * 2.0 ** -28 / (H * mesh.delta_x.astype(np.float64)) * scale
    assert np.allclose(got.flux_moments[0], term.sum(axis=0), rtol=2e-6, atol=0)
    assert np.allclose(got.flux_moments[1], (term ** 2).sum(axis=0), rtol=2e-6, atol=0)
    fund = 1.0 / (gens - skip + 1)
    assert np.allclose(got.flux, term.sum(axis=0) * fund, rtol=2e-6)
    err = got.flux_std_error(gens, skip)
    n = gens - skip
    want_err = fund * n * term.std(axis=0, ddof=1) / np.sqrt(n)
    ok = want_err > 0
    assert np.allclose(err[ok], want_err[ok], rtol=2e-3)  # f32 rounding of each term against a few-percent spread
    assert 0.002 < np.median(err[ok] / got.flux[ok]) < 0.1  # percent-level noise at 4e4 histories x 9 generations
    # the generation-level API accumulates the same sums
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=skip) as ctx:
        for g in range(gens):
            ctx.transport(g)
            ctx.finalize_generation(g)
        again = ctx.fetch()
    assert np.array_equal(again.flux_moments, got.flux_moments)


def test_memory_pool_reuse_and_trim():
    """Contexts draw their device buffers from a library-owned pool that stays mapped between contexts; results do not
    depend on whether a buffer is fresh or recycled, and nraps_mc_trim gives the memory back."""
    import torch
    v, xs, dx, mesh, fuel = load_case("b")
    first = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=3, histories=200_000, skip=1, want_tally=True)
    nb.trim(0)
    free_after_trim = torch.cuda.mem_get_info(0)[0]
    again = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=3, histories=200_000, skip=1, want_tally=True)
    cached = free_after_trim - torch.cuda.mem_get_info(0)[0]
    assert cached >= 200_000 * 32  # the birth-record buffer of the last run is still mapped ...
    third = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=3, histories=200_000, skip=1, want_tally=True,
                           tracking_mode="surface", threads_per_block=256, blocks_per_sm=4)
    nb.trim(0)
    assert torch.cuda.mem_get_info(0)[0] >= free_after_trim - (8 << 20)  # ... until it is trimmed
    for other in (again, third):
        assert np.array_equal(first.tally_fixed, other.tally_fixed) and np.array_equal(bits(first.k), bits(other.k))


def test_error_codes():
    v, xs, dx, mesh, fuel = load_case("a")
    with pytest.raises(_lib.NrapsError) as e:
        nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=3, histories=10, skip=3)
    assert e.value.code == 2  # skip >= generations: the reference panics at k_fund[skip]
    bad = nb.Mesh(mesh.matid.copy(), mesh.delta_x, mesh.mesh_left, mesh.mesh_right.copy())
    bad.mesh_right[10] += f32(1e-3)
    with pytest.raises(_lib.NrapsError) as e:
        nb.monte_carlo(v, xs, dx, bad, fuel, 1.0, generations=2, histories=10, skip=1)
    assert e.value.code == 3
    xs2 = nb.XSData(**{**xs.__dict__, "inv_sigtr": xs.inv_sigtr * f32(-1)})
    with pytest.raises(_lib.NrapsError) as e:
        nb.monte_carlo(v, xs2, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1)
    assert e.value.code == 4
    with pytest.raises(_lib.NrapsError) as e:
        nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1, device=99)
    assert e.value.code == 6


# ---------------------------------------------------------------- fission_bank source mode (new capability)
@pytest.mark.parametrize("case", ["b", "c"])
def test_fission_bank_mode_bit_exact(case):
    """Power iteration: bank sizes, bank contents (canonical order), entropy and every result equal the oracle's."""
    v, xs, dx, mesh, fuel = load_case(case)
    gens, H = 5, 60_000
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True,
                         source_mode="fission_bank")
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=8, want_tally=True,
                           source_mode="fission_bank")
    _assert_identical(got, want)
    assert np.array_equal(got.bank_sizes, want.bank_sizes) and got.counters["banked"] == want.counters["banked"]
    assert np.allclose(got.entropy, want.entropy, rtol=0, atol=1e-12)
    assert 0.9 * H < got.bank_sizes[-1] < 1.1 * H and got.bank_sizes[0] > 1.5 * H  # first bank is k0-normalised


def test_fission_bank_contents_and_sharding():
    v, xs, dx, mesh, fuel = load_case("c")
    H = 40_000
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=2, histories=H, skip=0, threads=8, source_mode="fission_bank", bank_gen=1)
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=H, skip=0, source_mode="fission_bank") as ctx:
        for gen in range(2):
            ctx.transport(gen)
            ctx.finalize_generation(gen)
            ctx.bank_compact(gen)
            if gen == 1:
                bank = ctx.read_bank()
            ctx.bank_advance(gen)
        res = ctx.fetch()
    assert np.array_equal(bank, want.bank_sites)
    assert np.array_equal(res.bank_sizes, want.bank_sizes)
    cells = (bank >> np.uint64(32)).astype(np.int64)
    assert np.isin(mesh.matid[cells], [0, 1]).all()  # sites only where nu*Sigma_f > 0


def test_fission_bank_shifts_k_like_the_survey_probe():
    """true source iteration raises k_C by ~2.4 % over the reference's flat fuel source."""
    v, xs, dx, mesh, fuel = load_case("c")
    flat = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=14, histories=200_000, skip=1)
    bank = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=14, histories=200_000, skip=1, source_mode="fission_bank")
    ratio = float(bank.k[6:].mean() / flat.k[6:].mean())
    assert 1.018 < ratio < 1.030, ratio
    assert bank.entropy[0] > bank.entropy[-1] > 7.5  # source settles from flat towards the fundamental mode


def _dist_worker(rank, world, port, out_dir, mode):
    import os

    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    v, xs, dx, mesh, fuel = load_case("c")
    res = nb.monte_carlo_distributed(v, xs, dx, mesh, fuel, 1.0, generations=4, histories=50_001, skip=1, device=rank,
                                     source_mode=mode)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), k=res.k, flux=res.flux, bank=res.bank_sizes, ent=res.entropy,
             coll=res.counters["collisions"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["uniform_fuel", "fission_bank"])
def test_two_gpus_equal_one_gpu_bit_for_bit(tmp_path, mode):
    """History sharding + NCCL all-reduce (uniform source: overlapped with the next generation; fission bank: sites
    loaded from the peers' banks over NVLink, nothing gathered): identical to the single-GPU run on every rank."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    mp.spawn(_dist_worker, args=(2, 29650 + (os.getpid() % 200), str(tmp_path), mode), nprocs=2, join=True)
    v, xs, dx, mesh, fuel = load_case("c")
    one = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=4, histories=50_001, skip=1, source_mode=mode)
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(bits(z["k"]), bits(one.k)) and np.array_equal(bits(z["flux"]), bits(one.flux))
        assert np.array_equal(z["bank"], one.bank_sizes) and np.allclose(z["ent"], one.entropy, atol=1e-12, rtol=0)
        assert int(z["coll"]) == one.counters["collisions"]


# ---------------------------------------------------------------- Woodcock delta tracking (north-star tracking mode)
@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_woodcock_replay_and_results_bit_exact(case):
    v, xs, dx, mesh, fuel = load_case(case)
    H = 100_000
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=3, histories=H, skip=0, tracking_mode="woodcock") as ctx:
        rec = ctx.trace(2, 0, H)
        tally, counters = ctx.read_tally()
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=3, histories=H, skip=0, threads=8, trace_gen=2, want_tally=True,
                           tracking_mode="woodcock")
    assert np.array_equal(rec, want.trace)
    assert np.array_equal(tally, want.tally_fixed[2])
    got, want = _both(case, generations=4, histories=80_000, tracking_mode="woodcock")
    _assert_identical(got, want)


def test_woodcock_variants_bit_exact():
    for kw in (dict(stale_xs=False), dict(scatter_mode="rust_pre182"), dict(source_mode="fission_bank")):
        got, want = _both("c", generations=3, histories=50_000, tracking_mode="woodcock", **kw)
        _assert_identical(got, want)
        assert np.array_equal(got.bank_sizes, want.bank_sizes)
    args = _with_bounds("a", 0.0, 0.5)
    got = nb.monte_carlo(*args, 1.0, generations=2, histories=50_000, skip=1, want_tally=True, tracking_mode="woodcock")
    deck, m = oracle_inputs(*args)
    want = orc.monte_carlo(deck, m, generations=2, histories=50_000, skip=1, threads=8, want_tally=True, tracking_mode="woodcock")
    _assert_identical(got, want)
    assert got.counters["leaks"] > 0 and got.counters["reflections"] == want.counters["reflections"] > 0
    v, xs, dx, mesh, fuel = load_case("c", mpfr=80, mpwr=40)  # fine mesh: cost independent of N
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=50_000, skip=1, want_tally=True, tracking_mode="woodcock")
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=2, histories=50_000, skip=1, threads=8, want_tally=True, tracking_mode="woodcock")
    _assert_identical(got, want)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_woodcock_agrees_with_reference_tracking_statistically(case):
    """north-star criterion: k within 3 sigma combined and per-bin chi-square against the surface-tracking
    (reference-semantics) solver, independent estimators on the same physics."""
    v, xs, dx, mesh, fuel = load_case(case)
    gens, H = 16, 200_000
    w = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True, tracking_mode="woodcock")
    s = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=gens, histories=H, skip=1, want_tally=True, seed=9, stream=9, stride=152917)
    kw, ks = w.k.astype(np.float64), s.k.astype(np.float64)
    sigma = np.hypot(kw.std(ddof=1), ks.std(ddof=1)) / np.sqrt(gens)
    assert abs(kw.mean() - ks.mean()) < 3 * sigma, (kw.mean(), ks.mean(), sigma)
    tw = w.tally_fixed.astype(np.float64).reshape(gens, -1)
    ts = s.tally_fixed.astype(np.float64).reshape(gens, -1)
    var = (tw.var(axis=0, ddof=1) + ts.var(axis=0, ddof=1)) / gens
    z2 = (tw.mean(axis=0) - ts.mean(axis=0)) ** 2 / var
    assert abs(z2.sum() - z2.size) < 6 * np.sqrt(2 * z2.size) * 1.2, (z2.sum(), z2.size)


# ---------------------------------------------------------------- event-based SoA-bank pipeline (alternative kernel variant)
@pytest.mark.parametrize("case", ["a", "c"])
def test_event_pipeline_equals_fused_woodcock_and_oracle(case):
    """source -> {advance, collide, compact}* over an HBM particle bank gives the same bins as the fused kernel."""
    got, want = _both(case, generations=3, histories=70_000, tracking_mode="woodcock", gpu_kw=dict(kernel_variant="event"))
    _assert_identical(got, want)
    v, xs, dx, mesh, fuel = load_case(case)
    with pytest.raises(_lib.NrapsError) as e:  # the event variant exists for Woodcock tracking only
        nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1, kernel_variant="event")
    assert e.value.code == 7


def test_event_pipeline_flight_cap_truncates_like_the_fused_kernel():
    """The event variant counts flights in 20 bits of a packed word: a small cap must cut the same histories at the same
    point as the fused kernel and the oracle (a cap it cannot count to is refused at creation, tests/test_host.py)."""
    got, want = _both("c", generations=2, histories=20_000, tracking_mode="woodcock", max_flights=7,
                      gpu_kw=dict(kernel_variant="event"))
    _assert_identical(got, want)
    assert got.counters["truncated"] > 0


@pytest.mark.parametrize("kw", [dict(), dict(tracking_mode="woodcock"), dict(tracking_mode="woodcock", source_mode="fission_bank")])
def test_analytic_k_infinity_deck_a_on_gpu(kw):
    """k = nu*Sigma_f / Sigma_a = 1.26 exactly for deck A once the stale-index quirk is off (see the CPU twin)."""
    v, xs, dx, mesh, fuel = load_case("a")
    r = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=12, histories=1_000_000, skip=1, stale_xs=False, **kw)
    k = r.k[1:].astype(np.float64)
    assert abs(k.mean() - 1.26) < 4 * k.std(ddof=1) / np.sqrt(len(k)) + 2e-5, (kw, k.mean(), k.std(ddof=1))


def test_nraps_driver_binary_writes_reference_csv_files(tmp_path):
    """The C++ `nraps` driver (deck in, three CSV files out) is the drop-in for the reference binary's pipeline."""
    import subprocess

    from tests.util import DECKS, ROOT

    exe = os.path.join(ROOT, "nraps_b200", "lib", "nraps")
    out = subprocess.run([exe, DECKS["a"], "--out", str(tmp_path), "--generations", "5", "--histories", "30000", "--skip", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("running MC code\n")  # src/mc_code.rs:292
    v, xs, dx, mesh, fuel = load_case("a")
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=5, histories=30000, skip=2, threads=8)
    files = ho.csv_files(want.flux, want.assembly_average, want.fission_source, want.k, want.k_fund, mesh.mesh_right[-1], len(mesh), 5)
    for name, text in files.items():
        assert (tmp_path / name).read_text() == text, name


# ---------------------------------------------------------------- shapes the shipped decks do not exercise
@pytest.mark.parametrize("M,G,pins,mpfr,mpwr,bl,br", [
    (2, 3, [1, 0, 1], 3, 2, 1.0, 1.0),              # generic-G kernel, fuel touching both walls, no moderator material
    (3, 5, [2, 0, 2, 1, 2, 0, 2], 5, 4, 1.0, 0.0),  # five groups, vacuum on the right
    (5, 8, [4, 0, 3, 1, 2, 0, 4], 4, 6, 0.7, 1.0),  # maximum G, five materials, albedo 0.7 on the left
    (2, 2, [0], 1, 0, 1.0, 1.0),                    # one single cell: N = 1 (no water cells at all)
    (3, 4, [2, 1, 2], 64, 2, 0.0, 0.0),             # one long fuel run (64 cells) between vacuum walls
])
@pytest.mark.parametrize("tracking", ["surface", "woodcock"])
def test_synthetic_shapes_bit_exact(M, G, pins, mpfr, mpwr, bl, br, tracking):
    from tests.util import synthetic_case

    args = synthetic_case(M, G, pins, mpfr, mpwr, seed=M * 10 + G, boundl=bl, boundr=br)
    for source in ("uniform_fuel", "fission_bank"):
        got = nb.monte_carlo(*args, 1.0, generations=3, histories=30_000, skip=1, want_tally=True, tracking_mode=tracking,
                             source_mode=source)
        deck, m = oracle_inputs(*args)
        want = orc.monte_carlo(deck, m, generations=3, histories=30_000, skip=1, threads=8, want_tally=True,
                               tracking_mode=tracking, source_mode=source)
        _assert_identical(got, want)
        assert np.array_equal(got.bank_sizes, want.bank_sizes)
        assert got.counters["reflections"] == want.counters["reflections"] or tracking == "surface"
        assert got.counters["truncated"] == 0 and np.isfinite(got.k).all() and got.k[-1] > 0


@pytest.mark.parametrize("extra", [[], ["--source", "fission_bank"], ["--tracking", "woodcock", "--source", "fission_bank"]])
def test_native_nccl_driver_two_gpus_equals_one(tmp_path, extra):
    """`nraps --gpus 2` (single process, one host thread per GPU, ncclAllReduce inside the C ABI library, fission bank
    read in place through peer access) writes byte-identical CSV files to the single-GPU run."""
    import subprocess

    import torch

    from tests.util import DECKS, ROOT

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = os.path.join(ROOT, "nraps_b200", "lib", "nraps")
    outs = []
    for gpus in (1, 2):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        run = subprocess.run([exe, DECKS["c"], "--out", str(d), "--generations", "4", "--histories", "60001", "--gpus", str(gpus)] + extra,
                             capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, run.stderr
        outs.append({n: (d / n).read_bytes() for n in ("k_eff.csv", "interface.csv", "vars.csv")})
    assert outs[0] == outs[1]


@pytest.mark.parametrize("tracking", ["surface", "woodcock"])
def test_mesh_larger_than_shared_memory_bit_exact(tracking):
    """N = 8160 cells x 4 groups needs 261 KB of tally bins: more than one SM's shared memory.  The kernels then
    read the mesh through L1/L2 and score straight into the global 64-bit bins (BIG mode); results are unchanged."""
    v, xs, dx, mesh, fuel = load_case("c", mpfr=160, mpwr=80)
    assert len(mesh) == 8160
    for source in ("uniform_fuel", "fission_bank"):
        got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=20_000, skip=1, want_tally=True,
                             tracking_mode=tracking, source_mode=source)
        deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
        want = orc.monte_carlo(deck, m, generations=2, histories=20_000, skip=1, threads=8, want_tally=True,
                               tracking_mode=tracking, source_mode=source)
        _assert_identical(got, want)
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1, tracking_mode=tracking) as ctx:
        assert ctx.launch_info()["smem_bytes"] < 48 * 1024
    with pytest.raises(_lib.NrapsError) as e:  # the event pipeline keeps its tally in shared memory only
        nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1, tracking_mode="woodcock", kernel_variant="event")
    assert e.value.code == 5
