"""Host side of the path on CPU: C-ABI symbols, deck parser, mesh_gen, CSV
output -- product C++ (through libnraps_b200.so) against the numpy restatement
in oracle/host_oracle.py and the structural fixtures of SURVEY section 8c."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nraps_b200 as nb
from nraps_b200 import _lib
from oracle import host_oracle as ho
from oracle import oracle as orc
from tests.util import DECKS, ROOT, bits, load_case, oracle_inputs

f32 = np.float32


def test_library_exports_every_declared_symbol():
    declared = set()
    for h in ("nraps_mc.h", "nraps_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(nraps_[a-z0-9_]+)\s*\(", src))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    hdr = open(os.path.join(ROOT, "include", "nraps_mc.h")).read()
    assert L.nraps_abi_version() == int(re.search(r"#define NRAPS_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION


def test_mc_entry_points_fail_loudly_without_arguments_or_gpu():
    L = _lib.lib()
    assert L.nraps_mc_run(None, None, None) == 1  # NRAPS_ERR_NULL, before any CUDA call
    h = C.c_void_p()
    assert L.nraps_mc_create(None, None, C.byref(h)) == 1
    assert L.nraps_strerror(5).decode().startswith("tables exceed")


@pytest.mark.parametrize("case", "abc")
def test_process_input_matches_numpy_restatement(case):
    v, xs, pins, dx, sol, solver = nb.process_input(DECKS[case])
    d = ho.process_input(DECKS[case])
    for name in ("analk", "mattypes", "energygroups", "generations", "histories", "skip", "numass", "numrods", "mpfr", "mpwr"):
        assert getattr(v, name) == getattr(d, name), name
    for name in ("roddia", "rodpitch", "boundl", "boundr"):
        assert f32(getattr(v, name)) == getattr(d, name), name
    assert f32(dx.fuel) == d.dx_fuel and f32(dx.water) == d.dx_water
    for a, b in [(xs.sigt, d.sigt), (xs.sigs, d.sigs), (xs.mu, d.mu), (xs.siga, d.siga), (xs.sigf, d.sigf),
                 (xs.nut, d.nut), (xs.chit, d.chit), (xs.scat_matrix, d.scat), (xs.inv_sigtr, d.inv_sigtr)]:
        assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(pins, d.matid) and sol == d.solution and solver == d.solver
    G, M = v.energygroups, v.mattypes
    assert len(xs.sigt) == M * G and len(xs.scat_matrix) == M * G * G and len(pins) == 70


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference decks only exist in the build container")
@pytest.mark.parametrize("case", "abc")
def test_fixture_decks_carry_the_reference_payload(case):
    ref = nb.process_input(f"/root/reference/TestCase{case.upper()}.txt")
    fix = nb.process_input(DECKS[case])
    assert ref[0] == fix[0] and ref[3] == fix[3] and ref[4:] == fix[4:]
    for name in ("sigt", "sigs", "mu", "siga", "sigf", "nut", "chit", "scat_matrix", "inv_sigtr"):
        assert np.array_equal(bits(getattr(ref[1], name)), bits(getattr(fix[1], name)))
    assert np.array_equal(ref[2], fix[2])


def test_parser_quirks(tmp_path):
    """src/process_input.rs:44-83: '#' kills a pending key=value; the byte after a comment line is
    never inspected; repeated keys concatenate; keys match on (last two chars, length) only."""
    base = open(DECKS["a"]).read()
    deck = tmp_path / "quirks.txt"
    text = base.replace("Histories = 100000\n", "Histories = 5 # trailing comment discards this line\nHistories = 777\n")
    # the first byte after a comment line is never inspected, so this '#' does not start a comment and the
    # 4-char key "#kip" lands in the Skip slot ("ip", 4)
    text = text.replace("Skip = 4\n", "# comment line\n#kip = 4\n")
    text = text.replace("Generations = 100\n", "xxxxxxxxxns = 100\n")      # 11 chars ending in "ns" == generations
    deck.write_text(text)
    v, xs, pins, dx, _, _ = nb.process_input(str(deck))
    d = ho.process_input(str(deck))
    assert v.histories == 777 == d.histories
    assert v.skip == 4 == d.skip
    assert v.generations == 100 == d.generations
    assert len(xs.scat_matrix) == 16 and len(pins) == 70  # four Scat lines, two MatID lines concatenated


def test_process_input_errors(tmp_path):
    with pytest.raises(_lib.NrapsError) as e:
        nb.process_input(str(tmp_path / "missing.txt"))
    assert e.value.code == 9
    bad = tmp_path / "bad.txt"
    bad.write_text(open(DECKS["a"]).read().replace("MPFR = 8", "MPFR = eight"))
    with pytest.raises(_lib.NrapsError):
        nb.process_input(str(bad))


@pytest.mark.parametrize("case,N,NF,L", [("a", 408, 272, 42.908), ("b", 400, 254, 41.598), ("c", 408, 272, 42.908)])
def test_mesh_gen_structure(case, N, NF, L):
    v, xs, dx, mesh, fuel = load_case(case)
    assert len(mesh) == N and len(fuel) == NF and abs(float(mesh.mesh_right[-1]) - L) < 2e-3
    assert mesh.mesh_left[0] == 0.0
    assert np.array_equal(bits(mesh.mesh_right[:-1]), bits(mesh.mesh_left[1:]))  # edges shared bit for bit
    runs = np.flatnonzero(np.diff(mesh.matid.astype(int))) + 1
    assert len(runs) + 1 == 69  # 69 material runs in all three decks
    assert list(mesh.matid[:2]) == [2, 2] and list(mesh.matid[-2:]) == [2, 2]  # edge water runs of 2 cells
    d = ho.process_input(DECKS[case])
    m = ho.mesh_gen(d.matid, d.mpfr, d.mpwr, d.numass, d.dx_fuel, d.dx_water)
    assert np.array_equal(mesh.matid, m[0]) and np.array_equal(fuel, m[4])
    for a, b in [(mesh.delta_x, m[1]), (mesh.mesh_left, m[2]), (mesh.mesh_right, m[3])]:
        assert np.array_equal(bits(a), bits(b))


def test_mesh_gen_centre_trim_with_control_rods():
    """in deck B the cut lands inside assembly 2: ... MOX x8, H2O x6, MOX x6, H2O x4 ..."""
    _, _, _, mesh, _ = load_case("b")
    m = mesh.matid
    change = np.flatnonzero(np.diff(m.astype(int))) + 1
    runs = [(int(m[s]), int(e - s)) for s, e in zip(np.r_[0, change], np.r_[change, len(m)])]
    mid = next(i for i, (mat, n) in enumerate(runs) if (mat, n) == (2, 6))
    assert runs[mid - 1] == (1, 8) and runs[mid + 1] == (1, 6) and runs[mid + 2] == (2, 4)
    assert sum(1 for mat, _ in runs if mat == 3) == 2  # two control-rod pins, meshed at water width (Q10)


def test_mesh_gen_fine_mesh_shape():
    _, _, _, mesh, fuel = load_case("c", mpfr=80, mpwr=40)
    assert len(mesh) == 4080 and len(fuel) == 2720 and abs(float(mesh.mesh_right[-1]) - 42.907) < 2e-3


def test_rust_float_display():
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 1.5102, 3.3e-7, 1e-10, 4.7727519e19, 123456789.0, 16777216.0, 1e-38, 3.4e38,
            float("inf"), float("-inf"), float("nan"), 42.908089, 0.30000001192]
    vals += list(rng.uniform(-2, 2, 300)) + list(10.0 ** rng.uniform(-30, 30, 300)) + list(rng.integers(0, 10**9, 100).astype(float))
    for v in vals:
        assert nb.format_f32(v) == ho.rust_f32_display(v), v
    for v in [42.90808868408203, 0.1, 1e22, 1e-7, 123456789.125, float(f32(41.598076))]:
        assert nb.format_f64(v) == ho.rust_f64_display(v), v
    assert nb.format_f32(4.7727519e19) == "47727517000000000000" and nb.format_f32(1.0) == "1" and nb.format_f32(0.0) == "0"


def test_plot_solution_files(tmp_path):
    rng = np.random.default_rng(1)
    G, N, gens = 4, 37, 9
    r = nb.SolutionResults(
        flux=(rng.random((G, N)) * 1e19).astype(f32), assembly_average=rng.random((G, N)).astype(f32),
        fission_source=np.r_[rng.random(N - 3), 0, 0, 0].astype(f32), k=rng.random(gens).astype(f32) + 1,
        k_fund=np.r_[0, rng.random(gens - 1) + 1].astype(f32),
    )
    L = float(f32(42.908089))
    nb.plot_solution(r, G, gens, N, L, str(tmp_path))
    want = ho.csv_files(r.flux, r.assembly_average, r.fission_source, r.k, r.k_fund, L, N, gens)
    for name, text in want.items():
        assert (tmp_path / name).read_text() == text, name
    rows = (tmp_path / "interface.csv").read_text().splitlines()
    assert len(rows) == 2 * G + 1 and all(len(row.split(",")) == N for row in rows)  # src/plot_solution.rs:43-52


def test_reference_plot_reader_reads_our_csv(tmp_path):
    """the CSV files must be the ones upstream's plot.py reads (plot.py:5-58).  The fixture holds what the
    READING part of the reference's own plot.py -- executed unmodified on files written by nraps_plot_solution, in the
    build container (tools/make_plot_reader_golden.py) -- bound to its variables: the live 2-group block
    (plot.py:27-37) and, for G = 4, the 4-group block upstream keeps commented out right below it (plot.py:39-56).
    Here: the writer still produces those very files, and every array plot.py read is the field we meant it to get."""
    import json
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_plot_reader_golden", os.path.join(ROOT, "tools", "make_plot_reader_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "plot_reader.json")))
    assert [c["name"] for c in fixture["cases"]] == [c["name"] for c in gen.CASES]
    for c in fixture["cases"]:
        G, N, gens = c["G"], c["N"], c["gens"]
        r = gen.results_for(G, N, gens, c["seed"])
        d = tmp_path / c["name"]
        d.mkdir()
        nb.plot_solution(r, G, gens, N, c["length"], str(d))
        for name, text in c["files"].items():
            assert (d / name).read_text() == text, (c["name"], name)
        read = c["read"]
        assert read["length"] == c["length"] and read["meshed"] == N and read["generations"] == gens
        same = lambda got, want: np.array_equal(np.asarray(got, np.float64).astype(f32), want)  # noqa: E731 -- shortest round-trip text
        assert same(read["k"], r.k) and same(read["k_fund"], r.k_fund) and same(read["fission"], r.fission_source)
        for g in range(G):
            assert same(read[f"flux{g}"], r.flux[g]) and same(read[f"average{g}"], r.assembly_average[g]), (c["name"], g)
        assert f"flux{G}" not in read and f"average{G}" not in read


def test_average_assembly_and_k_fund_match_oracle():
    rng = np.random.default_rng(2)
    fp = C.POINTER(C.c_float)
    for G, N, numass in [(2, 408, 2), (4, 400, 2), (3, 41, 3)]:
        flux = (rng.random((G, N)) * 1e18).astype(f32)
        a, b = np.zeros_like(flux), np.zeros_like(flux)
        _lib.lib().nraps_average_assembly(flux.ctypes.data_as(fp), G, N, numass, a.ctypes.data_as(fp))
        orc.lib().oracle_average_assembly(flux.ctypes.data_as(fp), G, N, numass, b.ctypes.data_as(fp))
        assert np.array_equal(bits(a), bits(b))
    for gens, skip in [(100, 1), (100, 4), (7, 6), (5, 0)]:
        k = (rng.random(gens) + 1).astype(f32)
        a, b = np.zeros(gens, f32), np.zeros(gens, f32)
        _lib.lib().nraps_k_fund(k.ctypes.data_as(fp), gens, skip, a.ctypes.data_as(fp))
        orc.lib().oracle_k_fund(k.ctypes.data_as(fp), gens, skip, b.ctypes.data_as(fp))
        assert np.array_equal(bits(a), bits(b))
        assert np.all(a[:skip] == 0) and a[skip] == k[skip]


def test_nccl_driver_library_exports_its_entry_point():
    """Loaded in a child process: the library links the system libnccl.so.2, and a process that has it mapped can no
    longer import torch (whose bundled, newer NCCL would be shadowed by the copy already loaded)."""
    import subprocess
    import sys

    src = open(os.path.join(ROOT, "include", "nraps_multi.h")).read()
    assert "nraps_mc_run_multi" in src
    path = os.path.join(ROOT, "nraps_b200", "lib", "libnraps_b200_nccl.so")
    child = ("import ctypes, sys\n"
             "try:\n"
             f"    L = ctypes.CDLL({path!r})\n"
             "except OSError as e:\n"
             "    print('SKIP', e); sys.exit(0)\n"
             "print('RC', L.nraps_mc_run_multi(None, None, None, 2, None))\n")
    out = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, check=True).stdout
    if out.startswith("SKIP"):  # libnccl.so.2 not installed on this host
        pytest.skip(out)
    assert out.split() == ["RC", "1"]  # NRAPS_ERR_NULL before any CUDA / NCCL call


# ---- diffusion solver (src/discrete.rs), SURVEY 8(f) rank 4: cross-check of the Monte Carlo path -------------------
# parity unpinned: the reference holds no test for this solver and inverts with nalgebra (f32 LU); the oracle inverts
# with LAPACK and the library eliminates the tridiagonal system in f64, so agreement is to the rounding noise of an f32
# dense inverse, stated below, not bit for bit.
DIFFUSION_FLUX_RTOL = 2e-4
DIFFUSION_K_RTOL = 5e-6


@pytest.mark.parametrize("case,bl,br", [("a", None, None), ("b", None, None), ("c", None, None), ("c", 0.0, 1.0),
                                        ("b", 0.5, 0.0), ("a", 0.3, 0.7)])
def test_diffusion_solver_matches_the_restatement(case, bl, br):
    import dataclasses

    from oracle import diffusion_oracle as dif
    v, xs, dx, mesh, fuel = load_case(case)
    if bl is not None:
        v = dataclasses.replace(v, boundl=bl, boundr=br)
    got = nb.nalgebra_method(xs, mesh, v.energygroups, v.mattypes, v.boundl, v.boundr, v.numass)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = dif.nalgebra_method(deck, m)
    # the stop test compares f32 noise with 1e-5 / 1e-6: the last iteration may fall on either side
    assert abs(got.counters["iterations"] - want.iterations) <= 1
    assert abs(got.k[0] / want.k - 1) < DIFFUSION_K_RTOL
    assert np.abs(got.flux / want.flux - 1).max() < DIFFUSION_FLUX_RTOL
    assert np.abs(got.assembly_average / want.assembly_average - 1).max() < DIFFUSION_FLUX_RTOL
    assert got.fission_source.size == 0 and got.k_fund.size == 0 and got.k.size == 1  # src/discrete.rs:349-355
    if case == "a" and bl is None:
        # deck A is one fuel material between reflecting walls with mu = 0: k = nu*Sigma_f / Sigma_a = 1.26 for any
        # flux shape -- the same analytic anchor the Monte Carlo path meets with the stale-index quirk off
        assert abs(got.k[0] - 1.26) < 5e-5


def test_diffusion_driver_and_csv(tmp_path):
    import subprocess

    exe = os.path.join(ROOT, "nraps_b200", "lib", "nraps")
    out = subprocess.run([exe, DECKS["c"], "--out", str(tmp_path), "--solution", "diffusion"],
                         check=True, capture_output=True, text=True)
    v, xs, dx, mesh, fuel = load_case("c")
    got = nb.nalgebra_method(xs, mesh, v.energygroups, v.mattypes, v.boundl, v.boundr, v.numass)
    assert out.stdout.splitlines()[0] == f"{got.k[0]:.10f}"  # println!("{:.10}", k), src/discrete.rs:347
    rows = open(tmp_path / "interface.csv").read().splitlines()
    assert len(rows) == 2 * v.energygroups and not (tmp_path / "k_eff.csv").exists()
    assert rows[0].split(",")[0] == nb.format_f32(float(got.flux[0, 0]))
    nb.plot_solution(got, v.energygroups, v.generations, len(mesh), float(mesh.mesh_right[-1]), str(tmp_path))
    assert open(tmp_path / "interface.csv").read().splitlines() == rows


def _mutate_deck(text: str, rng) -> str:
    """Random, format-preserving and format-breaking edits of a deck: the scanner's quirks (src/process_input.rs:44-83)
    decide what each one means, and both parsers must decide the same."""
    lines = text.split("\n")
    for _ in range(int(rng.integers(1, 7))):
        i = int(rng.integers(0, len(lines)))
        kind = int(rng.integers(0, 12))
        ln = lines[i]
        if kind == 0:
            lines.insert(i, "# " + "x" * int(rng.integers(0, 5)))                 # comment line (eats the next line's first byte)
        elif kind == 1:
            lines[i] = ln + "  # trailing"                                          # kills a pending key = value
        elif kind == 2:
            lines.insert(i, "")                                                     # blank line
        elif kind == 3 and "=" in ln:
            k, v = ln.split("=", 1)
            lines[i] = k.upper() + "=" + v if rng.integers(0, 2) else "  " + k.lower() + " =   " + v + "  "
        elif kind == 4 and "=" in ln and len(ln.split("=", 1)[1].split()) > 3:
            k, v = ln.split("=", 1)                                                 # split a list over a repeated key
            toks = v.split()
            cut = int(rng.integers(1, len(toks)))
            lines[i] = k + "= " + " ".join(toks[:cut])
            lines.insert(i + 1, k + "= " + " ".join(toks[cut:]))
        elif kind == 5:
            lines[i] = ln + "\r"                                                    # CRLF
        elif kind == 6 and "=" in ln:
            lines[i] = ln.replace(" =", "=", 1)                                     # key loses its last char (name_end = pos - 1)
        elif kind == 7 and "=" in ln:
            lines[i] = ln + " = 3"                                                  # second '=' moves key end and value start
        elif kind == 8:
            lines.insert(i, "NoSuchKey = 12 13")                                    # junk slot
        elif kind == 9 and "=" in ln:
            lines[i] = ln.split("=", 1)[0] + "="                                    # empty value
        elif kind == 10:
            lines.insert(i, "just some words without the sign")
        elif kind == 11 and "=" in ln:
            lines.insert(i, ln)                                                     # duplicated key: lists double, scalars break
    out = "\n".join(lines)
    if rng.integers(0, 6) == 0:
        out = out.rstrip("\n")                                                      # last line without a newline is never seen
    return out


def test_parser_fuzz_against_numpy_restatement(tmp_path):
    """300 randomly edited decks: the product parser and the numpy restatement of src/process_input.rs either both
    reject the deck or agree on every field, bit for bit."""
    rng = np.random.default_rng(2024)
    accepted = rejected = 0
    for n in range(300):
        base = open(DECKS["abc"[n % 3]]).read()
        text = _mutate_deck(base, rng)
        if text.rstrip("\n").rsplit("\n", 1)[-1].lstrip().startswith("#") and not text.endswith("\n"):
            continue  # a comment on an unterminated last line indexes past the buffer upstream (skip_line, :6-10)
        path = tmp_path / f"fuzz{n}.txt"
        path.write_bytes(text.encode())
        try:
            d = ho.process_input(str(path))
            # a pin list that is not u8, or a table shorter than SigT, is a panic upstream
            ok = all(0 <= int(m) <= 255 for m in d.matid)
            ok = ok and all(len(getattr(d, t)) >= len(d.sigt) for t in ("sigs", "mu", "siga", "sigf", "nut", "chit"))
            ok = ok and all(0 <= getattr(d, f) <= 255 for f in ("analk", "mattypes", "energygroups", "numass", "numrods", "solution"))
            ok = ok and all(getattr(d, f) >= 0 for f in ("generations", "histories", "skip", "mpfr", "mpwr"))
        except (ValueError, IndexError, OverflowError, ZeroDivisionError):
            ok = False
        if not ok:
            with pytest.raises(_lib.NrapsError):
                nb.process_input(str(path))
            rejected += 1
            continue
        v, xs, pins, dx, sol, solver = nb.process_input(str(path))
        for name in ("analk", "mattypes", "energygroups", "generations", "histories", "skip", "numass", "numrods", "mpfr", "mpwr"):
            assert getattr(v, name) == getattr(d, name), (n, name)
        for name in ("roddia", "rodpitch", "boundl", "boundr"):
            assert bits(getattr(v, name)) == bits(getattr(d, name)), (n, name)
        assert bits(dx.fuel) == bits(d.dx_fuel) and bits(dx.water) == bits(d.dx_water), n
        for a, b in [(xs.sigt, d.sigt), (xs.sigs, d.sigs), (xs.mu, d.mu), (xs.siga, d.siga), (xs.sigf, d.sigf),
                     (xs.nut, d.nut), (xs.chit, d.chit), (xs.inv_sigtr, d.inv_sigtr)]:
            assert len(a) == len(d.sigt) and np.array_equal(bits(a), bits(b[:len(a)])), n  # tables are handed out as [n_xs]
        assert np.array_equal(bits(xs.scat_matrix), bits(d.scat)), n
        assert np.array_equal(pins, d.matid) and sol == d.solution and solver == d.solver, n
        accepted += 1
    assert accepted > 60 and rejected > 60, (accepted, rejected)


def test_mesh_gen_fuzz_against_numpy_restatement():
    """Random pin lists (fuel, water, control rods), cells per pin and assembly counts: same cells, same fuel list, same
    f32 edges bit for bit -- or the same refusal (src/main.rs:85-143 panics where these return a shape error)."""
    rng = np.random.default_rng(77)
    built = refused = 0
    for _ in range(1500):
        numass, mpfr, mpwr = int(rng.integers(1, 4)), int(rng.integers(1, 10)), int(rng.integers(1, 9))
        pins = rng.integers(0, 4, int(rng.integers(3, 40))).astype(np.uint8)
        v = nb.Variables(analk=1, mattypes=4, energygroups=2, generations=2, histories=1, skip=1, numass=numass, numrods=len(pins),
                         roddia=0.94, rodpitch=0.322, mpfr=mpfr, mpwr=mpwr, boundl=1.0, boundr=1.0)
        dx = nb.DeltaX(fuel=float(f32(0.94) / f32(mpfr)), water=float(f32(0.322) / f32(mpwr)))
        try:
            m = ho.mesh_gen(pins, mpfr, mpwr, numass, f32(dx.fuel), f32(dx.water))
            ok = len(m[0]) > 0
        except IndexError:
            ok = False
        if not ok:
            with pytest.raises(_lib.NrapsError):
                nb.mesh_gen(pins, v, dx)
            refused += 1
            continue
        mesh, fuel = nb.mesh_gen(pins, v, dx)
        assert np.array_equal(mesh.matid, m[0]) and np.array_equal(fuel, m[4])
        for a, b in [(mesh.delta_x, m[1]), (mesh.mesh_left, m[2]), (mesh.mesh_right, m[3])]:
            assert np.array_equal(bits(a), bits(b))
        built += 1
    assert built > 1000 and refused > 0


def test_rust_float_display_random_bit_patterns():
    """Shortest round-trip digits in positional notation for arbitrary binary32 / binary64 bit patterns (subnormals,
    extremes, NaN payloads), against the restatement of Rust's Display (src/plot_solution.rs:14-34 uses to_string())."""
    rng = np.random.default_rng(9)
    pat = rng.integers(0, 1 << 32, 20000, dtype=np.uint64).astype(np.uint32)
    pat[:8] = [1, 0x007FFFFF, 0x00800000, 0x7F7FFFFF, 0x80000001, 0x3F800001, 0x4B800000, 0x4B7FFFFF]
    for v in pat.view(np.float32):
        assert nb.format_f32(float(v)) == ho.rust_f32_display(float(v)), v
    pat64 = rng.integers(0, 1 << 63, 5000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 5000, dtype=np.uint64)
    for v in pat64.view(np.float64):
        assert nb.format_f64(float(v)) == ho.rust_f64_display(float(v)), v


def test_validation_codes_come_before_any_cuda_call_and_there_is_no_fallback():
    """Every malformed problem is refused with its own code on a host without a GPU; a well-formed one gets as far as
    the CUDA runtime and fails with NRAPS_ERR_CUDA when no device exists -- the product never computes on the CPU."""
    v, xs, dx, mesh, fuel = load_case("a")

    def code(**kw):
        args = dict(variables=v, xsdata=xs, delta_x=dx, meshid=mesh, fuel_indices=fuel, generations=2, histories=10, skip=1)
        opts = {k: kw.pop(k) for k in list(kw) if k in ("scatter_mode", "tracking_mode", "kernel_variant", "source_mode", "bank_cap", "max_flights")}
        args.update(kw)
        with pytest.raises(_lib.NrapsError) as e:
            nb.monte_carlo(args["variables"], args["xsdata"], args["delta_x"], args["meshid"], args["fuel_indices"], args.get("k_new", 1.0),
                           generations=args["generations"], histories=args["histories"], skip=args["skip"], **opts)
        return e.value.code

    assert code(generations=3, skip=3) == 2                                   # k_fund[skip] out of range upstream
    assert code(histories=0) == 2
    assert code(k_new=0.0) == 2 and code(k_new=float("nan")) == 2
    bad = nb.Mesh(mesh.matid.copy(), mesh.delta_x, mesh.mesh_left, mesh.mesh_right.copy())
    bad.mesh_right[10] += f32(1e-3)
    assert code(meshid=bad) == 3                                              # right[i] != left[i+1]
    bad = nb.Mesh(mesh.matid.copy(), mesh.delta_x, mesh.mesh_left, mesh.mesh_right)
    bad.matid[5] = 9
    assert code(meshid=bad) == 3                                              # matid >= M
    assert code(fuel_indices=np.r_[fuel[:-1], np.uint64(len(mesh))]) == 3     # fuel index >= N
    assert code(xsdata=nb.XSData(**{**xs.__dict__, "inv_sigtr": xs.inv_sigtr * f32(-1)})) == 4
    assert code(xsdata=nb.XSData(**{**xs.__dict__, "inv_sigtr": xs.inv_sigtr * f32(np.inf)})) == 4
    assert code(kernel_variant="event") == 7                                  # the event pipeline is Woodcock-only
    assert code(kernel_variant="block_event") == 7                            # measured slower on a B200: compiled out of the product
    assert code(kernel_variant="event", tracking_mode="woodcock", max_flights=1 << 21) == 7  # its flight counter is 20 bits wide
    for b in (-0.1, 1.5, float("nan")):                                       # albedo outside [0, 1]: mu' = -mu * b leaves [-1, 1]
        assert code(variables=nb.Variables(**{**v.__dict__, "boundl": b})) == 2
        assert code(variables=nb.Variables(**{**v.__dict__, "boundr": b})) == 2
    assert code(bank_cap=300) == 7
    v1 = nb.Variables(**{**v.__dict__, "energygroups": 1})
    assert code(variables=v1) == 2                                            # G >= 2: nut[M*1] (src/mc_code.rs:356)
    if not os.path.exists("/dev/nvidiactl"):
        assert code() == 6                                                    # well-formed: NRAPS_ERR_CUDA, no CPU path
    L = _lib.lib()
    assert L.nraps_mc_select_lane(None, 0) == 1 and L.nraps_mc_transport(None, 0, 0, 1, None) == 1   # NRAPS_ERR_NULL: no context, no work


def test_python_mirror_refuses_arrays_shorter_than_the_abi_reads():
    """nraps_problem carries no array lengths (include/nraps_mc.h): the Python mirror checks them before the call."""
    v, xs, dx, mesh, fuel = load_case("c")
    short = nb.XSData(**{**xs.__dict__, "siga": xs.siga[:-1]})
    with pytest.raises(ValueError, match="siga"):
        nb.monte_carlo(v, short, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1)
    short = nb.XSData(**{**xs.__dict__, "scat_matrix": xs.scat_matrix[:60]})
    with pytest.raises(ValueError, match="scat_matrix"):
        nb.monte_carlo(v, short, dx, mesh, fuel, 1.0, generations=2, histories=10, skip=1)
    ragged = nb.Mesh(mesh.matid, mesh.delta_x, mesh.mesh_left[:-1], mesh.mesh_right)
    with pytest.raises(ValueError, match="left"):
        nb.monte_carlo(v, xs, dx, ragged, fuel, 1.0, generations=2, histories=10, skip=1)
    with pytest.raises(ValueError, match="sigt"):
        nb.nalgebra_method(nb.XSData(**{**xs.__dict__, "sigt": xs.sigt[:3]}), mesh, v.energygroups, v.mattypes, 1.0, 1.0, v.numass)


def test_bindings_mirror_the_header_field_for_field(tmp_path):
    """The three mirrors of include/nraps_mc.h -- the ctypes structures, the oracle's copy of nraps_problem and the
    #[repr(C)] structs of the Rust shim (rust/src/mc_code.rs, which no toolchain here can compile) -- must list the
    same fields in the same order with the same C types, and the ctypes layout must equal the C compiler's."""
    import re
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "nraps_mc.h")).read()

    def c_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(const )?(\w+) (.*)", decl)
            const, base, names = bool(m.group(1)), m.group(2), m.group(3)
            for name in names.split(","):
                name = name.strip()
                ptr = name.startswith("*")
                name = name.lstrip("*")
                arr = re.match(r"(\w+)\[(\w+)\]", name)
                if arr:
                    out.append((arr.group(1), base + "[]"))
                else:
                    out.append((name, ("const " if const and ptr else "") + base + ("*" if ptr else "")))
        return out

    rust = open(os.path.join(root, "rust", "src", "mc_code.rs")).read()
    rust_types = {"u32": "uint32_t", "u64": "uint64_t", "i32": "int32_t", "f32": "float", "f64": "double",
                  "*const f32": "const float*", "*const u8": "const uint8_t*", "*const u64": "const uint64_t*",
                  "*mut f32": "float*", "*mut u64": "uint64_t*", "*mut f64": "double*", "[u64; 8]": "uint64_t[]"}

    def rust_fields(struct):
        body = re.search(r"struct %s \{(.*?)\n\}" % struct, rust, re.S).group(1)
        return [(n.strip(), rust_types[t.strip()]) for n, t in re.findall(r"(\w+):\s*([^,]+?),", body)]

    ctypes_types = {C.c_uint32: "uint32_t", C.c_uint64: "uint64_t", C.c_int32: "int32_t", C.c_float: "float", C.c_double: "double"}

    def ct_fields(cls, drop_const=True):
        out = []
        for name, t in cls._fields_:
            if t in ctypes_types:
                out.append((name, ctypes_types[t]))
            elif hasattr(t, "_length_"):
                out.append((name, ctypes_types[t._type_] + "[]"))
            else:
                out.append((name, ctypes_types.get(t._type_, "uint8_t") + "*"))
        return out

    strip = lambda fields: [(n.lower(), t.replace("const ", "")) for n, t in fields]  # noqa: E731
    for cname, rname, pycls in (("nraps_problem", "NrapsProblem", _lib.Problem), ("nraps_options", "NrapsOptions", _lib.Options),
                                ("nraps_results", "NrapsResults", _lib.Results)):
        want = c_fields(cname)
        assert [(n.lower(), t) for n, t in rust_fields(rname)] == [(n.lower(), t) for n, t in want], rname
        assert strip(ct_fields(pycls)) == strip(want), pycls
    from oracle import oracle as orc
    assert strip(ct_fields(orc.Problem)) == strip(c_fields("nraps_problem"))
    # sizes and offsets as the C compiler lays them out
    probe = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nraps_mc.h"', "int main(void) {"]
    for cname in ("nraps_problem", "nraps_options", "nraps_results"):
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for name, _ in c_fields(cname):
            lines.append(f'printf("{cname}.{name} %zu\\n", offsetof({cname}, {name}));')
    lines.append("return 0; }")
    probe.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(probe), "-o", str(exe)], check=True, stdin=subprocess.DEVNULL)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, pycls in (("nraps_problem", _lib.Problem), ("nraps_options", _lib.Options), ("nraps_results", _lib.Results)):
        assert int(got[cname]) == C.sizeof(pycls)
        for (name, _), (pyname, _t) in zip(c_fields(cname), pycls._fields_):
            assert int(got[f"{cname}.{name}"]) == getattr(pycls, pyname).offset, (cname, name)


def test_walk_segments_of_the_surface_kernel():
    """nraps_walk_segments: the table the surface kernel walks by (include/nraps_host.h).  Against a numpy restatement on
    the three decks and two refinements: segments tile the mesh, never span a material boundary, hold cells of one
    binary32 width only and are maximal; the stop edges are the segment's outer neighbours, clipped so that the walk
    never enters a boundary cell; and the counts the design quotes (deck C: 76 segments for 69 material runs)."""
    import ctypes as C

    L = _lib.lib()
    for case, mpfr, mpwr in (("a", None, None), ("b", None, None), ("c", None, None), ("c", 80, 40), ("b", 17, 5)):
        v, xs, dx, mesh, fuel = load_case(case, mpfr=mpfr, mpwr=mpwr) if mpfr else load_case(case)
        N = len(mesh)
        stops, wbits, nseg = np.zeros(N, np.uint32), np.zeros(N, np.uint32), C.c_uint32(0)
        u32p = C.POINTER(C.c_uint32)
        rc = L.nraps_walk_segments(mesh.matid.ctypes.data_as(_lib._u8p), mesh.mesh_left.ctypes.data_as(_lib._fp),
                                   mesh.mesh_right.ctypes.data_as(_lib._fp), N, stops.ctypes.data_as(u32p), wbits.ctypes.data_as(u32p),
                                   C.byref(nseg))
        assert rc == 0
        w = (mesh.mesh_right - mesh.mesh_left).astype(f32).view(np.uint32)
        assert np.array_equal(wbits, w)
        cut = np.r_[True, (mesh.matid[1:] != mesh.matid[:-1]) | (w[1:] != w[:-1])]   # a segment starts here
        starts = np.flatnonzero(cut)
        ends = np.r_[starts[1:], N]
        assert nseg.value == len(starts)
        for a, b in zip(starts, ends):
            want = max(a - 1, 0) | ((min(b, N - 1) + 1) << 16)
            assert (stops[a:b] == want).all(), (case, a, b)
        runs = 1 + int((mesh.matid[1:] != mesh.matid[:-1]).sum())
        assert runs <= nseg.value <= runs + 12   # a run splits only where the position crosses a power of two
        if case == "c" and mpfr is None:
            assert (runs, nseg.value) == (69, 76)
    assert L.nraps_walk_segments(None, None, None, 4, None, None, None) == 1
