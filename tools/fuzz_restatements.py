"""Differential fuzz of the two CPU restatements (TEST INFRASTRUCTURE, CPU only).

oracle/oracle_mc.c (the checker of the CUDA path) against oracle/restatement_py.py on random slab problems: random
material / group counts, pin layouts, mesh refinements, wall albedos, every switch of SURVEY 9-B, both tracking modes
and both source modes.  Any difference in a per-history record, a tally bin, k, flux or a bank is printed with the
case's parameters, which `tests/test_oracle_restatement.py::test_random_problems` can then pin.

    python tools/fuzz_restatements.py --cases 200 --seed 1
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oracle as orc  # noqa: E402
from oracle import restatement_py as rp  # noqa: E402
from tests.util import bits, oracle_inputs, synthetic_case  # noqa: E402


def random_case(rng: np.random.Generator, fine: bool = False) -> dict:
    """Parameters of one problem; everything needed to rebuild it is in the returned dict.  fine = meshes with tens of
    cells per pin (the surface kernel's closed-form strides and its larger shared-memory images)."""
    M = int(rng.integers(2, 6))
    G = int(rng.integers(2, 9))
    n_pins = int(rng.integers(1, 8))
    pins = [int(p) for p in rng.integers(0, M, n_pins)]
    if not any(p < 2 for p in pins):  # materials 0 and 1 are the fuels: at least one pin must be able to fission
        pins[int(rng.integers(0, n_pins))] = int(rng.integers(0, 2))
    mpfr = int(rng.choice([1, 2, 3, 5, 8, 17]))
    mpwr = int(rng.choice([0, 1, 2, 4, 7])) if n_pins == 1 else int(rng.choice([1, 2, 4, 7]))
    if fine:
        mpfr = int(rng.choice([24, 40, 64, 100, 257]))
        mpwr = int(rng.choice([12, 20, 33, 50, 128]))
    walls = [0.0, 0.3, 1.0]
    tracking = str(rng.choice(["surface", "surface", "woodcock"]))
    source = str(rng.choice(["uniform_fuel", "uniform_fuel", "fission_bank"]))
    return dict(
        M=M, G=G, pins=pins, mpfr=mpfr, mpwr=mpwr, seed=int(rng.integers(0, 1 << 30)),
        bl=float(rng.choice(walls)), br=float(rng.choice(walls)), numass=int(rng.choice([1, 1, 2, 3])),
        threads=int(rng.choice([1, 1, 2, 5])), inclusive=bool(rng.integers(0, 4) == 0), f32_tally=bool(rng.integers(0, 3) == 0),
        scatter_mode=str(rng.choice(["single_xi", "rust_pre182", "rust_182"])), stale_xs=bool(rng.integers(0, 2)),
        tracking=tracking, source=source, H=int(rng.integers(20, 70)), gens=int(rng.integers(2, 4)),
        rng_seed=int(rng.integers(1, 1 << 40)), rng_seq=int(rng.integers(0, 1 << 20)), stride=int(rng.choice([1000, 152917, 40000])),
    )


class Unrunnable(Exception):
    """The reference itself panics on this input (mesh_gen trims past the ends, or no fuel cell is left to spawn in)."""


def run_case(c: dict) -> list[str]:
    """Both restatements on case `c`; returns the names of whatever differs (empty = identical)."""
    try:
        v, xs, dx, mesh, fuel = synthetic_case(c["M"], c["G"], c["pins"], c["mpfr"], c["mpwr"], seed=c["seed"], boundl=c["bl"],
                                               boundr=c["br"], numass=c["numass"])
    except Exception as e:  # src/main.rs:107-108: drain / truncate past the ends of the cell vector
        raise Unrunnable(str(e)) from e
    if len(fuel) == 0:      # src/mc_code.rs:46: gen_range(0..0) panics
        raise Unrunnable("no fuel cell left after the end trim")
    if len(mesh.matid) < c["numass"]:  # average_assembly would divide by a zero span (src/mc_code.rs:262)
        raise Unrunnable("fewer cells than assemblies")
    gens, H = c["gens"], c["H"]
    v.generations, v.histories, v.skip = gens, H, 1
    variables, xsdata, dxf, meshid, fi = rp.from_product_inputs(v, xs, dx, mesh, fuel)
    extended = c["tracking"] != "surface" or c["source"] != "uniform_fuel"
    # the reference's own worker split, inclusive ranges and per-worker f32 tallies exist for its own algorithm only
    threads = 1 if extended else c["threads"]
    inclusive = False if extended else c["inclusive"]
    exact = True if extended else not c["f32_tally"]
    sw = rp.Switches(scatter_mode=c["scatter_mode"], stale_xs=c["stale_xs"], seed=c["rng_seed"], seq=c["rng_seq"], stride=c["stride"],
                     threads=threads, inclusive_ranges=inclusive)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    okw = dict(scatter_mode=c["scatter_mode"], stale_xs=c["stale_xs"], seed=c["rng_seed"], seq=c["rng_seq"], stride=c["stride"])
    trace_gen = None if inclusive else gens - 1
    if extended:
        got = rp.monte_carlo_extended(variables, xsdata, dxf, meshid, fi, 1.0, sw, tracking=c["tracking"], source=c["source"],
                                      trace_gen=gens - 1)
        want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=1, want_tally=True, trace_gen=gens - 1,
                               tracking_mode=c["tracking"], source_mode=c["source"], bank_gen=gens - 1, **okw)
    else:
        got = rp.monte_carlo(variables, xsdata, dxf, meshid, fi, 1.0, sw, exact_tally=exact, trace_gen=trace_gen)
        want = orc.monte_carlo(deck, m, generations=gens, histories=H, skip=1, threads=threads, want_tally=exact, trace_gen=trace_gen,
                               tally_mode="fixed64" if exact else "f32_per_worker", inclusive_ranges=inclusive, **okw)
    bad = []
    if (extended or trace_gen is not None) and not np.array_equal(got["trace"], want.trace):
        bad.append("trace")
    if exact and not np.array_equal(got["tally_fixed"], want.tally_fixed):
        bad.append("tally_fixed")
    names = ("k", "flux", "fission_source") if extended else ("k", "k_fund", "flux", "fission_source", "assembly_average")
    for name in names:
        if not np.array_equal(bits(got[name]), bits(getattr(want, name))):
            bad.append(name)
    if c["source"] == "fission_bank":
        if not np.array_equal(got["bank_sizes"], want.bank_sizes):
            bad.append("bank_sizes")
        elif not np.array_equal(np.array(got["banks"][-1], np.uint64), want.bank_sites):
            bad.append("bank_sites")
        if not np.allclose(got["entropy"], want.entropy, rtol=0, atol=1e-12):
            bad.append("entropy")
    return bad


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    failures = skipped = 0
    for i in range(a.cases):
        c = random_case(rng)
        try:
            bad = run_case(c)
        except Unrunnable:
            skipped += 1
            continue
        except Exception as e:  # a crash in either restatement is a finding too
            bad = [f"exception {type(e).__name__}: {e}"]
        if bad:
            failures += 1
            print(f"case {i}: DIFFERS in {bad}\n  {c}", flush=True)
    print(f"{a.cases} random problems ({skipped} skipped: the reference panics on them), {failures} with differences, "
          f"{time.time() - t0:.0f} s")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
