"""Meshes around the shared-memory limit of one SM: histories/s of the surface and the Woodcock kernel (one GPU).

    python tools/r2_bigmesh.py [--histories H] [--refine 80 160 ...]    (refine = MPFR; MPWR = MPFR / 2; N = 51 * MPFR)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nraps_b200 as nb  # noqa: E402
from tests.util import load_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--histories", type=int, default=10_000_000)
ap.add_argument("--refine", type=int, nargs="*", default=[8, 80, 120, 160, 320])
a = ap.parse_args()
for mpfr in a.refine:
    args = load_case("c", mpfr=mpfr, mpwr=mpfr // 2)
    for tracking in ("surface", "woodcock"):
        r = nb.monte_carlo(*args, 1.0, generations=4, histories=a.histories, skip=1, tracking_mode=tracking)
        with nb.MonteCarloContext(*args, 1.0, generations=1, histories=1000, skip=0, tracking_mode=tracking) as c:
            c.transport(0)
            info = c.launch_info()
        print(f"N={len(args[3])} {tracking:9s} {a.histories * 4 / r.seconds_device:.4e} histories/s  k={r.k[1:].mean():.5f}  launch={info}", flush=True)
