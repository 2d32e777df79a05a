"""Run a few generations of one configuration (for ncu captures): python tools/run_generation.py [--tracking T] [--variant V] [--histories H] [--gens G] [--fine]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nraps_b200 as nb  # noqa: E402
from tests.util import load_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tracking", default="surface")
ap.add_argument("--variant", default="fused")
ap.add_argument("--source", default="uniform_fuel")
ap.add_argument("--histories", type=int, default=10_000_000)
ap.add_argument("--gens", type=int, default=2)
ap.add_argument("--fine", action="store_true")
a = ap.parse_args()
args = load_case("c", mpfr=80, mpwr=40) if a.fine else load_case("c")
r = nb.monte_carlo(*args, 1.0, generations=a.gens, histories=a.histories, skip=min(1, a.gens - 1), tracking_mode=a.tracking,
                   kernel_variant=a.variant, source_mode=a.source)
print("k", r.k.tolist(), "histories/s", a.histories * a.gens / r.seconds_device, r.counters)
