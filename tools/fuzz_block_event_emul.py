"""Fuzz of the block-event kernel's per-thread body (CPU-thread emulation, tests/emul) against the oracle on random
slab problems and random block geometries: same problem generator as tools/fuzz_restatements.py (G = 2..8, up to five
materials, random pin layouts, refinements, albedos, master streams, probe orders, the stale-index switch).

    python tools/fuzz_block_event_emul.py --cases 300 --seed 1
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tests.test_block_event_emul as t  # noqa: E402
from tests.util import synthetic_case  # noqa: E402
from tools.fuzz_restatements import random_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True, stdin=subprocess.DEVNULL)
    L = C.CDLL(t.LIB)
    L.bev_emul_generation.argtypes = [C.POINTER(t.orc.Problem)] + [C.c_uint64] * 6 + [C.c_int32] * 2 + [C.c_uint32] * 6 + [C.POINTER(C.c_uint64)] * 3
    rng = np.random.default_rng(a.seed)
    ran = bad = 0
    t0 = time.time()
    while ran < a.cases:
        c = random_case(rng)
        try:
            args = synthetic_case(c["M"], c["G"], c["pins"], c["mpfr"], c["mpwr"], seed=c["seed"], boundl=c["bl"], boundr=c["br"], numass=1)
        except Exception:
            continue  # the reference panics on this layout
        if len(args[4]) == 0:
            continue
        geo = dict(blocks=int(rng.integers(1, 4)), threads=int(rng.choice([1, 3, 8, 32])), slots=int(rng.choice([1, 5, 32, 100])),
                   chunk=int(rng.choice([1, 7, 64])), walk_cap=int(rng.choice([1, 3, 8, 24])))
        os.environ["BEV_EMUL_CLASS_T"] = str(int(rng.choice([0, 0, 2, 5])))
        try:
            t._run(L, args, H=int(rng.integers(200, 900)), gen=int(rng.integers(0, 3)), scatter_mode=c["scatter_mode"], stale_xs=c["stale_xs"],
                   seed=c["rng_seed"], seq=c["rng_seq"], stride=c["stride"], **geo)
        except AssertionError as e:
            bad += 1
            print(f"DIFFERS: {e}\n  problem {c}\n  geometry {geo} class_T {os.environ['BEV_EMUL_CLASS_T']}", flush=True)
        ran += 1
    print(f"{ran} random problems x random block geometries, {bad} with differences, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
