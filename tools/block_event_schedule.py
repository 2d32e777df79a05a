"""Schedule statistics of the block-level event pipeline from its CPU-thread emulation (tests/emul; no GPU).

For one emulated block: rounds, entries of each event list per round, and what a warp (32 consecutive list entries)
would see in the walk phase -- sum of crossings against 32 x the longest walk of the chunk, the lane utilisation of the
walk loop that ncu reports as 12-18 of 32 lanes for the lane kernel.  Statistics depend on the slot count, not on how
many CPU threads emulate the block.

    python tools/block_event_schedule.py [--case c] [--histories 40000] [--slots 1536] [--walk-cap 8]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_block_event_emul import LIB, SCATTER, _problem  # noqa: E402
from tests.util import load_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c")
    ap.add_argument("--histories", type=int, default=40000)
    ap.add_argument("--slots", type=int, default=1536)
    ap.add_argument("--threads", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--walk-cap", type=int, default=8)
    ap.add_argument("--mpfr", type=int, default=0)
    ap.add_argument("--mpwr", type=int, default=0)
    a = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True, stdin=subprocess.DEVNULL)
    L = C.CDLL(LIB)
    args = load_case(a.case, a.mpfr, a.mpwr) if a.mpfr else load_case(a.case)
    p, keep, deck, m = _problem(args, a.histories, 1)
    tally = np.zeros(p.G * p.N, np.uint64)
    counters = np.zeros(8, np.uint64)
    st = np.zeros(12, np.uint64)
    u64p = C.POINTER(C.c_uint64)
    L.bev_emul_generation.restype = C.c_int
    rc = L.bev_emul_generation(C.byref(p), C.c_uint64(0), C.c_uint64(0), C.c_uint64(a.histories), C.c_uint64(42), C.c_uint64(54),
                               C.c_uint64(152917), C.c_int32(SCATTER["single_xi"]), C.c_int32(1), C.c_uint32(a.walk_cap), C.c_uint32(1 << 24),
                               C.c_uint32(1), C.c_uint32(a.threads), C.c_uint32(a.slots), C.c_uint32(a.chunk), tally.ctypes.data_as(u64p),
                               counters.ctypes.data_as(u64p), st.ctypes.data_as(u64p))
    assert rc == 0
    rounds, coll, dead, fly, e0, e1, x0, x1, c0, c1, s0, s1 = (int(v) for v in st)
    H = a.histories
    print(f"deck {a.case}: {H} histories through one block of {a.slots} slots (walk cap {a.walk_cap})")
    print(f"  rounds {rounds} ({rounds * a.slots / H:.1f} slot-rounds per history); collisions/history {int(counters[1]) / H:.2f}, flights/history {int(counters[3]) / H:.2f}")
    print(f"  phase AB per round: collide {coll / rounds:.0f}, adopt {dead / rounds:.0f}, flight-only {fly / rounds:.0f} entries "
          f"({(coll + dead + fly) / rounds / a.slots:.2f} of the bank)")
    for cls, (e, x, c, s) in enumerate(((e0, x0, c0, s0), (e1, x1, c1, s1))):
        if not e:
            continue
        print(f"  walk class {cls}: {e / rounds:.0f} entries per round, {x / e:.2f} crossings per walk, warp chunks {c}, "
              f"crossing slots {s} -> {x / (32 * s):.3f} of the lanes busy in the walk loop ({32 * x / (32 * s):.1f} of 32)")
    tot_x, tot_s = x0 + x1, s0 + s1
    print(f"  walk loop overall: {tot_x / H:.1f} crossings per history, {tot_s / H:.2f} warp crossing slots per history "
          f"({32 * tot_x / (32 * tot_s):.1f} of 32 lanes busy; the lane kernel: 23.4 slots per history, ncu r1r)")


if __name__ == "__main__":
    main()
