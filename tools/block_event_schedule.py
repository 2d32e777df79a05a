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
    st = np.zeros(8, np.uint64)
    u64p = C.POINTER(C.c_uint64)
    L.bev_emul_generation.restype = C.c_int
    rc = L.bev_emul_generation(C.byref(p), C.c_uint64(0), C.c_uint64(0), C.c_uint64(a.histories), C.c_uint64(42), C.c_uint64(54),
                               C.c_uint64(152917), C.c_int32(SCATTER["single_xi"]), C.c_int32(1), C.c_uint32(a.walk_cap), C.c_uint32(1 << 24),
                               C.c_uint32(1), C.c_uint32(a.threads), C.c_uint32(a.slots), C.c_uint32(a.chunk), tally.ctypes.data_as(u64p),
                               counters.ctypes.data_as(u64p), st.ctypes.data_as(u64p))
    assert rc == 0
    rounds, coll, births, go, entries, crossings, chunks, slots = (int(v) for v in st)
    H = a.histories
    print(f"deck {a.case}: {H} histories through one block of {a.slots} records (walk cap {a.walk_cap}, "
          f"walk classes {'by run length' if not os.environ.get('BEV_EMUL_CLASS_T', '0') != '0' else 'by predicted crossings >= ' + os.environ['BEV_EMUL_CLASS_T']})")
    print(f"  rounds {rounds} ({entries / H:.1f} record visits of the walk phase per history); collisions/history {int(counters[1]) / H:.2f}, "
          f"flights/history {int(counters[3]) / H:.2f}")
    print(f"  phase AB per round: collide {coll / rounds:.0f}, births outside a collision {births / rounds:.0f}, flight-only or resumed {go / rounds:.0f} "
          f"entries ({(coll + births + go) / rounds / a.slots:.2f} of the bank)")
    print(f"  phase C per round: {entries / rounds:.0f} walks, {crossings / entries:.2f} crossings per walk")
    print(f"  walk loop: {crossings / H:.1f} crossings per history in {slots / H:.2f} warp crossing slots per history "
          f"({crossings / slots:.1f} of 32 lanes busy" + ("; the lane kernel on deck C as shipped: 23.4 slots per history, ncu r1r)" if not a.mpfr and a.case == "c" else ")"))


if __name__ == "__main__":
    main()
