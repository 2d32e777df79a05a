"""Where the time of one monte_carlo() call goes after other contexts have used the pool (as in bench.py): repeated calls,
NRAPS_TIMING split, with and without the three-generations-per-launch batching.  Throwaway diagnosis."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nraps_b200 as nb  # noqa: E402
from nraps_b200.dist import OverlappedReducer  # noqa: E402
from tests.util import load_case  # noqa: E402

args = load_case("c")
H = 10_000_000
torch.zeros(1, device="cuda")
os.environ["NRAPS_TIMING"] = "1"


def pipelined_context():
    with nb.MonteCarloContext(*args, 1.0, generations=8, histories=H, skip=1) as ctx:
        red = OverlappedReducer(ctx, 1, 0, torch.cuda.current_stream())
        for g in range(8):
            red.step(g, 0, H)
        red.drain()
        ctx.fetch(torch.cuda.current_stream().cuda_stream)


for tb in ("3", "1", "3"):
    os.environ["NRAPS_TAIL_BATCH"] = tb
    for rep in range(3):
        if rep == 1:
            pipelined_context()
        t0 = time.perf_counter()
        r = nb.monte_carlo(*args, 1.0, generations=20, histories=H, skip=1)
        t1 = time.perf_counter()
        print(f"tail_batch {tb} rep {rep}{' (after a two-lane context)' if rep == 1 else ''}: monte_carlo() {1e3 * (t1 - t0):.1f} ms, device {1e3 * r.seconds_device:.1f} ms", flush=True)
