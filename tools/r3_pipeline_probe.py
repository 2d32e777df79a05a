"""Probe: does the next generation's launch fill the tail of the previous one?  Two contexts (own buffers each) take the
even and the odd generations of config 3 on two streams; against one context on one stream.  Throwaway measurement."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nraps_b200 as nb  # noqa: E402
from tests.util import load_case  # noqa: E402

H, K, W = 10_000_000, 20, 4
args = load_case("c")
dev = "cuda:0"
torch.cuda.set_device(0)


def run(two_streams: bool, flush_l2: bool):
    ctxs = [nb.MonteCarloContext(*args, 1.0, generations=W + K, histories=H, skip=1, device=0) for _ in range(2 if two_streams else 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in ctxs]
    tallies = [torch.zeros(c.n_words, dtype=torch.int64, device=dev) for c in ctxs]
    for c, t in zip(ctxs, tallies):
        c.use_tally_tensor(t)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(g):
        i = g % len(ctxs)
        with torch.cuda.stream(streams[i]):
            if flush_l2:
                flush.zero_()
            ctxs[i].transport(g, 0, H, streams[i].cuda_stream)
            ctxs[i].finalize_generation(g, streams[i].cuda_stream)

    for g in range(W):
        step(g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g in range(W, W + K):
        step(g)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ks = [c.fetch(streams[i].cuda_stream).k for i, c in enumerate(ctxs)]
    for c in ctxs:
        c.close()
    return H * K / dt, dt / K * 1e3, ks


for two in (False, True, False, True):
    for fl in (True, False):
        r, ms, ks = run(two, fl)
        print(f"two_streams={two} l2_flush={fl}: {r:.4e} histories/s, {ms:.3f} ms per generation", flush=True)
