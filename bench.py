#!/usr/bin/env python
"""Throughput of the Monte Carlo transport path: neutron histories/s per generation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config4]

A "step" is one generation: source -> all events -> tally flush -> (all-reduce) -> k.
Workload at N = 1 is BASELINE config 3 (TestCaseC cross sections and geometry,
G = 4, N = 408 cells, 10^7 histories per generation, uniform-fuel source, PCG32
seed 42 / stream 54 / stride 152917).  For N > 1 (launched by torchrun, one rank
per GPU) every GPU keeps 10^7 histories per generation (weak scaling); the only
collective is the per-generation int64 all-reduce of the tally buffer.

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of
the reference algorithm (oracle/, all host threads) on bounded samples of the
same workload; it is the only other place this file touches oracle/ besides the
cpu_baseline leg.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "neutron histories/s per generation"
UNIT = "histories/s"
RECORD_BYTES = 24  # SURVEY 8d bank record: x, mu, cell, packed groups/flags, rng state


_json_fd = None


def reserve_stdout():
    """stdout carries the one JSON line and nothing else: from here on file descriptor 1 points at stderr, so that
    whatever a library prints there (NCCL's version banner when the box sets NCCL_DEBUG=VERSION, for one) cannot land
    in front of the line; emit() writes to the original stdout."""
    global _json_fd
    if _json_fd is None:
        sys.stdout.flush()
        _json_fd = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _json_fd is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_json_fd, data)


def workload(name: str):
    from tests.util import load_case

    if name == "config3":
        args = load_case("c")
        desc = "config3: TestCaseC XS+geometry (G=4, M=4, N=408 cells), 1e7 histories/generation/GPU, uniform_fuel source, PCG32 seed 42/stream 54/stride 152917"
        per_gpu = 10_000_000
    elif name == "config4":
        args = load_case("c", mpfr=80, mpwr=40)
        desc = "config4: TestCaseC XS, fine mesh MPFR=80/MPWR=40 (G=4, N=4080 cells), 1e8 histories/generation total (strong scaling)"
        per_gpu = 100_000_000
    elif name == "config5":
        args = load_case("c")
        desc = ("config5: TestCaseC XS+geometry (G=4, N=408), 1.25e8 histories/generation/GPU, fission_bank source "
                "(power iteration), NCCL all-gather of the bank each generation (weak scaling)")
        per_gpu = 125_000_000
    else:
        raise SystemExit(f"unknown workload {name}")
    return args, desc, per_gpu


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s; MEASURED_PEAKS.json absent)"


def ncu_facts(kernel: str) -> dict:
    """Numbers taken from the committed ncu captures of this kernel (profiles/traffic.json), not measured live."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)[kernel]
    except (OSError, KeyError, ValueError):
        return {}


def measured_traffic(kernel: str):
    return ncu_facts(kernel).get("bytes")


def cpu_sample(args, histories: int, generations: int, threads: int, faithful: bool = True):
    """Oracle in the reference's configuration: hardware_concurrency-1 workers, static ranges,
    per-worker f32 tallies, ordered reduction (src/mc_code.rs:302-338)."""
    from oracle import oracle as orc
    from tests.util import oracle_inputs

    deck, mesh = oracle_inputs(*args)
    r = orc.monte_carlo(deck, mesh, generations=generations, histories=histories, skip=0, threads=threads,
                        tally_mode="f32_per_worker" if faithful else "fixed64")
    return r


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reserve_stdout()
    args, desc, per_gpu = workload(a.workload)
    cores = os.cpu_count() or 2
    threads = max(1, cores - 1)
    sample = min(per_gpu, 1_000_000)
    t0 = time.perf_counter()
    cpu_sample(args, max(1000, sample // 20), 1, threads)  # page-in + first estimate
    est = (time.perf_counter() - t0) * 20
    while sample > 20_000 and est * (a.steps + a.warmup) > 200.0:
        sample //= 2
        est /= 2
    for _ in range(a.warmup):
        cpu_sample(args, sample, 1, threads)
    t0 = time.perf_counter()
    r = cpu_sample(args, sample, a.steps, threads)
    wall = time.perf_counter() - t0
    value = sample * a.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "histories_per_step_sampled": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{a.steps} generations x {sample} histories (bounded sample of the 1e7/generation workload), "
                                   f"C restatement of src/mc_code.rs (no Rust toolchain in this image), {threads} worker threads of {cores} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "k_mean": float(r.k.mean()),
    }
    emit(line)


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import nraps_b200 as nb
    from nraps_b200.dist import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun when called plainly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    reserve_stdout()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    args, desc, per_gpu = workload(a.workload)
    v, xs, dx, mesh, fuel = args
    H = a.histories if a.histories else (per_gpu if a.workload == "config4" else per_gpu * world)
    scaling = "strong" if a.workload == "config4" else "weak"
    source_mode = "fission_bank" if a.workload == "config5" else "uniform_fuel"
    K, W = a.steps, a.warmup
    gens_total = W + K

    base_opts = dict(device=local, threads_per_block=a.threads, blocks_per_sm=a.blocks_per_sm, chunk=a.chunk, source_mode=source_mode,
                     spawn_batch=a.spawn_batch, walk_cap=a.walk_cap, slots_per_thread=a.slots_per_thread)
    opts = dict(base_opts, tracking_mode=a.tracking, kernel_variant=a.variant)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    begin, count = shard_range(H, rank, world)
    from nraps_b200.dist import make_bank_callback

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_run(run_opts, sample_clocks):
        """W warm-up + K timed generations with everything resident on the device."""
        ctx = nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=gens_total, histories=H, skip=1, **run_opts)
        tally = torch.zeros(ctx.n_words, dtype=torch.int64, device=f"cuda:{local}")
        ctx.use_tally_tensor(tally)
        bank = make_bank_callback(ctx, world, local, stream) if source_mode == "fission_bank" else None

        def step(gen, ev=None):
            flush.zero_()
            if ev:
                ev[0].record()
            ctx.transport(gen, begin, count, stream)
            if ev:
                ev[1].record()
            if world > 1:
                dist.all_reduce(tally)
            ctx.finalize_generation(gen, stream)
            if bank is not None:
                bank(gen)

        for g in range(W):
            step(g)
        fence()
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
        k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record()
        for i in range(K):
            step(W + i, k_events[i])
        t_end.record()
        fence()
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        out = dict(ms_total=t_begin.elapsed_time(t_end), ms_kernel=sum(e0.elapsed_time(e1) for e0, e1 in k_events) / K,
                   res=ctx.fetch(stream), info=ctx.launch_info(), clocks=clocks, has_bank=bank is not None)
        ctx.close()
        if world > 1:
            t = torch.tensor([out["ms_total"], out["ms_kernel"]], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["ms_total"], out["ms_kernel"] = (float(x) for x in t.tolist())
        return out

    main_run = device_run(opts, True)
    ms_total, ms_kernel, res, info, clocks = (main_run[k] for k in ("ms_total", "ms_kernel", "res", "info", "clocks"))
    has_bank = main_run["has_bank"]
    coll_per_hist = res.counters["collisions"] / max(1, res.counters["histories"])
    # the other tracking mode on the same workload, reported beside the headline (not instead of it)
    other = "woodcock" if a.tracking == "surface" else "surface"
    other_run = None if a.no_variants else device_run(dict(base_opts, tracking_mode=other), False)

    # end to end through the public call: host arrays in, SolutionResults out (create + H2D + K generations + D2H)
    fence()
    t0 = time.perf_counter()
    if world > 1:
        e2e_res = nb.monte_carlo_distributed(v, xs, dx, mesh, fuel, 1.0, generations=K, histories=H, skip=1, **opts)
    else:
        e2e_res = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=K, histories=H, skip=1, **opts)
    fence()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        G, M, N, NF = v.energygroups, v.mattypes, len(mesh), len(fuel)
        h2d = 4 * (8 * M * G + M * G * G + 3 * N) + N + 8 * NF  # tables the call uploads, once per run
        d2h = 4 * (G * N + N + K) + 64                           # flux, fission source, k, counters
        peak, peak_src = peaks()
        b_hist = RECORD_BYTES + 2 * RECORD_BYTES * coll_per_hist
        if has_bank:
            b_hist += 12 + 12 * res.counters["banked"] / max(1, res.counters["histories"])  # source read + bank write, SURVEY 8d
        hist_per_launch = count
        achieved = b_hist * hist_per_launch / (ms_kernel * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": H * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "histories_per_generation": H, "histories_per_gpu": count, "generations_timed": K,
                       "source_mode": source_mode, "tracking_mode": a.tracking, "kernel_variant": a.variant, "parallelism": f"history-sharded x{world}, int64 tally all-reduce per generation",
                       "launch": info, "l2": "256 MiB device memset between steps inside the timed region (kernel inputs are ~12 KB of tables)"},
            "clocks": clocks,
            "e2e": {"value": H * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                    "wall_s": e2e_s, "device_s": float(getattr(e2e_res, "seconds_device", 0.0)),
                    "note": "one monte_carlo() call: context create + table upload + K generations + result download; bytes are per run / K"},
            "gpu_launches": (3 + (5 if has_bank else 0)) * K,  # source + transport + finalize (+ bank compaction / entropy)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("woodcock_kernel" if a.tracking == "woodcock" else "transport_kernel"),
                         "peak_source": peak_src, "kernel": ("block_event_kernel<4>" if a.variant == "block_event" else ("woodcock_kernel" if a.tracking == "woodcock" else "transport_kernel") + "<4,false,%s>" % ("true" if has_bank else "false")), "kernel_ms": ms_kernel,
                         "bytes_per_history": b_hist, "collisions_per_history": coll_per_hist,
                         "actual_limiter": {"what": "instruction issue (ncu, profiles/)", **{k: v for k, v in ncu_facts(
                             "woodcock_kernel" if a.tracking == "woodcock" else "transport_kernel").items() if k not in ("bytes", "source")}},
                         "note": "algorithmic bytes = 24 + 48*collisions/history (SURVEY 8d bank model); the fused kernel keeps "
                                 "particles in registers, so measured DRAM traffic (the 32-byte birth records) is ~2 % of that: "
                                 "the kernel is instruction-issue bound (profiles/)"},
            "k_mean": float(res.k[W:].mean()), "k_e2e_mean": float(e2e_res.k[1:].mean()) if K > 1 else float(e2e_res.k[0]),
        }
        if other_run is not None:
            o_coll = other_run["res"].counters["collisions"] / max(1, other_run["res"].counters["histories"])
            line["variants"] = {other: {
                "value": H * K / (other_run["ms_total"] * 1e-3), "unit": UNIT, "ms_per_step": other_run["ms_total"] / K,
                "k_mean": float(other_run["res"].k[W:].mean()), "collisions_per_history": o_coll,
                "roofline_frac": (RECORD_BYTES + 2 * RECORD_BYTES * o_coll) * count / (other_run["ms_kernel"] * 1e-3) / 1e9 / peak,
                "note": "same workload, same timing rules, other tracking mode: 'surface' follows the reference cell by cell "
                        "(bit-comparable with the CPU restatement); 'woodcock' is delta tracking with a collision-estimator tally "
                        "(statistically equivalent, 3 sigma / chi-square tested)"}}
        if world == 1 and not a.no_cpu:
            cores = os.cpu_count() or 2
            threads = max(1, cores - 1)
            sample_h, sample_g = min(H, 1_000_000), 3
            cpu_sample(args, 50_000, 1, threads)
            r = cpu_sample(args, sample_h, sample_g, threads)
            line["cpu_baseline"] = {
                "value": sample_h * sample_g / r.seconds_transport, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{sample_g} generations x {sample_h} histories of the same workload, C restatement of src/mc_code.rs "
                          f"threaded like the reference ({threads} workers of {cores} cores, per-worker f32 tallies)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4", "config5"])
    ap.add_argument("--histories", type=int, default=0, help="override histories per generation (total)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--spawn-batch", type=int, default=0)
    ap.add_argument("--walk-cap", type=int, default=0)
    ap.add_argument("--slots-per-thread", type=int, default=0, help="block_event variant: neutrons banked per thread")
    ap.add_argument("--tracking", default="surface", choices=["surface", "woodcock"],
                    help="surface = the reference's cell-by-cell tracking (headline, bit-comparable); woodcock = delta tracking")
    ap.add_argument("--variant", default="fused", choices=["fused", "event", "block_event"],
                    help="kernel variant (event = SoA-bank pipeline in HBM, woodcock only; block_event = experimental on-chip bank, surface only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the other-tracking-mode measurement")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
