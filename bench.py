#!/usr/bin/env python
"""Throughput of the Monte Carlo transport path: neutron histories/s per generation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config4|config5]

A "step" is one generation: births -> all flights and collisions -> tally -> (all-reduce) -> k.
The headline workload is BASELINE config 3 (TestCaseC cross sections and geometry, G = 4, N = 408 cells, 10^7
histories per generation and GPU, uniform-fuel source, PCG32 seed 42 / stream 54 / stride 152917), surface tracking
(the reference's own algorithm, bit-comparable with the CPU arm).  For N > 1 (launched by torchrun, one rank per GPU)
every GPU keeps 10^7 histories per generation (weak scaling); the only collective is the per-generation int64
all-reduce of the tally buffer, issued on a side stream while the next generation already transports.

Besides the headline the JSON line carries, under "configs", the two multi-GPU configurations BASELINE names:
config 4 (fine mesh N = 4080, 10^8 histories per generation in total: strong scaling) and config 5 (1.25e8 histories per
GPU and generation, fission-bank source: weak scaling; the bank stays where it was compacted and is read over NVLink),
each with a per-phase split, and "multi_gpu_bit_identical": the distributed result against a single-GPU run.

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference algorithm (oracle/, all
host threads) on the same configuration; it is the only other place this file touches oracle/ besides the
cpu_baseline leg, and it loads nothing of the product.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "neutron histories/s per generation"
UNIT = "histories/s"
RECORD_BYTES = 24  # SURVEY 8d bank record: x, mu, cell, packed groups/flags, rng state
DECK_C = os.path.join(ROOT, "tests", "golden", "decks", "case_c.txt")

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


WORKLOADS = {
    # name: (fine mesh?, histories per generation as a function of the world size, scaling, source mode, description)
    "config3": (False, lambda w: 10_000_000 * w, "weak", "uniform_fuel",
                "config3: TestCaseC XS+geometry (G=4, M=4, N=408 cells), 1e7 histories/generation/GPU, uniform_fuel source, "
                "PCG32 seed 42/stream 54/stride 152917"),
    "config4": (True, lambda w: 100_000_000, "strong", "uniform_fuel",
                "config4: TestCaseC XS, fine mesh MPFR=80/MPWR=40 (G=4, N=4080 cells), 1e8 histories/generation total (strong scaling)"),
    "config5": (False, lambda w: 125_000_000 * w, "weak", "fission_bank",
                "config5: TestCaseC XS+geometry (G=4, N=408), 1.25e8 histories/generation/GPU, fission_bank source "
                "(power iteration; every rank's bank read in place over NVLink) (weak scaling)"),
}


def workload_config(name: str, world: int, tracking: str = "surface") -> dict:
    """The `config` object of the JSON line: the same for our arm and the reference arm."""
    fine, hist, scaling, source, desc = WORKLOADS[name]
    return {"workload": desc, "histories_per_generation": hist(world), "source_mode": source, "tracking_mode": tracking}


def product_problem(name: str):
    """Solver inputs through the product's own host side (nraps_b200.process_input / mesh_gen)."""
    from tests.util import load_case

    return load_case("c", mpfr=80, mpwr=40) if WORKLOADS[name][0] else load_case("c")


def oracle_problem(name: str):
    """Solver inputs through the oracle's numpy host side only (the reference arm maps no product library)."""
    import numpy as np

    from oracle import host_oracle as ho

    deck = ho.process_input(DECK_C)
    if WORKLOADS[name][0]:
        deck.mpfr, deck.mpwr = 80, 40
        deck.dx_fuel = np.float32(deck.roddia / np.float32(80))
        deck.dx_water = np.float32(deck.rodpitch / np.float32(40))
    mesh = ho.mesh_gen(deck.matid, deck.mpfr, deck.mpwr, deck.numass, deck.dx_fuel, deck.dx_water)
    return deck, mesh


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


def cpu_run(problem, histories: int, generations: int, threads: int):
    """The oracle in the reference's configuration: hardware_concurrency-1 workers, static ranges, per-worker f32
    tallies, ordered reduction (src/mc_code.rs:302-338)."""
    from oracle import oracle as orc

    deck, mesh = problem
    return orc.monte_carlo(deck, mesh, generations=generations, histories=histories, skip=0, threads=threads,
                           tally_mode="f32_per_worker")


def cpu_sample_text(steps: int, sample: int, full: int, threads: int, cores: int) -> str:
    part = "the full generation" if sample == full else f"a bounded sample of the {full:.3g}-history generation"
    return (f"{steps} generations x {sample} histories ({part}), C restatement of src/mc_code.rs (no Rust toolchain in "
            f"this image), threaded like the reference: {threads} workers of {cores} cores, per-worker f32 tallies")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    reserve_stdout()
    problem = oracle_problem(a.workload)
    full = WORKLOADS[a.workload][1](max(world, a.gpus))
    cores = os.cpu_count() or 2
    threads = max(1, cores - 1)
    t0 = time.perf_counter()
    cpu_run(problem, 50_000, 1, threads)  # page-in + first estimate
    per_history = (time.perf_counter() - t0) / 50_000
    # every step is the full generation when K + W of them fit ~4 minutes (they do at N = 1: ~3 s each on 15 threads),
    # else a bounded sample of it -- the CPU rate does not depend on the sample size
    sample = full
    while sample > 100_000 and per_history * sample * (a.steps + a.warmup) > 240.0:
        sample //= 2
    for _ in range(a.warmup):
        cpu_run(problem, sample, 1, threads)
    t0 = time.perf_counter()
    r = cpu_run(problem, sample, a.steps, threads)
    wall = time.perf_counter() - t0
    value = sample * a.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": WORKLOADS[a.workload][2],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a.workload, max(world, a.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(a.steps, sample, full, threads, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "histories_per_step": sample, "k_mean": float(r.k.mean()),
    }
    emit(line)


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import nraps_b200 as nb
    from nraps_b200.dist import OverlappedReducer, setup_bank_peers, shard_range

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
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, W = a.steps, a.warmup
    tune = dict(threads_per_block=a.threads, blocks_per_sm=a.blocks_per_sm, chunk=a.chunk, spawn_batch=a.spawn_batch, walk_cap=a.walk_cap)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    main = torch.cuda.current_stream()

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def timed_run(name, tracking, k_steps, w_steps, *, variant="fused", sample_clocks=False, phases=False):
        """w_steps warm-up + k_steps timed generations of workload `name`, everything resident on the device.

        Uniform source: transport on the main stream, all-reduce + finalize of the same generation on a side stream
        (OverlappedReducer).  Fission bank: transport -> compact -> all-reduce -> finalize -> advance, in order."""
        fine, hist, scaling, source, desc = WORKLOADS[name]
        v, xs, dx, mesh, fuel = product_problem(name)
        H = a.histories if (a.histories and name == a.workload) else hist(world)
        begin, count = shard_range(H, rank, world)
        bank = source == "fission_bank"
        ctx = nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=w_steps + k_steps, histories=H, skip=1, device=local,
                                   source_mode=source, tracking_mode=tracking, kernel_variant=variant, profile_phases=phases, **tune)
        kernel_ev, coll_ev = [], []
        if bank:
            setup_bank_peers(ctx, rank, world, max(shard_range(H, r, world)[1] for r in range(world)))
            tally = torch.zeros(ctx.n_words, dtype=torch.int64, device=dev)
            ctx.use_tally_tensor(tally)

            def step(gen, timed):
                flush.zero_()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
                if timed:
                    e[0].record()
                ctx.transport(gen, begin, count, main.cuda_stream)
                if timed:
                    e[1].record()
                ctx.bank_compact(gen, main.cuda_stream)
                if timed:
                    e[2].record()
                if world > 1:
                    dist.all_reduce(tally)
                if timed:
                    e[3].record()
                    kernel_ev.append((e[0], e[1]))
                    coll_ev.append((e[2], e[3]))
                ctx.finalize_generation(gen, main.cuda_stream)
                ctx.bank_advance(gen, main.cuda_stream)

            drain = lambda: None  # noqa: E731
        else:
            red = OverlappedReducer(ctx, world, local, main)

            def step(gen, timed):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timed else None
                red.step(gen, begin, count, before_transport=lambda: (flush.zero_(), timed and e[0].record()),
                         after_transport=lambda: timed and e[1].record())
                if timed:
                    kernel_ev.append((e[0], e[1]))

            drain = red.drain

        for g in range(w_steps):
            step(g, False)
        drain()
        fence()
        phase0 = ctx.phase_ms() if phases else None
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record()
        for i in range(k_steps):
            step(w_steps + i, True)
        drain()
        t_end.record()
        fence()
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        ms_total = t_begin.elapsed_time(t_end)
        ms_kernel = sum(e0.elapsed_time(e1) for e0, e1 in kernel_ev) / k_steps
        if not bank and red.pipelined:
            # uniform source: the launches of consecutive generations are pipelined on two streams (OverlappedReducer), so
            # the events of one launch span the end of the launch before it; the kernel time of a generation is the
            # rate at which launches complete over the timed region
            ms_kernel = ms_total / k_steps
        ms_coll = sum(e0.elapsed_time(e1) for e0, e1 in coll_ev) / k_steps if coll_ev else 0.0
        out = dict(H=H, count=count, res=ctx.fetch(main.cuda_stream), info=ctx.launch_info(), clocks=clocks, bank=bank,
                   pipelined=(not bank) and red.pipelined,
                   scaling=scaling, desc=desc)
        if phases:
            p1 = ctx.phase_ms()
            out["phase_ms"] = {k: (p1[k] - phase0[k]) / k_steps for k in p1}
        if world > 1:
            dist.barrier()  # fission bank: peers may still hold mappings of this rank's bank buffers
        ctx.close()
        out["ms_total"], out["ms_kernel"], out["ms_coll"] = max_over_ranks(ms_total, ms_kernel, ms_coll)
        return out

    def config_entry(name, tracking, k_steps, w_steps):
        """One entry of "configs": throughput of the whole job and where a generation's time goes (max over ranks)."""
        r = timed_run(name, tracking, k_steps, w_steps, phases=True)
        ph = r["phase_ms"]
        names = list(ph)
        vals = max_over_ranks(*[ph[n] for n in names])
        ph = dict(zip(names, vals))
        ms_gen = r["ms_total"] / k_steps
        phases = {"source": ph["source"], "transport": ph["transport"], "tally_prefix": ph["prefix"], "bank_compaction": ph["compact"],
                  "all_reduce" + ("" if r["bank"] else " (side stream, overlapped)"): r["ms_coll"] if r["bank"] else None,
                  "finalize": ph["finalize"]}
        accounted = sum(v for v in phases.values() if v)
        phases["other (L2 flush memset, launch gaps, bank bookkeeping)"] = max(0.0, ms_gen - accounted) if r["bank"] else None
        timed = {k: v for k, v in phases.items() if v is not None}
        return {"value": r["H"] * k_steps / (r["ms_total"] * 1e-3), "unit": UNIT, "ms_per_generation": ms_gen, "scaling": r["scaling"],
                "histories_per_generation": r["H"], "histories_per_gpu": r["count"], "generations_timed": k_steps,
                "warmup": w_steps, "tracking_mode": tracking, "workload": r["desc"], "k_last": float(r["res"].k[-1]),
                "phases_ms_per_generation": phases, "slowest_phase": max(timed, key=timed.get), "launch": r["info"],
                "collisions_per_history": r["res"].counters["collisions"] / max(1, r["res"].counters["histories"])}

    # ---- the headline and the other tracking mode on the same workload
    head = timed_run(a.workload, a.tracking, K, W, variant=a.variant, sample_clocks=True)
    res, info, H, count = head["res"], head["info"], head["H"], head["count"]
    coll_per_hist = res.counters["collisions"] / max(1, res.counters["histories"])
    other = "woodcock" if a.tracking == "surface" else "surface"
    other_run = None if a.no_variants else timed_run(a.workload, other, K, W)

    # ---- end to end through the public call: host arrays in, SolutionResults out (create + H2D + K generations + D2H)
    v, xs, dx, mesh, fuel = product_problem(a.workload)
    e2e_opts = dict(tune, device=local, source_mode=WORKLOADS[a.workload][3], tracking_mode=a.tracking, kernel_variant=a.variant)
    fence()
    t0 = time.perf_counter()
    if world > 1:
        e2e_res = nb.monte_carlo_distributed(v, xs, dx, mesh, fuel, 1.0, generations=K, histories=H, skip=1, **e2e_opts)
    else:
        e2e_res = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=K, histories=H, skip=1, **e2e_opts)
    fence()
    (e2e_s,) = max_over_ranks(time.perf_counter() - t0)

    # ---- BASELINE configs 4 and 5, and the distributed result against one GPU
    configs = {}
    if not a.no_configs:
        kc, wc = max(2, min(4, K)), 2
        for key, name, tracking in (("config4_strong", "config4", "surface"), ("config4_strong_woodcock", "config4", "woodcock"),
                                    ("config5_weak", "config5", "surface"), ("config5_weak_woodcock", "config5", "woodcock")):
            try:
                configs[key] = config_entry(name, tracking, kc, wc)
            except Exception as e:  # noqa: BLE001 -- the headline must still be printed; a failure is reported in its place
                # (a setup failure -- e.g. the bank buffers cannot be shared between processes on this host -- hits every
                # rank alike, before any collective of the run)
                configs[key] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.synchronize()
    identical = None
    if not a.no_configs:
        vs, xss, dxs, meshs, fuels = product_problem("config3")
        identical = {}
        for mode in ("uniform_fuel", "fission_bank"):
            kw = dict(generations=4, histories=50_001, skip=1, source_mode=mode)
            try:
                many = nb.monte_carlo_distributed(vs, xss, dxs, meshs, fuels, 1.0, device=local, **kw) if world > 1 else None
            except Exception as e:  # noqa: BLE001 -- reported as "not identical" with the reason
                identical[mode] = False
                identical[mode + "_error"] = f"{type(e).__name__}: {e}"
                continue
            if rank == 0:
                one = nb.monte_carlo(vs, xss, dxs, meshs, fuels, 1.0, device=local, **kw)
                if world == 1:  # the generation-level route (what the multi-GPU launcher drives) against the batched call
                    with nb.MonteCarloContext(vs, xss, dxs, meshs, fuels, 1.0, device=local, **kw) as c1:
                        for g in range(4):
                            c1.transport(g)
                            if mode == "fission_bank":
                                c1.bank_compact(g)
                            c1.finalize_generation(g)
                            if mode == "fission_bank":
                                c1.bank_advance(g)
                        many = c1.fetch()
                identical[mode] = bool(np.array_equal(many.k.view(np.uint32), one.k.view(np.uint32))
                                       and np.array_equal(many.flux.view(np.uint32), one.flux.view(np.uint32))
                                       and np.array_equal(many.bank_sizes, one.bank_sizes))
        fence()

    if rank == 0:
        G, M, N, NF = v.energygroups, v.mattypes, len(mesh), len(fuel)
        h2d = 4 * (8 * M * G + M * G * G + 3 * N) + N + 8 * NF  # tables the call uploads, once per run
        d2h = 4 * (G * N + N + K) + 64                           # flux, fission source, k, counters
        peak, peak_src = peaks()
        b_hist = RECORD_BYTES + 2 * RECORD_BYTES * coll_per_hist
        if head["bank"]:
            b_hist += 12 + 12 * res.counters["banked"] / max(1, res.counters["histories"])  # source read + bank write, SURVEY 8d
        ms_kernel, ms_total = head["ms_kernel"], head["ms_total"]
        achieved = b_hist * count / (ms_kernel * 1e-3) / 1e9
        kname = "woodcock_kernel" if a.tracking == "woodcock" else "transport_kernel"
        facts = ncu_facts(kname)
        line = {
            "metric": METRIC, "value": H * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(a.workload, world, a.tracking),
            "details": {"histories_per_gpu": count, "generations_timed": K, "kernel_variant": a.variant,
                        "parallelism": f"history-sharded x{world}; int64 tally all-reduce + finalize per generation on a side stream, overlapped with the next generation" + ("; launches of consecutive generations alternate between two streams" if head["pipelined"] else ""),
                        "launch": info, "l2": "256 MiB device memset between steps inside the timed region (kernel inputs are ~12 KB of tables)"},
            "clocks": head["clocks"],
            "e2e": {"value": H * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                    "wall_s": e2e_s, "device_s": float(getattr(e2e_res, "seconds_device", 0.0)),
                    "note": "one monte_carlo() call: context create + table upload + K generations + result download; bytes are per run / K"
                            + ("; like `value`, the call launches consecutive uniform-source generations on two alternating streams, so "
                               "that the next launch fills the tail of the one before (DESIGN.md section 5)"
                               if not head["bank"] else "")},
            # births + transport + tally prefix sum + finalize (+ bank: 3 compaction kernels, histogram, entropy)
            "gpu_launches": ((4 if a.tracking == "surface" else 3) + (5 if head["bank"] else 0)) * K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": facts.get("bytes"), "peak_source": peak_src,
                         "kernel": kname + "<4,false,%s>" % ("true" if head["bank"] else "false"), "kernel_ms": ms_kernel,
                         "kernel_ms_note": ("births + transport + tally prefix sum of one generation (CUDA events on the launching stream); "
                                            "the transport kernel is > 98 % of it (profiles/ launch list)") if not head["pipelined"] else
                                           ("timed region / K: consecutive generations are launched on two streams so that the next launch fills "
                                            "the tail of the one before (uniform source: independent generations), per-launch events would overlap; "
                                            "births, L2 flush and tally prefix sum are inside, the transport kernel is > 98 % of it (profiles/ launch list)"),
                         "bytes_per_history": b_hist, "collisions_per_history": coll_per_hist,
                         "actual_limiter": {"what": "instruction issue (ncu, profiles/)", **{k: v for k, v in facts.items() if k not in ("bytes", "source", "round1")}},
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
        if configs:
            line["configs"] = configs
            line["multi_gpu_bit_identical"] = bool(identical and all(v for k, v in identical.items() if not k.endswith("_error")))
            line["multi_gpu_bit_identical_detail"] = {
                **(identical or {}), "what": "k, flux and bank sizes of a 4-generation 50 001-history TestCaseC run through this launch's "
                                            f"{world}-rank generation-level route against nraps_mc_run on one GPU, compared bit for bit"}
        if world == 1 and not a.no_cold:
            line["e2e_cold"] = cold_runs(local)
        if world == 1 and not a.no_cpu:
            from tests.util import oracle_inputs

            cores = os.cpu_count() or 2
            threads = max(1, cores - 1)
            problem = oracle_inputs(v, xs, dx, mesh, fuel)
            t0 = time.perf_counter()
            cpu_run(problem, 50_000, 1, threads)
            per_history = (time.perf_counter() - t0) / 50_000
            sample_h, sample_g = H, 3
            while sample_h > 100_000 and per_history * sample_h * sample_g > 30.0:
                sample_h //= 2
            t0 = time.perf_counter()
            cpu_run(problem, sample_h, sample_g, threads)
            wall = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample_h * sample_g / wall, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": cpu_sample_text(sample_g, sample_h, H, threads, cores)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cold_runs(device: int) -> dict:
    """Fresh-process runs of the drop-in binary on the three decks as shipped (BASELINE configs 1 and 2): wall clock of
    the whole process against the device time of its generations, with nraps_mc_run's own split of the difference."""
    exe = os.path.join(ROOT, "nraps_b200", "lib", "nraps")
    out = {}
    for case in "abc":
        deck = os.path.join(ROOT, "tests", "golden", "decks", f"case_{case}.txt")
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            run = subprocess.run([exe, deck, "--out", d, "--quiet", "--device", str(device)], capture_output=True, text=True,
                                 env=dict(os.environ, NRAPS_TIMING="1"), timeout=300)
            wall = time.perf_counter() - t0
        entry = {"process_wall_s": wall, "rc": run.returncode}
        for ln in run.stderr.splitlines():
            try:
                entry.update(json.loads(ln))
            except ValueError:
                pass
        out[f"deck_{case}"] = entry
    out["note"] = ("`nraps <deck>` as a new process per deck (generations x histories as the deck says: A 100 x 1e5, B and C 100 x 1e6): "
                   "process start, CUDA context, tables, 100 generations, CSV files.  nraps_mc_run_ms is the library call's own wall-clock split")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4", "config5"])
    ap.add_argument("--histories", type=int, default=0, help="override histories per generation (total) of the headline workload")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--spawn-batch", type=int, default=0)
    ap.add_argument("--walk-cap", type=int, default=0)
    ap.add_argument("--tracking", default="surface", choices=["surface", "woodcock"],
                    help="surface = the reference's cell-by-cell tracking (headline, bit-comparable); woodcock = delta tracking")
    ap.add_argument("--variant", default="fused", choices=["fused", "event"],
                    help="kernel variant (event = SoA-bank pipeline in HBM, woodcock only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the other-tracking-mode measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 4 / 5 and the bit-identity check")
    ap.add_argument("--no-cold", action="store_true", help="skip the fresh-process runs of the shipped decks")
    ap.add_argument("--quick", action="store_true", help="headline only: --no-cpu --no-variants --no-configs --no-cold")
    a = ap.parse_args()
    if a.quick:
        a.no_cpu = a.no_variants = a.no_configs = a.no_cold = True
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
