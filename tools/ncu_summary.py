"""Condense an .ncu-rep (one kernel launch, --set full --import-source on) into the text summary kept under profiles/.

Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep histories_per_launch > profiles/NAME.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, hist = sys.argv[1], float(sys.argv[2])
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    print(f"# ncu summary of {rep} ({hist:g} histories in this launch)")
    print("kernel:", vals[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:85s} {vals[i]:>18s} {units[i]}")
    stall = [(float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled")]
    for v, h in sorted(stall, reverse=True)[:8]:
        print(f"{h:85s} {v:18.3f}")
    src = page(rep, "source")
    sh, data = src[1], src[2:]
    ie, iav, isrc, ismp = sh.index("Instructions Executed"), sh.index("Avg. Threads Executed"), sh.index("Source"), sh.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    thr = sum(int(r[ie]) * float(r[iav]) for r in data)
    print(f"\nwarp instructions / history: {tot / hist:.1f}   thread instructions / history: {thr / hist:.1f}   SASS lines: {len(data)}")
    print("\n# hot SASS (>= 0.1% of executed warp instructions): index, warp-inst per history, avg active threads, stall samples, SASS")
    for n, r in enumerate(data):
        e = int(r[ie])
        if e > tot * 0.001:
            print(f"{n:5d} {e / hist:8.2f} {float(r[iav]):5.1f} {int(r[ismp]):7d}  {r[isrc][:100]}")


if __name__ == "__main__":
    main()
