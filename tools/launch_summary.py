"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.

Usage: python tools/launch_summary.py profiles/r1_launches_bench.csv > profiles/r1_launches_bench_summary.txt
"""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", name)
    if m and "at::" not in name:
        return m.group(1)
    m = re.search(r"at::native::(\w+)|at::(\w+)<", name)
    return "torch " + (m.group(1) or m.group(2)) if m else name[:60]


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        ms = float(r[v].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[u]]
        tot[short(r[k])] += ms
        cnt[short(r[k])] += 1
    total = sum(tot.values())
    print(f"# ncu launch list {path} (gpu__time_duration.sum, --clock-control none), summarised per kernel")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"# {sum(cnt.values())} launches, {total:.2f} ms total")
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name:62s} n={cnt[name]:4d} total={ms:9.3f} ms share={100 * ms / total:6.2f}% avg={ms / cnt[name]:8.4f} ms")


if __name__ == "__main__":
    main()
