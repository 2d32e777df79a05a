"""Static (no-GPU) report of every sm_100a kernel in libnraps_b200.so: registers, spills, barriers from the ptxas logs
the Makefile keeps (nraps_b200/lib/obj/*.ptxas.log) and the SASS instruction mix from `cuobjdump -sass`.

For the two persistent history kernels it also lists the innermost loop (the backward branch with the shortest span
that contains a shared-memory atomic): instructions per trip of the loop body and per tally score, which is what an
issue-bound kernel pays per cell crossing / tentative collision.

Usage: python tools/static_report.py > profiles/rN_static_ptxas_sass.txt     (after `make -C nraps_b200/csrc`)
"""
from __future__ import annotations

import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nraps_b200", "lib", "obj")
MIX = ["ATOMS", "RED", "ATOMG", "LDS", "STS", "LDG", "STG", "MUFU", "F2I", "I2F", "FFMA", "FADD", "FMUL", "IMAD", "BRA", "BSSY",
       "BSYNC", "WARPSYNC", "SHFL", "VOTE", "BAR"]


def demangle(names):
    if not names:  # cu++filt without arguments would wait on stdin
        return []
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=60).stdout.splitlines()
    short = []
    for s in out:
        s = s.replace("(bool)1", "1").replace("(bool)0", "0").replace("(int)", "")
        s = re.sub(r"\(anonymous namespace\)::|<unnamed>::|nraps::", "", s)
        s = re.sub(r"^void ", "", s)
        s = re.sub(r"\([^()]*\)$", "", s)  # the parameter list
        short.append(s)
    return short


def ptxas_info(log):
    """{mangled name: (registers, stack, spill stores, spill loads, barriers)}"""
    info, cur = {}, None
    text = open(log).read().splitlines()
    for i, line in enumerate(text):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = m.group(1)
            st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", text[i + 1])
            info[cur] = [0, int(st.group(1)), int(st.group(2)), int(st.group(3)), 0]
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?", line)
        if m and cur:
            info[cur][0] = int(m.group(1))
            info[cur][4] = int(m.group(2) or 0)
    return info


def sass_functions(obj):
    """{mangled name: [(address, mnemonic, text)]}"""
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=300).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            body = m.group(2).strip()
            toks = body.split()
            mn = toks[1] if toks[0].startswith("@") else toks[0]
            cur.append((int(m.group(1), 16), mn, body))
    return funcs


def inner_loop(ins):
    """Shortest backward branch span containing a shared atomic: (first index, last index) or None."""
    addr_to_idx = {a: i for i, (a, _, _) in enumerate(ins)}
    best = None
    for i, (a, mn, body) in enumerate(ins):
        if not mn.startswith("BRA"):
            continue
        m = re.search(r"0x([0-9a-f]+)\s*$", body)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_to_idx:
            continue
        j = addr_to_idx[tgt]
        span = ins[j:i + 1]
        if not any(x[1].startswith("ATOMS") for x in span):
            continue
        if best is None or (i - j) < (best[1] - best[0]):
            best = (j, i)
    return best


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(OBJ, "*.ptxas.log"))):
        stem = os.path.basename(log)[: -len(".ptxas.log")]
        info = ptxas_info(log)
        funcs = sass_functions(os.path.join(OBJ, stem + ".o"))
        names = [n for n in info if n in funcs]
        for n, short in zip(names, demangle(names)):
            ins = funcs[n]
            mix = collections.Counter()
            for _, mn, _ in ins:
                for key in MIX:
                    if mn == key or mn.startswith(key + "."):
                        mix[key] += 1
            rows.append((stem, short, info[n], len(ins), mix, inner_loop(ins), ins))
    print("# Static report of the sm_100a kernels (ptxas -v, cuobjdump -sass); produced without a GPU by tools/static_report.py")
    print("# nvcc flags: see nraps_b200/csrc/Makefile (-O3 -lineinfo -fmad=false, arch=compute_100a,code=sm_100a)")
    print()
    print(f"{'file':<13} {'kernel':<46} {'regs':>4} {'stack':>5} {'spill':>7} {'bar':>3} {'SASS':>5}  instruction mix (static counts)")
    for stem, short, (regs, stack, sst, sld, bar), n, mix, loop, _ in rows:
        mixs = " ".join(f"{k}={v}" for k, v in mix.items() if v)
        print(f"{stem:<13} {short:<46} {regs:>4} {stack:>5} {sst:>3}/{sld:<3} {bar:>3} {n:>5}  {mixs}")
    print()
    print("# Innermost loop of the persistent history kernels (shortest backward branch span with an ATOMS):")
    print("# body = SASS instructions between the loop head and its back branch (the rarely taken carry path of each score, 5")
    print("# instructions, included), scores = ATOMS returning a value (the low tally word)")
    for stem, short, _, n, _, loop, ins in rows:
        if loop is None or stem not in ("mc_transport", "mc_woodcock"):
            continue
        j, i = loop
        span = ins[j:i + 1]
        scores = sum(1 for _, mn, body in span if mn.startswith("ATOMS") and " RZ," not in body)
        per = f"{len(span) / scores:.1f} per score" if scores else ""
        print(f"{short:<46} body {len(span):>4} instructions @0x{span[0][0]:04x}-0x{span[-1][0]:04x}, {scores} scores  {per}")
    print()
    hot = [r for r in rows if r[0] == "mc_transport" and r[1].startswith("transport_kernel<4, 0, 0, 0>")]
    if hot and hot[0][5]:
        j, i = hot[0][5]
        print("# transport_kernel<4, 0, 0, 0> (the bench.py headline instantiation): the walk loop, unrolled by four")
        for a, _, body in hot[0][6][j:i + 1]:
            print(f"    /*{a:04x}*/  {body}")


if __name__ == "__main__":
    sys.exit(main())
