#!/bin/bash
# A/B of Woodcock-kernel builds: tools/r3_woodcock_ab.sh <tag> lib lib_x ...
set -u
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 20 --warmup 3 --quick --tracking woodcock > gpurun_out/${tag}_c3w_${v}_$rep.json 2>/dev/null
  NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 4 --warmup 3 --quick --tracking woodcock --workload config5 > gpurun_out/${tag}_c5w_${v}_$rep.json 2>/dev/null
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*w_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"]["block"])
    except Exception as e: print(f, "unreadable", e)
PY
