#!/bin/bash
# A/B of experiment builds of the library on one box: tools/r2_ab.sh <tag> lib lib_x ...  (C4=1: also config 4; TESTS=1: GPU suite first)
set -u
tag=$1; shift
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
for rep in 1 2; do
for v in "$@"; do
  NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 10 --warmup 3 --quick ${EXTRA:-} > gpurun_out/${tag}_c3_${v}_$rep.json 2>/dev/null
  if [ "${C4:-0}" = "1" ]; then
    NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 ${EXTRA:-} > gpurun_out/${tag}_c4_${v}_$rep.json 2>/dev/null
  fi
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"]["block"])
    except Exception as e: print(f, "unreadable", e)
PY
