#!/bin/bash
# bench config 3 (and config 4 with C4=1) with each experiment build of the library: tools/r2_variants.sh <tag> lib lib_x ...
set -u
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 10 --warmup 3 --quick ${EXTRA:-} > gpurun_out/${tag}_c3_$v.json 2>/dev/null; echo "c3 $v rc=$?"
  if [ "${C4:-0}" = "1" ]; then
    NRAPS_LIB_DIR=$PWD/nraps_b200/$v timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 ${EXTRA:-} > gpurun_out/${tag}_c4_$v.json 2>/dev/null; echo "c4 $v rc=$?"
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"]["block"])
    except Exception as e: print(f, "unreadable", e)
PY
