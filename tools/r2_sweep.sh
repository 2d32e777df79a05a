#!/bin/bash
# Sweep one bench.py option over values on config 3 (and config 4 with C4=1): tools/r2_sweep.sh <tag> <option> v1 v2 ...
set -u
tag=$1; opt=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick $opt $v ${EXTRA:-} > gpurun_out/${tag}_c3_$v.json 2>/dev/null; echo "c3 $opt $v rc=$?"
  if [ "${C4:-0}" = "1" ]; then
    timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 $opt $v ${EXTRA:-} > gpurun_out/${tag}_c4_$v.json 2>/dev/null; echo "c4 $opt $v rc=$?"
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"]["block"])
    except Exception as e: print(f, "unreadable", e)
PY
