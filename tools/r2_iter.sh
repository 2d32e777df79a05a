#!/bin/bash
# One GPU iteration of round 2 (run under gpurun from the repo root): parity first, then timing of the surface kernel.
#   gpurun --timeout 900 -- 'bash tools/r2_iter.sh <tag> [walk caps...]'
set -u
tag=${1:-iter}; shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest -m gpu: rc=$?"
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench config3: rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench config4: rc=$?"
for cap in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick --walk-cap $cap > gpurun_out/${tag}_bench_c3_cap$cap.json 2>/dev/null; echo "bench config3 cap $cap: rc=$?"
  timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 --walk-cap $cap > gpurun_out/${tag}_bench_c4_cap$cap.json 2>/dev/null; echo "bench config4 cap $cap: rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"])
    except Exception as e: print(f, "unreadable", e)
PY
