set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2g_c3.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 > gpurun_out/r2g_c4.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 --walk-cap -2 > gpurun_out/r2g_c4_nostride.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config5 > gpurun_out/r2g_c5.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_c*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"])
    except Exception as e: print(f, "unreadable", e)
PY
timeout 600 python tools/r2_bigmesh.py --refine 80 160 320 2>&1 | tail -8
