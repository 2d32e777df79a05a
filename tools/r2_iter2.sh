set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2f_c3.json 2>/dev/null; echo rc=$?
for tb in "640 2" "672 2" "448 3" "512 2" "576 2" "704 2"; do set -- $tb; timeout 300 python bench.py --steps 10 --warmup 3 --quick --threads $1 --blocks-per-sm $2 > gpurun_out/r2f_c3_t$1x$2.json 2>/dev/null; done
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 > gpurun_out/r2f_c4.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config5 > gpurun_out/r2f_c5.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_c*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"], d["details"]["launch"])
    except Exception as e: print(f, "unreadable", e)
PY
timeout 600 python tools/r2_bigmesh.py 2>&1 | tail -12
