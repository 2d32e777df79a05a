timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2p_c3_$i.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config5 > gpurun_out/r2p_c5_$i.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 > gpurun_out/r2p_c4_$i.json 2>/dev/null
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_c*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], d["details"]["launch"]["block"])
PY
