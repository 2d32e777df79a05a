set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -x -q -k "fine or config4 or larger_than or fuzz or synthetic" 2>&1 | tail -4
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 > gpurun_out/r2h_c4.json 2>/dev/null
timeout 300 python bench.py --steps 4 --warmup 3 --quick --workload config4 --walk-cap -2 > gpurun_out/r2h_c4_nostride.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_c*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], d["details"]["launch"])
PY
timeout 600 python tools/r2_bigmesh.py --refine 40 80 160 320 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:transport_kernel -c 1 -s 1 -o gpurun_out/prof_r2h_fine -f python tools/run_generation.py --gens 2 --fine --histories 2000000 > gpurun_out/prof_r2h_fine.log 2>&1; tail -2 gpurun_out/prof_r2h_fine.log
