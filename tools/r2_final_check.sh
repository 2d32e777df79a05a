#!/bin/bash
# What the driver does at round end, on one GPU: the -m gpu suite, smoke(), both bench arms
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_final_bench_n1.err
python - <<PY
import json
r=json.load(open("gpurun_out/r2_final_bench_reference.json")); d=json.load(open("gpurun_out/r2_final_bench_n1.json"))
print("reference arm: %.4e hist/s, %.1f ms/step, config equal: %s" % (r["value"], r["ms_per_step"], r["config"]==d["config"]))
print("ours: value %.4e  e2e %.4e  ms/step %.3f  frac %.3f  launches %d  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("ratio e2e/reference: %.1f" % (d["e2e"]["value"]/r["value"]))
for k,v in d["configs"].items(): print(k, "%.4e"%v["value"], "%.2f ms/gen"%v["ms_per_generation"], v["slowest_phase"])
print(d["multi_gpu_bit_identical"], {k:v["process_wall_s"] for k,v in d["e2e_cold"].items() if isinstance(v,dict)}, d["cpu_baseline"]["value"])
PY
