#!/bin/bash
# 2-GPU check of round 2 (gpurun --gpus 2): the multi-GPU parity tests, then bench.py under torchrun
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or native_nccl" > gpurun_out/r2e_pytest_2gpu.log 2>&1; echo "pytest 2 GPUs rc=$?"
tail -15 gpurun_out/r2e_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench rc=$?"
tail -15 gpurun_out/r2e_bench_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2e_bench_n2.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, "%.4e"%v["value"], v["ms_per_generation"], v["phases_ms_per_generation"])
print(d.get("multi_gpu_bit_identical_detail"))
PY
