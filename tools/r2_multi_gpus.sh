#!/bin/bash
# N-GPU run of round 2 (gpurun --gpus N): bench.py under torchrun as the driver launches it, NVLink byte counters around it
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/r2_multi_gpus.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json, re
d=json.load(open("gpurun_out/r2_bench_n$N.json")); print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, "%.4e"%v["value"], "%.3f ms/gen"%v["ms_per_generation"], {a:(round(b,3) if b is not None else None) for a,b in v["phases_ms_per_generation"].items()})
print(d.get("multi_gpu_bit_identical_detail"))
print("(nvidia-smi nvlink -gt d reports N/A for every link on these boxes: no NVLink byte counters; the births phase at N = 1 / 2 / 8 is the evidence)")
PY
