#!/bin/bash
# 2-GPU check of the shareable bank buffers: parity tests, then config 5 at growing footprints (remote bank 1 / 3.2 GB)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or native_nccl" 2>&1 | tail -6
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --quick --workload config5 --histories $2 > gpurun_out/r2_bank2_$1.json 2> gpurun_out/r2_bank2_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bank2_$1.json")); print("$1", "H=$2", "%.4e hist/s"%d["value"], "%.2f ms/gen"%d["ms_per_step"], "births+transport %.2f ms"%d["roofline"]["kernel_ms"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2_bank2_$1.err").read()[-1500:])
PY
}
run 2.5e8 250000000
run 8e8 800000000
