#!/bin/bash
# Why are the in-place bank reads slow on 8 GPUs?  (gpurun --gpus 8)
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1; head -12 gpurun_out/r2_topo_n$N.txt
python - <<PY
import torch
n=torch.cuda.device_count()
print("peer access matrix:", [[int(torch.cuda.can_device_access_peer(i,j)) if i!=j else 1 for j in range(n)] for i in range(n)])
PY
run() { # label, total histories, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --quick --workload config5 --histories $2 $3 > gpurun_out/r2_bank8_$1.json 2> gpurun_out/r2_bank8_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bank8_$1.json")); print("$1", "H=$2", "%.4e hist/s"%d["value"], "%.2f ms/gen"%d["ms_per_step"], "births+transport %.2f ms"%d["roofline"]["kernel_ms"])
except Exception as e: print("$1 failed", e)
PY
}
run torchrun_1e8 100000000 ""
run torchrun_2e8 200000000 ""
run torchrun_5e8 500000000 ""
run torchrun_1e9 1000000000 ""
run torchrun_1e9_woodcock 1000000000 "--tracking woodcock"
# one process, plain peer pointers (no IPC)
for H in 100000000 1000000000; do
  for src in fission_bank uniform_fuel; do
    t0=$(date +%s.%N)
    nraps_b200/lib/nraps tests/golden/decks/case_c.txt --out /tmp --quiet --gpus $N --source $src --histories $H --generations 4 2>&1 | tail -1 | cut -c1-200
    t1=$(date +%s.%N)
    echo "native route (one process, peer pointers) $src H=$H: $(echo "$t1 - $t0" | bc) s wall for 4 generations incl. start-up"
  done
done
