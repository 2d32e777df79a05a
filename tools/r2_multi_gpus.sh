#!/bin/bash
# N-GPU run of round 2 (gpurun --gpus N): bench.py under torchrun as the driver launches it, NVLink byte counters around it
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/r2_multi_gpus.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi nvlink -gt d > gpurun_out/r2_nvlink_before_n$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/r2_nvlink_after_n$N.txt 2>&1
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json, re
d=json.load(open("gpurun_out/r2_bench_n$N.json")); print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, "%.4e"%v["value"], "%.3f ms/gen"%v["ms_per_generation"], {a:(round(b,3) if b is not None else None) for a,b in v["phases_ms_per_generation"].items()})
print(d.get("multi_gpu_bit_identical_detail"))
def counters(path):
    out={}; gpu=None
    for ln in open(path):
        m=re.match(r"GPU (\d+):",ln)
        if m: gpu=int(m.group(1)); out[gpu]=[0,0]
        m=re.search(r"Link \d+: Data Tx: (\d+) KiB",ln)
        if m and gpu is not None: out[gpu][0]+=int(m.group(1))
        m=re.search(r"Link \d+: Data Rx: (\d+) KiB",ln)
        if m and gpu is not None: out[gpu][1]+=int(m.group(1))
    return out
try:
    b,a=counters("gpurun_out/r2_nvlink_before_n$N.txt"),counters("gpurun_out/r2_nvlink_after_n$N.txt")
    for g in sorted(a): print("GPU",g,"NVLink tx %.2f GiB rx %.2f GiB during the whole bench.py run"%((a[g][0]-b[g][0])/2**20,(a[g][1]-b[g][1])/2**20))
except Exception as e: print("nvlink counters unreadable:", e)
PY
