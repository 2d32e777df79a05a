#!/bin/bash
# kernel time against the size of the launch (config 3 physics): the intercept is the tail of the persistent kernel
mkdir -p gpurun_out
for h in 1250000 2500000 5000000 10000000 20000000 40000000; do
  timeout 300 python bench.py --steps 8 --warmup 3 --quick --histories $h ${EXTRA:-} > gpurun_out/${TAG:-scan}_$h.json 2>/dev/null
done
python - <<PY
import json,glob,os
tag=os.environ.get("TAG","scan")
for f in sorted(glob.glob(f"gpurun_out/{tag}_*.json"), key=lambda s:int(s.split('_')[-1].split('.')[0])):
    d=json.load(open(f)); h=int(f.split('_')[-1].split('.')[0]); print(h, "%.4e"%d["value"], "step %.3f ms"%d["ms_per_step"], "kernel %.3f ms"%d["roofline"]["kernel_ms"])
PY
