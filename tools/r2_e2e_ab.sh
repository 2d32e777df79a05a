#!/bin/bash
# full headline line (value + e2e) of config 3 with the generations of nraps_mc_run batched (default) and not
mkdir -p gpurun_out
for tb in 3 1 2; do
  NRAPS_TAIL_BATCH=$tb timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-variants --no-configs --no-cold > gpurun_out/${TAG:-e2e}_tb$tb.json 2>/dev/null
done
python - <<PY
import json,glob,os
tag=os.environ.get("TAG","e2e")
for f in sorted(glob.glob(f"gpurun_out/{tag}_tb*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "value %.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], "step %.3f ms"%d["ms_per_step"])
PY
