#!/bin/bash
# chunk / spawn-batch sweep of the surface kernel on config 3 (same box, two repetitions)
mkdir -p gpurun_out
for rep in 1 2; do
for c in 16 32 64 128; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick --chunk $c > gpurun_out/r3s_chunk${c}_$rep.json 2>/dev/null
done
for sb in 2 3 4 8; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick --spawn-batch $sb > gpurun_out/r3s_spawn${sb}_$rep.json 2>/dev/null
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r3s_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"])
    except Exception as e: print(f, "unreadable", e)
PY
