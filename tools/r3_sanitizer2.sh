#!/bin/bash
# compute-sanitizer initcheck + synccheck over the kernels of the final round-2 build
set -u
mkdir -p gpurun_out /tmp/san
export PYTHONUNBUFFERED=1
{
for tool in initcheck synccheck; do
  for args in "--tracking surface" "--tracking surface --source fission_bank" "--tracking surface --fine" "--tracking woodcock --source fission_bank"; do
    echo "== $tool: run_generation.py $args (30000 histories x 2 generations)"
    timeout 600 compute-sanitizer --tool $tool python tools/run_generation.py --histories 30000 --gens 2 $args 2>&1 | grep -E "^k |ERROR SUMMARY|Error|error|Uninitialized" | cut -c1-200 | head -8
  done
done
} > gpurun_out/r3_compute_sanitizer2.log 2>&1
cat gpurun_out/r3_compute_sanitizer2.log | cut -c1-200
