#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the kernels of the final round-2 build: surface (coarse, fine mesh, fission
# bank), Woodcock with the bank, the batched launches of the nraps driver and three large generations per launch.
set -u
mkdir -p gpurun_out /tmp/san
export PYTHONUNBUFFERED=1
{
for tool in memcheck racecheck; do
  for args in "--tracking surface" "--tracking surface --source fission_bank" "--tracking surface --fine" "--tracking woodcock --source fission_bank"; do
    echo "== $tool: run_generation.py $args (30000 histories x 2 generations)"
    timeout 600 compute-sanitizer --tool $tool python tools/run_generation.py --histories 30000 --gens 2 $args 2>&1 | grep -E "^k |ERROR SUMMARY|Error|error" | cut -c1-160 | head -5
  done
  echo "== $tool: nraps case_a batched (H=3000, 40 generations), surface"
  timeout 300 compute-sanitizer --tool $tool nraps_b200/lib/nraps tests/golden/decks/case_a.txt --histories 3000 --generations 40 --skip 2 --out /tmp/san --quiet 2>&1 | grep -E "k_fund|ERROR SUMMARY|Error|error" | head -5
done
} > gpurun_out/r3_compute_sanitizer.log 2>&1
cat gpurun_out/r3_compute_sanitizer.log | tail -40
