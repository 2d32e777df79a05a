#!/bin/bash
# Round-2 evidence for profiles/ (one GPU): compute-sanitizer over the new kernels, the ncu launch list of the bench
# command, ncu --set full of the transport kernel on config 3 and of the births in bank mode.
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
} > gpurun_out/r2_compute_sanitizer.log 2>&1
tail -4 gpurun_out/r2_compute_sanitizer.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2_launches_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:transport_kernel -c 1 -s 1 -o gpurun_out/prof_r2_final_c3 -f python tools/run_generation.py --gens 2 > gpurun_out/prof_r2_final_c3.log 2>&1; tail -1 gpurun_out/prof_r2_final_c3.log
ncu --set full --clock-control none --import-source on -k regex:source_kernel -c 1 -s 1 -o gpurun_out/prof_r2_source_bank -f python tools/run_generation.py --gens 2 --source fission_bank > gpurun_out/prof_r2_source_bank.log 2>&1; tail -1 gpurun_out/prof_r2_source_bank.log
