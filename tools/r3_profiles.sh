#!/bin/bash
# Round-2 (second session) evidence for profiles/ on one GPU: ncu launch list of the bench command, ncu --set full of the
# surface kernel on config 3 and on the fine mesh of config 4 (one generation per launch: NRAPS_TAIL_BATCH=1).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 NRAPS_TAIL_BATCH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches_bench.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r3_launches_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:transport_kernel -c 1 -s 1 -o gpurun_out/prof_r3_final_c3 -f python tools/run_generation.py --gens 2 > gpurun_out/prof_r3_final_c3.log 2>&1; tail -1 gpurun_out/prof_r3_final_c3.log
ncu --set full --clock-control none --import-source on -k regex:transport_kernel -c 1 -s 1 -o gpurun_out/prof_r3_final_c4 -f python tools/run_generation.py --gens 2 --fine --histories 5000000 > gpurun_out/prof_r3_final_c4.log 2>&1; tail -1 gpurun_out/prof_r3_final_c4.log
ls -la gpurun_out/*.ncu-rep
