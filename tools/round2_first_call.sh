#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root): everything that was written after round 1's GPU
# minutes were spent and has therefore only been verified on the CPU.  Each step runs under its own timeout so that a
# hang in the new kernel cannot hold the box; outputs go to gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# 1. the default suite must still be green on this box (sanity: the lane kernels' SASS did not change)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest -m gpu: rc=$?"
# 2. the block-level event pipeline: parity first (bit-exact against the oracle and the lane kernel), then speed
NRAPS_TEST_BLOCK_EVENT=1 timeout 300 python -m pytest tests/test_gpu_block_event.py -x -q > gpurun_out/r2_block_event_tests.log 2>&1
rc=$?; echo "block_event parity: rc=$rc"
if [ $rc -eq 0 ]; then
  for spt in 2 3 4; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants --variant block_event --slots-per-thread $spt \
        > gpurun_out/r2_bench_block_event_spt$spt.json 2> gpurun_out/r2_bench_block_event_spt$spt.err; echo "bench block_event spt=$spt: rc=$?"
  done
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants --variant block_event --spawn-batch 5 \
      > gpurun_out/r2_bench_block_event_T5.json 2> gpurun_out/r2_bench_block_event_T5.err; echo "bench block_event predicted classes T=5: rc=$?"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants --variant block_event --threads 1024 --blocks-per-sm 1 \
      > gpurun_out/r2_bench_block_event_1024.json 2> gpurun_out/r2_bench_block_event_1024.err; echo "bench block_event 1024x1: rc=$?"
fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench default: rc=$?"
# 3. GPU-vs-oracle fuzz on random problems (lane kernels, both tracking modes, both source modes)
NRAPS_GPU_FUZZ=60 timeout 600 python -m pytest tests/test_gpu_fuzz.py -x -q > gpurun_out/r2_gpu_fuzz.log 2>&1; echo "gpu fuzz: rc=$?"
tail -3 gpurun_out/r2_pytest_gpu.log gpurun_out/r2_block_event_tests.log gpurun_out/r2_gpu_fuzz.log
cat gpurun_out/r2_bench_*.json 2>/dev/null | cut -c1-300
