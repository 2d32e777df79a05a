#!/bin/bash
# Race check of the block-event kernel's per-thread body without a GPU: the CPU-thread emulation (tests/emul) built with
# ThreadSanitizer.  A missing or misplaced __syncthreads in mc_block_event.cuh is a data race between the emulation's
# pthreads, which TSan reports.  The detector is shown to have teeth by mutation: with any one of the three barriers
# of a round skipped (BEV_EMUL_DROP_SYNC=0|1|2) the runs go wrong (tallies differ from the oracle, or the run crashes)
# and TSan reports the races it gets to see (how many depends on the interleavings of the run); with all three in
# place it must report none and the tallies must be bit-exact.
#   bash tools/tsan_block_event.sh            -> profiles-style summary on stdout
set -u
cd "$(dirname "$0")/.."
mkdir -p tests/emul/_build
nvcc -gencode arch=compute_100a,code=sm_100a -Iinclude -DNRAPS_EMUL -O1 -g -std=c++17 -fmad=false -diag-suppress 20011,20014,177 \
    -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-pthread,-fsanitize=thread -shared -cudart static \
    -o tests/emul/_build/libbev_emul_tsan.so tests/emul/block_event_emul.cu -Xlinker -ltsan < /dev/null || exit 1
TSAN=$(ldd tests/emul/_build/libbev_emul_tsan.so | awk '/libtsan/ {print $3}')
cat > /tmp/tsan_bev_run.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
import tests.test_block_event_emul as t
from tests.util import load_case
L = C.CDLL("tests/emul/_build/libbev_emul_tsan.so")
L.bev_emul_generation.argtypes = [C.POINTER(t.orc.Problem)] + [C.c_uint64] * 6 + [C.c_int32] * 2 + [C.c_uint32] * 6 + [C.POINTER(C.c_uint64)] * 3
ok = True
for case, kw in (("c", dict(blocks=2, threads=8, slots=40, chunk=16)), ("b", dict(blocks=1, threads=6, slots=13, chunk=5, walk_cap=3))):
    v, xs, dx, mesh, fuel = load_case(case)
    v.boundl = 0.5
    try:
        t._run(L, (v, xs, dx, mesh, fuel), H=600, gen=0, **kw)
    except AssertionError:
        ok = False
print("results bit-exact" if ok else "results DIFFER from the oracle")
PY
for drop in none 0 1 2; do
  if [ "$drop" = none ]; then unset BEV_EMUL_DROP_SYNC; tries=3; else export BEV_EMUL_DROP_SYNC=$drop; tries=3; fi
  races=0; bad=0
  for try in $(seq $tries); do  # which interleavings occur (and so what TSan can see) varies from run to run
    out=$(TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 halt_on_error=0" LD_PRELOAD=$TSAN timeout 600 python /tmp/tsan_bev_run.py < /dev/null 2>&1)
    races=$((races + $(echo "$out" | grep -c "WARNING: ThreadSanitizer: data race")))
    echo "$out" | grep -q "results bit-exact" || bad=$((bad + 1))
  done
  echo "barrier dropped: $drop -> $tries runs: $races data-race reports, $bad runs with wrong tallies, a crash or a hang"
done
