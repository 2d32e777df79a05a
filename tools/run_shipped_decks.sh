# time-to-solution of the three decks exactly as shipped (BASELINE configs 1 and 2): GPU driver vs CPU oracle
for c in a b c; do
  mkdir -p /tmp/out_$c
  echo "nraps case_$c surface:"; nraps_b200/lib/nraps tests/golden/decks/case_$c.txt --out /tmp/out_$c --quiet 2>&1 | tail -2
  echo "nraps case_$c woodcock:"; nraps_b200/lib/nraps tests/golden/decks/case_$c.txt --out /tmp/out_$c --quiet --tracking woodcock 2>&1 | tail -2
done
exit 0
python - <<PY
import time, numpy as np
from oracle import oracle as orc
from tests.util import load_case, oracle_inputs
import os
for c in "abc":
    deck, mesh = oracle_inputs(*load_case(c))
    t=time.time(); r = orc.monte_carlo(deck, mesh, threads=max(1,(os.cpu_count() or 2)-1), tally_mode="f32_per_worker"); dt=time.time()-t
    print(f"oracle case_{c}: {dt:.1f} s  k_fund[-1]={r.k_fund[-1]:.5f}  ({deck.histories*deck.generations/dt:.3g} histories/s)")
PY
