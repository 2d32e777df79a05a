"""Generate tests/golden/mc_golden.json: small seeded runs of the CPU oracle on the three fixture decks.

The reference itself holds no end-to-end golden for the Monte Carlo path ("parity unpinned"),
so these vectors pin the ORACLE's behaviour (guarding it against regressions) and give the GPU parity tests a
committed target that does not depend on rebuilding the oracle.  Regenerate only when the numerics contract
 changes on purpose:  python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from tests.util import load_case, oracle_inputs  # noqa: E402

RUNS = [
    dict(case="a", generations=4, histories=20000, skip=1),
    dict(case="b", generations=3, histories=20000, skip=1),
    dict(case="c", generations=3, histories=20000, skip=1),
    dict(case="c", generations=3, histories=20000, skip=1, tracking_mode="woodcock"),
    dict(case="c", generations=4, histories=20000, skip=1, source_mode="fission_bank"),
    dict(case="b", generations=4, histories=20000, skip=1, source_mode="fission_bank", tracking_mode="woodcock"),
    dict(case="c", generations=2, histories=20000, skip=1, stale_xs=False),
    dict(case="c", generations=2, histories=20000, skip=1, scatter_mode="rust_pre182"),
]


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def record(run):
    kw = dict(run)
    case = kw.pop("case")
    deck, mesh = oracle_inputs(*load_case(case))
    r = orc.monte_carlo(deck, mesh, threads=4, want_tally=True, trace_gen=kw["generations"] - 1, **kw)
    return {
        "run": run,
        "k_bits": [int(v) for v in r.k.view(np.uint32)],
        "k_fund_bits": [int(v) for v in r.k_fund.view(np.uint32)],
        "tally_sha256": digest(r.tally_fixed),
        "tally_sum_per_generation": [int(v) for v in r.tally_fixed.reshape(r.tally_fixed.shape[0], -1).sum(axis=1)],
        "flux_sha256": digest(r.flux),
        "fission_source_sha256": digest(r.fission_source),
        "trace_sha256": digest(r.trace),
        "trace_head": r.trace[:8].tolist(),
        "counters": r.counters,
        "bank_sizes": [int(v) for v in r.bank_sizes],
    }


def main():
    out = {"format": 1, "tally_frac_bits": 28, "rng": {"seed": 42, "seq": 54, "stride": 152917}, "records": [record(r) for r in RUNS]}
    path = os.path.join(ROOT, "tests", "golden", "mc_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path, len(out["records"]), "records")


if __name__ == "__main__":
    main()
