"""Committed golden vectors (tests/golden/mc_golden.json, made by tools/make_golden.py from the CPU oracle):
the oracle must keep reproducing them (CPU) and the CUDA path must hit them bit for bit (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.util import ROOT, load_case, oracle_inputs

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "mc_golden.json")))


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check(rec, k, k_fund, tally, flux, fission, counters, bank_sizes, trace=None):
    assert [int(v) for v in k.view(np.uint32)] == rec["k_bits"]
    assert [int(v) for v in k_fund.view(np.uint32)] == rec["k_fund_bits"]
    assert _digest(tally) == rec["tally_sha256"]
    assert [int(v) for v in tally.reshape(tally.shape[0], -1).sum(axis=1)] == rec["tally_sum_per_generation"]
    assert _digest(flux) == rec["flux_sha256"] and _digest(fission) == rec["fission_source_sha256"]
    for name in ("histories", "collisions", "flights", "leaks", "truncated", "banked"):
        assert counters[name] == rec["counters"][name], name
    assert [int(v) for v in bank_sizes] == rec["bank_sizes"]
    if trace is not None:
        assert _digest(trace) == rec["trace_sha256"] and trace[:8].tolist() == rec["trace_head"]


@pytest.mark.parametrize("idx", range(len(GOLD["records"])))
def test_oracle_reproduces_golden(idx):
    from oracle import oracle as orc

    rec = GOLD["records"][idx]
    kw = dict(rec["run"])
    deck, mesh = oracle_inputs(*load_case(kw.pop("case")))
    r = orc.monte_carlo(deck, mesh, threads=3, want_tally=True, trace_gen=kw["generations"] - 1, **kw)
    _check(rec, r.k, r.k_fund, r.tally_fixed, r.flux, r.fission_source, r.counters, r.bank_sizes, r.trace)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(GOLD["records"])))
def test_gpu_reproduces_golden(idx):
    import nraps_b200 as nb

    rec = GOLD["records"][idx]
    kw = dict(rec["run"])
    args = load_case(kw.pop("case"))
    gens, H, skip = kw.pop("generations"), kw.pop("histories"), kw.pop("skip")
    r = nb.monte_carlo(*args, 1.0, generations=gens, histories=H, skip=skip, want_tally=True, **kw)
    _check(rec, r.k, r.k_fund, r.tally_fixed, r.flux, r.fission_source, r.counters, r.bank_sizes)
