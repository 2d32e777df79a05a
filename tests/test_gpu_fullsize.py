"""Full-size parity: the configurations BASELINE.json names, at the sizes it names, bit-compared with the oracle.

* configs 1 and 2: the three decks exactly as shipped (`TestCaseA.txt:34-36`: 100 generations x 10^5 histories,
  skip 4; `TestCaseB.txt:34-36`, `TestCaseC.txt:34-36`: 100 x 10^6, skip 1), through the drop-in call
  `monte_carlo(variables, xsdata, deltax, meshid, fuel_indices, 1.0)` with nothing overridden: every tally bin of
  every generation, k, k_fund, flux, fission source, assembly averages and the bytes of the three CSV files;
* one generation of config 3 (10^7 histories) and one of config 4's fine mesh (N = 4080, 10^6 histories): tally
  bins, k and flux.

The oracle runs on all host cores (about one minute in total on the GPU box); its fixed-point tally is an integer
sum, so the thread count does not change a bit of it.
"""
import os

import numpy as np
import pytest

import nraps_b200 as nb
from oracle import host_oracle as ho
from oracle import oracle as orc
from tests.util import bits, load_case, oracle_inputs

pytestmark = pytest.mark.gpu
THREADS = max(1, (os.cpu_count() or 2) - 1)


def _identical(got, want):
    assert np.array_equal(got.tally_fixed, want.tally_fixed)
    for name in ("k", "k_fund", "flux", "assembly_average", "fission_source"):
        assert np.array_equal(bits(getattr(got, name)), bits(getattr(want, name))), name
    for c in ("histories", "collisions", "flights", "leaks", "truncated"):
        assert got.counters[c] == want.counters[c], c


@pytest.mark.timeout(900)
@pytest.mark.parametrize("case,gens,H,skip", [("a", 100, 100_000, 4), ("b", 100, 1_000_000, 1), ("c", 100, 1_000_000, 1)])
def test_deck_as_shipped_bit_exact(case, gens, H, skip, tmp_path):
    v, xs, dx, mesh, fuel = load_case(case)
    assert (v.generations, v.histories, v.skip) == (gens, H, skip)  # the fixture decks carry the reference's run sizes
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, want_tally=True)
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, threads=THREADS, want_tally=True)
    _identical(got, want)
    assert got.k_fund[gens - 1] == want.k_fund[gens - 1] and np.isfinite(got.k).all()
    nb.plot_solution(got, v.energygroups, gens, len(mesh), float(mesh.mesh_right[-1]), str(tmp_path))
    files = ho.csv_files(want.flux, want.assembly_average, want.fission_source, want.k, want.k_fund, mesh.mesh_right[-1], len(mesh), gens)
    for name, text in files.items():
        assert (tmp_path / name).read_text() == text, name


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fine,H", [(False, 10_000_000), (True, 1_000_000)])
def test_one_generation_of_config3_and_config4_bit_exact(fine, H):
    """Generation 1 of config 3 (N = 408, 10^7 histories) and of config 4's mesh (N = 4080, 10^6 histories)."""
    v, xs, dx, mesh, fuel = load_case("c", mpfr=80, mpwr=40) if fine else load_case("c")
    assert len(mesh) == (4080 if fine else 408)
    with nb.MonteCarloContext(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=H, skip=0) as ctx:
        ctx.transport(1)
        tally, counters = ctx.read_tally()
    deck, m = oracle_inputs(v, xs, dx, mesh, fuel)
    want = orc.monte_carlo(deck, m, generations=2, histories=H, skip=0, threads=THREADS, want_tally=True)
    assert np.array_equal(tally, want.tally_fixed[1])
    # the oracle ran generations 0 and 1; its counters cover both, generation 1's share follows from the tally identity
    got = nb.monte_carlo(v, xs, dx, mesh, fuel, 1.0, generations=2, histories=H, skip=0, want_tally=True)
    _identical(got, want)
