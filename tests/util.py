"""Shared helpers for the test-suite (test-only; may import oracle/)."""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

import nraps_b200 as nb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = {c: os.path.join(ROOT, "tests", "golden", "decks", f"case_{c}.txt") for c in "abc"}


def load_case(case: str, mpfr: int | None = None, mpwr: int | None = None):
    """Product-side pipeline up to the solver inputs; optional fine-mesh override."""
    v, xs, pins, dx, _, _ = nb.process_input(DECKS[case])
    if mpfr is not None:
        v.mpfr, v.mpwr = mpfr, mpwr
        dx = nb.DeltaX(fuel=float(np.float32(v.roddia) / np.float32(mpfr)), water=float(np.float32(v.rodpitch) / np.float32(mpwr)))
    mesh, fuel = nb.mesh_gen(pins, v, dx)
    return v, xs, dx, mesh, fuel


def oracle_inputs(v, xs, dx, mesh, fuel):
    """Adapt product-side objects to oracle.monte_carlo(deck, mesh)."""
    deck = SimpleNamespace(
        energygroups=v.energygroups, mattypes=v.mattypes, numass=v.numass, generations=v.generations,
        histories=v.histories, skip=v.skip, boundl=v.boundl, boundr=v.boundr, dx_fuel=dx.fuel, dx_water=dx.water,
        sigt=xs.sigt, sigs=xs.sigs, mu=xs.mu, siga=xs.siga, sigf=xs.sigf, nut=xs.nut, chit=xs.chit,
        scat=xs.scat_matrix, inv_sigtr=xs.inv_sigtr,
    )
    return deck, (mesh.matid, mesh.delta_x, mesh.mesh_left, mesh.mesh_right, fuel)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
