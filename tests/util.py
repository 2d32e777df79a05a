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


def synthetic_case(M: int, G: int, pins, mpfr: int, mpwr: int, seed: int = 0, boundl: float = 1.0, boundr: float = 1.0,
                   numass: int = 1):
    """A physically plausible random M-material / G-group slab problem through the product-side mesh_gen.

    Materials 0 and 1 are fuels (fissile, emit in the top groups); the rest moderate.  Scatter matrices are
    lower-triangular-heavy (down-scatter) with a little up-scatter, rows sum to SigS, SigA = SigT - SigS."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    sigt = np.zeros((G, M), f32); sigs = np.zeros((G, M), f32); mu = np.zeros((G, M), f32)
    sigf = np.zeros((G, M), f32); nut = np.zeros((G, M), f32); chit = np.zeros((G, M), f32)
    scat = np.zeros((M, G, G), f32)
    for m in range(M):
        fuel = m < 2
        for g in range(G):
            st = f32(rng.uniform(0.2, 0.6) * (1 + 2.0 * g / max(1, G - 1)))
            absorb = f32(rng.uniform(0.02, 0.15) * (1 + g) if fuel else rng.uniform(0.001, 0.02))
            ss = f32(st - absorb)
            row = rng.uniform(0.0, 1.0, G) * np.exp(-1.5 * np.abs(np.arange(G) - g - 0.7))
            row[:g] *= 0.02  # little up-scatter
            row = (row / row.sum() * ss).astype(f32)
            scat[m, g] = row
            sigt[g, m], sigs[g, m] = st, ss
            mu[g, m] = f32(rng.uniform(0.0, 0.5))
            if fuel:
                sigf[g, m] = f32(absorb * rng.uniform(0.4, 0.8))
                nut[g, m] = f32(rng.uniform(2.3, 2.9))
        if fuel:
            c = rng.uniform(0, 1, G) * np.exp(-2.0 * np.arange(G))
            chit[:, m] = (c / c.sum()).astype(f32)
    siga = (sigt - sigs).astype(f32)
    inv_sigtr = (f32(1.0) / (sigt - (mu * sigs).astype(f32)).astype(f32)).astype(f32)
    xs = nb.XSData(sigt=sigt.reshape(-1), sigs=sigs.reshape(-1), mu=mu.reshape(-1), siga=siga.reshape(-1), sigf=sigf.reshape(-1),
                   nut=nut.reshape(-1), chit=chit.reshape(-1), scat_matrix=scat.reshape(-1), inv_sigtr=inv_sigtr.reshape(-1))
    v = nb.Variables(analk=1, mattypes=M, energygroups=G, generations=4, histories=1000, skip=1, numass=numass, numrods=len(pins),
                     roddia=0.94, rodpitch=float(f32(1.262) - f32(0.94)), mpfr=mpfr, mpwr=mpwr, boundl=boundl, boundr=boundr)
    dx = nb.DeltaX(fuel=float(f32(0.94) / f32(mpfr)), water=float(f32(v.rodpitch) / f32(mpwr)) if mpwr else 1.0)
    mesh, fuel = nb.mesh_gen(np.asarray(pins, np.uint8), v, dx)
    return v, xs, dx, mesh, fuel
