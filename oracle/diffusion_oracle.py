"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's diffusion solver.

Follows ``src/discrete.rs`` of the reference statement by statement (f32 arithmetic, dense
inverse, sequential sums, the scoping quirks of ``matrix_gen`` and the convergence test of
``nalgebra_method``).  It exists to check ``nraps_diffusion_run`` (SURVEY section 8(f) rank 4:
an independent physics cross-check of the Monte Carlo path); never imported by ``nraps_b200/``.

parity unpinned: the reference inverts with nalgebra 0.32.5 (LU, f32), this file with LAPACK
``sgetri``; the reference holds no test or golden vector for this solver, and its toolchain is
absent here, so agreement is to rounding of an f32 dense inverse (~1e-4), not bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32


def _seq_sum(a, axis=-1):
    """Rust's ``iter().sum::<f32>()``: left-to-right f32 accumulation (np.sum is pairwise)."""
    a = np.asarray(a, dtype=f32)
    if a.shape[axis] == 0:
        return f32(0.0) if a.ndim == 1 else np.zeros(a.shape[:axis] + a.shape[axis + 1:], f32)
    return np.take(np.cumsum(a, axis=axis, dtype=f32), -1, axis=axis)


def _beta(bound, d_next, d_curr):
    """src/discrete.rs:26-33 and :86-93."""
    bound = f32(bound)
    if bound == f32(1.0):
        return f32(1.0)
    if bound == f32(0.0):
        return f32(0.25)
    one, q = f32(1.0), f32(0.25)
    r = (one - bound) / (one + bound)
    return (one - q * (r * (one / d_next))) / (one + q * (r * (one / d_curr)))


def matrix_gen(n, deck, cell_mat, dx, g, boundl, boundr):
    """src/discrete.rs:4-107; returns the dense n x n matrix of group ``g``."""
    M, G = int(deck.mattypes), int(deck.energygroups)
    inv = np.asarray(deck.inv_sigtr, f32)
    sigt = np.asarray(deck.sigt, f32)
    scat = np.asarray(deck.scat, f32)
    third = f32(1.0) / f32(3.0)  # 3.0_f32.powi(-1)
    two = f32(2.0)
    a = np.zeros((n, n), f32)

    def d_mul(i):  # :17-19 and :47-55: (1/3) * inv_sigtr * dx.powi(-1)
        return (third * inv[int(cell_mat[i]) + M * g]) * (f32(1.0) / f32(dx[i]))

    def removal(i):  # dx * (sigt - scat[g -> g])
        m = int(cell_mat[i])
        return f32(dx[i]) * (sigt[m + M * g] - scat[((G + 1) * g + G * G * m) & 0xFF])

    d_curr, d_next = d_mul(0), d_mul(1)
    d_nextcurr0 = (two * d_curr * d_next) * (f32(1.0) / (d_curr + d_next))
    beta_l = _beta(boundl, d_next, d_curr)
    a[0, 0] = two * d_curr * (f32(1.0) - beta_l) + removal(0) + d_nextcurr0
    a[0, 1] = -d_nextcurr0
    for x in range(1, n - 1):
        dc, dp, dn = d_mul(x), d_mul(x - 1), d_mul(x + 1)
        d_prevcurr = (two * dc * dp) * (f32(1.0) / (dc + dp))
        d_nextcurr = (two * dc * dn) * (f32(1.0) / (dc + dn))
        a[x, x - 1] = -d_prevcurr
        a[x, x] = d_prevcurr + removal(x) + d_nextcurr
        a[x, x + 1] = -d_nextcurr
    # :74-105 -- the last row divides instead of multiplying by the reciprocal, and reads d_next / d_nextcurr of
    # the FIRST block (cells 0 and 1): the loop's bindings are out of scope here
    d_curr_e = (third * inv[int(cell_mat[n - 1]) + M * g]) / f32(dx[n - 1])
    d_prev_e = (third * inv[int(cell_mat[n - 2]) + M * g]) / f32(dx[n - 2])
    d_prevcurr_e = (two * d_curr_e * d_prev_e) / (d_curr_e + d_prev_e)
    beta_r = _beta(boundr, d_next, d_curr_e)
    a[n - 1, n - 2] = -d_prevcurr_e
    a[n - 1, n - 1] = two * d_curr_e * (f32(1.0) - beta_r) + removal(n - 1) + d_nextcurr0
    return a


def q_gen(deck, cell_mat, dx, flux):
    """src/discrete.rs:109-134."""
    M, G = int(deck.mattypes), int(deck.energygroups)
    nut, sigf, chit = (np.asarray(v, f32) for v in (deck.nut, deck.sigf, deck.chit))
    m = np.asarray(cell_mat, np.int64)
    terms = np.stack([(nut[m + M * x] * sigf[m + M * x]) * flux[x] for x in range(G)], axis=1)  # [n][G]
    prod = _seq_sum(terms, axis=1)
    return np.stack([(prod * np.asarray(dx, f32)) * chit[m + M * g] for g in range(G)]).astype(f32)


def scat_calc(deck, cell_mat, dx, flux, g):
    """src/discrete.rs:136-160 for every cell at once."""
    G = int(deck.energygroups)
    scat = np.asarray(deck.scat, f32)
    m = np.asarray(cell_mat, np.int64)
    acc = np.zeros(len(m), f32)
    for e in range(G):
        if e != g:
            acc = acc + (scat[G * G * m + G * e + g] * flux[e]) * np.asarray(dx, f32)
    return acc


@dataclass
class DiffusionOutput:
    flux: np.ndarray
    assembly_average: np.ndarray
    k: float
    iterations: int


def nalgebra_method(deck, mesh, max_iterations=100000) -> DiffusionOutput:
    """src/discrete.rs:181-356."""
    cell_mat, dx = mesh[0], np.asarray(mesh[1], f32)
    n = len(cell_mat)
    M, G = int(deck.mattypes), int(deck.energygroups)
    flux = np.ones((G, n), f32)
    q = q_gen(deck, cell_mat, dx, flux)
    k, delta_flux, delta_k = f32(1.0), f32(1.0), f32(1.0)
    a_inv = [np.linalg.inv(matrix_gen(n, deck, cell_mat, dx, g, deck.boundl, deck.boundr)).astype(f32) for g in range(G)]
    it = 0
    while delta_flux >= f32(1e-5) and delta_k >= f32(1e-6) and it < max_iterations:
        it += 1
        temp_q = q.copy()
        for g in range(G):
            scat = scat_calc(deck, cell_mat, dx, flux, g)
            rhs = (q[g] * (f32(1.0) / k)) + scat
            new = _seq_sum(a_inv[g] * rhs[None, :], axis=1)
            # :236-288 -- index 0 and n-1 take a max with the running value, the indices between overwrite it,
            # and all three divide by flux[g][0] (already replaced for every index but the first)
            delta_flux = max(abs((flux[g][0] - new[0]) / flux[g][0]), delta_flux)
            if n > 2:
                delta_flux = abs((flux[g][n - 2] - new[n - 2]) / new[0])
            delta_flux = max(abs((flux[g][n - 1] - new[n - 1]) / new[0]), delta_flux)
            flux[g] = new
        q = q_gen(deck, cell_mat, dx, flux)
        temp_k = k
        k = temp_k * (_seq_sum(q.reshape(-1)) / _seq_sum(temp_q.reshape(-1)))
        delta_k = abs((k - temp_k) / temp_k)
    nut, sigf = np.asarray(deck.nut, f32), np.asarray(deck.sigf, f32)
    m = np.asarray(cell_mat, np.int64)
    temp = np.stack([(flux[g] * nut[m + M * g]) * sigf[m + M * g] for g in range(G)]).astype(f32)
    # :325-329 -- chunks(G) of the group-major flattening: G consecutive cells of one group, not the G groups of a cell
    power_flux = _seq_sum(temp.reshape(-1).reshape(n, G), axis=1)
    s = _seq_sum(power_flux * dx)
    power_constant = f32(3565e6) / (f32(1.6022e-13) * f32(200.0) * s)
    flux = (flux * power_constant).astype(f32)
    mesh_assembly = n // int(deck.numass)
    avg = np.zeros_like(flux)
    for g in range(G):
        for a in range(int(deck.numass)):
            lo, hi = a * mesh_assembly, (a + 1) * mesh_assembly
            avg[g, lo:hi] = _seq_sum(flux[g, lo:hi]) / f32(mesh_assembly)
    return DiffusionOutput(flux=flux, assembly_average=avg, k=float(k), iterations=it)
