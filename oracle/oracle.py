"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_build/liboracle.so.

May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by nraps_b200/.  Build the library with ``make -C oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

TR_WORDS = 10
TR_NAMES = ["collisions", "crossings", "flights", "reflections", "rng_lo", "rng_hi", "cell", "xbits", "fate", "group"]
CT_WORDS = 8
CT_NAMES = ["histories", "collisions", "crossings", "flights", "reflections", "leaks", "truncated", "banked"]
TALLY_FRAC_BITS = 28

SCATTER_MODES = {"single_xi": 0, "rust_pre182": 1, "rust_182": 2}
TALLY_MODES = {"fixed64": 0, "f32_per_worker": 1}
SOURCE_MODES = {"uniform_fuel": 0, "fission_bank": 1}
TRACKING_MODES = {"surface": 0, "woodcock": 1}

_fp = C.POINTER(C.c_float)


class Problem(C.Structure):
    _fields_ = [
        ("M", C.c_uint32), ("G", C.c_uint32), ("N", C.c_uint32), ("NF", C.c_uint32), ("numass", C.c_uint32),
        ("generations", C.c_uint64), ("histories", C.c_uint64), ("skip", C.c_uint64),
        ("boundl", C.c_float), ("boundr", C.c_float), ("dx_fuel", C.c_float), ("dx_water", C.c_float),
        ("k0", C.c_float),
        ("sigt", _fp), ("sigs", _fp), ("mu", _fp), ("siga", _fp), ("sigf", _fp), ("nut", _fp), ("chit", _fp),
        ("inv_sigtr", _fp), ("scat", _fp),
        ("matid", C.POINTER(C.c_uint8)),
        ("dx", _fp), ("left", _fp), ("right", _fp),
        ("fuel_indices", C.POINTER(C.c_uint64)),
    ]


class Options(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("seq", C.c_uint64), ("stride", C.c_uint64),
        ("scatter_mode", C.c_int32), ("stale_xs", C.c_int32), ("tally_mode", C.c_int32),
        ("inclusive_ranges", C.c_int32), ("threads", C.c_int32), ("source_mode", C.c_int32),
        ("tracking_mode", C.c_int32), ("bank_cap", C.c_int32),
        ("hist_begin", C.c_uint64), ("hist_count", C.c_uint64), ("max_flights", C.c_uint64),
    ]


class Results(C.Structure):
    _fields_ = [
        ("flux", _fp), ("assembly_average", _fp), ("fission_source", _fp), ("k", _fp), ("k_fund", _fp),
        ("tally_fixed", C.POINTER(C.c_uint64)), ("trace", C.POINTER(C.c_uint32)), ("trace_gen", C.c_uint64),
        ("counters", C.c_uint64 * CT_WORDS), ("bank_sizes", C.POINTER(C.c_uint64)),
        ("bank_sites", C.POINTER(C.c_uint64)), ("bank_sites_cap", C.c_uint64), ("bank_gen", C.c_uint64),
        ("entropy", C.POINTER(C.c_double)),
        ("seconds_transport", C.c_double),
    ]


_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_monte_carlo.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Results)]
        L.oracle_monte_carlo.restype = C.c_int
        L.oracle_hit_boundary.argtypes = [C.c_float] * 5 + [_fp]
        L.oracle_cross_mesh.argtypes = [C.c_uint64] + [C.c_float] * 4 + [_fp, C.POINTER(C.c_uint64)]
        L.oracle_direction_f.argtypes = [C.c_float]
        L.oracle_direction_f.restype = C.c_float
        L.oracle_scat_mat_calc.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _fp, _fp]
        L.oracle_energy_search.argtypes = [_fp, C.c_uint32, C.c_float]
        L.oracle_energy_search.restype = C.c_uint32
        L.oracle_pcg32_demo.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_pcg32_state.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_logf_f.argtypes = [C.c_float]
        L.oracle_logf_f.restype = C.c_float
        L.oracle_unit_f.argtypes = [C.c_uint32]
        L.oracle_unit_f.restype = C.c_float
        L.oracle_logf_max_ulp.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_logf_max_ulp.restype = C.c_double
        L.oracle_average_assembly.argtypes = [_fp, C.c_uint32, C.c_uint32, C.c_uint32, _fp]
        L.oracle_k_fund.argtypes = [_fp, C.c_uint64, C.c_uint64, _fp]
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class OracleOutput:
    flux: np.ndarray
    assembly_average: np.ndarray
    fission_source: np.ndarray
    k: np.ndarray
    k_fund: np.ndarray
    tally_fixed: np.ndarray | None
    trace: np.ndarray | None
    counters: dict
    bank_sizes: np.ndarray | None
    seconds_transport: float
    threads: int = 0
    bank_sites: np.ndarray | None = None
    entropy: np.ndarray | None = None
    extra: dict = field(default_factory=dict)


def monte_carlo(deck, mesh, *, generations=None, histories=None, skip=None, k0=1.0,
                seed=42, seq=54, stride=152917, scatter_mode="single_xi", stale_xs=True,
                tally_mode="fixed64", inclusive_ranges=False, threads=1, source_mode="uniform_fuel",
                tracking_mode="surface", hist_begin=0, hist_count=0, max_flights=0,
                want_tally=False, trace_gen=None, bank_cap=0, bank_gen=None) -> OracleOutput:
    """Run the CPU oracle.  ``deck`` is a host_oracle.Deck-like object (attribute
    access), ``mesh`` the tuple returned by ``mesh_gen``."""
    cell_mat, dx, left, right, fuel = mesh
    gens = int(deck.generations if generations is None else generations)
    H = int(deck.histories if histories is None else histories)
    sk = int(deck.skip if skip is None else skip)
    G, M, N = int(deck.energygroups), int(deck.mattypes), len(cell_mat)
    keep = dict(
        sigt=_f(deck.sigt), sigs=_f(deck.sigs), mu=_f(deck.mu), siga=_f(deck.siga), sigf=_f(deck.sigf),
        nut=_f(deck.nut), chit=_f(deck.chit), inv_sigtr=_f(deck.inv_sigtr), scat=_f(deck.scat),
        matid=np.ascontiguousarray(cell_mat, dtype=np.uint8), dx=_f(dx), left=_f(left), right=_f(right),
        fuel=np.ascontiguousarray(fuel, dtype=np.uint64),
    )
    p = Problem(
        M=M, G=G, N=N, NF=len(fuel), numass=int(deck.numass), generations=gens, histories=H, skip=sk,
        boundl=float(deck.boundl), boundr=float(deck.boundr), dx_fuel=float(deck.dx_fuel),
        dx_water=float(deck.dx_water), k0=float(k0),
        sigt=_p(keep["sigt"]), sigs=_p(keep["sigs"]), mu=_p(keep["mu"]), siga=_p(keep["siga"]),
        sigf=_p(keep["sigf"]), nut=_p(keep["nut"]), chit=_p(keep["chit"]), inv_sigtr=_p(keep["inv_sigtr"]),
        scat=_p(keep["scat"]), matid=_p(keep["matid"], C.c_uint8), dx=_p(keep["dx"]), left=_p(keep["left"]),
        right=_p(keep["right"]), fuel_indices=_p(keep["fuel"], C.c_uint64),
    )
    o = Options(
        seed=seed, seq=seq, stride=stride, scatter_mode=SCATTER_MODES[scatter_mode], stale_xs=int(bool(stale_xs)),
        tally_mode=TALLY_MODES[tally_mode], inclusive_ranges=int(bool(inclusive_ranges)), threads=int(threads),
        source_mode=SOURCE_MODES[source_mode], tracking_mode=TRACKING_MODES[tracking_mode], bank_cap=int(bank_cap),
        hist_begin=int(hist_begin), hist_count=int(hist_count), max_flights=int(max_flights),
    )
    nh = int(hist_count) if hist_count else H
    flux = np.zeros((G, N), np.float32)
    avg = np.zeros((G, N), np.float32)
    fis = np.zeros(N, np.float32)
    k = np.zeros(gens, np.float32)
    kf = np.zeros(gens, np.float32)
    tally = np.zeros((gens, G, N), np.uint64) if want_tally else None
    trace = np.zeros((nh, TR_WORDS), np.uint32) if trace_gen is not None else None
    banks = np.zeros(gens, np.uint64)
    entropy = np.zeros(gens, np.float64)
    sites_cap = nh * (int(bank_cap) or 8) if bank_gen is not None else 0
    sites = np.zeros(max(1, sites_cap), np.uint64)
    r = Results(
        flux=_p(flux), assembly_average=_p(avg), fission_source=_p(fis), k=_p(k), k_fund=_p(kf),
        tally_fixed=_p(tally, C.c_uint64) if tally is not None else None,
        trace=_p(trace, C.c_uint32) if trace is not None else None,
        trace_gen=int(trace_gen or 0), bank_sizes=_p(banks, C.c_uint64),
        bank_sites=_p(sites, C.c_uint64) if bank_gen is not None else None, bank_sites_cap=sites_cap,
        bank_gen=int(bank_gen or 0), entropy=_p(entropy, C.c_double),
    )
    rc = lib().oracle_monte_carlo(C.byref(p), C.byref(o), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"oracle_monte_carlo failed: {rc}")
    return OracleOutput(
        flux=flux, assembly_average=avg, fission_source=fis, k=k, k_fund=kf, tally_fixed=tally, trace=trace,
        counters={n: int(r.counters[i]) for i, n in enumerate(CT_NAMES)}, bank_sizes=banks,
        seconds_transport=float(r.seconds_transport), threads=int(threads),
        bank_sites=sites[: int(banks[bank_gen])] if bank_gen is not None else None, entropy=entropy,
    )
