"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's host side.

Independent of the product's C++ host code (nraps_b200/csrc/host_*.cpp); the
tests compare the two.  Follows, relative to /root/reference:

* ``scan_ascii_chunk`` / ``get_index`` / ``process_input`` .. src/process_input.rs:6-175
* ``mesh_gen`` ............................................ src/main.rs:85-143
* CSV layout and float formatting ......................... src/plot_solution.rs:14-58

PARITY STATUS: the reference has no parser / mesh / CSV tests; the structural
fixtures asserted in tests/ (N, NF, L per deck) are SURVEY.md section 8c's.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32

_KEYS = {
    ("lk", 5): 0, ("es", 8): 1, ("ps", 12): 2, ("ns", 11): 3, ("es", 9): 4, ("ip", 4): 5,
    ("ss", 6): 6, ("ds", 7): 7, ("ia", 6): 8, ("ch", 8): 9, ("fr", 4): 10, ("wr", 4): 11,
    ("dl", 6): 12, ("dr", 6): 13, ("gt", 4): 14, ("gs", 4): 15, ("mu", 2): 16, ("ga", 4): 17,
    ("gf", 4): 18, ("ut", 3): 19, ("it", 4): 20, ("at", 4): 21, ("id", 5): 22, ("on", 8): 23,
    ("er", 6): 24,
}


def get_index(last2: str, length: int) -> int:
    """src/process_input.rs:13-42"""
    return _KEYS.get((last2, length), 25)


def scan_ascii_chunk(buf: bytes) -> list[str]:
    """src/process_input.rs:44-83, including the byte skipped after a comment line."""
    end = len(buf)
    temp = [""] * 26
    pos = line_start = name_end = val_start = 0
    while pos < end:
        c = buf[pos]
        if c == 0x23:  # '#'
            while pos < end and buf[pos] != 0x0A:
                pos += 1
            pos += 1
            line_start = pos
        elif c == 0x3D:  # '='
            name_end = pos - 1
            val_start = pos + 1
        elif c == 0x0A:
            if name_end > line_start:
                key = buf[line_start:name_end].decode("utf-8", "replace").strip().lower()
                value = buf[val_start:pos].decode("utf-8", "replace").strip()
                if len(key) < 2:
                    raise IndexError("key shorter than two characters: `&key[length - 2..length]` panics (:72)")
                temp[get_index(key[-2:], len(key))] += " " + value
            line_start = pos + 1
        pos += 1
    return temp


@dataclass
class Deck:
    analk: int
    mattypes: int
    energygroups: int
    generations: int
    histories: int
    skip: int
    numass: int
    numrods: int
    roddia: np.float32
    rodpitch: np.float32  # already RodPitch - RodDia (src/process_input.rs:102)
    mpfr: int
    mpwr: int
    boundl: np.float32
    boundr: np.float32
    dx_fuel: np.float32
    dx_water: np.float32
    sigt: np.ndarray
    sigs: np.ndarray
    mu: np.ndarray
    siga: np.ndarray
    sigf: np.ndarray
    nut: np.ndarray
    chit: np.ndarray
    scat: np.ndarray
    inv_sigtr: np.ndarray
    matid: np.ndarray
    solution: int
    solver: int


def _floats(s: str) -> np.ndarray:
    return np.array([f32(t) for t in s.split()], dtype=f32)


def process_input(path: str) -> Deck:
    """src/process_input.rs:85-175 (the deck path is an argument, not ./TestCaseC.txt)."""
    with open(path, "rb") as fh:
        t = scan_ascii_chunk(fh.read())
    roddia = f32(t[8].strip())
    rodpitch = f32(f32(t[9].strip()) - roddia)
    mpfr, mpwr = int(t[10]), int(t[11])
    sigt, sigs, mu = _floats(t[14]), _floats(t[15]), _floats(t[16])
    if len(mu) < len(sigt) or len(sigs) < len(sigt):
        raise IndexError("mu / sigs shorter than sigt: the loop at src/process_input.rs:152-156 indexes out of bounds")
    n = len(sigt)  # :152 iterates over sigt.len(); longer mu / sigs are simply not read
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_sigtr = (f32(1.0) / (sigt - (mu[:n] * sigs[:n]).astype(f32)).astype(f32)).astype(f32)
    return Deck(
        analk=int(t[0]), mattypes=int(t[1]), energygroups=int(t[2]), generations=int(t[3]),
        histories=int(t[4]), skip=int(t[5]), numass=int(t[6]), numrods=int(t[7]),
        roddia=roddia, rodpitch=rodpitch, mpfr=mpfr, mpwr=mpwr,
        boundl=f32(t[12].strip()), boundr=f32(t[13].strip()),
        dx_fuel=f32(roddia / f32(mpfr)), dx_water=f32(rodpitch / f32(mpwr)),
        sigt=sigt, sigs=sigs, mu=mu, siga=_floats(t[17]), sigf=_floats(t[18]), nut=_floats(t[19]),
        chit=_floats(t[20]), scat=_floats(t[21]), inv_sigtr=inv_sigtr,
        matid=np.array([int(v) for v in t[22].split()], dtype=np.uint8),
        solution=int(t[23]), solver={"1": 1, "2": 2, "3": 3}.get(t[24].strip(), 0),  # :168-173: anything else is LinAlg
    )


def mesh_gen(matid, mpfr: int, mpwr: int, numass: int, dx_fuel, dx_water):
    """src/main.rs:85-143 -> (cell matid u8[N], dx, left, right f32[N], fuel_indices u64[NF])."""
    temp: list[int] = []
    for m in matid:
        temp.extend([int(m)] * (mpfr if m in (0, 1) else mpwr))
    for index1 in range(1, numass):
        for _ in range(mpwr):
            del temp[(index1 * len(temp)) // numass]
    if len(temp) < mpwr // 2:
        raise IndexError("drain(0..mpwr/2) past the end panics (src/main.rs:107)")
    del temp[0 : mpwr // 2]
    if len(temp) < mpwr // 2:
        # `temp.len() - mpwr/2` underflows (:108): a panic with overflow checks, a no-op truncate without; no mesh
        # of a real deck gets here and the product rejects it as a shape error
        raise IndexError("mesh shorter than the edge trim")
    del temp[len(temp) - (mpwr // 2) :]
    fuel = np.array([i for i, m in enumerate(temp) if m in (0, 1)], dtype=np.uint64)
    n = len(temp)
    dx = np.empty(n, f32)
    left = np.empty(n, f32)
    right = np.empty(n, f32)
    mesh_left = f32(0.0)
    for i, m in enumerate(temp):
        d = f32(dx_fuel) if m in (0, 1) else f32(dx_water)
        dx[i] = d
        left[i] = mesh_left
        right[i] = f32(mesh_left + d)
        mesh_left = f32(mesh_left + d)
    return np.array(temp, dtype=np.uint8), dx, left, right, fuel


def rust_f32_display(v) -> str:
    """Rust ``f32::to_string``: shortest round-trip digits, never an exponent."""
    v = f32(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def rust_f64_display(v) -> str:
    v = np.float64(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def csv_files(flux, assembly_average, fission_source, k, k_fund, length_f32, n_mesh: int, gens: int):
    """src/plot_solution.rs:14-58 -> dict name -> file text."""
    rows = []
    for g in range(flux.shape[0]):
        rows.append(",".join(rust_f32_display(v) for v in flux[g]))
    for g in range(assembly_average.shape[0]):
        rows.append(",".join(rust_f32_display(v) for v in assembly_average[g]))
    rows.append(",".join(rust_f32_display(v) for v in fission_source))
    return {
        "vars.csv": f"{rust_f64_display(np.float64(f32(length_f32)))}\n{n_mesh}\n{gens}\n",
        "interface.csv": "\n".join(rows) + "\n",
        "k_eff.csv": ",".join(rust_f32_display(v) for v in k) + "\n" + ",".join(rust_f32_display(v) for v in k_fund) + "\n",
    }
