"""TEST INFRASTRUCTURE ONLY -- second, pure-Python restatement of the Monte Carlo path.

Purpose: the C oracle (oracle_mc.c) is the checker of the CUDA path, but nothing
in the reference pins its end-to-end output ("parity unpinned": no Rust
toolchain here, unseedable RNG at HEAD).  This file restates the reference a
second time, function by function with the reference's own decomposition and
names, so that tests/test_oracle_restatement.py can demand BIT-identical
per-history records, tallies, k and flux from two separately written
restatements.  A transcription slip in either one shows up as a mismatch.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Arithmetic is IEEE binary32 through numpy scalars; there is
no vectorisation and no shortcut, so it is only usable on a few hundred
histories.  Only tests/ may import it.

What is NOT from the reference (it has no seedable stream): the
random source handed to these functions.  `Stream` below is rand.rs's PCG32
(src/rand.rs:49-100) with an explicit seed, the Q2 uniform mapping, one
sub-stream per history (hid * stride draws into the master stream) and the
one-draw source-cell rule -- the same conventions DESIGN.md section 2 states for
the oracle and the product.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

f32 = np.float32
MASK64 = (1 << 64) - 1
PCG_MULT = 6364136223846793005

ZERO, ONE, TWO = f32(0.0), f32(1.0), f32(2.0)


# ---------------------------------------------------------------------------------------------
# random numbers
# ---------------------------------------------------------------------------------------------
class PCG32:
    """src/rand.rs:36-85; `new` takes the seed as an argument instead of the wall clock (:50-53)."""

    def __init__(self, seed: int, thread_id: int):
        self.state = 0                                   # :57
        self.inc = ((thread_id << 1) | 1) & MASK64       # :58
        self.next_u32()                                  # :62
        self.state = (seed + self.inc) & MASK64          # :65  (state == inc at this point upstream too)
        self.next_u32()                                  # :68

    def next_u32(self) -> int:                           # :74-85
        old = self.state
        self.state = (old * PCG_MULT + self.inc) & MASK64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF

    def advance(self, delta: int) -> None:
        """Jump `delta` draws ahead: the n-th power of the affine map s -> a s + c by repeated squaring
        (the published pcg_advance_lcg_64 algorithm; rand.rs has no jump-ahead)."""
        a, c = PCG_MULT, self.inc
        acc_a, acc_c = 1, 0
        while delta:
            if delta & 1:
                acc_a, acc_c = (acc_a * a) & MASK64, (acc_c * a + c) & MASK64
            a, c = (a * a) & MASK64, ((a + 1) * c) & MASK64
            delta >>= 1
        self.state = (acc_a * self.state + acc_c) & MASK64


def unit_from_u32(u: int) -> np.float32:
    """((u >> 9) + 0.5) * 2^-23, replaces rand 0.8.5's `random::<f32>()` (24-bit, can be 0 or 0.5)."""
    return f32((u >> 9) + 0.5) * f32(2.0 ** -23)


class Stream:
    """What `random::<f32>()` / `thread_rng().gen_range` (src/mc_code.rs:46-51,122,125,148,195,209) become."""

    def __init__(self, seed: int, seq: int, stride: int, hid: int):
        self.rng = PCG32(seed, seq)
        self.rng.advance(hid * stride)

    def random(self) -> np.float32:
        return unit_from_u32(self.rng.next_u32())

    def gen_range(self, n: int) -> int:
        return (self.rng.next_u32() * n) >> 32


def _fmaf(a, b, c) -> np.float32:
    """Correctly rounded binary32 fused multiply-add: exact product in binary64, round-to-odd sum, one final rounding."""
    p = float(a) * float(b)
    cc = float(c)
    s = p + cc
    bb = s - p
    err = (p - (s - bb)) + (cc - bb)
    if err != 0.0:
        (bits,) = struct.unpack("<q", struct.pack("<d", s))
        if (bits & 1) == 0:
            s = float(np.nextafter(s, np.inf if err > 0.0 else -np.inf))
    return f32(s)


def ln(x: np.float32) -> np.float32:
    """Stands in for `f32::ln` (src/mc_code.rs:148,209).  The reference defers to the platform libm, which is not
    bit-specified; oracle and product share one polynomial instead, restated here from its
    description: x = m 2^e with m in (sqrt(1/2), sqrt 2], degree-8 polynomial in f = m - 1, Cephes split of ln 2."""
    (ix,) = struct.unpack("<I", struct.pack("<f", float(x)))
    e = (ix >> 23) - 127
    m = f32(struct.unpack("<f", struct.pack("<I", (ix & 0x007FFFFF) | 0x3F800000))[0])
    if m > f32(1.41421356):
        m = m * f32(0.5)
        e += 1
    f = m - ONE
    z = f * f
    p = f32(7.0376836292e-2)
    for coef in (-1.1514610310e-1, 1.1676998740e-1, -1.2420140846e-1, 1.4249322787e-1, -1.6668057665e-1,
                 2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1):
        p = _fmaf(p, f, f32(coef))
    y = (f * z) * p
    fe = f32(e)
    y = _fmaf(fe, f32(-2.12194440e-4), y)
    y = _fmaf(f32(-0.5), z, y)
    r = f + y
    return _fmaf(fe, f32(0.693359375), r)


# ---------------------------------------------------------------------------------------------
# the reference's types (src/main.rs:22-83), as far as the path reads them
# ---------------------------------------------------------------------------------------------
@dataclass
class Variables:
    mattypes: int
    energygroups: int
    generations: int
    histories: int
    skip: int
    numass: int
    boundl: np.float32
    boundr: np.float32


@dataclass
class XSData:
    sigt: np.ndarray
    sigs: np.ndarray
    mu: np.ndarray
    siga: np.ndarray
    sigf: np.ndarray
    nut: np.ndarray
    chit: np.ndarray
    scat_matrix: np.ndarray
    inv_sigtr: np.ndarray


@dataclass
class Mesh:
    matid: int
    delta_x: np.float32
    mesh_left: np.float32
    mesh_right: np.float32


@dataclass
class Switches:
    """The result-changing readings of SURVEY 9-B that the oracle also exposes."""
    scatter_mode: str = "single_xi"   # Q3: single_xi | rust_pre182 | rust_182
    stale_xs: bool = True             # Q1
    inclusive_ranges: bool = False    # Q4 (the reference is inclusive; the parity set is not)
    threads: int = 1
    seed: int = 42
    seq: int = 54
    stride: int = 152917


# ---------------------------------------------------------------------------------------------
# src/mc_code.rs, function by function
# ---------------------------------------------------------------------------------------------
def partition_point_min(values, probe, mode: str) -> int:
    """`v.partition_point(|&x| x < probe()).min(v.len() - 1)` (src/mc_code.rs:31,124-126).

    `probe` is called once per comparison, as the closure at :125 is.  With a constant probe all three orders
    return the lower bound; with a fresh random per comparison (:125) the visiting order matters (Q3):
      rust_pre182 -- core::slice::binary_search_by up to rustc 1.81 (mid = left + size/2, halve towards the hit)
      rust_182    -- the branchless loop of rustc >= 1.82 (base/half, one final comparison)
      single_xi   -- one draw, then a plain lower bound (the intended sampling)
    """
    n = len(values)
    if mode == "single_xi":
        xi = probe()
        return partition_point_min(values, lambda: xi, "rust_pre182")
    if mode == "rust_pre182":
        size, left, right = n, 0, n
        while left < right:
            mid = left + size // 2
            if values[mid] < probe():
                left = mid + 1
            else:
                right = mid
            size = right - left
        return min(left, n - 1)
    if mode == "rust_182":
        size, base = n, 0
        while size > 1:
            half = size // 2
            mid = base + half
            if values[mid] < probe():
                base = mid
            size -= half
        return min(base + (1 if values[base] < probe() else 0), n - 1)
    raise ValueError(mode)


def energy(chi, index, variables: Variables, xsdata: XSData, meshid) -> int:
    """src/mc_code.rs:7-32"""
    skip, step = meshid[index].matid, variables.mattypes
    cumulative = ZERO
    chit = []
    for value in xsdata.chit[skip::step]:
        cumulative = cumulative + f32(value)
        chit.append(cumulative)
    return partition_point_min(chit, lambda: chi, "single_xi")


def direction(mu):
    """src/mc_code.rs:35-37"""
    return TWO * mu - ONE


def spawn_neutron(fuel_indices, variables, xsdata, meshid, rnd: Stream):
    """src/mc_code.rs:40-53; tuple fields are evaluated left to right: cell, position, mu, chi."""
    index = int(fuel_indices[rnd.gen_range(len(fuel_indices))])
    return index, rnd.random(), direction(rnd.random()), energy(rnd.random(), index, variables, xsdata, meshid)


def hit_boundary(mu, start_x, delta_s, bound, mesh_end):
    """src/mc_code.rs:56-62"""
    return mu * (-bound), (delta_s + (start_x - mesh_end)) * (-bound), mesh_end


def cross_mesh(mesh_index, mu, start_x, mesh_end, delta_s):
    """src/mc_code.rs:65-79"""
    mesh_index = mesh_index + 1 if mu >= ZERO else mesh_index - 1
    return delta_s + (start_x - mesh_end), mesh_end, mesh_index


def scat_mat_calc(energygroups, matid, neutron_energy, inv_sigs, scat_matrix):
    """src/mc_code.rs:82-111 (index arithmetic widened from u8, Q13)"""
    base_idx = energygroups ** 2 * matid + energygroups * neutron_energy
    cumulative = ZERO
    out = []
    for e in range(energygroups):
        cumulative = cumulative + f32(scat_matrix[base_idx + e])
        out.append(cumulative * inv_sigs)
    return out


def interaction(interaction_xi, scat_mat, xsdata, xs_index, neutron_energy, rnd: Stream, mode: str, after_draws=None):
    """src/mc_code.rs:114-132: mu and the group are drawn before the absorption test.  `after_draws` (fission_bank
    mode only, not in the reference) runs between the draws and the test, where the bank draw sits in the stream."""
    absorption = f32(xsdata.siga[xs_index]) / f32(xsdata.sigt[xs_index])
    mu = TWO * rnd.random() - ONE
    scatter_energy = partition_point_min(scat_mat, rnd.random, mode)
    if after_draws is not None:
        after_draws()
    if interaction_xi < absorption:
        return False, neutron_energy, ZERO
    return True, scatter_energy, mu


class Events:
    """Per-history bookkeeping the replay tests compare (not in the reference)."""

    def __init__(self):
        self.collisions = self.crossings = self.flights = self.reflections = 0
        self.fate = 0  # 1 absorbed, 2 leaked


def particle_travel(tally, meshid, mesh_index, neutron_energy, mu, start_x, mattypes, energygroups, boundr, boundl,
                    xsdata, rnd: Stream, sw: Switches, ev: Events, bank=None):
    """src/mc_code.rs:134-213.  `tally(g, cell, v)` stands for `tally[g][cell] += v`."""
    xs_index = meshid[mesh_index].matid + mattypes * neutron_energy                      # :147
    delta_s = mu * -ln(rnd.random()) * f32(xsdata.inv_sigtr[xs_index])                    # :148
    ev.flights += 1
    same_material = True
    while same_material:
        end_x = start_x + delta_s
        mesh_end = meshid[mesh_index].mesh_right if mu >= ZERO else meshid[mesh_index].mesh_left
        if (mu < ZERO and mesh_end > end_x and mesh_index == 0) or \
                (mu >= ZERO and end_x > mesh_end and mesh_index == len(meshid) - 1):       # :159-160
            tally(neutron_energy, mesh_index, abs((start_x - mesh_end) / mu))
            bound = boundr if mu >= ZERO else boundl
            if bound > ZERO:
                mu, delta_s, start_x = hit_boundary(mu, start_x, delta_s, bound, mesh_end)
                ev.reflections += 1
            else:
                ev.fate = 2
                return False, mesh_index, mu, neutron_energy, start_x
        elif abs(end_x - start_x) > abs(mesh_end - start_x):                               # :171
            tally(neutron_energy, mesh_index, abs((start_x - mesh_end) / mu))
            prev_mat = meshid[mesh_index].matid
            delta_s, start_x, mesh_index = cross_mesh(mesh_index, mu, start_x, mesh_end, delta_s)
            ev.crossings += 1
            if prev_mat != meshid[mesh_index].matid:
                same_material = False
        else:                                                                              # :182
            tally(neutron_energy, mesh_index, abs((start_x - end_x) / mu))
            ev.collisions += 1
            with np.errstate(divide="ignore", invalid="ignore"):  # SigS = 0 (control rod): 1/0 and 0*inf, as upstream
                scat_mat = scat_mat_calc(energygroups, meshid[mesh_index].matid, neutron_energy,
                                         ONE / f32(xsdata.sigs[xs_index]), xsdata.scat_matrix)
            hook = None
            if bank is not None:
                hook = lambda: bank(mesh_index, end_x, meshid[mesh_index].matid, neutron_energy, xs_index, rnd)  # noqa: E731
            alive, new_energy, new_mu = interaction(rnd.random(), scat_mat, xsdata, xs_index, neutron_energy, rnd,
                                                    sw.scatter_mode, hook)
            if not alive:
                ev.fate = 1
                return False, mesh_index, mu, neutron_energy, start_x
            start_x, neutron_energy, mu = end_x, new_energy, new_mu
            if not sw.stale_xs:  # Q1 switch: the reference never refreshes xs_index here
                xs_index = meshid[mesh_index].matid + mattypes * neutron_energy
            delta_s = mu * -ln(rnd.random()) * f32(xsdata.inv_sigtr[xs_index])             # :209
            ev.flights += 1
    return True, mesh_index, mu, neutron_energy, start_x


def particle_lifetime(xsdata, meshid, fuel_indices, variables, delta_x_fuel, start, end, gen, sw: Switches, trace):
    """src/mc_code.rs:215-257.  Returns (f32 tally [G][N], exact 2^-28 tally [G][N])."""
    G, N = variables.energygroups, len(meshid)
    tally32 = [[ZERO] * N for _ in range(G)]
    tally_fixed = [[0] * N for _ in range(G)]

    def score(g, cell, v):
        tally32[g][cell] = tally32[g][cell] + v
        tally_fixed[g][cell] += int(v * f32(2.0 ** 28))

    last = end if sw.inclusive_ranges else end - 1                                          # :226 is `start..=end`
    for y in range(start, last + 1):
        rnd = Stream(sw.seed, sw.seq, sw.stride, gen * variables.histories + y)
        mesh_index, spawn_sub_mesh, mu, neutron_energy = spawn_neutron(fuel_indices, variables, xsdata, meshid, rnd)
        start_x = meshid[mesh_index].mesh_left + (spawn_sub_mesh * delta_x_fuel)           # :230
        ev = Events()
        alive = True
        while alive:
            alive, mesh_index, mu, neutron_energy, start_x = particle_travel(
                score, meshid, mesh_index, neutron_energy, mu, start_x, variables.mattypes, G, variables.boundr,
                variables.boundl, xsdata, rnd, sw, ev)
        if trace is not None:
            (xbits,) = struct.unpack("<I", struct.pack("<f", float(start_x)))
            trace[y] = (ev.collisions, ev.crossings, ev.flights, ev.reflections, rnd.rng.state & 0xFFFFFFFF,
                        rnd.rng.state >> 32, mesh_index, xbits, ev.fate, neutron_energy)
    return tally32, tally_fixed


def average_assembly(flux, numass, energygroups):
    """src/mc_code.rs:259-274"""
    mesh_assembly = len(flux[0]) // numass
    out = [[ZERO] * len(flux[0]) for _ in range(energygroups)]
    for e in range(energygroups):
        for assembly in range(1, numass + 1):
            s = ZERO
            for x in range((assembly - 1) * mesh_assembly, assembly * mesh_assembly):
                s = s + flux[e][x]
            avg = s / f32(mesh_assembly)
            for index in range((assembly - 1) * mesh_assembly, assembly * mesh_assembly):
                out[e][index] = avg
    return out


def monte_carlo(variables: Variables, xsdata: XSData, delta_x_fuel, meshid, fuel_indices, k_new, sw: Switches,
                exact_tally: bool = False, trace_gen=None):
    """src/mc_code.rs:276-380.

    exact_tally=False: per-worker f32 tallies added in worker order (:224,331-338), the reference's arithmetic.
    exact_tally=True : Q15 -- every score truncated to 2^-28 and summed as integers, converted once per generation.
    """
    G, N, gens = variables.energygroups, len(meshid), variables.generations
    flux_out = [[ZERO] * N for _ in range(G)]
    fission_out = [ZERO] * N
    k_out = [ZERO] * gens
    k_fund = [ZERO] * gens
    fixed_out = np.zeros((gens, G, N), np.uint64)
    trace = {} if trace_gen is not None else None
    k_new = f32(k_new)

    for x in range(gens):
        tally = [[ZERO] * N for _ in range(G)]
        fixed = [[0] * N for _ in range(G)]
        k, k_new = k_new, ZERO
        threads = sw.threads                                                                # :302 (cores - 1 upstream)
        threaded_histories = variables.histories // threads
        starting_points = [t * threaded_histories for t in range(threads)]
        ending_points = starting_points[1:] + [variables.histories]
        for t in range(threads):                                                            # :309-338, in join order
            t32, tfx = particle_lifetime(xsdata, meshid, fuel_indices, variables, delta_x_fuel, starting_points[t],
                                         ending_points[t], x, sw, trace if trace_gen == x else None)
            for e in range(G):
                for i in range(N):
                    tally[e][i] = tally[e][i] + t32[e][i]
                    fixed[e][i] += tfx[e][i]
        fixed_out[x] = np.array(fixed, dtype=np.uint64)
        if exact_tally:
            tally = [[f32(float(fixed[e][i]) * (1.0 / 2.0 ** 28)) for i in range(N)] for e in range(G)]

        fund = ONE / f32((gens - (variables.skip - 1)) % (1 << 64))                          # :340
        for e in range(G):
            for i in range(N):
                dx, matid = meshid[i].delta_x, meshid[i].matid
                flux = tally[e][i] / (k * f32(variables.histories) * dx)                      # :346
                xi = matid + variables.mattypes * e
                fission_source = f32(xsdata.nut[xi]) * f32(xsdata.sigf[xi]) * flux            # :347-350
                k_new = k_new + k * dx * fission_source                                       # :351
                if x >= variables.skip:
                    conversion = (f32(3565e6) * k * f32(36.2)) / (
                        f32(200e6) * f32(1.602176634e-19) * f32(xsdata.nut[0 + variables.mattypes * 1])
                        * meshid[N - 1].mesh_right)                                           # :353-357
                    flux_out[e][i] = flux_out[e][i] + flux * conversion * fund                # :358
                    fission_out[i] = fission_out[i] + fission_source * fund                   # :359
        k_out[x] = k_new                                                                      # :363

    assembly_average = average_assembly(flux_out, variables.numass, G)                       # :366
    if variables.skip < gens:
        k_fund[variables.skip] = k_out[variables.skip]                                        # :368
        for g in range(variables.skip + 1, gens):                                             # :370-376
            s = ZERO
            for x in range(variables.skip, g + 1):
                s = s + k_out[x]
            k_fund[g] = s / f32(g - (variables.skip - 1))
    tr = None
    if trace is not None:
        tr = np.array([trace[y] for y in sorted(trace)], dtype=np.uint32)
    return dict(flux=np.array(flux_out, f32), assembly_average=np.array(assembly_average, f32),
                fission_source=np.array(fission_out, f32), k=np.array(k_out, f32), k_fund=np.array(k_fund, f32),
                tally_fixed=fixed_out, trace=tr)


def from_product_inputs(v, xs, dx, mesh, fuel):
    """Adapt the product-side objects of tests/util.load_case to this file's types."""
    variables = Variables(mattypes=int(v.mattypes), energygroups=int(v.energygroups), generations=int(v.generations),
                          histories=int(v.histories), skip=int(v.skip), numass=int(v.numass), boundl=f32(v.boundl),
                          boundr=f32(v.boundr))
    a = lambda t: np.ascontiguousarray(t, dtype=f32)  # noqa: E731
    xsdata = XSData(sigt=a(xs.sigt), sigs=a(xs.sigs), mu=a(xs.mu), siga=a(xs.siga), sigf=a(xs.sigf), nut=a(xs.nut),
                    chit=a(xs.chit), scat_matrix=a(xs.scat_matrix), inv_sigtr=a(xs.inv_sigtr))
    meshid = [Mesh(int(m), f32(d), f32(l), f32(r)) for m, d, l, r in
              zip(mesh.matid, mesh.delta_x, mesh.mesh_left, mesh.mesh_right)]
    return variables, xsdata, f32(dx.fuel), meshid, [int(i) for i in fuel]


# ---------------------------------------------------------------------------------------------
# The two capabilities the north star adds and the reference does not have: the fission-bank power
# iteration and Woodcock delta tracking.  There is no reference source to cite; what is restated is
# the definition in DESIGN.md section 5 ("Fission bank", "woodcock_kernel"), so that the C oracle's
# implementation of it -- the checker of the GPU kernels in those modes -- has a second opinion too.
# ---------------------------------------------------------------------------------------------
BANK_CAP = 8


def _bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def _from_bits(b: int) -> np.float32:
    return f32(struct.unpack("<f", struct.pack("<I", b))[0])


class FissionBank:
    """Sites banked by one generation: per history at most BANK_CAP of them, kept in (history, site) order."""

    def __init__(self, xsdata, mattypes, inv_k):
        self.xsdata, self.M, self.inv_k = xsdata, mattypes, inv_k
        self.rows = {}
        self.current = None

    def begin(self, y):
        self.current = self.rows.setdefault(y, [])
        self.produced = 0

    def __call__(self, cell, x, mat, g, xs_index, rnd):
        """n = floor(nu Sigma_f(mat, g) * inv_sigtr[xs] / k_prev + xi) sites at the collision point; the draw is only
        made in fissile material."""
        nusigf = f32(self.xsdata.nut[mat + self.M * g]) * f32(self.xsdata.sigf[mat + self.M * g])
        if nusigf > ZERO:
            wgt = nusigf * f32(self.xsdata.inv_sigtr[xs_index]) * self.inv_k
            n = int(wgt + rnd.random())
            for _ in range(n):
                if self.produced < BANK_CAP:
                    self.current.append((cell << 32) | _bits(x))
                self.produced += 1

    def dense(self):
        return [site for y in sorted(self.rows) for site in self.rows[y]]


def entropy_bits(bank, n_cells) -> float:
    """Shannon entropy of the bank over mesh cells, in bits, summed in cell order."""
    import math

    hist = [0] * n_cells
    for site in bank:
        hist[site >> 32] += 1
    e = 0.0
    for h in hist:
        if h:
            pr = h / len(bank)
            e -= pr * math.log2(pr)
    return e


def spawn_from(source_bank, fuel_indices, variables, xsdata, meshid, delta_x_fuel, rnd):
    """Birth of one history: from the bank (site index, mu, chi: no position draw) or, with an empty bank, the
    reference's flat fuel source (cell, position, mu, chi)."""
    if source_bank:
        site = source_bank[rnd.gen_range(len(source_bank))]
        mesh_index, start_x = site >> 32, _from_bits(site & 0xFFFFFFFF)
        mu = direction(rnd.random())
        return mesh_index, start_x, mu, energy(rnd.random(), mesh_index, variables, xsdata, meshid)
    mesh_index, sub, mu, g = spawn_neutron(fuel_indices, variables, xsdata, meshid, rnd)
    return mesh_index, meshid[mesh_index].mesh_left + (sub * delta_x_fuel), mu, g


def material_runs(meshid):
    lo, hi = [0] * len(meshid), [0] * len(meshid)
    i = 0
    while i < len(meshid):
        j = i
        while j < len(meshid) and meshid[j].matid == meshid[i].matid:
            j += 1
        for q in range(i, j):
            lo[q], hi[q] = i, j
        i = j
    return lo, hi


def woodcock_tables(variables, xsdata, meshid):
    """sigtr = sigt - mu*sigs (the collision density flights are sampled with, src/process_input.rs:152-156) and, for
    every (stale group, current group) pair, 1 / the largest sigtr of either group over the materials in the mesh."""
    M, G = variables.mattypes, variables.energygroups
    sigtr = [f32(xsdata.sigt[i]) - f32(xsdata.mu[i]) * f32(xsdata.sigs[i]) for i in range(M * G)]
    present = sorted({c.matid for c in meshid})
    inv_maj = {}
    for a in range(G):
        for b in range(G):
            mx = ZERO
            for m in present:
                mx = max(mx, sigtr[m + M * a], sigtr[m + M * b])
            inv_maj[(a, b)] = ONE / mx
    return sigtr, inv_maj


def locate(meshid, x) -> int:
    """Cell containing x: the number of interior edges <= x."""
    c = 0
    while c < len(meshid) - 1 and meshid[c].mesh_right <= x:
        c += 1
    return c


def woodcock_history(score, variables, xsdata, meshid, runs, tables, mesh_index, start_x, mu, g, rnd, sw, ev, bank):
    """Delta tracking: flights against the majorant, every tentative collision scores 1/Sigma_maj into the cell it
    lands in and is real with probability sigtr/Sigma_maj.  The group that sets the cross sections (Q1) is the one the
    neutron had when it entered its material run, until it is seen outside that run."""
    M, G, N = variables.mattypes, variables.energygroups, len(meshid)
    run_lo, run_hi = runs
    sigtr, inv_maj_of = tables
    length = meshid[N - 1].mesh_right
    x, cell, xsg = start_x, mesh_index, g
    home_lo, home_hi, left = run_lo[cell], run_hi[cell], False
    while True:
        inv_maj = inv_maj_of[(xsg, g)]
        xn = x + mu * -ln(rnd.random()) * inv_maj
        ev.flights += 1
        while xn < ZERO or xn > length:                       # albedo walls act on the direction cosine (Q8)
            lo_wall = xn < ZERO
            wall, b = (ZERO, variables.boundl) if lo_wall else (length, variables.boundr)
            if not b > ZERO:
                ev.fate = 2
                return cell, x, g
            rem = xn - wall
            mu = mu * (-b)
            xn = wall + rem * (-b)
            if (home_lo != 0) if lo_wall else (home_hi != N):
                left = True
            ev.reflections += 1
        cell, x = locate(meshid, xn), xn
        if cell < home_lo or cell >= home_hi:
            left = True
        mat = meshid[cell].matid
        g_eff = g if left else xsg
        xs_index = mat + M * g_eff
        score(g, cell, inv_maj)
        if rnd.random() < sigtr[xs_index] * inv_maj:
            ev.collisions += 1
            with np.errstate(divide="ignore", invalid="ignore"):
                scat_mat = scat_mat_calc(G, mat, g, ONE / f32(xsdata.sigs[xs_index]), xsdata.scat_matrix)
            hook = None
            if bank is not None:
                hook = lambda: bank(cell, x, mat, g, xs_index, rnd)  # noqa: E731
            alive, g_new, mu_new = interaction(rnd.random(), scat_mat, xsdata, xs_index, g, rnd, sw.scatter_mode, hook)
            if not alive:
                ev.fate = 1
                return cell, x, g
            g, mu = g_new, mu_new
            xsg = g_eff if sw.stale_xs else g
            home_lo, home_hi, left = run_lo[cell], run_hi[cell], False


def monte_carlo_extended(variables: Variables, xsdata: XSData, delta_x_fuel, meshid, fuel_indices, k_new, sw: Switches,
                         tracking: str = "surface", source: str = "uniform_fuel", trace_gen=None):
    """Generation loop of `monte_carlo` (exact tallies, one worker) with the two added modes.  Returns what
    `monte_carlo` returns plus bank_sizes, entropy and the dense bank of every generation."""
    G, N, gens = variables.energygroups, len(meshid), variables.generations
    flux_out = [[ZERO] * N for _ in range(G)]
    fission_out = [ZERO] * N
    k_out = [ZERO] * gens
    fixed_out = np.zeros((gens, G, N), np.uint64)
    trace_rows, bank_sizes, entropy, banks = [], [], [], []
    runs = material_runs(meshid)
    tables = woodcock_tables(variables, xsdata, meshid) if tracking == "woodcock" else None
    source_bank = []
    k_new = f32(k_new)
    for x in range(gens):
        fixed = [[0] * N for _ in range(G)]
        k, k_new = k_new, ZERO

        def score(g, cell, v):
            fixed[g][cell] += int(v * f32(2.0 ** 28))

        bank = FissionBank(xsdata, variables.mattypes, ONE / k) if source == "fission_bank" else None
        for y in range(variables.histories):
            rnd = Stream(sw.seed, sw.seq, sw.stride, x * variables.histories + y)
            mesh_index, start_x, mu, g = spawn_from(source_bank, fuel_indices, variables, xsdata, meshid, delta_x_fuel, rnd)
            ev = Events()
            if bank is not None:
                bank.begin(y)
            if tracking == "woodcock":
                mesh_index, start_x, g = woodcock_history(score, variables, xsdata, meshid, runs, tables, mesh_index,
                                                          start_x, mu, g, rnd, sw, ev, bank)
            else:
                alive = True
                while alive:
                    alive, mesh_index, mu, g, start_x = particle_travel(
                        score, meshid, mesh_index, g, mu, start_x, variables.mattypes, G, variables.boundr,
                        variables.boundl, xsdata, rnd, sw, ev, bank)
            if trace_gen == x:
                trace_rows.append((ev.collisions, ev.crossings, ev.flights, ev.reflections, rnd.rng.state & 0xFFFFFFFF,
                                   rnd.rng.state >> 32, mesh_index, _bits(start_x), ev.fate, g))
        if bank is not None:
            source_bank = bank.dense()
            banks.append(source_bank)
            bank_sizes.append(len(source_bank))
            entropy.append(entropy_bits(source_bank, N) if source_bank else 0.0)
        fixed_out[x] = np.array(fixed, dtype=np.uint64)
        tally = [[f32(float(fixed[e][i]) * (1.0 / 2.0 ** 28)) for i in range(N)] for e in range(G)]
        fund = ONE / f32((gens - (variables.skip - 1)) % (1 << 64))
        for e in range(G):                                                                   # src/mc_code.rs:342-362
            for i in range(N):
                dx, matid = meshid[i].delta_x, meshid[i].matid
                flux = tally[e][i] / (k * f32(variables.histories) * dx)
                xi = matid + variables.mattypes * e
                fission_source = f32(xsdata.nut[xi]) * f32(xsdata.sigf[xi]) * flux
                k_new = k_new + k * dx * fission_source
                if x >= variables.skip:
                    conversion = (f32(3565e6) * k * f32(36.2)) / (
                        f32(200e6) * f32(1.602176634e-19) * f32(xsdata.nut[0 + variables.mattypes * 1]) * meshid[N - 1].mesh_right)
                    flux_out[e][i] = flux_out[e][i] + flux * conversion * fund
                    fission_out[i] = fission_out[i] + fission_source * fund
        k_out[x] = k_new
    return dict(flux=np.array(flux_out, f32), fission_source=np.array(fission_out, f32), k=np.array(k_out, f32),
                tally_fixed=fixed_out, trace=np.array(trace_rows, dtype=np.uint32) if trace_rows else None,
                bank_sizes=np.array(bank_sizes, np.uint64), entropy=np.array(entropy), banks=banks)
