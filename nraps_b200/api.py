"""Host-side mirror of the reference's operator interface for the MC path.

Same names, argument meaning and pipeline order as the reference driver
(src/main.rs:332-364, commented out at HEAD):

    variables, xsdata, matid, deltax, solution, solver = process_input(path)
    meshid, fuel_indices = mesh_gen(matid, variables, deltax)
    results = monte_carlo(variables, xsdata, deltax, meshid, fuel_indices, 1.0)
    plot_solution(results, G, generations, N, L)

Everything numeric happens inside libnraps_b200.so through the C ABI of
include/nraps_mc.h / nraps_host.h; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import CT_NAMES, CT_WORDS, TR_NAMES, TR_WORDS, Options, Problem, Results, check, lib

SCATTER_MODES = {"single_xi": 0, "rust_pre182": 1, "rust_182": 2}
SOURCE_MODES = {"uniform_fuel": 0, "fission_bank": 1}
TRACKING_MODES = {"surface": 0, "woodcock": 1}
KERNEL_VARIANTS = {"fused": 0, "event": 1, "block_event": 2}


@dataclass
class Variables:  # src/main.rs:22-39
    analk: int
    mattypes: int
    energygroups: int
    generations: int
    histories: int
    skip: int
    numass: int
    numrods: int
    roddia: float
    rodpitch: float
    mpfr: int
    mpwr: int
    boundl: float
    boundr: float


@dataclass
class DeltaX:  # src/main.rs:41-44
    fuel: float
    water: float


@dataclass
class XSData:  # src/main.rs:46-56
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
class Mesh:  # Vec<Mesh> (src/main.rs:68-75) as structure-of-arrays
    matid: np.ndarray
    delta_x: np.ndarray
    mesh_left: np.ndarray
    mesh_right: np.ndarray

    def __len__(self) -> int:
        return len(self.matid)


@dataclass
class SolutionResults:  # src/main.rs:77-83
    flux: np.ndarray
    assembly_average: np.ndarray
    fission_source: np.ndarray
    k: np.ndarray
    k_fund: np.ndarray
    counters: dict = field(default_factory=dict)
    seconds_device: float = 0.0
    tally_fixed: np.ndarray | None = None
    bank_sizes: np.ndarray | None = None  # fission_bank mode: sites banked per generation
    entropy: np.ndarray | None = None     # fission_bank mode: Shannon entropy (bits) of each bank over cells
    flux_moments: np.ndarray | None = None  # extension, f64 [2][G][N]: sum and sum of squares over generations >= skip
                                            # of the per-generation term flux * conversion (src/mc_code.rs:358)

    def flux_std_error(self, generations: int, skip: int) -> np.ndarray:
        """Standard error of ``flux`` per bin from the spread between the accumulated generations, in the units and
        normalisation of ``flux`` (fund * sum of the per-generation terms, src/mc_code.rs:340,358)."""
        if self.flux_moments is None:
            raise ValueError("these results carry no flux_moments")
        n = generations - skip
        if n < 2:
            return np.full(self.flux.shape, np.nan)
        fund = 1.0 / float(np.uint64(generations - (skip - 1)))  # the reference's fund, not 1/n
        s1, s2 = self.flux_moments
        var = np.maximum(s2 - s1 * s1 / n, 0.0) / (n - 1)  # between generations
        return fund * np.sqrt(var * n)


def _np(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)


def process_input(path: str = "./TestCaseC.txt"):
    """src/process_input.rs:85-175 -> (Variables, XSData, matid, DeltaX, solution, solver)."""
    d = _lib.Deck()
    check(lib().nraps_process_input(os.fsencode(path), C.byref(d)), f"process_input({path})")
    try:
        v = Variables(
            analk=d.analk, mattypes=d.mattypes, energygroups=d.energygroups, generations=d.generations,
            histories=d.histories, skip=d.skip, numass=d.numass, numrods=d.numrods, roddia=d.roddia,
            rodpitch=d.rodpitch, mpfr=d.mpfr, mpwr=d.mpwr, boundl=d.boundl, boundr=d.boundr,
        )
        n = d.n_xs
        xs = XSData(
            sigt=_np(d.sigt, n, np.float32), sigs=_np(d.sigs, n, np.float32), mu=_np(d.mu, n, np.float32),
            siga=_np(d.siga, n, np.float32), sigf=_np(d.sigf, n, np.float32), nut=_np(d.nut, n, np.float32),
            chit=_np(d.chit, n, np.float32), scat_matrix=_np(d.scat, d.n_scat, np.float32),
            inv_sigtr=_np(d.inv_sigtr, n, np.float32),
        )
        matid = _np(d.matid, d.n_matid, np.uint8)
        dx = DeltaX(fuel=d.dx_fuel, water=d.dx_water)
        return v, xs, matid, dx, int(d.solution), int(d.solver)
    finally:
        lib().nraps_deck_free(C.byref(d))


def mesh_gen(matid, variables: Variables, deltax: DeltaX):
    """src/main.rs:85-143 -> (Mesh, fuel_indices)."""
    pins = np.ascontiguousarray(matid, dtype=np.uint8)
    m = _lib.Mesh()
    check(
        lib().nraps_mesh_gen(pins.ctypes.data_as(_lib._u8p), len(pins), variables.mpfr, variables.mpwr, variables.numass,
                             C.c_float(deltax.fuel), C.c_float(deltax.water), C.byref(m)),
        "mesh_gen",
    )
    try:
        mesh = Mesh(matid=_np(m.matid, m.N, np.uint8), delta_x=_np(m.dx, m.N, np.float32),
                    mesh_left=_np(m.left, m.N, np.float32), mesh_right=_np(m.right, m.N, np.float32))
        fuel = _np(m.fuel_indices, m.NF, np.uint64)
        return mesh, fuel
    finally:
        lib().nraps_mesh_free(C.byref(m))


class _DevArray:
    """__cuda_array_interface__ view of n int64 words at a raw device pointer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class _Marshalled:
    """Keeps the numpy buffers a Problem points into alive."""

    def __init__(self, variables, xsdata, deltax, meshid, fuel_indices, k_new, generations, histories, skip):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
        self.keep = dict(
            sigt=f(xsdata.sigt), sigs=f(xsdata.sigs), mu=f(xsdata.mu), siga=f(xsdata.siga), sigf=f(xsdata.sigf),
            nut=f(xsdata.nut), chit=f(xsdata.chit), inv_sigtr=f(xsdata.inv_sigtr), scat=f(xsdata.scat_matrix),
            matid=np.ascontiguousarray(meshid.matid, dtype=np.uint8), dx=f(meshid.delta_x),
            left=f(meshid.mesh_left), right=f(meshid.mesh_right),
            fuel=np.ascontiguousarray(fuel_indices, dtype=np.uint64),
        )
        k = self.keep
        p = lambda a, t=C.c_float: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
        self.generations = int(variables.generations if generations is None else generations)
        self.histories = int(variables.histories if histories is None else histories)
        self.skip = int(variables.skip if skip is None else skip)
        self.G, self.M, self.N = int(variables.energygroups), int(variables.mattypes), len(k["matid"])
        self.numass = int(variables.numass)
        # The C ABI takes bare pointers (the tables' extents follow from M, G and N), so a short array would be read
        # past its end; the reference indexes Vecs and panics instead.
        for name in ("sigt", "sigs", "mu", "siga", "sigf", "nut", "chit", "inv_sigtr"):
            if k[name].size < self.M * self.G:
                raise ValueError(f"xsdata.{name} holds {k[name].size} values, mattypes * energygroups = {self.M * self.G} are indexed")
        if k["scat"].size < self.M * self.G * self.G:
            raise ValueError(f"xsdata.scat_matrix holds {k['scat'].size} values, mattypes * energygroups^2 = {self.M * self.G * self.G} are indexed")
        for name in ("dx", "left", "right"):
            if k[name].size != self.N:
                raise ValueError(f"mesh arrays differ in length: matid has {self.N} cells, {name} has {k[name].size}")
        self.problem = Problem(
            M=self.M, G=self.G, N=self.N, NF=len(k["fuel"]), numass=self.numass,
            generations=self.generations, histories=self.histories, skip=self.skip,
            boundl=float(variables.boundl), boundr=float(variables.boundr), dx_fuel=float(deltax.fuel),
            dx_water=float(deltax.water), k0=float(k_new),
            sigt=p(k["sigt"]), sigs=p(k["sigs"]), mu=p(k["mu"]), siga=p(k["siga"]), sigf=p(k["sigf"]), nut=p(k["nut"]),
            chit=p(k["chit"]), inv_sigtr=p(k["inv_sigtr"]), scat=p(k["scat"]), matid=p(k["matid"], C.c_uint8),
            dx=p(k["dx"]), left=p(k["left"]), right=p(k["right"]), fuel_indices=p(k["fuel"], C.c_uint64),
        )


def make_options(*, seed=0, stream=0, stride=0, device=0, scatter_mode="single_xi", stale_xs=True,
                 source_mode="uniform_fuel", tracking_mode="surface", kernel_variant="fused", threads_per_block=0,
                 blocks_per_sm=0, chunk=0, quiet=True, max_flights=0, bank_cap=0, spawn_batch=0, walk_cap=0,
                 slots_per_thread=0, profile_phases=False) -> Options:
    return Options(
        seed=seed, stream=stream, stride=stride, device=device, scatter_mode=SCATTER_MODES[scatter_mode],
        stale_xs=int(bool(stale_xs)), source_mode=SOURCE_MODES[source_mode], tracking_mode=TRACKING_MODES[tracking_mode],
        kernel_variant=KERNEL_VARIANTS[kernel_variant], threads_per_block=threads_per_block,
        blocks_per_sm=blocks_per_sm, chunk=chunk, quiet=int(bool(quiet)), bank_cap=bank_cap, spawn_batch=spawn_batch, walk_cap=walk_cap,
        slots_per_thread=slots_per_thread,
        max_flights=max_flights, profile_phases=int(bool(profile_phases)),
    )


class _ResultBuffers:
    def __init__(self, G, N, gens, want_tally=False):
        self.flux = np.zeros((G, N), np.float32)
        self.avg = np.zeros((G, N), np.float32)
        self.fis = np.zeros(N, np.float32)
        self.k = np.zeros(gens, np.float32)
        self.kf = np.zeros(gens, np.float32)
        self.tally = np.zeros((gens, G, N), np.uint64) if want_tally else None
        self.bank_sizes = np.zeros(gens, np.uint64)
        self.entropy = np.zeros(gens, np.float64)
        self.moments = np.zeros((2, G, N), np.float64)
        p = lambda a, t=C.c_float: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
        self.c = Results(flux=p(self.flux), assembly_average=p(self.avg), fission_source=p(self.fis), k=p(self.k),
                         k_fund=p(self.kf), tally_fixed=p(self.tally, C.c_uint64) if want_tally else None,
                         bank_sizes=p(self.bank_sizes, C.c_uint64), entropy=p(self.entropy, C.c_double),
                         flux_moments=p(self.moments, C.c_double))

    def solution(self) -> SolutionResults:
        return SolutionResults(
            flux=self.flux, assembly_average=self.avg, fission_source=self.fis, k=self.k, k_fund=self.kf,
            counters={n: int(self.c.counters[i]) for i, n in enumerate(CT_NAMES)},
            seconds_device=float(self.c.seconds_device), tally_fixed=self.tally, bank_sizes=self.bank_sizes,
            entropy=self.entropy, flux_moments=self.moments,
        )


def monte_carlo(variables, xsdata, delta_x, meshid, fuel_indices, k_new: float = 1.0, *, generations=None,
                histories=None, skip=None, want_tally=False, **options) -> SolutionResults:
    """src/mc_code.rs:276-380 on one B200 (host buffers in, host buffers out)."""
    m = _Marshalled(variables, xsdata, delta_x, meshid, fuel_indices, k_new, generations, histories, skip)
    o = make_options(**options)
    rb = _ResultBuffers(m.G, m.N, m.generations, want_tally)
    check(lib().nraps_mc_run(C.byref(m.problem), C.byref(o), C.byref(rb.c)), "monte_carlo")
    return rb.solution()


def nalgebra_method(xsdata, meshid, energygroups: int, mattypes: int, boundl: float, boundr: float, numass: int,
                    *, max_iterations: int = 0) -> SolutionResults:
    """src/discrete.rs:181-356: the reference's finite-difference diffusion solver (host code), kept as a physics
    cross-check of the Monte Carlo path.  Returns flux, assembly_average and k = [k]; fission_source and k_fund are
    empty like the reference's.  ``counters["iterations"]`` holds the number of power iterations."""
    v = Variables(analk=0, mattypes=mattypes, energygroups=energygroups, generations=1, histories=1, skip=0,
                  numass=numass, numrods=0, roddia=0.0, rodpitch=0.0, mpfr=0, mpwr=0, boundl=boundl, boundr=boundr)
    m = _Marshalled(v, xsdata, DeltaX(0.0, 0.0), meshid, np.zeros(0, np.uint64), 1.0, None, None, None)
    rb = _ResultBuffers(m.G, m.N, 1)
    it = C.c_uint64(0)
    check(lib().nraps_diffusion_run(C.byref(m.problem), C.byref(rb.c), max_iterations, C.byref(it)), "nalgebra_method")
    return SolutionResults(flux=rb.flux, assembly_average=rb.avg, fission_source=np.zeros(0, np.float32), k=rb.k,
                           k_fund=np.zeros(0, np.float32), counters={"iterations": int(it.value)})


class MonteCarloContext:
    """Generation-level control of one GPU: transport -> (all-reduce) -> finalize."""

    def __init__(self, variables, xsdata, delta_x, meshid, fuel_indices, k_new: float = 1.0, *, generations=None,
                 histories=None, skip=None, **options):
        self._m = _Marshalled(variables, xsdata, delta_x, meshid, fuel_indices, k_new, generations, histories, skip)
        self._o = make_options(**options)
        self._h = C.c_void_p()
        check(lib().nraps_mc_create(C.byref(self._m.problem), C.byref(self._o), C.byref(self._h)), "nraps_mc_create")
        self.G, self.N, self.generations, self.histories = self._m.G, self._m.N, self._m.generations, self._m.histories
        self._ext = None
        self.n_words = self.tally_buffer()[1]  # G*N tally words + counters (+ N histogram words in fission_bank mode)

    def close(self):
        if self._h:
            lib().nraps_mc_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, k0: float = 1.0, stream=None):
        check(lib().nraps_mc_reset(self._h, C.c_float(k0), C.c_void_p(stream)), "nraps_mc_reset")

    def transport(self, gen: int, hist_begin: int = 0, hist_count: int | None = None, stream=None):
        n = self.histories - hist_begin if hist_count is None else hist_count
        check(lib().nraps_mc_transport(self._h, gen, hist_begin, n, C.c_void_p(stream)), "nraps_mc_transport")

    def select_lane(self, lane: int):
        """Scratch-buffer set (0 or 1) of the following transport calls: lets generation g+1 of the uniform source be
        launched on a second stream while the tail of generation g still runs (nraps_mc_select_lane)."""
        check(lib().nraps_mc_select_lane(self._h, int(lane)), "nraps_mc_select_lane")

    def finalize_generation(self, gen: int, stream=None):
        check(lib().nraps_mc_finalize_generation(self._h, gen, C.c_void_p(stream)), "nraps_mc_finalize_generation")

    def tally_buffer(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        check(lib().nraps_mc_tally_buffer(self._h, C.byref(ptr), C.byref(n)), "nraps_mc_tally_buffer")
        return ptr.value, n.value

    def use_tally_tensor(self, tensor):
        """Accumulate into a caller-owned int64 CUDA tensor (so torch.distributed can all-reduce it)."""
        assert tensor.is_cuda and tensor.numel() == self.n_words and tensor.element_size() == 8 and tensor.is_contiguous()
        self._ext = tensor
        check(lib().nraps_mc_set_tally_buffer(self._h, C.c_void_p(tensor.data_ptr())), "nraps_mc_set_tally_buffer")

    def read_tally(self, stream=None):
        words = np.zeros(self.n_words, np.uint64)
        check(lib().nraps_mc_read_tally(self._h, words.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_void_p(stream)),
              "nraps_mc_read_tally")
        tally = words[: self.G * self.N].reshape(self.G, self.N)
        return tally, {n: int(words[self.G * self.N + i]) for i, n in enumerate(CT_NAMES)}

    def trace(self, gen: int, hist_begin: int, hist_count: int, stream=None):
        rec = np.zeros((hist_count, TR_WORDS), np.uint32)
        check(lib().nraps_mc_trace(self._h, gen, hist_begin, hist_count, rec.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   C.c_void_p(stream)), "nraps_mc_trace")
        return rec

    # ---- fission_bank source mode
    def bank_compact(self, gen: int, stream=None):
        check(lib().nraps_mc_bank_compact(self._h, gen, C.c_void_p(stream)), "nraps_mc_bank_compact")

    def bank_local(self, stream=None):
        """(device pointer, count) of this rank's dense bank (synchronises the stream)."""
        ptr, n = C.c_void_p(), C.c_uint64()
        check(lib().nraps_mc_bank_local(self._h, C.byref(ptr), C.byref(n), C.c_void_p(stream)), "nraps_mc_bank_local")
        return ptr.value, n.value

    def bank_advance(self, gen: int, stream=None):
        """Record size / entropy of the bank of `gen`; generation gen+1 samples from it."""
        check(lib().nraps_mc_bank_advance(self._h, gen, C.c_void_p(stream)), "nraps_mc_bank_advance")

    def bank_reserve(self, shard_histories: int):
        """Allocate the two bank buffers so that other processes can map them (one process per GPU), sized for shards
        of up to `shard_histories`."""
        check(lib().nraps_mc_bank_reserve(self._h, shard_histories, None), "nraps_mc_bank_reserve")

    def bank_export(self) -> bytes:
        """Tickets of this rank's two bank buffers (2 x 64 bytes: pid + exported file descriptor + size)."""
        buf = C.create_string_buffer(2 * _lib.IPC_HANDLE_BYTES)
        check(lib().nraps_mc_bank_export(self._h, buf), "nraps_mc_bank_export")
        return buf.raw

    def bank_import(self, world: int, rank: int, handles: bytes):
        """Map the peers' bank buffers from every rank's exported tickets (rank order, 2 x 64 bytes each)."""
        assert len(handles) == world * 2 * _lib.IPC_HANDLE_BYTES
        check(lib().nraps_mc_bank_import(self._h, world, rank, C.c_char_p(handles)), "nraps_mc_bank_import")

    def read_bank(self, stream=None) -> np.ndarray:
        """Host copy of the local dense bank (tests)."""
        import torch

        ptr, n = self.bank_local(stream)
        if n == 0:
            return np.zeros(0, np.uint64)
        view = torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{self._o.device}")
        return view.cpu().numpy().view(np.uint64).copy()

    def fetch(self, stream=None) -> SolutionResults:
        rb = _ResultBuffers(self.G, self.N, self.generations)
        check(lib().nraps_mc_fetch(self._h, C.byref(rb.c), C.c_void_p(stream)), "nraps_mc_fetch")
        return rb.solution()

    def phase_ms(self) -> dict:
        """profile_phases=True: milliseconds per phase summed over the generations run so far (synchronizes)."""
        out = (C.c_double * len(_lib.PH_NAMES))()
        check(lib().nraps_mc_phase_ms(self._h, out), "nraps_mc_phase_ms")
        return dict(zip(_lib.PH_NAMES, [float(v) for v in out]))

    def launch_info(self) -> dict:
        out = (C.c_uint32 * 6)()
        check(lib().nraps_mc_launch_info(self._h, out), "nraps_mc_launch_info")
        return dict(zip(["grid", "block", "smem_bytes", "blocks_per_sm", "sm_count", "chunk"], [int(v) for v in out]))


def plot_solution(results: SolutionResults, energygroups: int, generations: int, number_meshes: int,
                  assembly_length: float, out_dir: str = ".") -> None:
    """src/plot_solution.rs:7-58: writes vars.csv, interface.csv, k_eff.csv (plot.py is not spawned).  Diffusion
    results (empty fission_source / k_fund) leave vars.csv and the 2G flux rows of interface.csv, like the reference."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    keep = [f(results.flux), f(results.assembly_average), f(results.fission_source), f(results.k), f(results.k_fund)]
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)) if a.size else None  # noqa: E731
    r = Results(flux=p(keep[0]), assembly_average=p(keep[1]), fission_source=p(keep[2]), k=p(keep[3]), k_fund=p(keep[4]))
    check(lib().nraps_plot_solution(C.byref(r), energygroups, generations, number_meshes, C.c_double(assembly_length),
                                    os.fsencode(out_dir)), "plot_solution")


def trim(device: int = 0) -> None:
    """Hand the device memory the library keeps cached between contexts back to the driver (nraps_mc_trim)."""
    check(lib().nraps_mc_trim(device), "trim")


def format_f32(v: float) -> str:
    buf = C.create_string_buffer(96)
    lib().nraps_format_f32(C.c_float(v), buf, 96)
    return buf.value.decode()


def format_f64(v: float) -> str:
    buf = C.create_string_buffer(400)
    lib().nraps_format_f64(C.c_double(v), buf, 400)
    return buf.value.decode()


def dev_logf(x: np.ndarray, device: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    check(lib().nraps_dev_logf(x.ctypes.data_as(_lib._fp), out.ctypes.data_as(_lib._fp), x.size, device), "nraps_dev_logf")
    return out


def dev_div(t: np.ndarray, mu: np.ndarray, device: int = 0):
    t = np.ascontiguousarray(t, dtype=np.float32)
    mu = np.ascontiguousarray(mu, dtype=np.float32)
    fast, ieee = np.empty_like(t), np.empty_like(t)
    p = lambda a: a.ctypes.data_as(_lib._fp)  # noqa: E731
    check(lib().nraps_dev_div(p(t), p(mu), p(fast), p(ieee), t.size, device), "nraps_dev_div")
    return fast, ieee


def dev_pcg32(seed: int, stream: int, stride: int, hid: int, n: int, device: int = 0):
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, np.float32)
    check(lib().nraps_dev_pcg32(seed, stream, stride, hid, n, u.ctypes.data_as(_lib._u32p), f.ctypes.data_as(_lib._fp),
                                device), "nraps_dev_pcg32")
    return u, f
