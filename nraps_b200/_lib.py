"""ctypes binding of nraps_b200/lib/libnraps_b200.so (include/nraps_mc.h, nraps_host.h).

The library is the product: if it is missing this module raises -- there is no
Python or CPU fallback for the transport path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRAPS_LIB_DIR: an experiment build of the same library (make LIBDIR=...); there is still no non-CUDA path behind it
LIB_PATH = os.path.join(os.environ.get("NRAPS_LIB_DIR") or os.path.join(_HERE, "lib"), "libnraps_b200.so")

CT_WORDS = 8
CT_NAMES = ["histories", "collisions", "crossings", "flights", "reflections", "leaks", "truncated", "banked"]
TR_WORDS = 10
TR_NAMES = ["collisions", "crossings", "flights", "reflections", "rng_lo", "rng_hi", "cell", "xbits", "fate", "group"]
TALLY_FRAC_BITS = 28
IPC_HANDLE_BYTES = 64
PH_NAMES = ["source", "transport", "prefix", "compact", "finalize"]
ABI_VERSION = 4  # NRAPS_ABI_VERSION of include/nraps_mc.h this binding was written against

_fp = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class Problem(C.Structure):
    _fields_ = [
        ("M", C.c_uint32), ("G", C.c_uint32), ("N", C.c_uint32), ("NF", C.c_uint32), ("numass", C.c_uint32),
        ("generations", C.c_uint64), ("histories", C.c_uint64), ("skip", C.c_uint64),
        ("boundl", C.c_float), ("boundr", C.c_float), ("dx_fuel", C.c_float), ("dx_water", C.c_float),
        ("k0", C.c_float),
        ("sigt", _fp), ("sigs", _fp), ("mu", _fp), ("siga", _fp), ("sigf", _fp), ("nut", _fp), ("chit", _fp),
        ("inv_sigtr", _fp), ("scat", _fp), ("matid", _u8p), ("dx", _fp), ("left", _fp), ("right", _fp),
        ("fuel_indices", _u64p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("stream", C.c_uint64), ("stride", C.c_uint64),
        ("device", C.c_int32), ("scatter_mode", C.c_int32), ("stale_xs", C.c_int32), ("source_mode", C.c_int32),
        ("tracking_mode", C.c_int32), ("kernel_variant", C.c_int32), ("threads_per_block", C.c_int32),
        ("blocks_per_sm", C.c_int32), ("chunk", C.c_int32), ("quiet", C.c_int32), ("bank_cap", C.c_int32),
        ("spawn_batch", C.c_int32), ("walk_cap", C.c_int32), ("slots_per_thread", C.c_int32), ("max_flights", C.c_uint64),
        ("profile_phases", C.c_int32), ("reserved0", C.c_int32),
    ]


class Results(C.Structure):
    _fields_ = [
        ("flux", _fp), ("assembly_average", _fp), ("fission_source", _fp), ("k", _fp), ("k_fund", _fp),
        ("tally_fixed", _u64p), ("counters", C.c_uint64 * CT_WORDS), ("seconds_device", C.c_double),
        ("bank_sizes", _u64p), ("entropy", C.POINTER(C.c_double)), ("flux_moments", C.POINTER(C.c_double)),
    ]


class Deck(C.Structure):
    _fields_ = [
        ("analk", C.c_uint32), ("mattypes", C.c_uint32), ("energygroups", C.c_uint32), ("numass", C.c_uint32),
        ("numrods", C.c_uint32),
        ("generations", C.c_uint64), ("histories", C.c_uint64), ("skip", C.c_uint64), ("mpfr", C.c_uint64),
        ("mpwr", C.c_uint64),
        ("roddia", C.c_float), ("rodpitch", C.c_float), ("boundl", C.c_float), ("boundr", C.c_float),
        ("dx_fuel", C.c_float), ("dx_water", C.c_float),
        ("n_xs", C.c_uint32), ("n_scat", C.c_uint32), ("n_matid", C.c_uint32),
        ("sigt", _fp), ("sigs", _fp), ("mu", _fp), ("siga", _fp), ("sigf", _fp), ("nut", _fp), ("chit", _fp),
        ("inv_sigtr", _fp), ("scat", _fp), ("matid", _u8p),
        ("solution", C.c_int32), ("solver", C.c_int32),
    ]


class Mesh(C.Structure):
    _fields_ = [
        ("N", C.c_uint32), ("NF", C.c_uint32), ("matid", _u8p), ("dx", _fp), ("left", _fp), ("right", _fp),
        ("fuel_indices", _u64p),
    ]


# every symbol include/nraps_mc.h and include/nraps_host.h declare
EXPORTS = [
    "nraps_mc_run", "nraps_mc_create", "nraps_mc_destroy", "nraps_mc_trim", "nraps_mc_reset", "nraps_mc_transport",
    "nraps_mc_finalize_generation", "nraps_mc_tally_buffer", "nraps_mc_set_tally_buffer", "nraps_mc_select_lane", "nraps_mc_read_tally",
    "nraps_mc_fetch", "nraps_mc_trace", "nraps_mc_launch_info", "nraps_mc_bank_compact", "nraps_mc_bank_local",
    "nraps_mc_bank_advance", "nraps_mc_bank_reserve", "nraps_mc_bank_export", "nraps_mc_bank_import", "nraps_mc_bank_peers",
    "nraps_mc_phase_ms", "nraps_dev_logf", "nraps_dev_div", "nraps_dev_pcg32",
    "nraps_strerror", "nraps_last_cuda_error", "nraps_abi_version", "nraps_options_default",
    "nraps_process_input", "nraps_deck_free", "nraps_mesh_gen", "nraps_mesh_free", "nraps_problem_from",
    "nraps_walk_segments", "nraps_format_f32", "nraps_format_f64", "nraps_plot_solution", "nraps_average_assembly", "nraps_k_fund",
    "nraps_diffusion_run",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C nraps_b200/csrc). nraps_b200 has no fallback path."
        )
    L = C.CDLL(LIB_PATH)
    if L.nraps_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH} has ABI version {L.nraps_abi_version()}, this binding needs {ABI_VERSION}: rebuild it")
    vp = C.c_void_p
    L.nraps_mc_run.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Results)]
    L.nraps_mc_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(vp)]
    L.nraps_mc_destroy.argtypes = [vp]
    L.nraps_mc_trim.argtypes = [C.c_int32]
    L.nraps_mc_reset.argtypes = [vp, C.c_float, vp]
    L.nraps_mc_transport.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.nraps_mc_finalize_generation.argtypes = [vp, C.c_uint64, vp]
    L.nraps_mc_tally_buffer.argtypes = [vp, C.POINTER(vp), _u64p]
    L.nraps_mc_set_tally_buffer.argtypes = [vp, vp]
    L.nraps_mc_select_lane.argtypes = [vp, C.c_int32]
    L.nraps_mc_read_tally.argtypes = [vp, _u64p, vp]
    L.nraps_mc_fetch.argtypes = [vp, C.POINTER(Results), vp]
    L.nraps_mc_trace.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, _u32p, vp]
    L.nraps_mc_launch_info.argtypes = [vp, _u32p]
    L.nraps_mc_phase_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.nraps_mc_bank_compact.argtypes = [vp, C.c_uint64, vp]
    L.nraps_mc_bank_local.argtypes = [vp, C.POINTER(vp), _u64p, vp]
    L.nraps_mc_bank_advance.argtypes = [vp, C.c_uint64, vp]
    L.nraps_mc_bank_reserve.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.nraps_mc_bank_export.argtypes = [vp, vp]
    L.nraps_mc_bank_import.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.nraps_mc_bank_peers.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.nraps_dev_logf.argtypes = [_fp, _fp, C.c_uint32, C.c_int32]
    L.nraps_dev_div.argtypes = [_fp, _fp, _fp, _fp, C.c_uint32, C.c_int32]
    L.nraps_dev_pcg32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _u32p, _fp, C.c_int32]
    L.nraps_options_default.argtypes = [C.POINTER(Options)]
    L.nraps_options_default.restype = None
    L.nraps_strerror.argtypes = [C.c_int]
    L.nraps_strerror.restype = C.c_char_p
    L.nraps_last_cuda_error.restype = C.c_char_p
    L.nraps_process_input.argtypes = [C.c_char_p, C.POINTER(Deck)]
    L.nraps_deck_free.argtypes = [C.POINTER(Deck)]
    L.nraps_deck_free.restype = None
    L.nraps_mesh_gen.argtypes = [_u8p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_float, C.c_float, C.POINTER(Mesh)]
    L.nraps_mesh_free.argtypes = [C.POINTER(Mesh)]
    L.nraps_mesh_free.restype = None
    L.nraps_walk_segments.argtypes = [_u8p, _fp, _fp, C.c_uint32, _u32p, _u32p, _u32p]
    L.nraps_problem_from.argtypes = [C.POINTER(Deck), C.POINTER(Mesh), C.c_float, C.POINTER(Problem)]
    L.nraps_format_f32.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
    L.nraps_format_f32.restype = C.c_size_t
    L.nraps_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
    L.nraps_format_f64.restype = C.c_size_t
    L.nraps_plot_solution.argtypes = [C.POINTER(Results), C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, C.c_char_p]
    L.nraps_average_assembly.argtypes = [_fp, C.c_uint32, C.c_uint32, C.c_uint32, _fp]
    L.nraps_average_assembly.restype = None
    L.nraps_k_fund.argtypes = [_fp, C.c_uint64, C.c_uint64, _fp]
    L.nraps_k_fund.restype = None
    L.nraps_diffusion_run.argtypes = [C.POINTER(Problem), C.POINTER(Results), C.c_uint64, C.POINTER(C.c_uint64)]
    _lib = L
    return L


class NrapsError(RuntimeError):
    def __init__(self, code: int, where: str):
        L = lib()
        msg = L.nraps_strerror(code).decode()
        if code == 6:
            msg += " -- " + L.nraps_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} (code {code})")
        self.code = code


def check(code: int, where: str) -> None:
    if code != 0:
        raise NrapsError(code, where)
