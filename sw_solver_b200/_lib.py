"""ctypes binding of libswb200.so (include/swb200.h).  Thin and literal: one Python
callable per exported symbol, argument types declared, non-zero return codes turned into
``SwbError``.  There is no fallback: if the shared object is missing it is built with
nvcc, and if that is impossible the import of the solver fails loudly."""

import ctypes
import os
import warnings

from . import build as _build

c_int32, c_int64, c_double, c_void_p = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)

SWB_F64, SWB_F32 = 0, 1
SWB_STRICT, SWB_FAST, SWB_FAST_GUARDED = 0, 1, 2
MODES = {"strict": SWB_STRICT, "fast": SWB_FAST, "fast_guarded": SWB_FAST_GUARDED}
DTYPES = {"float64": SWB_F64, "float32": SWB_F32}
SWB_E_TIMEOUT, SWB_E_DEPTH = -4, -6
IPC_HANDLE_BYTES = 64


class SwbError(RuntimeError):
    """A libswb200 call failed (code > 0: cudaError_t, code < 0: SWB_E_*)."""

    def __init__(self, code, message):
        super().__init__(f"libswb200 error {code}: {message}")
        self.code = code


class SwbConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "struct_bytes", "M", "N", "j_begin", "n_local", "j_first", "j_last", "pitch", "dtype", "mode",
        "use_diffusion", "f_is_2d", "has_hs", "device", "log_capacity", "rank", "world", "halo", "f_sep", "reserved0",
    )] + [(n, c_double) for n in (
        "courant", "final_time", "dxmin", "dymin", "g", "a", "nu", "dphi", "fsep_scale", "fsep_sa",
    )]


TABLE_FIELDS = ("phi", "c", "tg", "ac", "f_lat", "tgMidy", "cMidy", "dy", "dy1", "dyc", "dy1c", "Ay", "By", "Cy",
                "fsep_lon", "fsep_lat_b", "fsep_lat_d")


class SwbTables(ctypes.Structure):
    _fields_ = [(n, _dp) for n in TABLE_FIELDS]


class SwbStatus(ctypes.Structure):
    _fields_ = [("time", c_double), ("eigenx", c_double), ("eigeny", c_double), ("last_dt", c_double),
                ("steps", c_int64), ("done", c_int32), ("current", c_int32), ("error", c_int32),
                ("reserved", c_int32)]


#: every symbol include/swb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "swb_abi_version": (ctypes.c_int, []),
    "swb_last_error": (ctypes.c_char_p, []),
    "swb_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(SwbConfig)]),
    "swb_destroy": (ctypes.c_int, [c_void_p]),
    "swb_set_tables": (ctypes.c_int, [c_void_p, ctypes.POINTER(SwbTables)]),
    "swb_required_pitch": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "swb_bind_state": (ctypes.c_int, [c_void_p] + [c_void_p] * 8),
    "swb_to_device_layout": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p]),
    "swb_from_device_layout": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, c_void_p]),
    "swb_cfl_init": (ctypes.c_int, [c_void_p, c_void_p]),
    "swb_enqueue_steps": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p]),
    "swb_step_fixed_dt": (ctypes.c_int, [c_void_p, c_double, c_void_p]),
    "swb_laplacian": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "swb_max_speed": (ctypes.c_int, [c_void_p, _dp, c_void_p]),
    "swb_max_speed_async": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p]),
    "swb_laplacian_ext": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p]),
    "swb_set_spin_limit": (ctypes.c_int, [c_void_p, c_double]),
    "swb_status_async": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "swb_status_peek": (ctypes.c_int, [c_void_p, ctypes.POINTER(SwbStatus)]),
    "swb_get_status": (ctypes.c_int, [c_void_p, ctypes.POINTER(SwbStatus), c_void_p]),
    "swb_read_log": (ctypes.c_int, [c_void_p, _dp, _dp, ctypes.c_int, c_void_p]),
    "swb_launch_count": (c_int64, [c_void_p]),
    "swb_resident": (ctypes.c_int, [c_void_p]),
    "swb_enable_resident_bands": (ctypes.c_int, [c_void_p]),
    "swb_ipc_export": (ctypes.c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64)]),
    "swb_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "swb_ipc_close": (ctypes.c_int, [c_void_p]),
    "swb_exchange_block": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_uint64)]),
    "swb_link_peers": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.c_int,
                                      ctypes.c_int, ctypes.POINTER(c_void_p), ctypes.c_int, ctypes.c_int]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) libswb200.so and declare its prototypes."""
    global _lib
    if _lib is None:
        path = _build.LIB
        alt = os.environ.get("SWB_LIB")  # dev: an alternative build of the same sources (e.g. -DSWB_RES_TIMING)
        if alt:
            if not os.path.exists(alt):
                raise ImportError(f"SWB_LIB={alt} does not exist")
            path = alt
        elif _build.needs_build():
            try:
                path = _build.build()
            except _build.NvccMissing:
                # no toolkit on this box: a prebuilt library is acceptable, but say that it is older
                # than the sources (a compile ERROR is never swallowed -- stale kernels must not run
                # silently after an edit)
                if not os.path.exists(path):
                    raise
                warnings.warn(f"{path} is older than its sources and nvcc is not available to rebuild it", RuntimeWarning)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing and cannot be built: the CUDA path is the only path")
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here means header and library disagree
            fn.restype, fn.argtypes = res, args
        if L.swb_abi_version() != 3:
            raise ImportError("libswb200.so ABI version mismatch")
        _lib = L
    return _lib


def check(code: int):
    if code != 0:
        raise SwbError(code, load().swb_last_error().decode("utf-8", "replace"))
