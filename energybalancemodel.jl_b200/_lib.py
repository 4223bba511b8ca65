"""ctypes binding of libebm_cuda.so (include/ebm_cuda.h) -- the same symbols the Julia extension ccalls.

There is no CPU fallback: if the shared library is missing, or it reports an error (e.g. no CUDA
device), the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EBM_CUDA_LIB", os.path.join(_HERE, "lib", "libebm_cuda.so"))

EBM_OK, EBM_ERR_INVALID, EBM_ERR_CUDA, EBM_ERR_OOM, EBM_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
CLASSIC_NPAR, MIZ_NPAR, NFORCING = 15, 22, 10
CLASSIC_NVAR, MIZ_NVAR, NSEASON, NDIAG = 3, 10, 3, 4

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)

# every symbol include/ebm_cuda.h declares
EXPORTED_SYMBOLS = (
    "ebm_version", "ebm_last_error", "ebm_device_count", "ebm_launch_count", "ebm_shutdown",
    "ebm_classic_run", "ebm_classic_run_device", "ebm_classic_step", "ebm_classic_step_debug",
    "ebm_miz_run", "ebm_miz_run_device", "ebm_miz_step",
    "ebm_transpose_device", "ebm_fp64_peak", "ebm_classic_run_multi", "ebm_miz_run_multi",
)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nt", C.c_int32), ("dur", C.c_int32), ("grid_kind", C.c_int32),
                ("winter_inx", C.c_int32), ("summer_inx", C.c_int32), ("x", _dp), ("t", _dp)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("lastonly", C.c_int32), ("field_stride", C.c_int32), ("strict", C.c_int32),
                ("years_per_launch", C.c_int32), ("newton_maxit", C.c_int32), ("newton_tol", C.c_double),
                ("step_limit", C.c_int32), ("start_year", C.c_int32), ("classic_stencil", C.c_int32)]


class ClassicOutputs(C.Structure):
    _fields_ = [("diag", _dp), ("seasonal", _dp), ("raw", _dp), ("E_final", _dp), ("Tg_final", _dp), ("flags", _i32p)]


class MizOutputs(C.Structure):
    _fields_ = [("diag", _dp), ("seasonal", _dp), ("raw", _dp), ("Ei_final", _dp), ("Ew_final", _dp), ("h_final", _dp),
                ("D_final", _dp), ("phi_final", _dp), ("T0_final", _dp), ("newton_iters", _i64p), ("nonconv", _i64p),
                ("flags", _i32p)]


class ClassicDeviceArgs(C.Structure):
    _fields_ = [("nmem", C.c_int64), ("par", C.c_void_p), ("forc", C.c_void_p), ("E", C.c_void_p), ("Tg", C.c_void_p),
                ("diag", C.c_void_p), ("seasonal", C.c_void_p), ("raw", C.c_void_p), ("flags", C.c_void_p),
                ("member_index", C.c_void_p)]


class MizDeviceArgs(C.Structure):
    _fields_ = [("nmem", C.c_int64), ("par", C.c_void_p), ("forc", C.c_void_p), ("Ei", C.c_void_p), ("Ew", C.c_void_p),
                ("h", C.c_void_p), ("D", C.c_void_p), ("phi", C.c_void_p), ("T0", C.c_void_p), ("diag", C.c_void_p),
                ("seasonal", C.c_void_p), ("raw", C.c_void_p), ("newton_iters", C.c_void_p), ("nonconv", C.c_void_p),
                ("flags", C.c_void_p)]


class Multi(C.Structure):
    _fields_ = [("ndevices", C.c_int32), ("diag_device", C.c_int32), ("packet", C.c_int32), ("reserved", C.c_int32),
                ("devices", _i32p)]


class EBMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libebm_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen libebm_cuda.so and declare prototypes.  Fails loudly if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -m ebm_b200.build` "
                          "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    lib.ebm_version.restype = C.c_char_p
    lib.ebm_last_error.restype = C.c_char_p
    lib.ebm_device_count.restype = C.c_int32
    lib.ebm_launch_count.restype = C.c_int64
    lib.ebm_shutdown.restype = C.c_int32
    lib.ebm_classic_run.restype = C.c_int32
    lib.ebm_classic_run.argtypes = [C.POINTER(Grid), C.c_int64, _dp, _dp, _dp, _dp, C.POINTER(Options), C.POINTER(ClassicOutputs)]
    lib.ebm_classic_run_device.restype = C.c_int32
    lib.ebm_classic_run_device.argtypes = [C.POINTER(Grid), C.POINTER(ClassicDeviceArgs), C.POINTER(Options), C.c_void_p]
    lib.ebm_classic_step.restype = C.c_int32
    lib.ebm_classic_step.argtypes = [C.POINTER(Grid), _dp, C.c_int32, C.c_double, _dp, _dp, _dp, _dp]
    lib.ebm_classic_step_debug.restype = C.c_int32
    lib.ebm_classic_step_debug.argtypes = [C.POINTER(Grid), _dp, C.c_int32, C.c_double, _dp, _dp, _dp, _dp, C.c_int32, _dp]
    lib.ebm_miz_run.restype = C.c_int32
    lib.ebm_miz_run.argtypes = [C.POINTER(Grid), C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(Options), C.POINTER(MizOutputs)]
    lib.ebm_miz_run_device.restype = C.c_int32
    lib.ebm_miz_run_device.argtypes = [C.POINTER(Grid), C.POINTER(MizDeviceArgs), C.POINTER(Options), C.c_void_p]
    lib.ebm_miz_step.restype = C.c_int32
    lib.ebm_miz_step.argtypes = [C.POINTER(Grid), _dp, C.c_int32, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _i32p]
    lib.ebm_classic_run_multi.restype = C.c_int32
    lib.ebm_classic_run_multi.argtypes = [C.POINTER(Grid), C.c_int64, _dp, _dp, _dp, _dp, C.POINTER(Options), C.POINTER(Multi),
                                          C.POINTER(ClassicOutputs)]
    lib.ebm_miz_run_multi.restype = C.c_int32
    lib.ebm_miz_run_multi.argtypes = [C.POINTER(Grid), C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(Options),
                                      C.POINTER(Multi), C.POINTER(MizOutputs)]
    lib.ebm_transpose_device.restype = C.c_int32
    lib.ebm_transpose_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.ebm_fp64_peak.restype = C.c_int32
    lib.ebm_fp64_peak.argtypes = [C.c_int32, _dp, _dp]
    _lib = lib
    return lib


def check(rc: int):
    if rc != EBM_OK:
        raise EBMError(rc, load().ebm_last_error().decode("utf-8", "replace"))


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def make_grid(st) -> Grid:
    """Marshal a SpaceTime verbatim (x, t, season indices) -- nothing is re-derived on the C side."""
    g = Grid(st.nx, st.nt, st.dur, st.grid_kind, st.winter.inx, st.summer.inx, dptr(st.x), dptr(st.t))
    g._keep = (st.x, st.t)  # keep the arrays alive as long as the struct
    return g


def make_multi(ndevices=0, devices=None, diag_device=-1, packet=0) -> Multi:
    """ebm_multi_t: GPUs of one *_run_multi call (0 = all visible), where the diagnostics go, packet size."""
    arr = None
    if devices is not None:
        if len(set(devices)) != len(devices) or min(devices, default=0) < 0:
            raise ValueError(f"devices must be distinct non-negative CUDA ordinals, got {list(devices)}")
        arr = (C.c_int32 * len(devices))(*devices)
        ndevices = len(devices)
    m = Multi(ndevices, diag_device, packet, 0, C.cast(arr, _i32p) if arr is not None else None)
    m._keep = arr
    return m


def make_options(device=-1, lastonly=True, field_stride=0, strict=False, years_per_launch=0, newton_maxit=0,
                 newton_tol=0.0, step_limit=0, start_year=0, classic_stencil=0) -> Options:
    return Options(device, int(lastonly), field_stride, int(strict), years_per_launch, newton_maxit, newton_tol,
                   step_limit, start_year, int(classic_stencil))
