"""_lib.py — ctypes binding of libqups_b200.so (include/qups_b200.h).

The product path: there is NO CPU fallback.  If the CUDA library is missing or
cannot be loaded, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUPS_B200_LIB") or os.path.join(_HERE, "libqups_b200.so")  # override: tuning builds

F32, F16, F64 = 0, 1, 2
NEAREST, LINEAR, CUBIC, LANCZOS3 = 0, 1, 2, 3
FLAG_KEEP_RX, FLAG_KEEP_TX, FLAG_TRANSPOSE = 8, 16, 32
PATH_AUTO, PATH_GENERIC, PATH_TILED = 0, 1, 2
INTERP = {"nearest": NEAREST, "linear": LINEAR, "cubic": CUBIC, "lanczos3": LANCZOS3}

EXPORTS = (
    "qups_das", "qups_delays", "qups_das_host", "qups_modulate", "qups_wsinterpd2", "qups_wsinterpd",
    "qups_das_fused", "qups_das_cohfac", "qups_apod_generate", "qups_chd_prep", "qups_aperture", "qups_greens", "qups_convd", "qups_pwznxcorr", "qups_refocus", "qups_host_release", "qups_last_error", "qups_version", "qups_launch_count", "qups_last_das_kernel", "qups_last_ws2_kernel",
)


class QupsError(RuntimeError):
    """Raised for a negative qups_status; mirrors MATLAB error() at the reference's call sites."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[qups_b200 {code}] {msg}")
        self.code = code


class DasParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32),
        ("I1", C.c_uint64), ("I2", C.c_uint64), ("I3", C.c_uint64),
        ("N", C.c_uint64), ("M", C.c_uint64), ("T", C.c_uint64),
        ("F", C.c_uint64), ("S", C.c_uint64),
        ("flag", C.c_int32), ("vs", C.c_int32), ("dv", C.c_int32),
        ("apod_real", C.c_int32), ("y_f32", C.c_int32), ("path", C.c_int32),
        ("accumulate", C.c_int32), ("host_chunks", C.c_int32), ("y_device", C.c_int32), ("reserved_", C.c_int32),
        ("fs", C.c_double), ("fmod", C.c_double),
        ("x_frame_stride", C.c_uint64), ("y_frame_stride", C.c_uint64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
        ("pitch_hint", C.c_double * 2), ("c_hint", C.c_double),
    ]


class Ws2Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32),
        ("T", C.c_uint64),
        ("D", C.c_uint32), ("interp", C.c_int32), ("w_real", C.c_int32), ("y_f32", C.c_int32),
        ("omega", C.c_double),
        ("sizes", C.c_uint64 * 8),
        ("dstride", C.c_uint64 * 40),
    ]


class GreensParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32),
        ("I", C.c_uint64), ("S", C.c_uint64), ("T", C.c_uint64), ("N", C.c_uint64), ("M", C.c_uint64),
        ("E", C.c_uint64),
        ("n0", C.c_int64),
        ("interp", C.c_int32), ("y_f32", C.c_int32),
        ("t0x", C.c_double), ("fs", C.c_double), ("fsr", C.c_double), ("c0", C.c_double), ("R0", C.c_double),
    ]


class ConvdParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("is_complex", C.c_int32), ("shape", C.c_int32),
        ("C", C.c_uint64), ("S", C.c_uint64), ("Lx", C.c_uint64), ("Ly", C.c_uint64), ("yC", C.c_uint64), ("yS", C.c_uint64),
    ]


class ApodFused(C.Structure):
    """qups_apod_fused (include/qups_b200.h): closed-form apodization evaluated inside the DAS kernel."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("rx_kind", C.c_int32), ("tx_kind", C.c_int32), ("lat_dim", C.c_int32),
        ("rx_p", C.c_float * 4), ("tx_p", C.c_float * 4),
        ("rx_aux", C.c_void_p), ("tx_aux", C.c_void_p), ("lat", C.c_void_p),
    ]


class PrepParams(C.Structure):
    """qups_prep_params: fused ChannelData pre-processing (zeropad -> hilbert -> downmix -> cast)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("in_dtype", C.c_int32), ("out_dtype", C.c_int32), ("hilbert", C.c_int32),
        ("T", C.c_uint64), ("K", C.c_uint64), ("B", C.c_uint64), ("A", C.c_uint64),
        ("traces_per_t0", C.c_uint64), ("n_t0", C.c_uint64),
        ("fs", C.c_double), ("fmix", C.c_double),
    ]


IN_REAL_F32, IN_CPLX_F32, IN_REAL_I16, IN_REAL_F64 = range(4)


class XcorrParams(C.Structure):
    """qups_xcorr_params: pair-wise windowed zero-normalised cross-correlation (kern/pwznxcorr.m)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("is_complex", C.c_int32), ("ref", C.c_int32),
        ("zero", C.c_int32), ("norm", C.c_int32), ("pad", C.c_int32), ("stride", C.c_uint32),
        ("L", C.c_uint32), ("W", C.c_uint32),
        ("T", C.c_uint64), ("N", C.c_uint64), ("F", C.c_uint64), ("x0N", C.c_uint64), ("x0F", C.c_uint64),
    ]


XC_NEIGHBOR, XC_CENTER, XC_X0 = range(3)


class RefocusParams(C.Structure):
    """qups_refocus_params: REFoCUS decode (src/UltrasoundSystem.m:3729-3757)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32),
        ("T", C.c_uint64), ("N", C.c_uint64), ("V", C.c_uint64), ("E", C.c_uint64),
        ("n_t0", C.c_uint32), ("reserved_", C.c_uint32), ("fs", C.c_double),
    ]


class ApertureParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("op", C.c_int32), ("nlags", C.c_uint32),
        ("C", C.c_uint64), ("A", C.c_uint64), ("S", C.c_uint64), ("gamma", C.c_double),
    ]


APD_COHFAC, APD_DMAS, APD_PCF, APD_SLSC_AVERAGE, APD_SLSC_ENSEMBLE = range(5)
AP_RX_NONE, AP_RX_ACCEPTANCE_ANGLE, AP_RX_COSINE_ANGLE, AP_RX_APERTURE_GROWTH, AP_RX_TRANSLATING = range(5)
AP_TX_NONE, AP_TX_SCANLINE, AP_TX_TRANSLATING, AP_TX_PARALLELOGRAM = range(4)

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly when it is absent (no CPU fallback on this path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m qups_b200.build` (nvcc, sm_100a). "
            "qups_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.qups_das.argtypes = [C.POINTER(DasParams), vp, vp, vp, vp, vp, vp, vp, u64p, vp, vp]
    L.qups_das_fused.argtypes = [C.POINTER(DasParams), C.POINTER(ApodFused), vp, vp, vp, vp, vp, vp, vp, u64p, vp, vp]
    L.qups_das_cohfac.argtypes = [C.POINTER(DasParams), vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.qups_das_cohfac.restype = C.c_int
    L.qups_apod_generate.argtypes = [C.POINTER(ApodFused), C.c_int32, vp, C.c_int32, vp, vp, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_uint64, vp]
    L.qups_chd_prep.argtypes = [C.POINTER(PrepParams), vp, vp, vp, vp]
    L.qups_aperture.argtypes = [C.POINTER(ApertureParams), vp, vp, vp, C.POINTER(C.c_uint32), vp]
    L.qups_delays.argtypes = [C.POINTER(DasParams), vp, vp, vp, vp, vp, vp, u64p, vp]
    L.qups_das_host.argtypes = [C.POINTER(DasParams), vp, vp, vp, vp, vp, vp, C.c_uint64, vp, C.c_uint64, u64p, vp,
                                C.c_int]
    L.qups_modulate.argtypes = [C.c_int32, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_double,
                                C.c_double, vp]
    L.qups_wsinterpd2.argtypes = [C.POINTER(Ws2Params), vp, vp, vp, vp, vp, vp]
    L.qups_wsinterpd.argtypes = [C.POINTER(Ws2Params), vp, vp, vp, vp, vp]
    L.qups_greens.argtypes = [C.POINTER(GreensParams), vp, vp, vp, vp, vp, vp, vp]
    L.qups_convd.argtypes = [C.POINTER(ConvdParams), vp, vp, vp, vp]
    L.qups_convd.restype = C.c_int
    L.qups_pwznxcorr.argtypes = [C.POINTER(XcorrParams), vp, vp, vp, vp, C.POINTER(C.c_int32), vp]
    L.qups_pwznxcorr.restype = C.c_int
    L.qups_refocus.argtypes = [C.POINTER(RefocusParams), vp, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
    L.qups_refocus.restype = C.c_int
    for f in ("qups_das", "qups_das_fused", "qups_das_cohfac", "qups_apod_generate", "qups_chd_prep", "qups_aperture", "qups_delays", "qups_das_host", "qups_modulate", "qups_wsinterpd2", "qups_wsinterpd",
              "qups_greens", "qups_version"):
        getattr(L, f).restype = C.c_int
    L.qups_last_error.restype = C.c_char_p
    L.qups_last_das_kernel.restype = C.c_char_p
    L.qups_last_ws2_kernel.restype = C.c_char_p
    L.qups_launch_count.restype = C.c_uint64
    L.qups_launch_count.argtypes = [C.c_int]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise QupsError(rc, lib().qups_last_error().decode("utf-8", "replace"))


def launch_count(reset: bool = False) -> int:
    return int(lib().qups_launch_count(1 if reset else 0))


def last_das_kernel() -> str:
    return lib().qups_last_das_kernel().decode()


def last_ws2_kernel() -> str:
    return lib().qups_last_ws2_kernel().decode()
