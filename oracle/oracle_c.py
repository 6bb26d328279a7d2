"""oracle_c.py — ctypes loader for the C oracle (oracle/libqups_oracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/qups_oracle.h): the checker for tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
FUN = {"DAS": 0, "SYN": 1, "MUL": 2, "BF": 3, "delays": 4}
INTERP = {"nearest": 0, "linear": 1, "cubic": 2, "lanczos3": 3}
MAX_APOD = 8


class DasArgs(C.Structure):
    _fields_ = [
        ("fun", C.c_int32), ("interp", C.c_int32), ("VS", C.c_int32), ("DV", C.c_int32),
        ("tpose", C.c_int32), ("S", C.c_int32), ("apod_complex", C.c_int32), ("pad_", C.c_int32),
        ("I", C.c_uint64 * 3), ("N", C.c_uint64), ("M", C.c_uint64), ("T", C.c_uint64), ("F", C.c_uint64),
        ("fs", C.c_double), ("fmod", C.c_double),
        ("Pi", C.c_void_p), ("Pr", C.c_void_p), ("Pv", C.c_void_p), ("Nv", C.c_void_p),
        ("x", C.c_void_p), ("t0", C.c_void_p), ("cinv", C.c_void_p),
        ("csz", C.c_uint64 * 5),
        ("apod", C.c_void_p * MAX_APOD),
        ("asz", (C.c_uint64 * 5) * MAX_APOD),
    ]


class Ws2Args(C.Structure):
    _fields_ = [
        ("interp", C.c_int32), ("sum_n", C.c_int32), ("sum_m", C.c_int32), ("w_complex", C.c_int32),
        ("I", C.c_uint64), ("N", C.c_uint64), ("M", C.c_uint64), ("T", C.c_uint64),
        ("omega", C.c_double),
        ("x", C.c_void_p),
        ("t1", C.c_void_p), ("t1sz", C.c_uint64 * 3),
        ("t2", C.c_void_p), ("t2sz", C.c_uint64 * 3),
        ("w", C.c_void_p), ("wsz", C.c_uint64 * 3),
    ]


class GreensArgs(C.Structure):
    _fields_ = [
        ("interp", C.c_int32), ("pad_", C.c_int32),
        ("S", C.c_uint64), ("N", C.c_uint64), ("M", C.c_uint64), ("T", C.c_uint64),
        ("K", C.c_uint64), ("E", C.c_uint64),
        ("n0", C.c_int64),
        ("c0", C.c_double), ("fs", C.c_double), ("fsr", C.c_double), ("R0", C.c_double), ("wv_t0", C.c_double),
        ("ps", C.c_void_p), ("amp", C.c_void_p), ("pn", C.c_void_p), ("pv", C.c_void_p), ("kern", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref PTX)."""
    so = os.path.join(_HERE, "libqups_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("qups_oracle.c", "oracle_body.inc", "qups_oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for sfx in ("_f", "_d"):
            getattr(_LIB, "oracle_das" + sfx).argtypes = [C.POINTER(DasArgs), C.c_void_p]
            getattr(_LIB, "oracle_wsinterpd2" + sfx).argtypes = [C.POINTER(Ws2Args), C.c_void_p]
            getattr(_LIB, "oracle_greens" + sfx).argtypes = [C.POINTER(GreensArgs), C.c_void_p]
        _LIB.oracle_interp1_f.argtypes = [C.c_void_p, C.c_long, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.oracle_interp1_d.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.oracle_interp1_f.restype = None
        _LIB.oracle_interp1_d.restype = None
    return _LIB


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def _f(a, dt):
    return np.asfortranarray(np.asarray(a, dtype=dt))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _sz5(a, Isz, N, M):
    shp = list(a.shape) + [1] * (5 - a.ndim)
    full = list(Isz) + [N, M]
    for d in range(5):
        if shp[d] not in (1, full[d]):
            raise ValueError("size inconsistent with I1 x I2 x I3 x N x M")
    return shp


def interp1(v, xq, method="linear", dtype=np.float32):
    rdt = np.dtype(dtype)
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    v = np.ascontiguousarray(np.asarray(v, dtype=cdt))
    xq = np.asarray(xq, dtype=rdt)
    out = np.empty(xq.shape, dtype=cdt)
    fn = lib().oracle_interp1_f if rdt == np.float32 else lib().oracle_interp1_d
    re = rdt.type(0)
    yr = (C.c_float if rdt == np.float32 else C.c_double)()
    yi = (C.c_float if rdt == np.float32 else C.c_double)()
    flat = out.reshape(-1)
    for j, q in enumerate(xq.reshape(-1)):
        fn(_ptr(v), v.shape[0], q.item(), INTERP[method], C.byref(yr), C.byref(yi))
        flat[j] = complex(yr.value, yi.value)
    del re
    return out


def das_spec(fun, Pi, Pr, Pv, Nv, x, t0, fs, c=1540.0, *, interp="linear", apod=(), VS=True, DV=False,
             fmod=0.0, tpose=False, dtype=np.float32):
    """Same contract as oracle_np.das_spec, computed by the C oracle."""
    rdt = np.dtype(dtype)
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    Pi = np.asarray(Pi, dtype=rdt)
    Isz = tuple(Pi.shape[1:]) + (1,) * (4 - Pi.ndim)
    I = int(np.prod(Isz))
    P = _f(Pi.reshape(3, I, order="F"), rdt)
    Pr = _f(np.asarray(Pr, rdt).reshape(3, -1), rdt)
    Pv = _f(np.asarray(Pv, rdt).reshape(3, -1), rdt)
    Nv = _f(np.asarray(Nv, rdt).reshape(3, -1), rdt)
    if fun != "delays":
        x = np.asarray(x)
        if x.ndim == 3:
            x = x[..., None]
        xf = _f(x, cdt)
        T, d2, d3, F = xf.shape
        N, M = (d3, d2) if tpose else (d2, d3)
    else:
        xf = np.zeros((1,), dtype=cdt)
        N, M, F, T = Pr.shape[1], max(Pv.shape[1], Nv.shape[1]), 1, 0
    if Pv.shape[1] == 1:
        Pv = _f(np.repeat(Pv, M, axis=1), rdt)
    if Nv.shape[1] == 1:
        Nv = _f(np.repeat(Nv, M, axis=1), rdt)
    if Pr.shape[1] == 1:
        Pr = _f(np.repeat(Pr, N, axis=1), rdt)
    cinv = _f((rdt.type(1) / np.asarray(c, dtype=rdt)).astype(rdt), rdt)
    t0v = _f(np.broadcast_to(np.asarray(t0, dtype=rdt).reshape(-1), (M,)), rdt)
    A = DasArgs()
    A.fun, A.interp, A.VS, A.DV, A.tpose = FUN[fun], INTERP[interp], int(VS), int(DV), int(tpose)
    A.I[0], A.I[1], A.I[2] = Isz
    A.N, A.M, A.T, A.F = N, M, T, F
    A.fs, A.fmod = float(fs), float(fmod)
    A.Pi, A.Pr, A.Pv, A.Nv, A.x, A.t0, A.cinv = map(_ptr, (P, Pr, Pv, Nv, xf, t0v, cinv))
    for d, s in enumerate(_sz5(cinv, Isz, N, M)):
        A.csz[d] = s
    keep = []
    apod = list(apod)
    A.S = len(apod)
    A.apod_complex = int(any(np.iscomplexobj(a) for a in apod))
    for k, a in enumerate(apod):
        a = _f(a, cdt if A.apod_complex else rdt)
        keep.append(a)
        A.apod[k] = a.ctypes.data
        for d, s in enumerate(_sz5(a, Isz, N, M)):
            A.asz[k][d] = s
    if fun == "delays":
        y = np.empty((I, N, M), dtype=rdt, order="F")
        osz = Isz + (N, M)
    else:
        On = N if fun in ("SYN", "BF") else 1
        Om = M if fun in ("MUL", "BF") else 1
        y = np.empty((I, On, Om, F), dtype=cdt, order="F")
        osz = Isz + (On, Om, F)
    fn = lib().oracle_das_f if rdt == np.float32 else lib().oracle_das_d
    rc = fn(C.byref(A), _ptr(y))
    if rc != 0:
        raise RuntimeError(f"oracle_das failed: {rc}")
    return y.reshape(osz, order="F")


def wsinterpd2_inm(x, t1, t2, w=None, *, sum_n=True, sum_m=True, interp="linear", omega=0.0, dtype=np.float32):
    """Canonical (I,N,M) wsinterpd2: x (T,N,M); t1,t2,w broadcastable to (I,N,M)."""
    rdt = np.dtype(dtype)
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    xf = _f(x, cdt)
    T, N, M = xf.shape
    t1 = _f(np.asarray(t1, rdt), rdt)
    t2 = _f(np.asarray(t2, rdt), rdt)
    t1 = t1.reshape(t1.shape + (1,) * (3 - t1.ndim), order="F")
    t2 = t2.reshape(t2.shape + (1,) * (3 - t2.ndim), order="F")
    I = max(t1.shape[0], t2.shape[0])
    if w is None:
        w = np.ones((1, 1, 1), dtype=rdt)
    wc = np.iscomplexobj(w)
    w = _f(np.asarray(w), cdt if wc else rdt)
    w = w.reshape(w.shape + (1,) * (3 - w.ndim), order="F")
    I = max(I, w.shape[0])
    A = Ws2Args()
    A.interp, A.sum_n, A.sum_m, A.w_complex = INTERP[interp], int(sum_n), int(sum_m), int(wc)
    A.I, A.N, A.M, A.T = I, N, M, T
    A.omega = float(omega)
    A.x, A.t1, A.t2, A.w = map(_ptr, (xf, t1, t2, w))
    for d in range(3):
        A.t1sz[d], A.t2sz[d], A.wsz[d] = t1.shape[d], t2.shape[d], w.shape[d]
    On, Om = (1 if sum_n else N), (1 if sum_m else M)
    y = np.empty((I, On, Om), dtype=cdt, order="F")
    fn = lib().oracle_wsinterpd2_f if rdt == np.float32 else lib().oracle_wsinterpd2_d
    rc = fn(C.byref(A), _ptr(y))
    if rc != 0:
        raise RuntimeError(f"oracle_wsinterpd2 failed: {rc}")
    return y


def greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wv_t0, fsr=1.0, R0=0.0, interp="cubic", dtype=np.float32, E=1):
    """E > 1: pn / pv hold E sub-element positions per element, column n + N*en (3 x N x E flattened, src/UltrasoundSystem.m:785-790)."""
    rdt = np.dtype(dtype)
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    ps, pn, pv = (_f(np.asarray(a, rdt).reshape(3, -1), rdt) for a in (ps, pn, pv))
    amp = _f(amp, rdt)
    kern = _f(kern, cdt)
    A = GreensArgs()
    A.interp = INTERP[interp]
    A.S, A.N, A.M, A.T, A.K, A.E = ps.shape[1], pn.shape[1] // E, pv.shape[1] // E, T, kern.shape[0], E
    A.n0 = int(n0)
    A.c0, A.fs, A.fsr, A.R0, A.wv_t0 = float(c0), float(fs), float(fsr), float(R0), float(wv_t0)
    A.ps, A.amp, A.pn, A.pv, A.kern = map(_ptr, (ps, amp, pn, pv, kern))
    x = np.empty((T, A.N, A.M), dtype=cdt, order="F")
    fn = lib().oracle_greens_f if rdt == np.float32 else lib().oracle_greens_d
    rc = fn(C.byref(A), _ptr(x))
    if rc != 0:
        raise RuntimeError(f"oracle_greens failed: {rc}")
    return x
