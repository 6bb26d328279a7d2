"""kern.py — host-side mirror of the reference's L2 "kern" functions for the hot path.

Same names, option strings, argument meaning and error behaviour as

    kern/das_spec.m     das_spec(fun, Pi, Pr, Pv, Nv, x, t0, fs, c, varargin)
    kern/wsinterpd2.m   wsinterpd2(x, t1, t2, dim, w, sdim, interp, extrapval, omega)
    kern/wsinterpd.m    wsinterpd(x, t, dim, w, sdim, interp, extrapval, omega)

so the parity tests read like the reference's own.  The reference host language
is MATLAB (absent from this image); this module is the Python stand-in for the
MATLAB glue in matlab/ + mex/ and calls the SAME C ABI (include/qups_b200.h).
All computation happens in libqups_b200.so on the GPU; there is no CPU path
here ('device', 0 raises).  PyTorch is used only for device memory and streams.

Arrays use MATLAB's logical shapes (x is T x N x M x F..., Pi is 3 x I1 x I2 x I3)
and may be NumPy arrays (copied to the GPU, result returned as NumPy) or CUDA
torch tensors (result is a CUDA tensor).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DasParams, Ws2Params, QupsError

_RT = {"single": torch.float32, "double": torch.float64, "halfT": torch.float32}
_CT = {"single": torch.complex64, "double": torch.complex128}
_DT = {"single": _lib.F32, "double": _lib.F64, "halfT": _lib.F16}


def _as_tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(a) if np.ndim(a) == 0 else np.asarray(a))


def _colmajor(t: torch.Tensor, dtype, device) -> torch.Tensor:
    """Contiguous device tensor whose memory is the column-major image of `t` (shape reversed)."""
    if t.ndim > 1:
        t = t.permute(*reversed(range(t.ndim)))
    return t.to(device=device, dtype=dtype).contiguous()


def _from_colmajor(buf: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    v = buf.reshape(tuple(reversed(shape)))
    return v.permute(*reversed(range(len(shape)))) if len(shape) > 1 else v


def _cplx_buf(t: torch.Tensor, prec: str, device) -> torch.Tensor:
    """Column-major interleaved complex buffer in the precision's storage type."""
    if prec == "halfT":
        c = _colmajor(t if t.is_complex() else t.to(torch.complex64), torch.complex64, device)
        return torch.view_as_real(c).to(torch.float16).contiguous()
    c = _colmajor(t if t.is_complex() else t.to(_CT[prec]), _CT[prec], device)
    return c


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _mod_size(P: torch.Tensor) -> torch.Tensor:
    """parse_inputs/modSize: coordinates in the 1st dimension (kern/das_spec.m:587-597)."""
    if P.ndim == 1:
        P = P.reshape(-1, 1)
    if P.shape[0] <= 4:
        return P
    if P.ndim == 2 and P.shape[1] <= 4:
        return P.T
    return P


def _mod_dim(P: torch.Tensor) -> torch.Tensor:
    """expand_inputs/modDim: lift 1-D / 2-D / 4-D coordinates to 3-D (kern/das_spec.m:656-669)."""
    d = P.shape[0]
    z = torch.zeros_like(P[:1])
    if d == 1:
        return torch.cat([P, z, z], 0)
    if d == 2:
        return torch.cat([P[:1], z, P[1:2]], 0)
    if d == 3:
        return P
    if d == 4:
        return P[:3] / P[3:4]
    raise ValueError("Improper coordinate dimension.")


def _stride5(sz):
    """[1; cumprod(sz(1:end-1))] .* (sz ~= 1)   (kern/das_spec.m:259-260)."""
    st, acc = [], 1
    for s in sz:
        st.append(acc if s != 1 else 0)
        acc *= s
    return st


class FusedApod:
    """Closed-form apodization: what the reference's ap* generators
    (src/UltrasoundSystem.m:4892-5429) return as dense ND masks, kept as a parameter block and evaluated inside the
    DAS kernel.  Pass it wherever an 'apod' array goes; `dense()` materialises the array the reference would return
    (on the GPU, by the same device functions).  rx_* weights depend on (pixel, receive), tx_* on (pixel, transmit)."""

    def __init__(self, rx_kind=0, rx_p=(), rx_aux=None, tx_kind=0, tx_p=(), tx_aux=None, lat=None, lat_dim=2, name=""):
        self.rx_kind, self.rx_p, self.rx_aux = int(rx_kind), tuple(float(v) for v in rx_p), rx_aux
        self.tx_kind, self.tx_p, self.tx_aux = int(tx_kind), tuple(float(v) for v in tx_p), tx_aux
        self.lat, self.lat_dim, self.name = lat, int(lat_dim), name

    def merged(self, o: "FusedApod") -> "FusedApod":
        """Product of two closed-form apodizations; each side (receive / transmit) can be closed-form only once."""
        if (self.rx_kind and o.rx_kind) or (self.tx_kind and o.tx_kind):
            raise QupsError(-3, "two closed-form apodizations on the same aperture: pass one of them as dense()")
        if self.lat is not None and o.lat is not None and (self.lat_dim != o.lat_dim or not np.array_equal(self.lat, o.lat)):
            raise QupsError(-1, "inconsistent lateral pixel coordinates")
        a, b = (self, o) if self.rx_kind else (o, self)
        t = self if self.tx_kind else o
        l = self if self.lat is not None else o
        return FusedApod(a.rx_kind, a.rx_p, a.rx_aux, t.tx_kind, t.tx_p, t.tx_aux, l.lat, l.lat_dim, self.name + "*" + o.name)

    def _struct(self, dev, keep: list):
        f = _lib.ApodFused()
        f.struct_size = C.sizeof(_lib.ApodFused)
        f.rx_kind, f.tx_kind, f.lat_dim = self.rx_kind, self.tx_kind, self.lat_dim
        for k, v in enumerate(self.rx_p[:4]): f.rx_p[k] = v
        for k, v in enumerate(self.tx_p[:4]): f.tx_p[k] = v
        for nm in ("rx_aux", "tx_aux", "lat"):
            v = getattr(self, nm)
            if v is not None:  # MATLAB-shaped (rows x count) -> column-major fp32 on the device
                t = _colmajor(_as_tensor(np.asarray(v, np.float64) if not isinstance(v, torch.Tensor) else v), torch.float32, dev)
                keep.append(t)
                setattr(f, nm, t.data_ptr())
        return f

    def dense(self, Pi, Pr=None, M=None, which="rx", complex_=False):
        """The dense array the reference generator returns: I1 x I2 x I3 x N (which='rx') or I1 x I2 x I3 x 1 x M ('tx')."""
        dev = torch.device("cuda", torch.cuda.current_device())
        Pi_t = _as_tensor(Pi)
        Pi_t = Pi_t.reshape(tuple(Pi_t.shape) + (1,) * (4 - Pi_t.ndim))
        Isz = tuple(int(v) for v in Pi_t.shape[1:4])
        keep = []
        f = self._struct(dev, keep)
        dPi = _colmajor(_mod_dim(Pi_t), torch.float32, dev)
        if which == "rx":
            Pr_t = _mod_dim(_mod_size(_as_tensor(Pr)))
            NM = int(Pr_t.shape[1])
            dPr = _colmajor(Pr_t, torch.float32, dev)
        else:
            NM, dPr = int(M), None
        out = torch.empty(int(np.prod(Isz)) * NM, dtype=torch.complex64 if complex_ else torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().qups_apod_generate(C.byref(f), 0 if which == "rx" else 1, _ptr(out), int(complex_), _ptr(dPi),
                                                     _ptr(dPr), Isz[0], Isz[1], Isz[2], NM, _stream(dev)))
        y = _from_colmajor(out, Isz + ((NM,) if which == "rx" else (1, NM)))
        numpy_out = not (isinstance(Pi, torch.Tensor) and Pi.is_cuda)
        return np.asfortranarray(y.cpu().numpy()) if numpy_out else y


def das_spec(fun, Pi, Pr, Pv, Nv, x, t0, fs=None, c=1540.0, *varargin, _path=_lib.PATH_AUTO, _y_f32=False):
    """Specialised delay-and-sum beamformer — mirror of ``kern/das_spec.m:1``.

    fun in {'DAS','SYN','MUL','BF','delays'} — plus 'DAS+cohfac' (not a reference fun: the sequence
    ``b = das_spec('SYN', ...); y = sum(b, 4); r = cohfac(b, 4)`` of kern/cohfac.m as ONE call that never materialises the I x N
    cube; returns (y, r)); options (strings, as the reference :113-148):
    'plane-waves' | 'virtual-source' | 'diverging-waves' | 'focused-waves' |
    'input-precision', {'single','double','halfT'} | 'device', d | 'interp', method |
    'apod', A (repeatable) | 'modulation', fmod | 'transpose', tf.
    Returns I1 x I2 x I3 x [1|N] x [1|M] x F x ...
    """
    if fun not in ("DAS", "SYN", "BF", "MUL", "delays", "DAS+cohfac"):
        raise ValueError("Invalid beamformer.")
    VS, DV, interp_type, apod, fmod, tpose, device = True, False, "linear", [], 0.0, False, -1
    fused = None  # closed-form apodization (FusedApod) passed through 'apod' 
    xt = _as_tensor(x) if x is not None else torch.zeros((0,))
    if xt.dtype in (torch.float64, torch.complex128):
        prec = "double"
    elif xt.dtype in (torch.float16,):
        prec = "halfT"
    else:
        prec = "single" if xt.dtype in (torch.float32, torch.complex64) else "double"
    n, args = 0, list(varargin)
    while n < len(args):
        o = args[n]
        if o == "plane-waves": VS = False
        elif o == "virtual-source": VS = True
        elif o == "diverging-waves": DV = True
        elif o == "focused-waves": DV = False
        elif o == "input-precision": n += 1; prec = str(args[n])
        elif o == "device": n += 1; device = int(args[n])
        elif o == "interp": n += 1; interp_type = str(args[n])
        elif o == "apod":
            n += 1
            if isinstance(args[n], FusedApod): fused = args[n] if fused is None else fused.merged(args[n])
            else: apod.append(args[n])
        elif o == "modulation": n += 1; fmod = float(args[n])
        elif o == "transpose": n += 1; tpose = bool(args[n])
        else:
            raise ValueError("Unrecognized option")
        n += 1
    if prec not in _DT:
        raise ValueError(f"unknown input-precision {prec!r}")
    if device == 0:
        raise QupsError(-3, "qups_b200 implements only the GPU branch of das_spec; 'device', 0 (native interp1) "
                            "is the reference's own CPU path")
    if interp_type not in _lib.INTERP:  # QUPS:das_spec:UnrecognizedInput  (kern/das_spec.m:203-207)
        raise ValueError(f"Unrecognized interpolation of type {interp_type}: must be one of "
                         "{'nearest', 'linear', 'cubic', 'lanczos3'}.")
    if fs is None:
        if fun == "delays":
            fs = 1.0
        else:
            t0a = np.asarray(t0, dtype=np.float64).reshape(-1)
            if t0a.size < 2:
                raise ValueError("Undefined sampling rate.")
            fs, t0 = float(np.mean(np.diff(t0a))), float(t0a.min())
    numpy_out = not any(isinstance(v, torch.Tensor) and v.is_cuda for v in (Pi, Pr, Pv, Nv, x))
    dev = torch.device("cuda", torch.cuda.current_device() if device < 0 else device - 1)
    L = _lib.lib()
    rt = _RT[prec]

    Pi_t, Pr_t, Pv_t, Nv_t = (_mod_size(_as_tensor(P)) for P in (Pi, Pr, Pv, Nv))
    Pi_t = Pi_t.reshape(tuple(Pi_t.shape) + (1,) * (4 - Pi_t.ndim))
    Isz = tuple(int(s) for s in Pi_t.shape[1:4])
    I = Isz[0] * Isz[1] * Isz[2]
    if fun == "delays":
        T = 0
        N = Pr_t.shape[1]
        M = max(Pv_t.shape[1], Nv_t.shape[1])
        fsz = ()
    else:
        xs = tuple(xt.shape) + (1,) * (3 - xt.ndim)
        T, d2, d3 = xs[:3]
        N, M = (d3, d2) if tpose else (d2, d3)
        fsz = tuple(xs[3:])
    F = int(np.prod(fsz)) if fsz else 1
    # expand_inputs (kern/das_spec.m:602-670)
    if Pv_t.shape[1] == 1: Pv_t = Pv_t.expand(Pv_t.shape[0], M)
    if Nv_t.shape[1] == 1: Nv_t = Nv_t.expand(Nv_t.shape[0], M)
    if Pr_t.shape[1] == 1: Pr_t = Pr_t.expand(Pr_t.shape[0], N)
    if not (Pv_t.shape[1] == M and Nv_t.shape[1] == M):
        raise AssertionError("Inconsistent transmitter data size.")
    if Pr_t.shape[1] != N:
        raise AssertionError("Inconsistent receiver data size.")
    cinv_t = 1.0 / _as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c).to(rt)  # cinv = 1./c  (:170)
    csz = tuple(cinv_t.shape) + (1,) * (5 - cinv_t.ndim)
    full = Isz + (N, M)
    if not all(csz[d] in (1, full[d]) for d in range(3)):
        raise AssertionError("Sound speed data size inconsistent with pixel data size")
    if csz[3] not in (1, N): raise AssertionError("Sound speed data size inconsistent with receiver data size")
    if csz[4] not in (1, M): raise AssertionError("Sound speed data size inconsistent with transmit data size")
    apod_t = [_as_tensor(a) for a in apod]
    apod_real = all(not a.is_complex() for a in apod_t)
    asz = []
    for a in apod_t:
        s = tuple(a.shape) + (1,) * (5 - a.ndim)
        if len(s) > 5 or not all(s[d] in (1, full[d]) for d in range(3)):
            raise AssertionError("Apodization data size inconsistent with pixel data size")
        if s[3] not in (1, N): raise AssertionError("Apodization data size inconsistent with receiver data size")
        if s[4] not in (1, M): raise AssertionError("Apodization data size inconsistent with transmit data size")
        asz.append(s)
    S = len(apod_t)

    # device buffers (column-major)
    # pixels: keep (d, I1, I2, I3) logical shape -> column-major buffer is 3 x I with I1 fastest
    dPi = _colmajor(_mod_dim(Pi_t), rt, dev)
    dPr = _colmajor(_mod_dim(Pr_t), rt, dev)
    dNv = _colmajor(_mod_dim(Nv_t), rt, dev)
    Pv3 = _mod_dim(Pv_t).to(rt)
    t0v = _as_tensor(np.asarray(t0, dtype=np.float64) if not isinstance(t0, torch.Tensor) else t0).reshape(-1).to(rt)
    if t0v.numel() not in (1, M):
        raise AssertionError("t0 must be a scalar or have one entry per transmit")
    Pv4 = torch.cat([Pv3, t0v.to(Pv3.device).expand(M).reshape(1, M)], 0)  # Pv(4,:) = t0   (:361)
    dPv = _colmajor(Pv4, rt, dev)
    dC = _colmajor(cinv_t.reshape(csz), rt, dev)
    if S:
        if prec == "halfT":
            parts = [(_cplx_buf(a.reshape(s), prec, dev).reshape(-1) if not apod_real else
                      _colmajor(a.reshape(s), torch.float16, dev).reshape(-1)) for a, s in zip(apod_t, asz)]
        else:
            parts = [(_colmajor(a.reshape(s), rt, dev) if apod_real else _cplx_buf(a.reshape(s), prec, dev)).reshape(-1)
                     for a, s in zip(apod_t, asz)]
        dA = torch.cat(parts)
    else:
        dA = None
    # [cstride, astride]   (kern/das_spec.m:256-260)
    acs = _stride5(csz) + [0]
    base = 0
    for s in asz:
        acs += _stride5(s) + [base]
        base += int(np.prod(s))
    acs_c = (C.c_uint64 * len(acs))(*acs)

    p = DasParams()
    p.struct_size = C.sizeof(DasParams)
    p.dtype = _DT[prec]
    p.I1, p.I2, p.I3 = Isz
    p.N, p.M, p.T, p.F, p.S = N, M, T, F, S
    keep_rx, keep_tx = fun in ("SYN", "BF"), fun in ("MUL", "BF")
    p.flag = _lib.INTERP[interp_type] + 8 * keep_rx + 16 * keep_tx + 32 * tpose  # :209-213
    p.vs, p.dv = int(VS), int(DV)
    p.apod_real, p.y_f32, p.path = int(apod_real), int(_y_f32), int(_path)
    p.fs, p.fmod = float(fs), float(fmod)
    if not Pi_t.is_cuda and cinv_t.numel() == 1 and Isz[0] > 1 and Isz[1] > 1 and Pi_t.shape[0] == 3:
        # geometry still on the host: hand the launcher the pixel pitch / sound speed so it never reads them back from the device
        P0 = Pi_t[:, 0, 0, 0].double()
        p.pitch_hint[0] = float(torch.linalg.norm(Pi_t[:, 1, 0, 0].double() - P0))
        p.pitch_hint[1] = float(torch.linalg.norm(Pi_t[:, 0, 1, 0].double() - P0))
        p.c_hint = float(1.0 / cinv_t.reshape(-1)[0])
    st = _stream(dev)

    with torch.cuda.device(dev):
        if fun == "delays":
            tau = torch.empty(I * N * M, dtype=rt, device=dev)
            _lib.check(L.qups_delays(C.byref(p), _ptr(tau), _ptr(dPi), _ptr(dPr), _ptr(dPv), _ptr(dNv), _ptr(dC),
                                     acs_c, st))
            y = _from_colmajor(tau, Isz + (N, M))
        else:
            # frames collapse in COLUMN-major order (f = f1 + F1*f2 + ..., as MATLAB's x(:,:,:,:) does): the output is
            # unfolded column-major below, so a row-major collapse would hand the frames back permuted
            xf = xt.reshape(xs)
            if len(xs) > 4:
                xf = xf.permute(0, 1, 2, *reversed(range(3, len(xs))))
            dX = _cplx_buf(xf.reshape(xs[:3] + (F,)), prec, dev)
            On, Om = (N if keep_rx else 1), (M if keep_tx else 1)
            if prec == "halfT" and not _y_f32:
                yb = torch.empty((F * Om * On * I, 2), dtype=torch.float16, device=dev)
            else:
                yb = torch.empty(F * Om * On * I, dtype=torch.complex64 if prec != "double" else torch.complex128, device=dev)
            if fun == "DAS+cohfac":
                if fused is not None or S != 0 or prec != "single" or F != 1:
                    raise _lib.QupsError(-3, "DAS+cohfac: plain weights only (no apodization, single precision, one frame)")
                cf = torch.empty(I, dtype=torch.float32, device=dev)
                _lib.check(L.qups_das_cohfac(C.byref(p), _ptr(yb), _ptr(cf), _ptr(dPi), _ptr(dPr), _ptr(dPv), _ptr(dNv), _ptr(dC),
                                             _ptr(dX), st))
                y, r = _from_colmajor(yb, Isz), _from_colmajor(cf, Isz)
                return (np.asfortranarray(y.cpu().numpy()), np.asfortranarray(r.cpu().numpy())) if numpy_out else (y, r)
            if fused is not None:
                keep = []
                fz = fused._struct(dev, keep)
                _lib.check(L.qups_das_fused(C.byref(p), C.byref(fz), _ptr(yb), _ptr(dPi), _ptr(dPr), _ptr(dPv), _ptr(dNv),
                                            _ptr(dA), _ptr(dC), acs_c, _ptr(dX), st))
            else:
                _lib.check(L.qups_das(C.byref(p), _ptr(yb), _ptr(dPi), _ptr(dPr), _ptr(dPv), _ptr(dNv), _ptr(dA), _ptr(dC),
                                      acs_c, _ptr(dX), st))
            if yb.dtype == torch.float16:
                yb = torch.view_as_complex(yb.float())
            y = _from_colmajor(yb, Isz + (On, Om) + (fsz if fsz else ()))
    if numpy_out:
        return np.asfortranarray(y.cpu().numpy())
    return y


def _sz(t, nd):
    return tuple(t.shape) + (1,) * (nd - t.ndim)


def _swapdim(t, a, b):
    nd = max(t.ndim, a + 1, b + 1)
    t = t.reshape(_sz(t, nd))
    return t.transpose(a, b) if a != b else t


def wsinterpd2(x, t1, t2, dim=1, w=1, sdim=(), interp="linear", extrapval=0, omega=0, _prec=None, _y_f32=False):
    """Weighted-sum interpolation with separable delays — mirror of ``kern/wsinterpd2.m:1``.

    y = sum_{sdim} w .* exp(omega .* (t1+t2)) .* interp1(x, 1 + t1 + t2, interp, 0)
    `dim`/`sdim` are 1-based as in MATLAB.  Only extrapval = 0 (what every hot-path caller passes:
    src/ChannelData.m:1445, src/UltrasoundSystem.m:843) is implemented.
    _prec='halfT' selects the half kernels (wsinterpd2h, src/interpd.cu:451-458): data, weights AND delay tables are passed as
    fp16 (torch has no complex-half type, so the mirror narrows here); the result comes back widened to complex64.
    """
    if interp not in _lib.INTERP:
        raise ValueError("Interp option not recognized: " + str(interp))
    if not (extrapval == 0):
        raise QupsError(-3, "qups_b200.wsinterpd2 implements extrapval = 0 only")
    if np.real(omega) != 0:
        raise QupsError(-3, "real(omega) ~= 0 is handled by the reference's native path only (kern/wsinterpd2.m:101)")
    xt, t1t = _as_tensor(x), _as_tensor(t1)
    t2t = _as_tensor(t2) if t2 is not None else None
    wt = _as_tensor(np.asarray(w)) if not isinstance(w, torch.Tensor) else w
    if t1t.is_complex() or (t2t is not None and t2t.is_complex()):
        raise AssertionError("Sample indices must be real.")
    numpy_out = not any(isinstance(v, torch.Tensor) and v.is_cuda for v in (x, t1, t2, w))
    dev = torch.device("cuda", torch.cuda.current_device())
    prec = "double" if xt.dtype in (torch.float64, torch.complex128) else "single"
    half = _prec == "halfT"
    rt, ct = _RT[prec], _CT[prec]
    d0 = dim - 1
    sd = [s - 1 for s in (sdim if np.ndim(sdim) else [sdim])]
    sd = [d0 if s == 0 else (0 if s == d0 else s) for s in sd]  # swap with the moved dim (:67)
    nd = max(xt.ndim, t1t.ndim, t2t.ndim if t2t is not None else 1, wt.ndim, dim, max(sd, default=0) + 1)
    xt, t1t, wt = (_swapdim(v, d0, 0) for v in (xt, t1t, wt))
    nd = max(nd, xt.ndim, t1t.ndim, wt.ndim)
    xs, s1, ws = _sz(xt, nd), _sz(t1t, nd), _sz(wt, nd)
    if t2t is not None:
        t2t = _swapdim(t2t, d0, 0)
        s2 = _sz(t2t, nd)
    else:
        s2 = (1,) * nd
    T = xs[0]
    dsz = [max(s1[0], s2[0])] + [max(s1[k], s2[k], xs[k]) for k in range(1, nd)]
    for k in range(nd):
        for nm, s in (("t1", s1), ("t2", s2), ("w", ws)) + ((("x", xs),) if k else ()):
            if s[k] not in (1, dsz[k]):
                raise AssertionError(f"size of {nm} incompatible in dim {k + 1}")
    osz = [1 if k in sd else dsz[k] for k in range(nd)]
    if nd > 8:
        raise QupsError(-3, "at most 8 dims")
    strides = [_stride5(ws), _stride5(osz), _stride5(s1), _stride5(s2), _stride5((1,) + tuple(xs[1:]))]
    p = Ws2Params()
    p.struct_size = C.sizeof(Ws2Params)
    p.dtype = _DT["halfT"] if half else _DT[prec]
    p.y_f32 = int(bool(_y_f32))
    p.T, p.D, p.interp = T, nd, _lib.INTERP[interp]
    w_real = not wt.is_complex()
    p.w_real = int(w_real)
    p.omega = float(np.imag(omega))
    for k in range(nd):
        p.sizes[k] = dsz[k]
        for r in range(5):
            p.dstride[r + 5 * k] = strides[r][k]
    tt = torch.float16 if half else rt
    dX = _cplx_buf(xt.reshape(xs), "halfT" if half else prec, dev)
    d1 = _colmajor(t1t.reshape(s1), tt, dev)
    d2 = _colmajor(t2t.reshape(s2), tt, dev) if t2t is not None else None
    dW = _colmajor(wt.reshape(ws), tt, dev) if w_real else _cplx_buf(wt.reshape(ws), "halfT" if half else prec, dev)
    if half and not _y_f32:
        yb = torch.zeros((int(np.prod(osz)), 2), dtype=torch.float16, device=dev)
    else:
        yb = torch.zeros(int(np.prod(osz)), dtype=ct, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qups_wsinterpd2(C.byref(p), _ptr(yb), _ptr(dW), _ptr(dX), _ptr(d1), _ptr(d2), _stream(dev)))
    if yb.dtype == torch.float16:
        yb = torch.view_as_complex(yb.float())
    y = _swapdim(_from_colmajor(yb, osz), 0, d0)
    return np.asfortranarray(y.cpu().numpy()) if numpy_out else y


def wsinterpd(x, t, dim=1, w=1, sdim=(), interp="linear", extrapval=0, omega=0, _prec=None, _y_f32=False):
    """Mirror of ``kern/wsinterpd.m:1`` (single delay table)."""
    return wsinterpd2(x, t, None, dim, w, sdim, interp, extrapval, omega, _prec=_prec, _y_f32=_y_f32)


def convd(x, y=None, dim=None, shape="full", _half=False):
    """Batched 1-D convolution along one dimension — mirror of ``kern/convd.m:1`` (GPU branch :135-201).

    C = convd(A, B, dim, shape); B defaults to conj(flip(A)) (auto-correlation); dim (1-based) defaults to the
    first non-singleton dimension; shape in {'full','same','valid'}.  Returns (C, lags) with lags as :102-110.
    B must match A outside `dim`, or be singleton in all dims before and/or all dims after `dim`.
    """
    if shape not in ("full", "same", "valid"):
        raise ValueError("shape must be one of {'full', 'same', 'valid'}")
    xt = _as_tensor(x)
    numpy_out = not (isinstance(x, torch.Tensor) and x.is_cuda)
    if dim is None:
        dim = next((d + 1 for d, s in enumerate(xt.shape) if s != 1), 1)
    d0 = dim - 1
    nd = max(xt.ndim, dim)
    xs = _sz(xt, nd)
    yt = torch.conj(torch.flip(xt.reshape(xs), [d0])).resolve_conj() if y is None else _as_tensor(y)
    ys = _sz(yt, nd)
    cplx = xt.is_complex() or yt.is_complex()
    dbl = xt.dtype in (torch.float64, torch.complex128) or yt.dtype in (torch.float64, torch.complex128)
    dt = (torch.complex128 if dbl else torch.complex64) if cplx else (torch.float64 if dbl else torch.float32)
    C_, S_ = int(np.prod(xs[:d0])) if d0 else 1, int(np.prod(xs[d0 + 1:])) if d0 + 1 < nd else 1
    yC, yS = int(np.prod(ys[:d0])) if d0 else 1, int(np.prod(ys[d0 + 1:])) if d0 + 1 < nd else 1
    if not ((yC in (1, C_)) and (yS in (1, S_)) and (yC == 1 or tuple(ys[:d0]) == tuple(xs[:d0]))
            and (yS == 1 or tuple(ys[d0 + 1:]) == tuple(xs[d0 + 1:]))):
        raise AssertionError("A and B must have compatible dimensions")
    Lx, Ly = xs[d0], ys[d0]
    if shape == "full":
        Lz, lags = Lx + Ly - 1, np.arange(-(Ly - 1), Lx)
    elif shape == "same":
        Lz, lags = Lx, np.arange(0, Lx) - (Ly - 1) // 2
    else:
        Lz, lags = max(Lx - Ly + 1, 0), np.arange(0, max(Lx - Ly + 1, 0))
    dev = torch.device("cuda", torch.cuda.current_device())
    dX, dY = _colmajor(xt.reshape(xs), dt, dev), _colmajor(yt.reshape(ys), dt, dev)
    osz = tuple(xs[:d0]) + (Lz,) + tuple(xs[d0 + 1:])
    z = torch.zeros(int(np.prod(osz)), dtype=dt, device=dev)
    if _half:  # convh / convch (src/convd.cu:141,153): half storage; the mirror narrows / widens around the call
        if dbl: raise QupsError(-3, "half convolution of double data")
        nar = lambda t: (torch.view_as_real(t).to(torch.float16).contiguous() if t.is_complex() else t.to(torch.float16).contiguous())
        dX, dY = nar(dX), nar(dY)
        z = torch.zeros((z.numel(), 2) if cplx else (z.numel(),), dtype=torch.float16, device=dev)
    p = _lib.ConvdParams()
    p.struct_size = C.sizeof(_lib.ConvdParams)
    p.dtype, p.is_complex = (_lib.F16 if _half else (_lib.F64 if dbl else _lib.F32)), int(cplx)
    p.shape = {"full": 0, "same": 1, "valid": 2}[shape]
    p.C, p.S, p.Lx, p.Ly, p.yC, p.yS = C_, S_, Lx, Ly, yC, yS
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qups_convd(C.byref(p), _ptr(z), _ptr(dX), _ptr(dY), _stream(dev)))
    if _half:
        z = torch.view_as_complex(z.float()) if cplx else z.float()
    out = _from_colmajor(z, osz)
    lag_shape = [1] * nd
    lag_shape[d0] = -1
    lags = lags.reshape(lag_shape)
    return (np.asfortranarray(out.cpu().numpy()) if numpy_out else out), lags


# ---- aperture-domain post-processing (kern/cohfac.m, kern/dmas.m, kern/pcf.m, kern/slsc.m) ---------------------------
def _aperture(op, b, dim, lags=(), gamma=1.0, two=False):
    bt = _as_tensor(b)
    numpy_out = not (isinstance(b, torch.Tensor) and b.is_cuda)
    dev = torch.device("cuda", torch.cuda.current_device())
    dbl = bt.dtype in (torch.float64, torch.complex128)
    ct, rt = (torch.complex128, torch.float64) if dbl else (torch.complex64, torch.float32)
    nd = max(bt.ndim, dim)
    sz = _sz(bt, nd)
    d0 = dim - 1
    Cn, A, Sn = int(np.prod(sz[:d0])) if d0 else 1, sz[d0], int(np.prod(sz[d0 + 1:])) if d0 + 1 < nd else 1
    dB = _colmajor(bt.reshape(sz).to(ct), ct, dev)
    cplx_out = op in (_lib.APD_DMAS, _lib.APD_SLSC_AVERAGE, _lib.APD_SLSC_ENSEMBLE)
    out = torch.empty(Cn * Sn, dtype=ct if cplx_out else rt, device=dev)
    out2 = torch.empty(Cn * Sn, dtype=rt, device=dev) if two else None
    p = _lib.ApertureParams()
    p.struct_size = C.sizeof(_lib.ApertureParams)
    p.dtype, p.op, p.nlags = (_lib.F64 if dbl else _lib.F32), op, len(lags)
    p.C, p.A, p.S, p.gamma = Cn, A, Sn, float(gamma)
    lg = (C.c_uint32 * max(1, len(lags)))(*[int(v) for v in lags])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qups_aperture(C.byref(p), _ptr(out), _ptr(out2), _ptr(dB), lg, _stream(dev)))
    osz = tuple(sz[:d0]) + (1,) + tuple(sz[d0 + 1:])
    cv = lambda t: (np.asfortranarray(_from_colmajor(t, osz).cpu().numpy()) if numpy_out else _from_colmajor(t, osz))
    return (cv(out), cv(out2)) if two else cv(out)


def _last_nonsingleton(x):
    sh = tuple(np.shape(x)) if not isinstance(x, torch.Tensor) else tuple(x.shape)
    nz = [d + 1 for d, s in enumerate(sh) if s != 1]
    return nz[-1] if nz else 1


def cohfac(b, dim=None):
    """r = |sum(b,dim)|^2 ./ sum(|b|^2,dim) / size(b,dim) — mirror of ``kern/cohfac.m``."""
    return _aperture(_lib.APD_COHFAC, b, _last_nonsingleton(b) if dim is None else int(dim))


def dmas(bn, dim=None, L=None):
    """Delay-multiply-and-sum across the aperture — mirror of ``kern/dmas.m`` (scalar L -> lags 1:L)."""
    dim = _last_nonsingleton(bn) if dim is None else int(dim)
    N = (tuple(bn.shape) + (1,) * dim)[dim - 1]
    lags = range(1, N) if L is None else (range(1, int(L) + 1) if np.ndim(L) == 0 else [int(v) for v in np.ravel(L)])
    return _aperture(_lib.APD_DMAS, bn, dim, [v for v in lags if 1 <= v < N])


def pcf(b, dim=None, gamma=1.0):
    """Phase coherence factor [w, sf] — mirror of ``kern/pcf.m`` (auxiliary unwrap)."""
    bt = _as_tensor(b)
    if not bt.is_complex():
        raise ValueError("Input must be complex.")  # QUPS:pcf:realInput
    return _aperture(_lib.APD_PCF, b, max(1, _last_nonsingleton(b)) if dim is None else int(dim), gamma=gamma, two=True)


def slsc(x, dim=None, L=None, method="average"):
    """Short-lag spatial coherence — mirror of ``kern/slsc.m`` with a singleton time-sample dimension (kdim)."""
    if method not in ("average", "ensemble"):
        raise ValueError("method must be 'average' or 'ensemble'")
    dim = _last_nonsingleton(x) if dim is None else int(dim)
    A = (tuple(x.shape) + (1,) * dim)[dim - 1]
    L = max(1, A // 4) if L is None else L
    lags = list(range(1, int(L) + 1)) if np.ndim(L) == 0 else [int(v) for v in np.ravel(L)]
    return _aperture(_lib.APD_SLSC_AVERAGE if method == "average" else _lib.APD_SLSC_ENSEMBLE, x, dim, lags)


# ---- pair-wise windowed zero-normalised cross-correlation (kern/pwznxcorr.m) ------------------------------------------
def pwznxcorr(x, lags, W=None, U=1, *, pad=True, zero=True, norm=True, ref="neighbor", stride=1, x0=None, tdim=1, ndim=2,
              ldim=None, multi=False):
    """y = pwznxcorr(x, lags, W, U, ...) — mirror of ``kern/pwznxcorr.m:1`` (argument block :113-131, native branch).

    x: N-D data, time along `tdim`, channels along `ndim` (1-based).  lags: integer lags (a scalar L means -L:L, :142-143);
    W: window length (ones(W), unscaled, :147-150) or a weight vector along time; default max(ceil(max|lags|/2), 1).
    Returns the correlation with the channel dimension N - stride ('neighbor') or N ('center' / 'x0') long and the lags
    along dimension `ldim` (default ndims(x) + 1).  Fractional lags, U > 1 and multi = true are the reference's
    interpd / resample / convd branches and are not on this path (rejected).
    """
    if U != 1: raise _lib.QupsError(-3, "pwznxcorr: upsampling (U > 1) is not supported")
    if multi: raise _lib.QupsError(-3, "pwznxcorr: multi = true is not supported")
    if ref not in ("neighbor", "center", "x0"): raise ValueError("ref must be one of {'neighbor', 'center', 'x0'}")
    lg = np.atleast_1d(np.asarray(lags, dtype=np.float64)).ravel()
    if not np.all(np.isfinite(lg)): raise ValueError("lags must be finite")
    if not np.all(lg == np.floor(lg)): raise _lib.QupsError(-3, "pwznxcorr: fractional lags are not supported")
    if lg.size == 1: lg = np.arange(-lg[0], lg[0] + 1)
    lg = lg.astype(np.int64)
    if W is None: W = max(int(np.ceil(np.max(np.abs(lg)) / 2)), 1)
    xt = _as_tensor(x)
    numpy_out = not (isinstance(x, torch.Tensor) and x.is_cuda)
    dev = torch.device("cuda", torch.cuda.current_device())
    dbl = xt.dtype in (torch.float64, torch.complex128)
    ct, rt = (torch.complex128, torch.float64) if dbl else (torch.complex64, torch.float32)
    if np.ndim(W) == 0:
        wv = torch.ones(int(W), dtype=rt)
    else:
        Wa = np.asarray(W)
        wsz = list(Wa.shape) + [1] * max(0, max(tdim, ndim) - Wa.ndim)
        if any(s != 1 for d, s in enumerate(wsz) if d != tdim - 1 and d != ndim - 1):
            raise ValueError("The filter weights w must be scalar in all dimensions except time (%d) and channel (%d)." % (tdim, ndim))  # QUPS:pwznxcorr:incompatibleWeightSize
        if wsz[ndim - 1] != 1 and Wa.ndim >= ndim: raise _lib.QupsError(-3, "pwznxcorr: channel-dimension weights (multi) are not supported")
        wv = torch.as_tensor(Wa.reshape(-1), dtype=rt)
    nd = max(xt.ndim, tdim, ndim)
    xs = _sz(xt, nd)
    xv = xt.reshape(xs)
    order = [tdim - 1, ndim - 1] + [d for d in range(nd) if d not in (tdim - 1, ndim - 1)]
    xp = xv.permute(order)
    T, N = xp.shape[0], xp.shape[1]
    rest = tuple(xp.shape[2:])
    F = int(np.prod(rest)) if rest else 1
    cplx = xt.is_complex()
    dX = _colmajor(xp.reshape(T, N, F).to(ct if cplx else rt), ct if cplx else rt, dev)
    dX0, x0N, x0F = None, 1, 1
    if ref == "x0":
        if x0 is None: raise ValueError("ref = 'x0' needs the reference data x0")
        x0t = _as_tensor(x0)
        x0s = _sz(x0t, nd)
        x0p = x0t.reshape(x0s).permute(order)
        if x0p.shape[0] != T: raise _lib.QupsError(-3, "pwznxcorr: x0 must have the time extent of x")
        x0N, x0F = x0p.shape[1], int(np.prod(x0p.shape[2:])) if nd > 2 else 1
        if x0N not in (1, N) or x0F not in (1, F): raise AssertionError("x and x0 must have compatible dimensions")
        dX0 = _colmajor(x0p.reshape(T, x0N, x0F).to(ct), ct, dev)
    Nout = N - int(stride) if ref == "neighbor" else N
    L = int(lg.size)
    y = torch.empty(T * Nout * F * L, dtype=ct, device=dev)
    p = _lib.XcorrParams()
    p.struct_size = C.sizeof(_lib.XcorrParams)
    p.dtype, p.is_complex = (_lib.F64 if dbl else _lib.F32), int(cplx)
    p.ref = {"neighbor": _lib.XC_NEIGHBOR, "center": _lib.XC_CENTER, "x0": _lib.XC_X0}[ref]
    p.zero, p.norm, p.pad, p.stride = int(bool(zero)), int(bool(norm)), int(bool(pad)), int(stride)
    p.L, p.W, p.T, p.N, p.F, p.x0N, p.x0F = L, int(wv.numel()), T, N, F, x0N, x0F
    dW = wv.to(dev).contiguous()
    lgc = (C.c_int32 * L)(*[int(v) for v in lg])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qups_pwznxcorr(C.byref(p), _ptr(y), _ptr(dX), _ptr(dX0), _ptr(dW), lgc, _stream(dev)))
    out = _from_colmajor(y, (T, Nout, F, L)).reshape((T, Nout) + rest + (L,))   # frames were flattened row-major above
    # back to the caller's dimension order, lags along ldim
    inv = [0] * nd
    for i, d in enumerate(order): inv[d] = i
    out = out.permute(inv + [nd])                                   # x's dims, then lags
    ld = (nd + 1) if ldim is None else int(ldim)
    if ld <= nd:
        if out.shape[ld - 1] != 1: raise AssertionError("the lag dimension must be a singleton dimension of x")
        out = out.squeeze(ld - 1).movedim(-1, ld - 1)
    elif ld > nd + 1:
        out = out.reshape(tuple(out.shape[:-1]) + (1,) * (ld - nd - 1) + (L,))
    if not cplx and not norm: out = out.real
    return np.asfortranarray(out.cpu().numpy()) if numpy_out else out
