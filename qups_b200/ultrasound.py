"""ultrasound.py — thin mirror of the reference's L4 boundary for the hot path.

    UltrasoundSystem.DAS      src/UltrasoundSystem.m:3172-3372   -> das_spec
    UltrasoundSystem.bfDAS    src/UltrasoundSystem.m:4334-4474   -> bfDASLUT -> ChannelData.sample2sep -> wsinterpd2
    UltrasoundSystem.greens   src/UltrasoundSystem.m:463-882     -> greens kernel (+ focusTx identity for FSA)
    ChannelData               src/ChannelData.m:36-61            (data T x N x M x F, t0, fs)

Only the argument assembly the reference performs at these call sites is restated (geometry providers are the
minimal ones in synth.py); all arithmetic runs in libqups_b200.so.  MATLAB scripts keep using the reference's
own classes — this module exists so that the parity tests and bench can drive the same boundary from Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, kern, synth
from ._lib import GreensParams


@dataclass
class ChannelData:
    """T x N x M (x F) data cube with start time t0 (scalar or one per transmit) and sampling frequency fs."""
    data: object
    t0: object = 0.0
    fs: float = 1.0

    @property
    def T(self): return self.data.shape[0]
    @property
    def N(self): return self.data.shape[1]
    @property
    def M(self): return self.data.shape[2]

    # ---- pre-processing on the device: mirrors of the ChannelData methods scripts call before DAS ----
    # Every method is ONE pass of qups_chd_prep over the cube; prep() fuses the whole chain into a single pass.
    def prep(self, B=0, A=0, hilbert=False, fmix=0.0, out="single"):
        """zeropad(B, A) -> hilbert -> downmix(fmix) -> singleT/halfT in one kernel (1 read + 1 write of the cube).
        Returns a ChannelData whose data is a CUDA tensor (complex64, or (.., 2) float16 for out='halfT')."""
        x = self.data if isinstance(self.data, torch.Tensor) else torch.from_numpy(np.asarray(self.data))
        dev = torch.device("cuda", torch.cuda.current_device())
        shp = tuple(x.shape) + (1,) * (3 - x.ndim)
        T, K = shp[0], int(np.prod(shp[1:]))
        kinds = {torch.float32: _lib.IN_REAL_F32, torch.complex64: _lib.IN_CPLX_F32, torch.int16: _lib.IN_REAL_I16,
                 torch.float64: _lib.IN_REAL_F64}
        if x.dtype == torch.complex128: x = x.to(torch.complex64)
        if x.dtype not in kinds:
            raise _lib.QupsError(-3, f"ChannelData.prep: unsupported data type {x.dtype}")
        xin = kern._colmajor(x.reshape(shp), x.dtype, dev)
        t0 = torch.as_tensor(np.asarray(self.t0, np.float64).reshape(-1)).to(device=dev, dtype=torch.float32)
        if t0.numel() not in (1, shp[2]):
            raise AssertionError("t0 must be a scalar or have one entry per transmit")
        p = _lib.PrepParams()
        p.struct_size = C.sizeof(_lib.PrepParams)
        p.in_dtype, p.out_dtype, p.hilbert = kinds[x.dtype], (_lib.F16 if out == "halfT" else _lib.F32), int(bool(hilbert))
        p.T, p.K, p.B, p.A = T, K, int(B), int(A)
        p.traces_per_t0, p.n_t0 = shp[1], t0.numel()
        p.fs, p.fmix = float(self.fs), float(fmix)
        L = int(B) + T + int(A)
        if out == "halfT":
            yb = torch.empty((K * L, 2), dtype=torch.float16, device=dev)
        else:
            yb = torch.empty(K * L, dtype=torch.complex64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().qups_chd_prep(C.byref(p), kern._ptr(yb), kern._ptr(xin), kern._ptr(t0), kern._stream(dev)))
        osz = (L,) + shp[1:]
        # the C ABI writes half2; torch has no usable complex-half type, so the mirror widens the (fp16-rounded) values
        y = kern._from_colmajor(yb, osz) if out != "halfT" else kern._from_colmajor(torch.view_as_complex(yb.float()), osz)
        t0n = np.asarray(self.t0, np.float64) - int(B) / float(self.fs)
        return ChannelData(y, t0n if t0n.ndim else float(t0n), self.fs)

    def zeropad(self, B=0, A=0):
        """src/ChannelData.m:1153-1183."""
        if B < 0 or A < 0: raise AssertionError("Data append or prepend size must be positive.")
        return self.prep(B=B, A=A)

    def hilbert(self, N=None):
        """src/ChannelData.m:935-966 (N > T zero-pads the transform, as MATLAB's hilbert(x, N))."""
        N = self.T if N is None else int(N)
        if N < self.T: raise _lib.QupsError(-3, "hilbert with N < T (truncation) is not implemented")
        out = self.prep(A=N - self.T, hilbert=True)
        return out

    def downmix(self, fc):
        """src/ChannelData.m:757-807."""
        return self.prep(fmix=float(fc))

    def singleT(self): return self.prep(out="single")
    def halfT(self): return self.prep(out="halfT")

    # ---- sampling (src/ChannelData.m:1230-1447): delays in seconds -> wsinterpd / wsinterpd2 ------------------------------
    def _lifted(self, apdim):
        """Data and t0 with the receive / transmit dimensions moved to apdim (1-based), frames behind max(apdim)
        (swapdimD(chd, [ndim, mdim, 4:D], [apdim, max(apdim) + (1:D-3)]), src/ChannelData.m:1304-1305)."""
        x = self.data
        is_t = isinstance(x, torch.Tensor)
        shp = tuple(x.shape) + (1,) * (3 - x.ndim)
        D = len(shp)
        nd = max(apdim) + (D - 3)
        new = [1] * nd
        new[0], new[apdim[0] - 1], new[apdim[1] - 1] = shp[0], shp[1], shp[2]
        for k in range(3, D):
            new[max(apdim) + (k - 3)] = shp[k]
        # T, N, M, F.. keep their relative memory order only if apdim is increasing; build by explicit permutation
        src_of = {0: 0, apdim[0] - 1: 1, apdim[1] - 1: 2}
        for k in range(3, D):
            src_of[max(apdim) + (k - 3)] = k
        xr = x.reshape(shp) if is_t else np.asarray(x).reshape(shp)
        extra = nd - D
        xr = xr.reshape(shp + (1,) * extra)
        free = iter(range(D, nd))
        perm = [src_of[d] if d in src_of else next(free) for d in range(nd)]
        xl = xr.permute(*perm) if is_t else np.transpose(xr, perm)
        t0 = np.asarray(self.t0, np.float64).reshape(-1)
        t0s = [1] * nd
        if t0.size > 1:
            t0s[apdim[1] - 1] = t0.size
        return xl, t0.reshape(t0s)

    def sample(self, tau, interp="linear", w=1, sdim=(), fmod=0.0, apdim=(2, 3)):
        """y = sample(chd, tau, interp, w, sdim, fmod, apdim): ntau = (tau - t0) .* fs; wsinterpd(data, ntau, 1, w, sdim,
        interp, 0, 2i*pi*fmod/fs)   (src/ChannelData.m:1230-1336).  tau: time along dim 1, singleton-or-matching elsewhere."""
        xl, t0 = self._lifted(tuple(apdim))
        tau = np.asarray(tau.cpu() if isinstance(tau, torch.Tensor) else tau)
        nd = max(tau.ndim, xl.ndim)
        ts, xs = tuple(tau.shape) + (1,) * (nd - tau.ndim), tuple(xl.shape) + (1,) * (nd - xl.ndim)
        for d in range(1, nd):
            if not (ts[d] == xs[d] or ts[d] == 1 or xs[d] == 1):
                raise AssertionError(f"Delay size must match the data size ({xs[d]}) or be singleton in dimension {d + 1}.")
        rt = np.float32 if (xl.dtype in (torch.complex64, torch.float32) if isinstance(xl, torch.Tensor) else xl.dtype in (np.complex64, np.float32)) else np.float64
        ntau = ((tau.reshape(ts) - t0.reshape(tuple(t0.shape) + (1,) * (nd - t0.ndim))) * float(self.fs)).astype(rt)
        return kern.wsinterpd(xl, ntau, 1, w, sdim, interp, 0, 2j * np.pi * fmod / float(self.fs))

    def sample2sep(self, tau1, tau2, interp="linear", w=1, sdim=(), fmod=0.0, apdim=(2, 3)):
        """Separable delays tau = tau1 + tau2 (src/ChannelData.m:1338-1447): t0 is subtracted from the table that gives the
        smaller broadcast, then wsinterpd2(data, ntau1, ntau2, 1, w, sdim, interp, 0, 2i*pi*fmod/fs)."""
        xl, t0 = self._lifted(tuple(apdim))
        t1, t2 = (np.asarray(t.cpu() if isinstance(t, torch.Tensor) else t) for t in (tau1, tau2))
        nd = max(t1.ndim, t2.ndim, xl.ndim, t0.ndim)
        pad = lambda a: a.reshape(tuple(a.shape) + (1,) * (nd - a.ndim))
        t1, t2, t0 = pad(t1), pad(t2), pad(t0)
        xs = tuple(xl.shape) + (1,) * (nd - xl.ndim)
        for tt in (t1, t2):
            for d in range(1, nd):
                if not (tt.shape[d] == xs[d] or tt.shape[d] == 1 or xs[d] == 1):
                    raise AssertionError(f"Delay size must match the data size ({xs[d]}) or be singleton in dimension {d + 1}.")
        rt = np.float32 if (xl.dtype in (torch.complex64, torch.float32) if isinstance(xl, torch.Tensor) else xl.dtype in (np.complex64, np.float32)) else np.float64
        fs = float(self.fs)
        bsz = lambda a: int(np.prod(np.maximum(a.shape, t0.shape)))
        if bsz(t1) < bsz(t2):
            n1, n2 = ((t1 - t0) * fs).astype(rt), (t2 * fs).astype(rt)
        else:
            n1, n2 = (t1 * fs).astype(rt), ((t2 - t0) * fs).astype(rt)
        return kern.wsinterpd2(xl, n1, n2, 1, w, sdim, interp, 0, 2j * np.pi * fmod / fs)

    def rectifyt0(self, interp="cubic", t0_=None):
        """Collapse t0 to a scalar by resampling every transmit onto one time axis (src/ChannelData.m:1205-1228)."""
        t0 = np.asarray(self.t0, np.float64).reshape(-1)
        if t0.size <= 1:
            return ChannelData(self.data, float(t0[0]) if t0.size else 0.0, self.fs)
        t0_ = float(t0.min()) if t0_ is None else float(t0_)
        npad = int(np.ceil((t0 - t0_).max() * float(self.fs)))
        T = self.T + npad
        tau = (t0_ + np.arange(T) / float(self.fs)).reshape(-1, 1, 1)
        y = self.sample(tau, interp)        # zeropad(chd, 0, npad) is implicit: samples past the end are 0 (extrapval)
        return ChannelData(y, t0_, self.fs)


@dataclass
class Sequence:
    """type in {'FSA','PW','FC','VS','DV'}; focus = 3 x M foci (FC/VS/DV) or unit normals (PW)."""
    type: str = "FSA"
    focus: Optional[np.ndarray] = None
    c0: float = 1540.0
    apd: Optional[np.ndarray] = None   # elements x pulses apodization matrix (Sequence.apodization_, e.g. hadamard(M)); None = default


@dataclass
class UltrasoundSystem:
    tx: np.ndarray                     # 3 x M transmit element positions
    rx: np.ndarray                     # 3 x N receive element positions
    seq: Sequence
    scan: np.ndarray                   # 3 x I1 x I2 x I3 pixel positions (Scan.positions())
    fs: float
    fc: float = 5e6
    bw_frac: float = 0.6
    tx_offset: np.ndarray = field(default_factory=lambda: np.zeros((3, 1)))
    tx_normal: np.ndarray = field(default_factory=lambda: np.array([[0.0], [0.0], [1.0]]))
    rx_normal: Optional[np.ndarray] = None   # 3 x N element normals (3rd output of Transducer.orientations); default +z
    rx_angle: Optional[np.ndarray] = None    # N element azimuth angles in degrees (1st output of orientations); default 0
    # ScanPolar-style lateral coordinates for apScanline / apTranslatingAperture (src/UltrasoundSystem.m:4951-4953, 5103-5109):
    # pixel angle per index along scan_lat_dim (scan.a), transmit angles (seq.angles); None -> Cartesian x coordinates
    scan_lat: Optional[np.ndarray] = None
    scan_lat_dim: int = 2
    tx_lat: Optional[np.ndarray] = None

    # ---- apodization generators (src/UltrasoundSystem.m:4892-5429) --------------------------------
    # Each returns a kern.FusedApod: pass it to DAS as an apodization argument (evaluated inside the kernel), or call
    # .dense(...) for the ND array the reference returns.  ScanCartesian geometry: the lateral pixel coordinate is x.
    def _rx_normals(self):
        n = self.rx.shape[1]
        return np.broadcast_to(np.array([[0.0], [0.0], [1.0]]), (3, n)) if self.rx_normal is None else np.asarray(self.rx_normal, np.float64)

    def _tx_lateral(self):
        return np.asarray(self.tx_lat, np.float64).reshape(-1) if self.tx_lat is not None else np.asarray(self.seq.focus, np.float64)[0]

    def apAcceptanceAngle(self, theta=45.0):
        """apod = (n . (Pi - Pn)/|Pi - Pn|) >= cosd(theta)   (:5303-5375)."""
        if not theta > 0: raise ValueError("theta must be positive")
        return kern.FusedApod(rx_kind=_lib.AP_RX_ACCEPTANCE_ANGLE, rx_p=(np.float32(np.cos(np.deg2rad(float(theta)))),),
                              rx_aux=self._rx_normals(), name="apAcceptanceAngle")

    def apCosineAngle(self, theta=45.0):
        """apod = cosd(min(90, (90/theta) acosd(n . r)))   (:5377-5429)."""
        if not theta > 0: raise ValueError("theta must be positive")
        return kern.FusedApod(rx_kind=_lib.AP_RX_COSINE_ANGLE, rx_p=(np.float32(90.0 / float(theta)),), rx_aux=self._rx_normals(),
                              name="apCosineAngle")

    def apApertureGrowth(self, f=1.5, Dmax=np.inf):
        """apod = (z > f |2d|) & (|2d| < Dmax), d/z in the element's frame for non-planar arrays   (:5165-5267)."""
        if not (f > 0 and Dmax > 0): raise ValueError("f and Dmax must be positive")
        ae = None if self.rx_angle is None else np.asarray(self.rx_angle, np.float64).reshape(-1)
        nonplanar = ae is not None and bool(np.any(ae != 0))
        aux = np.stack([np.cos(np.deg2rad(ae)), np.sin(np.deg2rad(ae))]) if nonplanar else None
        return kern.FusedApod(rx_kind=_lib.AP_RX_APERTURE_GROWTH, rx_p=(f, Dmax, float(nonplanar)), rx_aux=aux, name="apApertureGrowth")

    def apScanline(self, tol):
        """apod = |x_i - x_focus(m)| < tol   (:4892-4968)."""
        if not tol > 0: raise ValueError("tol must be positive")
        return kern.FusedApod(tx_kind=_lib.AP_TX_SCANLINE, tx_p=(np.float32(tol),), tx_aux=self._tx_lateral(), lat=self.scan_lat,
                              lat_dim=self.scan_lat_dim, name="apScanline")

    def apTranslatingAperture(self, tol):
        """apod = |x_i - x_focus(m)| <= tol(1) & |x_i - x_n| <= tol(end)   (:5074-5163)."""
        tol = np.atleast_1d(np.asarray(tol, np.float64))
        if not np.all(tol > 0): raise ValueError("tol must be positive")
        polar = self.scan_lat is not None  # ScanPolar: receivers are compared by their orientation angle (:5108)
        rx_lat = np.asarray(self.rx_angle, np.float64).reshape(-1) if polar and self.rx_angle is not None else np.asarray(self.rx, np.float64)[0]
        return kern.FusedApod(rx_kind=_lib.AP_RX_TRANSLATING, rx_p=(np.float32(tol[-1]),), rx_aux=rx_lat,
                              tx_kind=_lib.AP_TX_TRANSLATING, tx_p=(np.float32(tol[0]),), tx_aux=self._tx_lateral(),
                              lat=self.scan_lat, lat_dim=self.scan_lat_dim, name="apTranslatingAperture")

    def apTxParallelogram(self, theta=None, phi=0.0, bounds=None):
        """Pixels whose projection along the (tilted) transmit direction lands on the aperture   (:5269-5301)."""
        fo = np.asarray(self.seq.focus, np.float64)
        theta = np.rad2deg(np.arctan2(fo[0], fo[2])) if theta is None else np.asarray(theta, np.float64).reshape(-1)
        phi = np.atleast_1d(np.asarray(phi, np.float64))
        if bounds is None:  # us.xdc.bounds(): lateral extent of the element positions
            bounds = (float(self.rx[0].min()), float(self.rx[0].max()))
        a1, a2 = np.deg2rad(phi[0] + theta), np.deg2rad(phi[-1] + theta)
        aux = np.stack([np.sin(a1), np.cos(a1), np.sin(a2), np.cos(a2)])
        return kern.FusedApod(tx_kind=_lib.AP_TX_PARALLELOGRAM, tx_p=(np.float32(bounds[0]), np.float32(bounds[1])), tx_aux=aux, name="apTxParallelogram")

    def apMultiline(self):
        """Linear weights between the two transmits straddling each scan line (:4970-5072): a small 1 x I2 x 1 x 1 x M
        matrix, host logic exactly as the reference (find last-left / first-right transmit); passed to DAS as an array."""
        x = np.asarray(self.scan, np.float64)[0, 0, :, 0] if self.scan.ndim == 4 else np.asarray(self.scan, np.float64)[0, 0, :]
        xv = np.asarray(self.seq.focus, np.float64)[0]
        A = np.zeros((x.size, xv.size))
        for i, xi in enumerate(x):
            da = xi - xv
            l, r = np.nonzero(da >= 0)[0], np.nonzero(da <= 0)[0]
            if l.size == 0 or r.size == 0:
                continue
            li, ri = l[-1], r[0]
            dlr = abs(xv[li] - xv[ri])
            al, ar = (1.0, 0.0) if dlr == 0 else (1 - abs(xv[li] - xi) / dlr, 1 - abs(xv[ri] - xi) / dlr)
            A[i, li] += al
            A[i, ri] += ar
        return A.reshape(1, x.size, 1, 1, xv.size)

    # ---- DAS ------------------------------------------------------------------------------
    def _pos_args(self):
        """src/UltrasoundSystem.m:3341-3351 (DAS) / :4436-4440 (bfDAS)."""
        t = self.seq.type
        if t == "FSA":
            return self.tx, np.broadcast_to(self.tx_normal, self.tx.shape), ("diverging-waves",)
        if t == "PW":
            return np.zeros((3, 1)), np.asarray(self.seq.focus), ("plane-waves",)
        if t in ("VS", "FC", "DV"):
            nf = np.asarray(self.seq.focus) - self.tx_offset
            nv = nf / np.linalg.norm(nf, 2)  # matrix 2-norm, as the reference (harmless: only the sign is used)
            return np.asarray(self.seq.focus), nv, (("diverging-waves",) if t == "DV" else ())
        raise ValueError(f"unknown sequence type {t}")

    def DAS(self, chd: ChannelData, *apod, c0=None, fmod=0.0, interp="cubic", keep_tx=False, keep_rx=False,
            prec="single", **kw):
        """b = DAS(us, chd, A1..An, 'c0', 'fmod', 'prec', 'interp', 'keep_tx', 'keep_rx')  (:3172-3295)."""
        fun = {(False, False): "DAS", (True, False): "SYN", (False, True): "MUL", (True, True): "BF"}[(keep_rx, keep_tx)]
        Pv, Nv, ext = self._pos_args()
        rt = np.float64 if prec == "double" else np.float32
        opts = list(ext) + ["interp", interp, "input-precision", prec]
        for a in apod:
            opts += ["apod", a]
        if fmod:
            opts += ["modulation", float(fmod)]
        c = self.seq.c0 if c0 is None else c0
        f = lambda v: np.asarray(v, dtype=rt)
        b = kern.das_spec(fun, f(self.scan), f(self.rx), f(Pv), f(Nv), chd.data, chd.t0, chd.fs, c, *opts, **kw)
        # permute(b, [1:3, 6:D+2, 4:5]) -> I1 x I2 x I3 x F... x [N] x [M]   (:3361)
        nd = b.ndim
        order = [0, 1, 2] + list(range(5, nd)) + [3, 4]
        return b.permute(*order) if isinstance(b, torch.Tensor) else np.transpose(b, order)

    # ---- bfDAS (separable delay tables) -------------------------------------------------------
    def bfDAS(self, chd: ChannelData, c0=None, fmod=0.0, interp="cubic", apod=1, keep_tx=False, keep_rx=False):
        """tau_rx = |Pi-Pr|/c0 (I x N), tau_tx = dv/c0 (I x 1 x M) -> sample2sep -> wsinterpd2  (:4431-4473)."""
        c = self.seq.c0 if c0 is None else c0
        Pv, Nv, ext = self._pos_args()
        Isz = tuple(self.scan.shape[1:]) + (1,) * (4 - self.scan.ndim)
        Pi = np.asarray(self.scan, np.float64).reshape(3, -1, order="F")
        Pr = np.asarray(self.rx, np.float64)
        Pv = np.broadcast_to(np.asarray(Pv, np.float64), (3, chd.M))
        rv = Pi[:, :, None] - Pv[:, None, :]
        t = self.seq.type
        if t in ("DV", "FSA"):
            dv = np.linalg.norm(rv, axis=0)
        elif t == "PW":
            dv = (rv * np.asarray(Nv, np.float64)[:, None, :]).sum(0)
        else:  # VS / FC: normalize(focus - offset, 1, "norm") per column (:4439)
            nf = np.asarray(self.seq.focus, np.float64) - self.tx_offset
            nf = nf / np.linalg.norm(nf, axis=0, keepdims=True)
            dv = np.linalg.norm(rv, axis=0) * np.sign((rv * nf[:, None, :]).sum(0))
        dr = np.linalg.norm(Pi[:, :, None] - Pr[:, None, :], axis=0)
        rt = np.float32 if np.asarray(chd.data).dtype in (np.complex64, np.float32) else np.float64
        tau_rx = (dr / c).astype(rt).reshape(Isz + (chd.N, 1), order="F")
        tau_tx = (dv / c).astype(rt).reshape(Isz + (1, chd.M), order="F")
        return self.bfDASLUT(chd, tau_rx, tau_tx, apod, fmod=fmod, interp=interp, keep_tx=keep_tx, keep_rx=keep_rx)

    def bfDASLUT(self, chd: ChannelData, tau_rx, tau_tx=None, *apod, fmod=0.0, interp="cubic", keep_tx=False, keep_rx=False,
                 bsize=None):
        """b = bfDASLUT(us, chd, tau_rx, tau_tx, A1..An, 'fmod', 'interp', 'keep_tx', 'keep_rx', 'bsize')
        (src/UltrasoundSystem.m:4476-4673): look-up-table delay-and-sum.  Validates / reshapes the delay tables (same error
        identifiers), splices the transmits in blocks of bsize, reduces the apodization per block and hands each block to
        ChannelData.sample2sep -> wsinterpd2 (src/ChannelData.m:1338-1447, kern/wsinterpd2.m); blocks are summed (or
        concatenated when keep_tx).  Output I1 x I2 x I3 x [1|N] x [1|M]."""
        Isz = tuple(self.scan.shape[1:]) + (1,) * (4 - self.scan.ndim)
        nPix, N, M = int(np.prod(Isz)), chd.N, chd.M
        tau_rx = np.asarray(tau_rx)
        tau_tx = np.swapaxes(tau_rx.reshape(tau_rx.shape + (1,) * (5 - tau_rx.ndim)), 3, 4) if tau_tx is None else np.asarray(tau_tx)
        sz = lambda a, k: tuple(a.shape) + (1,) * (k - a.ndim)
        if sz(tau_rx, 4)[:4] != Isz + (N,) or tau_rx.ndim > 4 and any(v != 1 for v in tau_rx.shape[4:]):
            D = tau_rx.ndim
            I_, L_ = int(np.prod(tau_rx.shape[:D - 1])), tau_rx.shape[D - 1]
            if I_ == nPix and L_ == N:
                tau_rx = tau_rx.reshape(Isz + (N,), order="F")
            else:
                raise _lib.QupsError(-1, "QUPS:UltrasoundSystem:bfDASLUT:incompatibleReceiveDelayTable: Expected a table with "
                                     f"{nPix} pixels and {N} receives but instead there are {I_} pixels and {L_} receives.")
        if sz(tau_tx, 5)[:5] != Isz + (1, M):
            D = tau_tx.ndim
            I_, L_ = int(np.prod(tau_tx.shape[:D - 1])), tau_tx.shape[D - 1]
            if I_ == nPix and L_ == M:
                tau_tx = tau_tx.reshape(Isz + (1, M), order="F")
            else:
                raise _lib.QupsError(-1, "QUPS:UltrasoundSystem:bfDASLUT:incompatibleTransmitDelayTable: Expected a table with "
                                     f"{nPix} pixels and {M} transmits but instead there are {I_} pixels and {L_} transmits.")
        tau_rx = tau_rx.reshape(Isz + (N, 1), order="F")
        tau_tx = tau_tx.reshape(Isz + (1, M), order="F")
        rt = np.float32 if np.asarray(chd.data).dtype in (np.complex64, np.float32) else np.float64
        tau_rx, tau_tx = tau_rx.astype(rt), tau_tx.astype(rt)
        apods = [np.asarray(a) for a in (apod if apod else (1,))]
        apods = [a.reshape(tuple(a.shape) + (1,) * (5 - a.ndim), order="F") if a.ndim else a.reshape((1,) * 5) for a in apods]
        for a in apods:
            if not all(a.shape[d] in (1, (Isz + (N, M))[d]) for d in range(5)):
                raise AssertionError("Apodization data size inconsistent with the scan / receive / transmit sizes")
        t0 = np.asarray(chd.t0, np.float64).reshape(-1)
        x = chd.data
        sdim = tuple(d for d, keep in ((4, keep_rx), (5, keep_tx)) if not keep)   # dims to sum after apodization (:4630-4632)
        bsize = M if bsize is None else int(bsize)
        if bsize < 1: raise ValueError("bsize must be a positive integer")
        a0 = 1
        for a in apods:  # apodization common to all transmits is reduced once (:4643-4644)
            if a.shape[4] == 1: a0 = a0 * a
        out, acc = [], 0
        for m0 in range(0, M, bsize):  # [chds, im] = splice(chd, chd.mdim, bsize) (:4641)
            ms = slice(m0, min(M, m0 + bsize))
            a = a0
            for ap_ in apods:  # reduce apodization per tx block (:4646-4649)
                if ap_.shape[4] != 1: a = a * ap_[:, :, :, :, ms]
            chdm = ChannelData(x[:, :, ms], t0[ms] if t0.size > 1 else float(t0[0]) if t0.size else 0.0, chd.fs)
            # bim = sample2sep(chds(m), tau_txm, tau_rx, interp, a, sdim, fmod, [4, 5])   (:4651)
            y = chdm.sample2sep(tau_tx[:, :, :, :, ms], tau_rx, interp, np.asarray(a), sdim, fmod, (4, 5))
            if keep_tx: out.append(y)
            else: acc = acc + y
        if keep_tx:
            return np.concatenate(out, axis=4) if isinstance(out[0], np.ndarray) else torch.cat(out, dim=4)
        return acc

    # ---- greens ---------------------------------------------------------------------------------
    def greens(self, scat_pos, scat_amp, c0=None, interp="cubic", R0=None, fsk=None, device=None, sort=True):
        """chd = greens(us, scat)  (src/UltrasoundSystem.m:463-882), FSA simulation on the GPU.

        Returns ChannelData with data T x N x M truncated to its non-zero time support (:871-874).
        """
        c0 = self.seq.c0 if c0 is None else c0
        fs = float(self.fs)
        fsk = fs if fsk is None else float(fsk)
        R0 = (c0 / self.fc) if R0 is None else R0      # default max(us.lambda)  (:550)
        ps = np.asarray(scat_pos, np.float64).reshape(3, -1)
        amp = np.asarray(scat_amp, np.float64).reshape(-1)
        pv, pn = np.asarray(self.tx, np.float64), np.asarray(self.rx, np.float64)
        kern_s, wv_t0, wv_tend = synth.greens_kernel(self.fc, self.bw_frac, fsk)
        # time bounds from the aperture bounding boxes (:567-580, 608-615)
        def box(p):
            lo, hi = p.min(1), p.max(1)
            c = np.array([[(hi if (i >> d) & 1 else lo)[d] for i in range(8)] for d in range(3)])
            return c, np.linalg.norm(hi - lo)
        txb, txr = box(pv)
        rxb, rxr = box(pn)
        dist = lambda b: np.linalg.norm(ps[:, :, None] - b[:, None, :], axis=0)
        dtx, drx = dist(txb), dist(rxb)
        taumax = (dtx.max() + drx.max() + txr + rxr) / c0
        taumin = (dtx.min() + drx.min() - txr - rxr) / c0
        tmin = taumin + wv_t0 - (wv_tend - wv_t0)
        tmax = taumax + wv_tend
        n0, ne = int(np.floor(tmin * fs)), int(np.ceil(tmax * fs))
        T = ne - n0 + 1
        if sort:  # sort scatterers by rmin*rmax (:630-647)
            rs = np.linalg.norm(ps[:, :, None] - pv[:, None, :], axis=0)
            rmin, rmax = 2 * rs.min(1), 2 * rs.max(1)
            if not np.array_equal(pv, pn):
                rr = np.linalg.norm(ps[:, :, None] - pn[:, None, :], axis=0)
                rmin, rmax = rs.min(1) + rr.min(1), rs.max(1) + rr.max(1)
            order = np.argsort(rmin * rmax, kind="stable")
            ps, amp = ps[:, order], amp[order]
        x = greens_raw(ps, amp, pn, pv, kern_s, n0, T, fs, c0, wv_t0, fsk / fs, R0, interp, device=device)
        # truncate the all-zero head/tail (:871-874)
        nz = (x != 0).reshape(T, -1).any(dim=1)
        idx = torch.nonzero(nz).reshape(-1)
        if idx.numel():
            a, b = int(idx[0]), int(idx[-1])
            x, n0 = x[a:b + 1], n0 + a
        chd = ChannelData(x, n0 / fs, fs)
        # synthesize the requested sequence from the FSA data (:877); identity for a pure FSA sequence
        return focusTx(chd, self.seq, pv, interp=interp)


def seq_delays(seq: "Sequence", tx: np.ndarray) -> np.ndarray:
    """Sequence.delays(tx): M x S transmit delays (src/Sequence.m:888-925)."""
    p = np.asarray(tx, np.float64)                       # 3 x M elements
    M = p.shape[1]
    if seq.type == "FSA":
        return np.zeros((M, M))
    f = np.asarray(seq.focus, np.float64)                # 3 x S
    if seq.type == "PW":
        return -(f[:, None, :] * p[:, :, None]).sum(0) / seq.c0
    v = f[:, None, :] - p[:, :, None]                    # element -> focus
    tau = np.sqrt((v ** 2).sum(0)) / seq.c0
    if seq.type == "FC":
        sgn = 1.0
    elif seq.type == "DV":
        sgn = -1.0
    else:  # VS: negative when the focus is behind the transducer
        sgn = np.where((f[2][None, :] > p[2][:, None]).all(0), 1.0, -1.0)[None, :]
    return tau * sgn


def seq_apodization(seq: "Sequence", tx: np.ndarray) -> np.ndarray:
    """Sequence.apodization(tx) default (src/Sequence.m:951-960)."""
    M = np.asarray(tx).shape[1]
    if getattr(seq, "apd", None) is not None:
        return np.asarray(seq.apd)
    if seq.type == "FSA":
        return np.eye(M)
    return np.ones((M, np.asarray(seq.focus).shape[1]))


def focusTx(chd: "ChannelData", seq: "Sequence", tx: np.ndarray, interp="cubic", buffer=0, ws2=None) -> "ChannelData":
    """chd' = focusTx(us, chd, seq): synthesise the transmits of `seq` from full-synthetic-aperture data,
    chd'(t,n,m') = sum_m apd(m,m') * chd(t - tau(m,m'), n, m)    (src/UltrasoundSystem.m:3374-3503) via
    sample2sep over the transmit dimension (:3498 -> src/ChannelData.m:1338 -> wsinterpd2)."""
    ws2 = kern.wsinterpd2 if ws2 is None else ws2
    tau = -seq_delays(seq, tx)                            # M x M'   (:3448)
    apd = seq_apodization(seq, tx)
    if seq.type == "FSA" and not np.any(tau) and np.array_equal(apd, np.eye(apd.shape[0])):
        return chd                                        # identity (:3453-3455)
    fs = float(chd.fs)
    nz = (apd != 0) | np.zeros(tau.shape, bool)
    nmin = int(np.floor(tau[nz].min() * fs))
    nmax = int(np.ceil(tau[nz].max() * fs))
    t0 = np.asarray(chd.t0, np.float64) + nmin / fs       # (:3460)
    tau = tau - nmin / fs
    x = chd.data
    pad = (nmax - nmin) + int(buffer)                     # zeropad(chd, 0, ...)  (:3462)
    is_t = isinstance(x, torch.Tensor)
    if is_t:
        x = torch.cat([x, torch.zeros((pad,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)], 0) if pad else x
        rt = np.float32 if x.dtype == torch.complex64 else np.float64
    else:
        x = np.concatenate([x, np.zeros((pad,) + x.shape[1:], x.dtype)], 0) if pad else x
        rt = np.float32 if x.dtype == np.complex64 else np.float64
    Tn, N, M = x.shape[:3]
    # sample2sep(chd.time, -tau, interp, apd, mdim): ntau1 = (time - t0) fs = 0..T'-1 ; ntau2 = -tau fs
    t1 = np.arange(Tn, dtype=rt).reshape(Tn, 1, 1, 1)
    t2 = (-(tau * fs)).astype(rt).reshape(1, 1, M, -1)
    w = apd.astype(rt).reshape(1, 1, apd.shape[0], -1)
    x4 = x.reshape(tuple(x.shape[:3]) + (1,)) if is_t else x.reshape(x.shape[:3] + (1,), order="F")
    y = ws2(x4, t1, t2, 1, w, (3,), interp, 0, 0)         # T' x N x 1 x M'
    y = y[:, :, 0, :]
    return ChannelData(y, t0 if np.ndim(t0) else float(t0), fs)


def refocus(chd: "ChannelData", seq: "Sequence", tx: np.ndarray, method="tikhonov", gamma=None):
    r"""[chd, Hi] = refocus(us, chd, seq, 'method', method, 'gamma', gamma) — mirror of src/UltrasoundSystem.m:3505-3768.

    Host logic as in the reference (:3690-3727, evaluated in float64): encoding matrix H = a.' .* exp(-2j*pi*f.*tau.')
    per frequency, weights w = pagenorm(H,2)^-2, decoder Hi = (H'H + gamma w I) \ H.' | H.' w | w pinv(H), NaN -> 0.
    The data path (:3729-3757: fft, time-alignment phase, tensor-times-matrix over the transmit dimension, ifft) is ONE
    C-ABI call, qups_refocus.  Returns (ChannelData with t0 = min(t0), Hi as an E x V x T array)."""
    if method not in ("tikhonov", "adjoint", "pinv"):
        raise ValueError("method must be one of {'tikhonov', 'adjoint', 'pinv'}")
    x = chd.data
    xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
    if xt.ndim != 3: raise ValueError("refocus: data must be T x N x M")
    T, N, V = (int(v) for v in xt.shape)
    fs = float(chd.fs)
    tau = np.asarray(seq_delays(seq, tx), np.float64)                     # E x V  (:3690)
    a = seq_apodization(seq, tx) * np.ones(tau.shape)                     # E x V  (:3691)
    E = tau.shape[0]
    if tau.shape[1] != V: raise ValueError("refocus: the sequence has %d pulses, the data %d transmits" % (tau.shape[1], V))
    if gamma is None: gamma = 10.0 * (N / 10.0) ** 2                       # (:3683)
    f = np.arange(T) * fs / T                                              # chd.fftaxis (src/ChannelData.m:1491)
    H = a.T[None] * np.exp(-2j * np.pi * f[:, None, None] * tau.T[None])   # pages first: T x V x E  (:3697)
    with np.errstate(divide="ignore"):
        w = np.linalg.norm(H, 2, axis=(1, 2)) ** -2.0                      # pagenorm(H, 2).^-2  (:3702)
    if method == "tikhonov":
        if N != E: raise ValueError("Arrays have incompatible sizes for this operation.")  # gamma .* w .* eye(chd.N) + H'H (:3709-3710)
        A = np.conj(np.swapaxes(H, 1, 2)) @ H + (float(gamma) * w)[:, None, None] * np.eye(E)[None]
        Hi = np.zeros((T, E, V), np.complex128)
        for k in range(T):
            if np.all(np.isfinite(A[k])):
                try: Hi[k] = np.linalg.solve(A[k], H[k].T)                 # pagemldivide(A, pagetranspose(H))  (:3712)
                except np.linalg.LinAlgError: Hi[k] = np.nan
            else: Hi[k] = np.nan
    elif method == "adjoint":
        Hi = np.swapaxes(H, 1, 2) * w[:, None, None]                        # (:3715)
    else:
        Hi = w[:, None, None] * np.linalg.pinv(H)                          # (:3719)
    Hi = np.where(np.isnan(Hi), 0, Hi)                                      # (:3727)
    Hi = np.ascontiguousarray(np.transpose(Hi, (1, 2, 0)))                  # E x V x T
    dev = torch.device("cuda", torch.cuda.current_device())
    dX = kern._colmajor(xt.to(torch.complex64), torch.complex64, dev)
    dH = kern._colmajor(torch.from_numpy(Hi), torch.complex64, dev)
    y = torch.empty(T * N * E, dtype=torch.complex64, device=dev)
    t0 = np.atleast_1d(np.asarray(chd.t0, np.float64)).reshape(-1)
    if t0.size not in (1, V): raise ValueError("refocus: t0 must be scalar or per transmit")
    p = _lib.RefocusParams()
    p.struct_size = C.sizeof(_lib.RefocusParams)
    p.dtype, p.T, p.N, p.V, p.E, p.n_t0, p.fs = _lib.F32, T, N, V, E, int(t0.size), fs
    t0c = (C.c_double * int(t0.size))(*[float(v) for v in t0])
    t0o = C.c_double(0.0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qups_refocus(C.byref(p), kern._ptr(y), kern._ptr(dX), kern._ptr(dH), t0c, C.byref(t0o), kern._stream(dev)))
    out = kern._from_colmajor(y, (T, N, E))
    if not (isinstance(x, torch.Tensor) and x.is_cuda): out = np.asfortranarray(out.cpu().numpy())
    return ChannelData(out, float(t0o.value), fs), Hi


def greens_raw(ps, amp, pn, pv, kern_s, n0, T, fs, c0, wv_t0, fsr=1.0, R0=0.0, interp="cubic", device=None,
               dtype=np.float32, E=1):
    """Direct call of the qups_greens C ABI; returns a CUDA tensor of logical shape (T, N, M).
    dtype=np.float16 selects the half variant (greensh, src/greens.cu:113-122): half2 waveform in, half2 traces out
    (returned widened to complex64), fp32 geometry.  E > 1: pn / pv hold E sub-element positions per element, column
    n + N*en (3 x N x E flattened; element sub-divisions, src/UltrasoundSystem.m:785-790)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    half = np.dtype(dtype) == np.float16
    rt = torch.float32 if np.dtype(dtype) in (np.float32, np.float16) else torch.float64
    ct = torch.complex64 if rt == torch.float32 else torch.complex128
    col = lambda a_: torch.from_numpy(np.ascontiguousarray(np.asarray(a_, np.float64).T)).to(dev, rt).contiguous()
    dPs, dPn, dPv = col(ps), col(pn), col(pv)
    dA = torch.from_numpy(np.asarray(amp, np.float64)).to(dev, rt).contiguous()
    dK = torch.from_numpy(np.asarray(kern_s, np.complex128)).to(dev, ct).contiguous()
    N, M = dPn.shape[0] // E, dPv.shape[0] // E
    y = torch.empty((M, N, T), dtype=ct, device=dev)
    if half:
        dK = torch.view_as_real(dK).to(torch.float16).contiguous()
        y = torch.empty((M, N, T, 2), dtype=torch.float16, device=dev)
    p = GreensParams()
    p.struct_size = C.sizeof(GreensParams)
    p.dtype = _lib.F16 if half else (_lib.F32 if rt == torch.float32 else _lib.F64)
    p.I, p.S, p.T, p.N, p.M, p.E = dPs.shape[0], T, dK.shape[0], N, M, E
    p.n0, p.interp = int(n0), _lib.INTERP[interp]
    p.t0x, p.fs, p.fsr, p.c0, p.R0 = float(wv_t0), float(fs), float(fsr), float(c0), float(R0)
    vp = lambda t_: C.c_void_p(t_.data_ptr())
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().qups_greens(C.byref(p), vp(y), vp(dPs), vp(dA), vp(dPn), vp(dPv), vp(dK), st))
    if half:
        y = torch.view_as_complex(y.float())
    return y.permute(2, 1, 0)
