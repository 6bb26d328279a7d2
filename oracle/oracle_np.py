"""oracle_np.py — independent NumPy restatement of the QUPS DAS hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg as the *checker*; never by the product path
(qups_b200/), which must fail loudly without its CUDA library.

PARITY STATUS: "parity unpinned" for exact image values (see
oracle/qups_oracle.h). This module is the second, array-style restatement of

  * kern/das_spec.m:391-561    (CPU branch: 'DAS' | 'SYN' | 'MUL' | 'BF' | 'delays')
  * MATLAB interp1(v, xq, method, 0) as called at kern/das_spec.m:477
  * kern/wsinterpd2.m:240-308  (CPU branch, general N-D broadcasting)
  * kern/wsinterpd.m:230-270
  * src/UltrasoundSystem.m:720-863 (greens CPU math)

used to cross-check the C oracle (oracle/qups_oracle.c), and as the fp64
arbiter.  NumPy rounds every float32 operation individually, so with
dtype=float32 it performs the canonical fp32 sequence of SURVEY.md §8(c).
"""
from __future__ import annotations

import numpy as np

INTERP = {"nearest": 0, "linear": 1, "cubic": 2, "lanczos3": 3}


def _cdtype(rdtype):
    return np.complex64 if np.dtype(rdtype) == np.float32 else np.complex128


def interp1(v: np.ndarray, xq: np.ndarray, method: str = "linear", extrapval=0) -> np.ndarray:
    """MATLAB ``interp1(v, xq, method, extrapval)`` on the implicit grid 1..T.

    v is a (T,) real/complex vector, xq any shape (1-based sample positions).
    Semantics per SURVEY.md §8(c); call sites kern/das_spec.m:477,506,536,554.
    """
    v = np.asarray(v)
    T = v.shape[0]
    xq = np.asarray(xq)
    rdt = xq.dtype
    out = np.full(xq.shape, extrapval, dtype=np.result_type(v.dtype, np.float32 if rdt == np.float32 else rdt))
    ok = (xq >= 1) & (xq <= T)  # NaN -> False
    if not ok.any():
        return out
    q = xq[ok]
    one = rdt.type(1)
    if method == "nearest":
        # round half away from zero (q >= 1 so floor(q + .5) with exact tie handling)
        k = np.floor(q).astype(np.int64)
        k = k + ((q - k.astype(rdt)) >= rdt.type(0.5))
        k = np.clip(k, 1, T)
        out[ok] = v[k - 1]
    elif method == "linear":
        k = np.clip(np.floor(q).astype(np.int64), 1, T - 1)
        s = (q - k.astype(rdt)).astype(rdt)
        v0, v1 = v[k - 1], v[k]
        d = v1 - v0
        d = s * d
        out[ok] = v0 + d
    elif method == "cubic":
        # R2020b+ 'cubic' == cubic convolution (Keys, a=-1/2) with end padding
        vp = np.empty(T + 2, dtype=v.dtype)
        vp[1:-1] = v
        three = v.real.dtype.type(3)
        vp[0] = (three * v[0] - three * v[1]) + v[2]
        vp[-1] = (three * v[-1] - three * v[-2]) + v[-3]
        k = np.clip(np.floor(q).astype(np.int64), 1, T - 1)
        s = (q - k.astype(rdt)).astype(rdt)
        s2 = s * s
        s3 = s2 * s
        w0 = (-one * s3 + rdt.type(2) * s2) - s
        w1 = (rdt.type(3) * s3 - rdt.type(5) * s2) + rdt.type(2)
        w2 = (rdt.type(-3) * s3 + rdt.type(4) * s2) + s
        w3 = s3 - s2
        acc = ((w0 * vp[k - 1] + w1 * vp[k]) + w2 * vp[k + 1]) + w3 * vp[k + 2]
        out[ok] = rdt.type(0.5) * acc
    elif method == "lanczos3":
        # GPU-only sampler of the reference: src/interpd.cu:118-150 (a = 2)
        tau = q - one
        kf = np.floor(tau)
        ti = kf.astype(np.int64)
        u = (tau - kf).astype(rdt)
        good = (ti - 1 >= 0) & (ti + 2 < T)
        acc = np.zeros(q.shape, dtype=out.dtype)
        tig = np.clip(ti, 1, T - 3)
        for j in (-1, 0, 1, 2):
            uu = (u - rdt.type(j)).astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                w = 2.0 * np.sin(np.pi * uu) * np.sin(np.pi * uu / 2.0) / (np.pi * np.pi * uu * uu)
            w = np.where(uu == 0, 1.0, w).astype(rdt)
            acc = acc + w * v[tig + j]
        out[ok] = np.where(good, acc, 0)
    else:
        raise ValueError(f"unknown interp method {method!r}")
    return out


def _norm3(r):
    """vecnorm(r, 2, 1) for a (3, ...) array with individually rounded ops."""
    q = r[0] * r[0]
    q = q + r[1] * r[1]
    q = q + r[2] * r[2]
    return np.sqrt(q)


def _dot3(a, b):
    q = a[0] * b[0]
    q = q + a[1] * b[1]
    q = q + a[2] * b[2]
    return q


def tx_rx_distances(Pi, Pr, Pv, Nv, VS=True, DV=False):
    """kern/das_spec.m:427-436. Pi (3,I), Pr (3,N), Pv/Nv (3,M) -> dv (I,M), dr (I,N)."""
    rv = Pi[:, :, None] - Pv[:, None, :]  # 3 x I x M
    if VS:
        d = _norm3(rv)
        if DV:
            dv = d
        else:
            s = np.sign(_dot3(rv, Nv[:, None, :]))
            dv = d * s.astype(d.dtype)
    else:
        dv = _dot3(rv, Nv[:, None, :])
    dr = _norm3(Pi[:, :, None] - Pr[:, None, :])  # I x N
    return dv, dr


def _bcast5(a, Isz, N, M, rdt=None):
    """Broadcast-index helper: view ``a`` as 5-D (I1,I2,I3,N,M) with singleton dims kept."""
    a = np.asarray(a)
    shp = list(a.shape) + [1] * (5 - a.ndim)
    full = list(Isz) + [N, M]
    for d in range(5):
        if shp[d] not in (1, full[d]):
            raise ValueError("size inconsistent with I1 x I2 x I3 x N x M")
    return a.reshape(shp, order="F")


def das_spec(fun, Pi, Pr, Pv, Nv, x, t0, fs, c=1540.0, *, interp="linear", apod=(), VS=True, DV=False,
             fmod=0.0, tpose=False, dtype=np.float32):
    """CPU branch of ``das_spec`` (kern/das_spec.m:391-561).

    Pi is (3, I1, I2, I3) (Fortran pixel order i = i1 + I1*(i2 + I2*i3)), x is
    (T, N, M[, F]) stored so that x[:, n, m] is a trace (transposed: (T, M, N)).
    Returns (I1, I2, I3, [1|N], [1|M][, F]) complex, or real delays.
    """
    rdt = np.dtype(dtype)
    cdt = _cdtype(rdt)
    Pi = np.asarray(Pi, dtype=rdt)
    Isz = tuple(Pi.shape[1:]) + (1,) * (4 - Pi.ndim)
    I = int(np.prod(Isz))
    P = Pi.reshape(3, I, order="F")
    Pr = np.asarray(Pr, dtype=rdt).reshape(3, -1)
    Pv = np.asarray(Pv, dtype=rdt).reshape(3, -1)
    Nv = np.asarray(Nv, dtype=rdt).reshape(3, -1)
    x = np.asarray(x)
    if fun != "delays":
        x = x.astype(cdt)
        if x.ndim == 3:
            x = x[..., None]
        T, d2, d3, F = x.shape
        N, M = (d3, d2) if tpose else (d2, d3)
    else:
        N, M, F, T = Pr.shape[1], max(Pv.shape[1], Nv.shape[1]), 1, 0
    # expand_inputs: kern/das_spec.m:636-641
    if Pv.shape[1] == 1:
        Pv = np.repeat(Pv, M, axis=1)
    if Nv.shape[1] == 1:
        Nv = np.repeat(Nv, M, axis=1)
    if Pr.shape[1] == 1:
        Pr = np.repeat(Pr, N, axis=1)
    assert Pv.shape[1] == M and Nv.shape[1] == M and Pr.shape[1] == N
    cinv = (rdt.type(1) / np.asarray(c, dtype=rdt)).astype(rdt)  # cinv = 1./c  :170
    cinv5 = _bcast5(cinv, Isz, N, M)
    t0v = np.broadcast_to(np.asarray(t0, dtype=rdt).reshape(-1), (M,)) if np.size(t0) in (1, M) else None
    assert t0v is not None, "t0 must be scalar or one per transmit"
    fs = rdt.type(fs)

    dv, dr = tx_rx_distances(P, Pr, Pv, Nv, VS, DV)  # (I,M), (I,N)

    def cinv_nm(n, m):
        sl = cinv5[:, :, :, n if cinv5.shape[3] > 1 else 0, m if cinv5.shape[4] > 1 else 0]
        return np.broadcast_to(sl, Isz).reshape(I, order="F")

    if fun == "delays":  # :448-449
        out = np.empty((I, N, M), dtype=rdt)
        for m in range(M):
            for n in range(N):
                out[:, n, m] = cinv_nm(n, m) * (dv[:, m] + dr[:, n])
        return out.reshape(Isz + (N, M), order="F")

    # apply (de)modulation to the data: :413-417
    if fmod:
        j = np.arange(T, dtype=np.float64).astype(rdt)
        tj = t0v[None, :] + (j / fs)[:, None]  # T x M
        w = rdt.type(2.0 * np.pi * fmod)
        th = w * tj
        ph = (np.cos(th) + 1j * np.sin(th)).astype(cdt)  # T x M
        x = x * (ph[:, :, None, None] if tpose else ph[:, None, :, None])

    apods = [_bcast5(np.asarray(a), Isz, N, M) for a in apod]

    def apod_nm(a5, n, m):
        sl = a5[:, :, :, n if a5.shape[3] > 1 else 0, m if a5.shape[4] > 1 else 0]
        sl = np.broadcast_to(sl, Isz).reshape(I, order="F")
        return sl.astype(cdt if np.iscomplexobj(sl) else rdt)

    keep_rx = fun in ("SYN", "BF")
    keep_tx = fun in ("MUL", "BF")
    y = np.zeros((I, N if keep_rx else 1, M if keep_tx else 1, F), dtype=cdt)
    one = rdt.type(1)
    for m in range(M):
        yn = np.zeros((I, F), dtype=cdt)
        for n in range(N):
            tau = cinv_nm(n, m) * (dv[:, m] + dr[:, n])
            tau = tau - t0v[m]
            xq = tau * fs
            xq = one + xq
            for f in range(F):
                tr = x[:, m, n, f] if tpose else x[:, n, m, f]
                v = interp1(tr, xq, interp, 0).astype(cdt)
                if fun == "BF":
                    for a5 in apods:  # :558
                        v = v * apod_nm(a5, n, m)
                    y[:, n, m, f] = v
                else:
                    if apods:  # a = asn{end}; for s=1:S-1, a = a.*asn{s}   :473
                        a = apod_nm(apods[-1], n, m)
                        for a5 in apods[:-1]:
                            a = a * apod_nm(a5, n, m)
                        v = a * v
                    if fun == "DAS":
                        yn[:, f] = yn[:, f] + v
                    elif fun == "SYN":
                        y[:, n, 0, f] = y[:, n, 0, f] + v
                    elif fun == "MUL":
                        y[:, 0, m, f] = y[:, 0, m, f] + v
        if fun == "DAS":
            y[:, 0, 0, :] = y[:, 0, 0, :] + yn
    return y.reshape(Isz + y.shape[1:], order="F")


def wsinterpd2(x, t1, t2, dim=1, w=1, sdim=(), interp="linear", extrapval=0, omega=0):
    """General N-D ``wsinterpd2`` (kern/wsinterpd2.m:1-320, CPU branch :240-308).

    y = sum_{sdim} w .* exp(omega .* (t1+t2)) .* interp1(x, 1 + t1 + t2, interp, extrapval)
    x has time along ``dim`` (1-based); t1, t2 have the sample-index dimension at
    ``dim``; all other dims broadcast (MATLAB implicit expansion). ``sdim`` are
    1-based dims of the broadcast result to sum (kept as singleton).
    """
    x = np.asarray(x)
    t1 = np.asarray(t1)
    t2 = np.asarray(t2)
    w = np.asarray(w)
    nd = max(x.ndim, t1.ndim, t2.ndim, w.ndim, dim, max(sdim, default=1))

    def lift(a):
        return a.reshape(a.shape + (1,) * (nd - a.ndim))

    x, t1, t2, w = lift(x), lift(t1), lift(t2), lift(w)
    ax = dim - 1
    t = t1 + t2
    rdt = t.dtype
    xm = np.moveaxis(x, ax, 0)  # T, rest_x
    tm = np.moveaxis(t, ax, 0)  # K, rest_t
    rest = np.broadcast_shapes(xm.shape[1:], tm.shape[1:])
    xb = np.broadcast_to(xm, (xm.shape[0],) + rest)
    tb = np.broadcast_to(tm, (tm.shape[0],) + rest)
    T = xb.shape[0]
    cdt = np.result_type(x.dtype, np.complex64 if rdt == np.float32 else np.complex128) if (
        np.iscomplexobj(x) or np.iscomplexobj(w) or omega != 0) else x.dtype
    out = np.empty(tb.shape, dtype=cdt)
    xf = xb.reshape(T, -1)
    tf = tb.reshape(tb.shape[0], -1)
    of = out.reshape(tb.shape[0], -1)
    for j in range(xf.shape[1]):
        of[:, j] = interp1(xf[:, j], rdt.type(1) + tf[:, j], interp, extrapval)
    out = np.moveaxis(of.reshape(tb.shape), 0, ax)
    if omega != 0:
        th = (rdt.type(np.imag(omega)) * t).astype(rdt)
        out = (np.cos(th) + 1j * np.sin(th)).astype(cdt) * out
    out = w * out
    if sdim:
        out = np.nansum(out, axis=tuple(s - 1 for s in sdim), keepdims=True)
    return out


def wsinterpd(x, t, dim=1, w=1, sdim=(), interp="linear", extrapval=0, omega=0):
    """``wsinterpd`` (kern/wsinterpd.m:1-281): single delay table variant."""
    t = np.asarray(t)
    return wsinterpd2(x, t, np.zeros((1,) * t.ndim, dtype=t.dtype), dim, w, sdim, interp, extrapval, omega)


def greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wv_t0, fsr=1.0, R0=0.0, interp="cubic", dtype=np.float32):
    """CPU math of ``UltrasoundSystem.greens`` (src/UltrasoundSystem.m:778-851).

    ps (3,S), amp (S,), pn (3,N), pv (3,M), kern (K,) complex, output time axis
    t = n0 : n0+T-1 (integer samples, :613-615). Returns x (T,N,M) complex.
    """
    rdt = np.dtype(dtype)
    cdt = _cdtype(rdt)
    ps = np.asarray(ps, rdt)
    pn = np.asarray(pn, rdt)
    pv = np.asarray(pv, rdt)
    amp = np.asarray(amp, rdt)
    kern = np.asarray(kern, cdt)
    S, N, M = ps.shape[1], pn.shape[1], pv.shape[1]
    c0, fs, fsr, R0 = rdt.type(c0), rdt.type(fs), rdt.type(fsr), rdt.type(R0)
    t0 = rdt.type(wv_t0) * fs
    tvec = np.arange(n0, n0 + T, dtype=np.int64).astype(rdt)  # T
    r_rx = _norm3(ps[:, :, None] - pn[:, None, :])  # S x N
    r_tx = _norm3(ps[:, :, None] - pv[:, None, :])  # S x M
    tau_rx = (r_rx / c0) * fs
    tau_tx = (r_tx / c0) * fs
    x = np.zeros((T, N, M), dtype=cdt)
    for m in range(M):
        for n in range(N):
            if R0:
                att = amp / (np.maximum(r_rx[:, n], R0) * np.maximum(r_tx[:, m], R0))
            else:
                att = amp
            wgt = att / fsr
            t1 = fsr * ((tvec[None, :] - tau_tx[:, m, None]) - t0)  # S x T
            t2 = (-fsr) * tau_rx[:, n]  # S
            tau = t1 + t2[:, None]
            v = interp1(kern, rdt.type(1) + tau, interp, 0).astype(cdt)  # S x T
            v = wgt[:, None] * v
            acc = np.zeros(T, dtype=cdt)
            for s in range(S):  # sequential sum over scatterers, in order
                acc = acc + v[s]
            x[:, n, m] = acc
    return x
