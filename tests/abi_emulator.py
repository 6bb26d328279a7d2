"""CPU emulation of the qups_wsinterpd2 C-ABI call (TEST INFRASTRUCTURE ONLY).

Lets the CPU suite run the REAL host-side mirror (qups_b200/kern.py::wsinterpd2: dim moves, broadcast sizes, the 5 x D
stride matrix, column-major buffers) without a GPU: the library call is replaced by a Python interpreter of exactly the
arguments the mirror hands to the C ABI (include/qups_b200.h: qups_ws2_params + device pointers), evaluated with the
NumPy oracle's interp1.  What is under test is the packing, not the kernel."""
import contextlib
import ctypes as C
import itertools

import numpy as np
import torch

from oracle import oracle_np


def _buf(ptr, n, dtype):
    if not ptr or n == 0:
        return None
    return np.ctypeslib.as_array((C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)).view(dtype)


class FakeLib:
    """Stands in for ctypes.CDLL(libqups_b200.so) for the wsinterpd2 / wsinterpd entry points."""

    def __init__(self):
        self.calls = 0

    def qups_last_error(self):
        return b"emulator"

    def qups_wsinterpd2(self, p_ref, y, w, x, t1, t2, stream):
        p = p_ref._obj
        self.calls += 1
        D, T = int(p.D), int(p.T)
        sizes = [int(p.sizes[k]) for k in range(D)]
        st = np.array([[int(p.dstride[r + 5 * k]) for k in range(D)] for r in range(5)])  # rows: w, y, t1, t2, x-trace
        span = lambda r: 1 + int(sum((sizes[k] - 1) * st[r][k] for k in range(D)))
        rdt, cdt = (np.float64, np.complex128) if p.dtype == 2 else (np.float32, np.complex64)
        val = lambda v: v.value if hasattr(v, "value") else v
        W = _buf(val(w), span(0), rdt if p.w_real else cdt)
        Y = _buf(val(y), span(1), cdt)
        T1 = _buf(val(t1), span(2), rdt)
        T2 = _buf(val(t2), span(3), rdt) if val(t2) else None
        X = _buf(val(x), span(4) * T, cdt)
        method = {0: "nearest", 1: "linear", 2: "cubic"}[int(p.interp)]
        Y[:] = 0
        if int(np.prod(sizes)) > 20000:
            # large calls: rebuild the dense column-major arrays from the stride matrix and let the vectorised NumPy oracle do
            # the arithmetic (the element-by-element interpreter below stays the reference for small calls)
            def dense(buf, r, lead=None):
                shp = tuple(sizes[k] if st[r][k] else 1 for k in range(D))
                acc = 1
                for k in range(D):
                    if shp[k] != 1:
                        assert st[r][k] == acc, ("non-dense stride", r, st[r], shp)
                        acc *= shp[k]
                if lead is None:
                    return buf[:acc].reshape(shp, order="F")
                return buf[:acc * lead].reshape((lead,) + shp[1:], order="F")   # x: T samples along dim 1, traces behind
            assert st[4][0] == 0, "x must not vary along the sampling dimension"
            xd = dense(X, 4, lead=T)
            t1d = dense(T1, 2)
            t2d = dense(T2, 3) if T2 is not None else np.zeros((1,) * D, rdt)
            wd = dense(W, 0)
            sdim = tuple(k + 1 for k in range(D) if st[1][k] == 0 and sizes[k] > 1)
            out = oracle_np.wsinterpd2(xd, t1d, t2d, 1, wd, sdim, method, 0, 1j * p.omega)
            Y[:out.size] = np.asarray(out, cdt).reshape(-1, order="F")
            return 0
        for idx in itertools.product(*[range(s) for s in sizes]):
            off = [int(sum(i * s for i, s in zip(idx, st[r]))) for r in range(5)]
            t = rdt(T1[off[2]]) + (rdt(T2[off[3]]) if T2 is not None else rdt(0))
            v = oracle_np.interp1(X[off[4] * T:(off[4] + 1) * T], np.array([1 + t]), method, 0)[0]
            Y[off[1]] += np.exp(1j * p.omega * t) * W[off[0]] * v
        return 0

    def qups_wsinterpd(self, p_ref, y, w, x, t, stream):
        return self.qups_wsinterpd2(p_ref, y, w, x, t, None, stream)


@contextlib.contextmanager
def emulated(monkeypatch):
    """Route qups_b200.kern's device plumbing to the CPU and its library handle to FakeLib."""
    from qups_b200 import kern, _lib
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "check", lambda rc: None if rc == 0 else (_ for _ in ()).throw(RuntimeError(rc)))
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    real_device = torch.device
    monkeypatch.setattr(kern.torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(kern, "_stream", lambda dev: C.c_void_p(0))
    yield fake


# ---- qups_das / qups_das_fused -------------------------------------------------------------------------------------------
def _val(v):
    return v.value if hasattr(v, "value") else v


def _das_emulated(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, fused=None):
    """Interpret the qups_das argument list (include/qups_b200.h) with the C oracle: rebuild the MATLAB-shaped arrays from the
    column-major buffers, sizes and strides the mirror packed, run oracle_c.das_spec, write y back column-major."""
    from oracle import oracle_c, apod_np
    assert p.dtype == 0, "emulator: fp32 calls only"
    I1, I2, I3, N, M, T = (int(v) for v in (p.I1, p.I2, p.I3, p.N, p.M, p.T))
    F, S, I = int(p.F) or 1, int(p.S), int(p.I1 * p.I2 * p.I3)
    full = (I1, I2, I3, N, M)
    f32, c64 = np.float32, np.complex64
    Pi_ = _buf(_val(Pi), 3 * I, f32).reshape((3, I1, I2, I3), order="F")
    Pr_ = _buf(_val(Pr), 3 * N, f32).reshape((3, N), order="F")
    Pv4_ = _buf(_val(Pv4), 4 * M, f32).reshape((4, M), order="F")
    Nv_ = _buf(_val(Nv), 3 * M, f32).reshape((3, M), order="F")
    tpose = bool(p.flag & 32)
    xs = (T, M, N, F) if tpose else (T, N, M, F)
    x_ = _buf(_val(x), T * N * M * F, c64).reshape(xs, order="F")
    acs = [int(acstride[k]) for k in range(6 + 6 * S)]

    def strided(ptr, st6, dtype):
        shp = tuple(full[d] if st6[d] else 1 for d in range(5))
        n = int(np.prod(shp))
        # the mirror packs every array densely in column-major order, so non-zero strides must be the dense ones
        acc = 1
        for d in range(5):
            if shp[d] != 1:
                assert st6[d] == acc, ("non-dense stride", st6, shp)
                acc *= shp[d]
        b = _buf(_val(ptr) + st6[5] * np.dtype(dtype).itemsize, n, dtype)
        return b.reshape(shp, order="F")

    c_arr = strided(cinv, acs[:5] + [0], f32)
    apods = [strided(apod, acs[6 + 6 * s: 12 + 6 * s], f32 if p.apod_real else c64) for s in range(S)]
    if fused is not None:
        fa = fused
        aux = lambda ptr, n: None if not ptr else _buf(ptr, n, f32).copy()
        lat = None if not fa.lat else _buf(fa.lat, full[fa.lat_dim - 1], f32).astype(np.float64)
        Pi64, Pn64 = Pi_.astype(np.float64), Pr_.astype(np.float64)
        if fa.rx_kind in (1, 2):
            nn = aux(fa.rx_aux, 3 * N).reshape((3, N), order="F").astype(np.float64)
            th = np.rad2deg(np.arccos(np.float64(fa.rx_p[0]))) if fa.rx_kind == 1 else 90.0 / np.float64(fa.rx_p[0])
            gen = apod_np.apAcceptanceAngle if fa.rx_kind == 1 else apod_np.apCosineAngle
            a = gen(Pi64, Pn64, nn, th, literal=False)
            if fa.rx_kind == 1:  # threshold exactly as passed (cosd(theta) was rounded to fp32 by the caller)
                a = (apod_np._dircos(Pi64, Pn64, nn, False) >= f32(fa.rx_p[0])).astype(f32)
            apods.append(a.astype(f32))
        elif fa.rx_kind == 3:
            ae = None
            if fa.rx_p[2]:
                cs = aux(fa.rx_aux, 2 * N).reshape((2, N), order="F")
                ae = np.rad2deg(np.arctan2(cs[1].astype(np.float64), cs[0].astype(np.float64)))
            apods.append(apod_np.apApertureGrowth(Pi64, Pn64, ae=ae, f=fa.rx_p[0], Dmax=fa.rx_p[1], literal=False).astype(f32))
        elif fa.rx_kind == 4:
            xn = aux(fa.rx_aux, N).astype(np.float64)
            xi = apod_np._lateral(Pi64, lat, fa.lat_dim, f32)[..., None]
            apods.append((np.abs(xi - xn.astype(f32).reshape(1, 1, 1, -1)) <= f32(fa.rx_p[0])).astype(f32))
        if fa.tx_kind in (1, 2):
            xv = aux(fa.tx_aux, M)
            xi = apod_np._lateral(Pi64, lat, fa.lat_dim, f32)[..., None, None]
            d = np.abs(xi - xv.reshape(1, 1, 1, 1, -1))
            apods.append(((d < f32(fa.tx_p[0])) if fa.tx_kind == 1 else (d <= f32(fa.tx_p[0]))).astype(f32))
        elif fa.tx_kind == 3:
            q = aux(fa.tx_aux, 4 * M).reshape((4, M), order="F")
            P = Pi_[..., None]
            x0 = P[0] - q[0] * (P[2] / q[1])
            x1 = P[0] - q[2] * (P[2] / q[3])
            lo, hi = f32(fa.tx_p[0]), f32(fa.tx_p[1])
            apods.append((((lo < x0) | (lo < x1)) & ((x0 <= hi) | (x1 <= hi))).astype(f32)[:, :, :, None, :])
    keep_rx, keep_tx = bool(p.flag & 8), bool(p.flag & 16)
    fun = {(False, False): "DAS", (True, False): "SYN", (False, True): "MUL", (True, True): "BF"}[(keep_rx, keep_tx)]
    interp = {0: "nearest", 1: "linear", 2: "cubic", 3: "lanczos3"}[p.flag & 7]
    with np.errstate(divide="ignore"):
        c = (f32(1) / c_arr).astype(f32)
    out = oracle_c.das_spec(fun, Pi_, Pr_, Pv4_[:3], Nv_, x_, Pv4_[3], p.fs, c, interp=interp, apod=apods, VS=bool(p.vs), DV=bool(p.dv),
                            fmod=p.fmod, tpose=tpose)
    On, Om = (N if keep_rx else 1), (M if keep_tx else 1)
    yb = _buf(_val(y), I * On * Om * F, c64)
    yb[:] = np.asarray(out, c64).reshape(-1, order="F")
    return 0


def _install_das(fake):
    fake.last = "none"

    def qups_das(p_ref, y, Pi, Pr, Pv4, Nv, apod, cinv, acs, x, stream):
        fake.calls += 1
        fake.last = "das"
        return _das_emulated(p_ref._obj, y, Pi, Pr, Pv4, Nv, apod, cinv, acs, x)

    def qups_das_fused(p_ref, f_ref, y, Pi, Pr, Pv4, Nv, apod, cinv, acs, x, stream):
        fake.calls += 1
        fake.last = "das_fused"
        return _das_emulated(p_ref._obj, y, Pi, Pr, Pv4, Nv, apod, cinv, acs, x, fused=f_ref._obj)

    fake.qups_das, fake.qups_das_fused = qups_das, qups_das_fused
    fake.qups_last_das_kernel = lambda: b"emulated"


_orig_emulated = emulated


@contextlib.contextmanager
def emulated(monkeypatch):  # noqa: F811  (extends the context above with the DAS entry points)
    with _orig_emulated(monkeypatch) as fake:
        _install_das(fake)
        yield fake


# ---- qups_greens -------------------------------------------------------------------------------------------------------------
def _install_greens(fake):
    def qups_greens(p_ref, y, Pi, a, Pr, Pv, kern, stream):
        from oracle import oracle_c
        p = p_ref._obj
        fake.calls += 1
        assert p.dtype == 0 and int(p.E) in (0, 1), "emulator: fp32, E = 1"
        I, S, K, N, M = (int(v) for v in (p.I, p.S, p.T, p.N, p.M))
        f32, c64 = np.float32, np.complex64
        ps = _buf(_val(Pi), 3 * I, f32).reshape((3, I), order="F")
        amp = _buf(_val(a), I, f32)
        pn = _buf(_val(Pr), 3 * N, f32).reshape((3, N), order="F")
        pv = _buf(_val(Pv), 3 * M, f32).reshape((3, M), order="F")
        kn = _buf(_val(kern), K, c64)
        interp = {0: "nearest", 1: "linear", 2: "cubic"}[int(p.interp)]
        out = oracle_c.greens(ps, amp, pn, pv, kn, int(p.n0), S, p.fs, p.c0, p.t0x, p.fsr, p.R0, interp)   # S x N x M
        _buf(_val(y), S * N * M, c64)[:] = np.asarray(out, c64).reshape(-1, order="F")
        return 0

    fake.qups_greens = qups_greens


_emulated_das = emulated


@contextlib.contextmanager
def emulated(monkeypatch):  # noqa: F811  (adds the greens entry point and the device plumbing of ultrasound.py)
    from qups_b200 import ultrasound
    with _emulated_das(monkeypatch) as fake:
        _install_greens(fake)
        real_device = type(torch.zeros(1).device)
        monkeypatch.setattr(ultrasound.torch, "device", lambda *a, **k: real_device("cpu"))
        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 0})())
        yield fake


# ---- qups_chd_prep / qups_aperture / qups_apod_generate ------------------------------------------------------------------------
def _install_misc(fake):
    def qups_chd_prep(p_ref, out, inp, t0, stream):
        from oracle import prep_np
        p = p_ref._obj
        fake.calls += 1
        T, K, B, A = (int(v) for v in (p.T, p.K, p.B, p.A))
        L = B + T + A
        dt = {0: np.float32, 1: np.complex64, 2: np.int16, 3: np.float64}[int(p.in_dtype)]
        x = _buf(_val(inp), T * K, dt).reshape((T, K), order="F")
        n_t0, per = int(p.n_t0) or 1, int(p.traces_per_t0) or 1
        t0v = _buf(_val(t0), n_t0, np.float32).astype(np.float64) if _val(t0) else np.zeros(1)
        # trace k uses t0[(k / traces_per_t0) % n_t0]: evaluate the oracle per distinct t0
        y = np.zeros((L, K), np.complex64)
        for j in range(n_t0):
            sel = [k for k in range(K) if (k // per) % n_t0 == j]
            if not sel:
                continue
            xs = x[:, sel]
            xs = xs if np.iscomplexobj(xs) and not p.hilbert else np.real(xs).astype(np.float64) if p.hilbert else xs
            yy, _ = prep_np.prep(xs.reshape(T, len(sel), 1), float(t0v[j]), p.fs, B=B, A=A, hilbert=bool(p.hilbert), fmix=p.fmix)
            y[:, sel] = yy[:, :, 0]
        if int(p.out_dtype) == 1:   # half2 storage
            ob = _buf(_val(out), 2 * L * K, np.float16)
            ob[0::2] = y.real.reshape(-1, order="F").astype(np.float16)
            ob[1::2] = y.imag.reshape(-1, order="F").astype(np.float16)
        else:
            _buf(_val(out), L * K, np.complex64)[:] = y.reshape(-1, order="F")
        return 0

    def qups_aperture(p_ref, out, out2, b, lags, stream):
        from oracle import aperture_np as apd
        p = p_ref._obj
        fake.calls += 1
        Cn, A, Sn = int(p.C), int(p.A), int(p.S)
        cdt, rdt = (np.complex128, np.float64) if p.dtype == 2 else (np.complex64, np.float32)
        x = _buf(_val(b), Cn * A * Sn, cdt).reshape((Cn, A, Sn), order="F")
        lg = [int(lags[k]) for k in range(int(p.nlags))]
        op = int(p.op)
        if op == 0: r = apd.cohfac(x, 2)
        elif op == 1: r = apd.dmas(x, 2, lg if lg else [A + 1])
        elif op == 2: r, sf = apd.pcf(x, 2, p.gamma)
        else: r = apd.slsc(x, 2, lg, "average" if op == 3 else "ensemble")
        cplx = op in (1, 3, 4)
        _buf(_val(out), Cn * Sn, cdt if cplx else rdt)[:] = np.asarray(r).reshape(-1, order="F").astype(cdt if cplx else rdt)
        if op == 2 and _val(out2):
            _buf(_val(out2), Cn * Sn, rdt)[:] = np.asarray(sf).reshape(-1, order="F").astype(rdt)
        return 0

    fake.qups_chd_prep, fake.qups_aperture = qups_chd_prep, qups_aperture


_emulated_greens = emulated


@contextlib.contextmanager
def emulated(monkeypatch):  # noqa: F811  (adds the pre-processing and aperture entry points)
    with _emulated_greens(monkeypatch) as fake:
        _install_misc(fake)
        yield fake


# ---- qups_apod_generate ----------------------------------------------------------------------------------------------------------
def _install_apodgen(fake):
    def qups_apod_generate(f_ref, which, out, as_complex, Pi, Pr, I1, I2, I3, NM, stream):
        """Dense image of a closed-form apodization, from the canonical-fp32 oracle (same reconstruction as _das_emulated)."""
        fa = f_ref._obj
        fake.calls += 1
        I1, I2, I3, NM = int(I1), int(I2), int(I3), int(NM)
        I = I1 * I2 * I3
        f32 = np.float32
        Pi_ = _buf(_val(Pi), 3 * I, f32).reshape((3, I1, I2, I3), order="F")
        from oracle import apod_np
        Pi64 = Pi_.astype(np.float64)
        lat = None if not fa.lat else _buf(fa.lat, (I1, I2, I3)[fa.lat_dim - 1], f32).astype(np.float64)
        if int(which) == 0:
            Pn = _buf(_val(Pr), 3 * NM, f32).reshape((3, NM), order="F").astype(np.float64)
            if fa.rx_kind in (1, 2):
                nn = _buf(fa.rx_aux, 3 * NM, f32).reshape((3, NM), order="F").astype(np.float64)
                if fa.rx_kind == 1:
                    a = (apod_np._dircos(Pi64, Pn, nn, False) >= f32(fa.rx_p[0])).astype(f32)
                else:
                    a = apod_np.apCosineAngle(Pi64, Pn, nn, 90.0 / np.float64(fa.rx_p[0]), literal=False)
            elif fa.rx_kind == 3:
                ae = None
                if fa.rx_p[2]:
                    cs = _buf(fa.rx_aux, 2 * NM, f32).reshape((2, NM), order="F")
                    ae = np.rad2deg(np.arctan2(cs[1].astype(np.float64), cs[0].astype(np.float64)))
                a = apod_np.apApertureGrowth(Pi64, Pn, ae=ae, f=fa.rx_p[0], Dmax=fa.rx_p[1], literal=False)
            elif fa.rx_kind == 4:
                xn = _buf(fa.rx_aux, NM, f32)
                xi = apod_np._lateral(Pi64, lat, fa.lat_dim, f32)[..., None]
                a = (np.abs(xi - xn.reshape(1, 1, 1, -1)) <= f32(fa.rx_p[0]))
            else:
                a = np.ones((I1, I2, I3, NM))
        else:
            if fa.tx_kind in (1, 2):
                xv = _buf(fa.tx_aux, NM, f32)
                xi = apod_np._lateral(Pi64, lat, fa.lat_dim, f32)[..., None]
                d = np.abs(xi - xv.reshape(1, 1, 1, -1))
                a = (d < f32(fa.tx_p[0])) if fa.tx_kind == 1 else (d <= f32(fa.tx_p[0]))
            elif fa.tx_kind == 3:
                q = _buf(fa.tx_aux, 4 * NM, f32).reshape((4, NM), order="F")
                P_ = Pi_[..., None]
                x0, x1 = P_[0] - q[0] * (P_[2] / q[1]), P_[0] - q[2] * (P_[2] / q[3])
                lo, hi = f32(fa.tx_p[0]), f32(fa.tx_p[1])
                a = ((lo < x0) | (lo < x1)) & ((x0 <= hi) | (x1 <= hi))
            else:
                a = np.ones((I1, I2, I3, NM))
        a = np.asarray(a, f32).reshape(-1, order="F")
        if as_complex:
            _buf(_val(out), I * NM, np.complex64)[:] = a
        else:
            _buf(_val(out), I * NM, f32)[:] = a
        return 0

    fake.qups_apod_generate = qups_apod_generate


_emulated_misc = emulated


@contextlib.contextmanager
def emulated(monkeypatch):  # noqa: F811  (adds the dense apodization generator)
    with _emulated_misc(monkeypatch) as fake:
        _install_apodgen(fake)
        yield fake
