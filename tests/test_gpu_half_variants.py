"""GPU parity of the half-precision variants of the non-DAS path kernels — wsinterpd2h / wsinterpdh (src/interpd.cu:422-429,
451-458), greensh (src/greens.cu:113-122), convh / convch (src/convd.cu:141,153).  fp16 parity definition:
the oracle on the fp16-ROUNDED inputs with fp32 math; with fp32 output (y_f32) the comparison holds at the fp32 bar, with half
output up to one half rounding of the result."""
import numpy as np
import pytest

from tests.util import rel_linf

pytestmark = pytest.mark.gpu
f32, f16 = np.float32, np.float16


def _r16(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return (a.real.astype(f16).astype(f32) + 1j * a.imag.astype(f16).astype(f32)).astype(np.complex64)
    return a.astype(f16).astype(f32)


@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_wsinterpd2h_generic_and_staged(oracle_c, interp, monkeypatch):
    import qups_b200
    rng = np.random.default_rng(4)
    T, N, M, I = 160, 12, 5, 301
    x = _r16(rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M)))
    x[:4] = 0
    x[-4:] = 0
    base = np.linspace(10, 60, I)[:, None]
    # half tables: keep them on a coarse grid so that fp16 represents them exactly enough to stay off nearest ties
    tn = _r16(np.round((base + rng.uniform(0, 20, (1, N))) * 8) / 8 + 0.0625)
    tm = _r16(np.round((0.5 * base + rng.uniform(0, 20, (1, M))) * 8) / 8)
    w = f32(0.75)
    ref = oracle_c.wsinterpd2_inm(x, tn.reshape(I, N, 1), tm.reshape(I, 1, M), interp=interp).reshape(-1) * w
    args = (x, tn.reshape(I, N, 1), tm.reshape(I, 1, M), 1, w, (2, 3), interp)
    got = np.asarray(qups_b200.wsinterpd2(*args, _prec="halfT", _y_f32=True)).reshape(-1)
    assert qups_b200.last_ws2_kernel() == "ws2_tiled"
    assert rel_linf(got, ref) < 1e-5, rel_linf(got, ref)
    got16 = np.asarray(qups_b200.wsinterpd2(*args, _prec="halfT")).reshape(-1)
    assert rel_linf(got16, ref) < 2e-3
    monkeypatch.setenv("QUPS_B200_WS2_GENERIC", "1")
    gen = np.asarray(qups_b200.wsinterpd2(*args, _prec="halfT", _y_f32=True)).reshape(-1)
    assert qups_b200.last_ws2_kernel() == "wsinterpd2"
    assert rel_linf(gen, ref) < 1e-5, rel_linf(gen, ref)
    # single-table variant with a weight array and a kept aperture (wsinterpdh): generic kernel
    tau = _r16(rng.uniform(-2, T + 1, (I, N, M)))
    wa = _r16(rng.uniform(0, 1, (I, N, M)))
    ref1 = oracle_c.wsinterpd2_inm(x, tau, np.zeros((1, 1, 1), f32), wa, sum_m=False, interp=interp)
    got1 = np.asarray(qups_b200.wsinterpd(x, tau, 1, wa, (2,), interp, _prec="halfT", _y_f32=True))
    assert rel_linf(got1.reshape(ref1.shape), ref1) < 1e-5


def test_greensh_matches_fp32_kernel_on_rounded_waveform(oracle_c):
    from qups_b200 import synth
    from qups_b200.ultrasound import greens_raw
    fs, fc, c0 = 20e6, 5e6, 1500.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    kern = _r16(kern / np.abs(kern).max())
    pn = synth.linear_array(7, 0.3e-3)
    rng = np.random.default_rng(2)
    S = 400
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), np.zeros(S), rng.uniform(3e-3, 12e-3, S)], 0)
    amp = rng.standard_normal(S) * 1e-4          # keeps the half2 output inside fp16 range (1/r^2 ~ 1e4..1e5)
    n0, T = 40, 700
    ref = oracle_c.greens(ps, amp, pn, pn, kern, n0, T, fs, c0, wt0, 1.0, 3e-4, "cubic")
    g32 = greens_raw(ps, amp, pn, pn, kern, n0, T, fs, c0, wt0, 1.0, 3e-4, "cubic").cpu().numpy()
    g16 = greens_raw(ps, amp, pn, pn, kern, n0, T, fs, c0, wt0, 1.0, 3e-4, "cubic", dtype=np.float16).cpu().numpy()
    assert np.abs(ref).max() > 0 and np.abs(g16).max() < 6e4
    assert rel_linf(g32, ref) < 1e-3
    assert np.array_equal(g16, _r16(g32))         # same fp32 sum, ONE rounding to half at the end


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", ["full", "same", "valid"])
def test_convh_convch(cplx, shape):
    import qups_b200
    rng = np.random.default_rng(6)
    mk = lambda *s: _r16(rng.standard_normal(s) + (1j * rng.standard_normal(s) if cplx else 0))
    x, y = mk(5, 40, 3), mk(5, 9, 3)
    ref, lags = qups_b200.convd(x, y, 2, shape)                 # fp32 kernel on the same (rounded) inputs
    got, lags2 = qups_b200.convd(x, y, 2, shape, _half=True)
    assert np.array_equal(lags, lags2)
    assert np.array_equal(np.asarray(got), _r16(np.asarray(ref)))   # fp32 accumulation, one rounding to half per output
