"""GPU parity tests of the fused ChannelData pre-processing pass (src/ChannelData.m:757-807, 935-966,
1153-1183): qups_chd_prep vs the NumPy restatement.  Tolerances: zeropad / cast exact; downmix 2e-6 (fp32 complex
multiply + sincosf); hilbert 1e-5 of max|x| (fp32 FFT vs a float64 FFT)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _chd(T, N, M, kind="real", seed=0, t0=None, fs=20e6):
    from qups_b200 import ultrasound as U
    rng = np.random.default_rng(seed)
    if kind == "real": x = rng.standard_normal((T, N, M)).astype(np.float32)
    elif kind == "i16": x = rng.integers(-2000, 2000, (T, N, M)).astype(np.int16)
    elif kind == "f64": x = rng.standard_normal((T, N, M))
    else: x = (rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M))).astype(np.complex64)
    t0 = 1.3e-6 if t0 is None else t0
    return U.ChannelData(x, t0, fs), x


def _np(y):
    return y.data.cpu().numpy() if hasattr(y.data, "cpu") else np.asarray(y.data)


def test_zeropad_exact_and_t0():
    chd, x = _chd(100, 5, 3, "cplx")
    y = chd.zeropad(7, 12)
    Y = _np(y)
    assert Y.shape == (119, 5, 3)
    assert np.array_equal(Y[7:107], x) and not Y[:7].any() and not Y[107:].any()
    assert y.t0 == pytest.approx(1.3e-6 - 7 / 20e6)


@pytest.mark.parametrize("kind", ["real", "i16", "f64", "cplx"])
def test_casts(kind):
    chd, x = _chd(64, 4, 3, kind)
    Y = _np(chd.singleT())
    assert Y.dtype == np.complex64 and np.array_equal(Y, x.astype(np.complex64))
    H = _np(chd.halfT())
    want = x.astype(np.complex64)
    assert np.allclose(H.astype(np.complex64), want, rtol=1e-3, atol=1e-3 * np.abs(want).max())


@pytest.mark.parametrize("T", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 100, 777, 1500, 4096, 8192])
def test_hilbert_matches_float64_fft(T):
    from oracle import prep_np
    chd, x = _chd(T, 3, 2, "real", seed=T)
    Y = _np(chd.hilbert())
    ref, _ = prep_np.prep(x, chd.t0, chd.fs, hilbert=True)
    assert Y.shape == ref.shape
    assert np.max(np.abs(Y - ref)) < 1e-5 * np.max(np.abs(ref))
    assert np.max(np.abs(Y.real - x)) < 1e-5 * np.max(np.abs(x))   # the analytic signal keeps the input as its real part


def test_hilbert_zero_padded_length_and_limits():
    import qups_b200
    from oracle import prep_np
    chd, x = _chd(300, 2, 2, "real")
    Y = _np(chd.hilbert(512))
    ref, _ = prep_np.prep(x, chd.t0, chd.fs, A=212, hilbert=True)
    assert np.max(np.abs(Y - ref)) < 1e-5 * np.max(np.abs(ref))
    big, _ = _chd(5000, 1, 1, "real")
    with pytest.raises(qups_b200.QupsError):
        big.hilbert()


def test_downmix_per_transmit_t0():
    from oracle import prep_np
    t0 = np.array([1.0e-6, 1.7e-6, 2.9e-6])
    chd, x = _chd(500, 4, 3, "cplx", t0=t0)
    Y = _np(chd.downmix(5e6))
    ref, _ = prep_np.prep(x, t0, chd.fs, fmix=5e6)
    assert np.max(np.abs(Y - ref)) < 2e-6 * np.max(np.abs(ref))


def test_fused_chain_equals_the_chain_of_methods():
    from oracle import prep_np
    chd, x = _chd(1000, 6, 4, "i16", t0=np.array([0.0, 1e-7, 2e-7, 3e-7]))
    fused = chd.prep(B=8, A=16, hilbert=True, fmix=7.5e6)
    ref, t0r = prep_np.prep(x.astype(np.float64), chd.t0, chd.fs, B=8, A=16, hilbert=True, fmix=7.5e6)
    Y = _np(fused)
    assert Y.shape == (1024, 6, 4)
    assert np.max(np.abs(Y - ref)) < 1e-5 * np.max(np.abs(ref))
    assert np.allclose(fused.t0, t0r)
    chain = chd.zeropad(8, 16).hilbert().downmix(7.5e6)
    assert np.max(np.abs(_np(chain) - Y)) < 1e-5 * np.max(np.abs(ref))


def test_prep_feeds_das(oracle_c):
    """Pre-processed cube straight into DAS: same image as DAS on the oracle-prepared cube."""
    import qups_b200
    from oracle import prep_np
    from qups_b200 import ultrasound as U
    from tests.util import small_problem, oracle_kwargs, rel_linf
    P = small_problem("FC", nz=40, nx=36, N=12, M=5, T=256)
    rf = np.asfortranarray(P["x"].real.astype(np.float32))
    chd = U.ChannelData(rf, 0.0, P["fs"]).prep(hilbert=True)
    xr, _ = prep_np.prep(rf, 0.0, P["fs"], hilbert=True)
    f32 = np.float32
    got = qups_b200.das_spec("DAS", P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), chd.data,
                             0.0, P["fs"], P["c"], *P["opts"], "interp", "cubic").cpu().numpy()
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], np.asfortranarray(xr), 0.0, P["fs"], P["c"], interp="cubic",
                            **oracle_kwargs(P["opts"]))[..., 0]
    assert rel_linf(got, ref) < 2e-5
