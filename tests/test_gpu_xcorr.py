"""GPU parity of qups_pwznxcorr (csrc/xcorr.cu) vs the NumPy restatement of kern/pwznxcorr.m (oracle/xcorr_np.py).
Tolerance: 1e-5 of the largest magnitude for fp32 (moving sums in the oracle's tap order; FMA contraction only), 1e-12 fp64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(shape, seed=0, dtype=np.complex64):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.complexfloating):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def _close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    a, b = np.nan_to_num(a), np.nan_to_num(b)
    assert np.max(np.abs(a - b)) <= tol * max(1e-30, np.max(np.abs(b))), np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 1e-5), (np.complex128, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("ref", ["neighbor", "center", "x0"])
@pytest.mark.parametrize("zero,norm", [(True, True), (False, True), (True, False), (False, False)])
def test_pwznxcorr_matches_oracle(dtype, tol, ref, zero, norm):
    import qups_b200
    from oracle.xcorr_np import pwznxcorr
    x = _data((700, 6, 3), 1, dtype)  # two time tiles of 512
    x0 = _data((700, 1, 3), 2, np.complex128 if dtype == np.complex128 else np.complex64) if ref == "x0" else None
    lags = [-5, -1, 0, 2, 7]
    kw = dict(zero=zero, norm=norm, ref=ref, x0=x0)
    got = qups_b200.pwznxcorr(x, lags, 12, **kw)
    want = pwznxcorr(x, lags, 12, **kw)
    if not norm and not np.iscomplexobj(x):
        want = want.real
    _close(got, want, tol if norm or not zero else tol * 20)  # debiasing subtracts the unscaled window SUM: cancellation


def test_pwznxcorr_scalar_lag_weights_stride_nopad_and_dims():
    import qups_b200
    from oracle.xcorr_np import pwznxcorr
    x = _data((300, 7, 2, 2), 3)
    w = np.hanning(9).astype(np.float32) + 0.1
    _close(qups_b200.pwznxcorr(x, 3, w, stride=2), pwznxcorr(x, 3, w, stride=2), 1e-5)
    _close(qups_b200.pwznxcorr(x, [4, -4], 6, pad=False, zero=False), pwznxcorr(x, [4, -4], 6, pad=False, zero=False), 1e-5)
    # time along dim 3, channels along dim 1, lags in the (singleton) 2nd dimension
    xt = np.ascontiguousarray(np.transpose(x[:, :, :1, :], (1, 2, 0, 3)))  # (N, 1, T, F)
    got = qups_b200.pwznxcorr(xt, [0, 1, 2], 5, tdim=3, ndim=1, ldim=2)
    want = np.transpose(pwznxcorr(x[:, :, 0, :], [0, 1, 2], 5), (1, 3, 0, 2))  # (T, N', F, L) -> (N', L, T, F)
    _close(got, want, 1e-5)
    # the default window W = 1 debiases every sample to exactly zero: 0/0 = NaN, as in the reference
    assert np.all(np.isnan(qups_b200.pwznxcorr(x[:64, :3, 0, 0], 1)))


def test_pwznxcorr_rejects_what_is_off_path():
    import qups_b200
    x = _data((64, 4), 4)
    with pytest.raises(qups_b200.QupsError):
        qups_b200.pwznxcorr(x, [0.5], 4)
    with pytest.raises(qups_b200.QupsError):
        qups_b200.pwznxcorr(x, [1], 4, 2)
    with pytest.raises(qups_b200.QupsError):
        qups_b200.pwznxcorr(x, [1], 4, multi=True)
    with pytest.raises(qups_b200.QupsError):
        qups_b200.pwznxcorr(x, [1], 4, stride=4)
