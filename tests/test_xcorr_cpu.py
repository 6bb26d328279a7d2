"""CPU checks of the pwznxcorr restatement (oracle/xcorr_np.py) against the pseudo-code of the reference's help text
(kern/pwznxcorr.m:9-20) evaluated with explicit loops, and against the properties the reference documents."""
import numpy as np
import pytest

from oracle.xcorr_np import pwznxcorr, conv_same


def _data(shape, seed=0, dtype=np.complex64):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)


def test_conv_same_is_matlab_same():
    rng = np.random.default_rng(1)
    z = rng.standard_normal(37)
    for W in (1, 2, 5, 8):
        w = rng.standard_normal(W)
        full = np.convolve(z, w, "full")
        assert np.allclose(conv_same(z, w), full[W // 2: W // 2 + len(z)])  # MATLAB: central part starting at floor(W/2)


@pytest.mark.parametrize("W", [3, 4])
def test_plain_windowed_inner_product_matches_help_text_loop(W):
    T, N = 48, 4
    x = _data((T, N), 2)
    lags = [-3, 0, 2]
    y = pwznxcorr(x, lags, W, zero=False, norm=False)
    P, c = 3, W // 2
    xp = np.concatenate([x, np.zeros((P, N), x.dtype)])
    ref = np.zeros((T, N - 1, len(lags)), np.complex128)
    for li, l in enumerate(lags):
        for n in range(N - 1):
            for t in range(T):
                for k in range(W):
                    s = t + c - k
                    if 0 <= s < T + P:
                        ref[t, n, li] += xp[s, n].astype(np.complex128) * np.conj(xp[(s + l) % (T + P), n + 1])
    assert np.max(np.abs(y - ref)) < 1e-5 * np.max(np.abs(ref))


def test_identical_channels_zero_lag_normalise_to_one():
    x = np.repeat(_data((96, 1), 3), 4, 1)
    y = pwznxcorr(x, [0], 9)
    assert np.max(np.abs(y[9:-9] - 1)) < 1e-5


def test_shifted_copy_peaks_at_its_lag():
    T = 200
    s = _data((T + 8,), 4)
    x = np.stack([s[4:T + 4], s[6:T + 6]], 1)  # channel 1 leads by 2 samples: x1[t] = x0[t + 2]
    y = pwznxcorr(x, 4, 16, zero=False)
    lag = np.arange(-4, 5)
    pk = lag[np.argmax(np.abs(y[40:160, 0, :]).mean(0))]
    assert pk == -2  # v = x(t + lag, n + 1) matches u = x(t, n) at lag = -2


def test_center_and_x0_references_and_shapes():
    x = _data((40, 5, 3), 5)
    yc = pwznxcorr(x, [0, 1], 6, ref="center")
    assert yc.shape == (40, 5, 3, 2)
    y0 = pwznxcorr(x, [0, 1], 6, ref="x0", x0=x[:, 2:3])  # N = 5: the median channel is channel 3 (1-based)
    assert np.allclose(yc, y0, atol=1e-6)
    x4 = _data((40, 4), 6)
    yc4 = pwznxcorr(x4, [1], 6, ref="center")
    y04 = pwznxcorr(x4, [1], 6, ref="x0", x0=x4[:, 1:3].mean(1, keepdims=True))
    assert np.allclose(yc4, y04, atol=1e-6)
    assert pwznxcorr(x, 2, 4, stride=2).shape == (40, 3, 3, 5)
