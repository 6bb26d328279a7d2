"""prep_np.py — CPU restatement of the ChannelData pre-processing chain (TEST INFRASTRUCTURE ONLY; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).

    zeropad  src/ChannelData.m:1153-1183    hilbert  src/ChannelData.m:935-966 (MATLAB hilbert: fft, [1 2..2 1 0..0], ifft)
    downmix  src/ChannelData.m:757-807      time     src/ChannelData.m:1667

The FFT runs in float64 (the arbiter); the downmix phase follows the single-precision sequence the reference
executes on single data: t = fl32(t0' + fl32(j/fs)), theta = fl32(fl32(-2*pi*fc) * t).  Parity unpinned (no MATLAB here).
"""
import numpy as np

f32 = np.float32


def hilbert_weights(L):
    nd2 = L // 2
    w = np.zeros(L)
    w[0] = 1
    w[1:nd2] = 2
    if L > 1:
        w[nd2] = 1 + (L % 2)
    return w


def prep(x, t0, fs, B=0, A=0, hilbert=False, fmix=0.0):
    """x: T x N x M (real or complex); t0 scalar or M; returns (y complex64 (B+T+A) x N x M, t0')."""
    x = np.asarray(x)
    T = x.shape[0]
    L = B + T + A
    y = np.zeros((L,) + x.shape[1:], np.complex128)
    y[B:B + T] = x
    if hilbert:
        X = np.fft.fft(y.real, axis=0)
        y = np.fft.ifft(X * hilbert_weights(L).reshape((-1,) + (1,) * (x.ndim - 1)), axis=0)
    y = y.astype(np.complex64)
    t0 = np.asarray(t0, np.float64)
    t0p = (t0.astype(f32) - f32(B) / f32(fs)).astype(f32)
    if fmix:
        j = np.arange(L, dtype=f32)
        tt = (j / f32(fs)).astype(f32)
        t = (t0p.reshape((1, 1, -1)) + tt.reshape(-1, 1, 1)).astype(f32)        # L x 1 x (1|M)
        th = (f32(-2.0 * np.pi * fmix) * t).astype(f32)
        ph = (np.cos(th.astype(np.float64)) + 1j * np.sin(th.astype(np.float64))).astype(np.complex64)
        y = (y * ph).astype(np.complex64)
    return y, t0 - B / fs
