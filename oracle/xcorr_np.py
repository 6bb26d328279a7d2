"""xcorr_np.py — CPU restatement of kern/pwznxcorr.m (TEST INFRASTRUCTURE ONLY; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import it).

Pair-wise windowed zero-normalised cross-correlation, restated statement by statement from the reference's native branch
(`iflt = false`, i.e. no Image Processing Toolbox; integer lags; U = 1; `multi = false`), kern/pwznxcorr.m:142-262:

    :167-169  kernfun(z) = convn(z, w, 'same')         zero-padded moving sum along time with the weights w
    :175-177  integer lags sample by circshift(x, -l)  (circular over the PADDED length)
    :181-191  pad: P = ceil(max|lags|) zeros appended along time
    :194-214  reference channel: neighbor (n vs n+S) | center (mean of the median channel(s)) | x0
    :225-226  xlz = xl - kernfun(xl) (note: the window SUM is subtracted when W is a scalar — w = ones(W), :147-150;
              the reference does not scale the window, and neither does this restatement);  xln = kernfun(|xlz|^2)
    :229-260  per lag: xr_l = conj(shift(xr)); xrz_l = xr_l - kernfun(xr_l); y = kernfun(xlz .* xrz_l);
              y ./ (sqrt(xln) .* sqrt(kernfun(|xrz_l|^2)) .* sqrt(Wn)),  Wn = 1
    :266      the padding is cropped
Arithmetic stays in the precision of the data (MATLAB single stays single).  Parity unpinned (no MATLAB here); the
reference's own test only asserts a non-empty result (test/KernTest.m:267-268)."""
import numpy as np


def conv_same(z, w):
    """convn(z, w, 'same') along axis 0 for a weight vector w (MATLAB centring: same[t] = full[t + floor(W/2)])."""
    W = len(w)
    c = W // 2
    Tp = z.shape[0]
    out = np.zeros_like(z)
    for k in range(W):  # out[t] += w[k] * z[t + c - k]
        sh = c - k
        lo, hi = max(0, -sh), min(Tp, Tp - sh)
        if hi > lo:
            out[lo:hi] += w[k] * z[lo + sh:hi + sh]
    return out


def pwznxcorr(x, lags, W=None, *, zero=True, norm=True, ref="neighbor", stride=1, x0=None, pad=True):
    """x: (T, N, F...) real or complex, time first, channels second (tdim = 1, ndim = 2).  Returns (T, N', F..., L)."""
    x = np.asarray(x)
    rdt = np.float64 if x.dtype in (np.float64, np.complex128) else np.float32
    cdt = np.complex128 if rdt is np.float64 else np.complex64
    lags = np.atleast_1d(np.asarray(lags, dtype=np.float64))
    if not np.all(lags == np.floor(lags)):
        raise NotImplementedError("fractional lags (interpd branch) are not restated")
    if lags.size == 1:
        lags = np.arange(-lags[0], lags[0] + 1)           # :142-143
    lags = lags.astype(np.int64)
    if W is None:
        W = max(int(np.ceil(np.max(np.abs(lags)) / 2)), 1)  # default window (arguments block)
    w = np.ones(int(W), rdt) if np.ndim(W) == 0 else np.asarray(W, rdt).ravel()
    T, N = x.shape[0], x.shape[1]
    rest = x.shape[2:]
    xs = x.reshape(T, N, -1).astype(cdt)
    F = xs.shape[2]
    P = int(np.ceil(np.max(np.abs(lags)))) if pad else 0
    z0 = np.zeros((P, N, F), cdt)
    xp = np.concatenate([xs, z0], 0)
    Tp = T + P
    if ref == "neighbor":
        S = int(stride)
        xl, xr = xp[:, :N - S], xp[:, S:]
    elif ref == "center":
        mid = (N + 1 - 1 + 1) / 2                            # C = 1 channel (multi = false)
        n = sorted({int(np.floor(mid)), int(np.ceil(mid))})  # 1-based
        xl = xp
        xr = xp[:, [i - 1 for i in n]].mean(1, keepdims=True).astype(cdt)
    elif ref == "x0":
        x0 = np.asarray(x0)
        x0s = x0.reshape(x0.shape[0], x0.shape[1] if x0.ndim > 1 else 1, -1).astype(cdt)
        xl = xp
        xr = np.concatenate([x0s, np.zeros((P,) + x0s.shape[1:], cdt)], 0)
    else:
        raise ValueError(ref)
    kern = lambda z: conv_same(z, w)
    xlz = xl - kern(xl) if zero else xl
    xln = kern((xlz * np.conj(xlz)).real.astype(rdt)) if norm else None
    out = []
    for l in lags:
        xr_l = np.conj(np.roll(xr, -int(l), axis=0))
        xrz = xr_l - kern(xr_l) if zero else xr_l
        y = kern((xlz * xrz).astype(cdt))
        if norm:
            xrn = kern((xrz * np.conj(xrz)).real.astype(rdt))
            with np.errstate(invalid="ignore", divide="ignore"):
                r = (np.sqrt(xln) * np.sqrt(xrn)).astype(rdt)
                y = (y / r).astype(cdt)
        out.append(y[:T])
    y = np.stack(out, -1)                                    # (T, N', F, L)
    return y.reshape((T, y.shape[1]) + rest + (len(lags),))
