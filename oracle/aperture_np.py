"""aperture_np.py — CPU restatement of the aperture-domain post-processing functions (TEST INFRASTRUCTURE ONLY; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).  Whole-array NumPy float64 versions of
kern/cohfac.m, kern/dmas.m, kern/pcf.m and kern/slsc.m (native branch, kdim singleton), written as the reference's
expressions.  `dim` is 1-based as in MATLAB.  Parity unpinned against MATLAB (absent here); pinned against an independent scalar-loop
restatement of the same reference statements in tests/test_aperture_cpu.py."""
import numpy as np


def cohfac(b, dim):
    b = np.asarray(b)
    ax = dim - 1
    return np.abs(b.sum(ax, keepdims=True)) ** 2 / (np.abs(b) ** 2).sum(ax, keepdims=True) / b.shape[ax]


def dmas(bn, dim, L=None):
    bn = np.asarray(bn)
    ax = dim - 1
    N = bn.shape[ax]
    if L is None:
        lags = range(1, N)
    elif np.ndim(L) == 0:
        lags = range(1, int(L) + 1)
    else:
        lags = sorted(set(range(1, N)) & set(int(v) for v in np.ravel(L)))
    b = 0
    for i in lags:
        if N - i <= 0:
            continue
        b = b + (np.take(bn, range(0, N - i), ax) * np.take(bn, range(i, N), ax)).sum(ax, keepdims=True)
    b = b + np.zeros(bn.sum(ax, keepdims=True).shape, bn.dtype)
    return np.exp(1j * np.angle(b)) * np.sqrt(np.abs(b))


def pcf(b, dim, gamma=1.0):
    b = np.asarray(b)
    ax = dim - 1
    phi = np.angle(b)
    s0 = np.nanstd(phi, axis=ax, keepdims=True)          # std(phi, 1, dim, "omitnan")
    sa = np.nanstd(phi - np.pi * np.sign(phi), axis=ax, keepdims=True)
    sf = np.fmin(s0, sa)
    return np.maximum(0, 1 - (gamma / np.sqrt(np.pi / 3)) * sf), sf


def slsc(x, dim, L=None, method="average"):
    x = np.asarray(x, np.complex128)
    ax = dim - 1
    A = x.shape[ax]
    L = max(1, A // 4) if L is None else L
    lags = list(range(1, int(L) + 1)) if np.ndim(L) == 0 else [int(v) for v in np.ravel(L)]
    m, n = np.meshgrid(np.arange(A), np.arange(A), indexing="ij")
    H = np.abs(m - n)
    S = np.isin(H, lags)
    nl = len(lags)
    xm = np.moveaxis(x, ax, -1)                              # ... x A
    if method == "average":
        with np.errstate(invalid="ignore", divide="ignore"):
            xn = np.nan_to_num(xm / np.abs(xm))
        W = S / (A - H) / 2 / nl
        z = np.einsum("...i,ij,...j->...", np.conj(xn), W, xn)
    else:
        z = np.einsum("...i,ij,...j->...", np.conj(xm), S.astype(float), xm)
        a = np.einsum("ij,...j->...", S.astype(float), np.abs(xm) ** 2)
        b = np.einsum("ij,...i->...", S.astype(float), np.abs(xm) ** 2)
        with np.errstate(invalid="ignore", divide="ignore"):
            z = z * np.nan_to_num(1 / np.sqrt(a) / np.sqrt(b), posinf=0.0)
    return np.moveaxis(z[..., None], -1, ax)
