"""Shared helpers for the parity tests."""
import numpy as np

from qups_b200 import synth


def small_problem(kind="FC", nz=24, nx=20, ny=1, N=12, M=5, T=160, seed=0, fs=20e6, t0=None, int_data=False, F=1,
                  pitch=0.3e-3, zlim=(2e-3, 9e-3), pad=4):
    """A small DAS problem whose delays stay inside the T-sample traces (except where a test wants edges)."""
    rng = np.random.default_rng(seed)
    Pr = synth.linear_array(N, pitch)
    xs = np.linspace(-2e-3, 2e-3, nx)
    ys = np.linspace(-0.5e-3, 0.5e-3, ny) if ny > 1 else (0.0,)
    Pi = synth.scan_cartesian(xs, np.linspace(*zlim, nz), ys)
    opts = ()
    if kind == "PW":
        th = np.deg2rad(np.linspace(-10, 10, M))
        Pv = np.zeros((3, 1))
        Nv = np.stack([np.sin(th), 0 * th, np.cos(th)], 0)
        opts = ("plane-waves",)
    elif kind == "FSA":
        Pv = synth.linear_array(M, pitch * N / M)
        Nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, M))
        opts = ("diverging-waves",)
    elif kind == "DV":
        Pv = np.stack([np.linspace(-1e-3, 1e-3, M), np.zeros(M), np.full(M, -3e-3)], 0)
        Nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, M))
        opts = ("diverging-waves",)
    else:  # FC: focused, foci inside the image so the sign flips across the tile
        Pv = np.stack([np.linspace(-1e-3, 1e-3, M), np.zeros(M), np.full(M, 5e-3)], 0)
        Nv = Pv / np.linalg.norm(Pv, 2)
    if int_data:
        x = (rng.integers(-8, 9, (T, N, M, F)) + 1j * rng.integers(-8, 9, (T, N, M, F))).astype(np.complex64)
    else:
        x = (rng.standard_normal((T, N, M, F)) + 1j * rng.standard_normal((T, N, M, F))).astype(np.complex64)
    if pad:
        x[:pad] = 0
        x[T - pad:] = 0
    if F == 1:
        x = x[..., 0]
    x = np.asfortranarray(x)
    if t0 is None:
        t0 = 0.0
    return dict(Pi=Pi, Pr=Pr, Pv=Pv, Nv=Nv, x=x, t0=t0, fs=fs, c=1540.0, opts=opts)


def oracle_kwargs(opts):
    kw = dict(VS=True, DV=False)
    if "plane-waves" in opts:
        kw["VS"] = False
    if "diverging-waves" in opts:
        kw["DV"] = True
    return kw


def rel_linf(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
