"""Generates tests/golden/ext/*.npz — ORACLE-generated regression pins for the SURVEY §8f rows and greens (NOT reference
outputs: the reference is MATLAB and cannot run here).  Run from the repo root: python tests/golden/make_golden_ext.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_c, apod_np, prep_np, aperture_np, xcorr_np, refocus_np  # noqa: E402
from qups_b200 import synth  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ext")
f32 = np.float32


def apod():
    Pn = synth.linear_array(9, 0.3e-3)
    nn = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, 9))
    Pi = synth.scan_cartesian(np.linspace(-2.4e-3, 2.4e-3, 13), np.linspace(1e-3, 9e-3, 17))
    xv = np.linspace(-1.5e-3, 1.5e-3, 5)
    th = np.linspace(-10, 10, 5)
    out = dict(Pi=Pi.astype(f32), Pn=Pn.astype(f32), nn=nn.astype(f32), xv=xv, th=th,
               acc=apod_np.apAcceptanceAngle(Pi, Pn, nn, 30.0, literal=False), cos=apod_np.apCosineAngle(Pi, Pn, nn, 40.0, literal=False),
               grow=apod_np.apApertureGrowth(Pi, Pn, f=1.3, Dmax=2e-3, literal=False),
               scan=apod_np.apScanline(Pi, xv, 0.41e-3, literal=False),
               trans=apod_np.apTranslatingAperture(Pi, xv, Pn[0], (0.41e-3, 0.9e-3), literal=False),
               para=apod_np.apTxParallelogram(Pi, th, (-3.0, 3.0), (-1.2e-3, 1.2e-3), literal=False))
    np.savez_compressed(os.path.join(HERE, "apod.npz"), **out)


def prep():
    rng = np.random.default_rng(3)
    x = rng.integers(-1500, 1500, (100, 3, 2)).astype(np.int16)
    t0 = np.array([1.1e-6, 1.45e-6])
    y, t0p = prep_np.prep(x.astype(np.float64), t0, 20e6, B=5, A=23, hilbert=True, fmix=5e6)
    y2, _ = prep_np.prep(x.astype(np.float64), t0, 20e6, B=0, A=0, hilbert=True)           # L = 100: Bluestein on the device
    np.savez_compressed(os.path.join(HERE, "prep.npz"), x=x, t0=t0, fs=20e6, B=5, A=23, fmix=5e6, y=y, t0p=t0p, y_hilbert100=y2)


def aperture():
    rng = np.random.default_rng(5)
    b = (rng.standard_normal((6, 10, 4)) + 1j * rng.standard_normal((6, 10, 4))).astype(np.complex64)
    w, sf = aperture_np.pcf(b, 2, 0.8)
    np.savez_compressed(os.path.join(HERE, "aperture.npz"), b=b, cohfac=aperture_np.cohfac(b, 2), dmas3=aperture_np.dmas(b, 2, 3),
                        pcf_w=w, pcf_sf=sf, slsc_avg=aperture_np.slsc(b, 2, 3, "average"), slsc_ens=aperture_np.slsc(b, 2, 3, "ensemble"))


def greens():
    fs, fc, c0 = 20e6, 5e6, 1540.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    pn = synth.linear_array(4, 0.3e-3).astype(f32)
    pv = synth.linear_array(3, 0.4e-3).astype(f32)
    rng = np.random.default_rng(9)
    S = 200
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), rng.uniform(-1e-3, 1e-3, S), rng.uniform(3e-3, 11e-3, S)], 0).astype(f32)
    amp = rng.standard_normal(S).astype(f32)
    n0, T = 50, 700
    y32 = oracle_c.greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, 2e-4, "cubic")
    y64 = oracle_c.greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, 2e-4, "cubic", dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "greens.npz"), ps=ps, amp=amp, pn=pn, pv=pv, kern=kern, n0=n0, T=T, fs=fs, c0=c0, wt0=wt0,
                        R0=2e-4, y32=y32, y64=y64)


def xcorr():
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((160, 5, 2)) + 1j * rng.standard_normal((160, 5, 2))).astype(np.complex64)
    w = (np.hanning(7) + 0.2).astype(f32)
    np.savez_compressed(os.path.join(HERE, "xcorr.npz"), x=x, w=w, lags=np.array([-3, 0, 2]),
                        y_neighbor=xcorr_np.pwznxcorr(x, [-3, 0, 2], w), y_center_nonorm=xcorr_np.pwznxcorr(x, [-3, 0, 2], 6, ref="center", norm=False),
                        y_x0=xcorr_np.pwznxcorr(x, 2, 8, ref="x0", x0=x[:, 1:2], zero=False))


def refocus():
    rng = np.random.default_rng(13)
    T, N, V, fs, c0 = 64, 8, 6, 20e6, 1540.0
    xe = (np.arange(N) - (N - 1) / 2) * 0.3e-3
    th = np.deg2rad(np.linspace(-9, 9, V))
    tau = -(np.sin(th)[None, :] * xe[:, None]) / c0                   # plane-wave delays, elements x pulses
    apd = np.ones((N, V))
    x = (rng.standard_normal((T, N, V)) + 1j * rng.standard_normal((T, N, V))).astype(np.complex64)
    t0 = np.linspace(1e-6, 1.4e-6, V)
    out = dict(x=x, tau=tau, apd=apd, t0=t0, fs=fs, angles=np.rad2deg(th), c0=c0)
    for m in ("tikhonov", "adjoint"):
        y, t0m, Hi = refocus_np.refocus(x, t0, fs, tau, apd, m)
        out["y_" + m], out["Hi_" + m], out["t0_out"] = y, Hi, t0m
    np.savez_compressed(os.path.join(HERE, "refocus.npz"), **out)


if __name__ == "__main__":
    os.makedirs(HERE, exist_ok=True)
    apod(); prep(); aperture(); greens(); xcorr(); refocus()
    print("wrote", sorted(os.listdir(HERE)), sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE)), "bytes")
