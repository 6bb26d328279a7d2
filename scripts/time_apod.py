"""Timing of DAS with closed-form apodization fused into the kernel vs the same mask passed as a dense array
(SURVEY.md §8f-1).  C2 geometry (1024^2, 256 x 256, cubic).  Not the bench contract (see bench.py)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200  # noqa: E402
from qups_b200 import synth, ultrasound as U  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nz", type=int, default=1024)
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--M", type=int, default=256)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
P = synth.config_c2(a.nz, a.nx, a.N, a.M, 2048, "cubic")
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
f32 = np.float32
dev = lambda v: torch.from_numpy(np.asarray(v, f32)).cuda()
Pi, Pr, Pv, Nv = dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv)
us = U.UltrasoundSystem(tx=P.Pr, rx=P.Pr, seq=U.Sequence("FC", P.Pv), scan=P.Pi, fs=P.fs)
pitch = float(abs(P.Pv[0, 1] - P.Pv[0, 0]))


def run(label, *apod):
    extra = sum((("apod", v) for v in apod), ())
    ms = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = qups_b200.das_spec("DAS", Pi, Pr, Pv, Nv, x, 0.0, P.fs, P.c0, *P.opts, "interp", "cubic", *extra)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"{label:58s} {min(ms):8.2f} ms  {P.I/min(ms)/1e3:7.2f} Mpix/s  kernel {qups_b200.last_das_kernel()}  |y| {float(y.abs().sum()):.6e}", flush=True)
    return y


run("apod = 1")
for name, spec in (("apAcceptanceAngle(45)", us.apAcceptanceAngle(45.0)), ("apApertureGrowth(f=1.5)", us.apApertureGrowth(1.5)),
                   ("apCosineAngle(45)", us.apCosineAngle(45.0)), ("apScanline(0.6 pitch)", us.apScanline(0.6 * pitch)),
                   ("apTranslatingAperture(0.6 pitch, 64 pitch)", us.apTranslatingAperture((0.6 * pitch, 64 * pitch))),
                   ("apScanline * apApertureGrowth", us.apScanline(0.6 * pitch).merged(us.apApertureGrowth(1.5)))):
    yf = run("FUSED " + name, spec)
    t = time.time()
    arrs = []
    if spec.rx_kind: arrs.append(spec.dense(Pi, Pr, which="rx"))
    if spec.tx_kind: arrs.append(spec.dense(Pi, M=P.M, which="tx"))
    torch.cuda.synchronize()
    gb = sum(v.numel() * 4 for v in arrs) / 1e9
    yd = run(f"dense array(s) {gb:.2f} GB (generated in {1e3*(time.time()-t):.0f} ms)", *arrs)
    print(f"    fused vs dense: max|diff|/max = {float((yf-yd).abs().max()/yd.abs().max()):.2e}", flush=True)
    del arrs, yd, yf
    torch.cuda.empty_cache()
