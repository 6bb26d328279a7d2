"""Timing of the non-headline kernels: generic DAS (SYN with a dense apodization), greens, wsinterpd2 (bfDAS)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200
from qups_b200 import synth, _lib, ultrasound

def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

f32 = np.float32
dev = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).cuda()
P = synth.config_c2(256, 256, 64, 64, 1024)
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
g = (dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv))
pairs = P.I * P.N * P.M
t = ev_time(lambda: qups_b200.das_spec("DAS", *g, x, 0.0, P.fs, P.c0, "interp", "cubic"))
print(f"tiled   DAS 256x256 px 64x64: {t:.2f} ms  {pairs/t/1e6:.1f} Gpair/s")
t = ev_time(lambda: qups_b200.das_spec("DAS", *g, x, 0.0, P.fs, P.c0, "interp", "cubic", _path=_lib.PATH_GENERIC))
print(f"generic DAS 256x256 px 64x64: {t:.2f} ms  {pairs/t/1e6:.1f} Gpair/s")
apod = torch.rand((256, 256, 1, 64, 1), device="cuda")
t = ev_time(lambda: qups_b200.das_spec("DAS", *g, x, 0.0, P.fs, P.c0, "interp", "cubic", "apod", apod))
print(f"generic DAS + dense IxN real apod: {t:.2f} ms  {pairs/t/1e6:.1f} Gpair/s  ({qups_b200.last_das_kernel()})")
t = ev_time(lambda: qups_b200.das_spec("SYN", *g, x, 0.0, P.fs, P.c0, "interp", "cubic"))
print(f"generic SYN: {t:.2f} ms  {pairs/t/1e6:.1f} Gpair/s")
# greens: 10k scatterers, 64 x 64 elements
rng = np.random.default_rng(1)
S = 10000
ps = np.stack([rng.uniform(-25e-3, 25e-3, S), np.zeros(S), rng.uniform(1e-3, 51e-3, S)], 0)
amp = rng.standard_normal(S)
pn = synth.linear_array(64, 0.2e-3 * 4)
kern, wt0, wtend = synth.greens_kernel(7.5e6, 0.6, P.fs)
r = np.linalg.norm(ps[:, :, None] - pn[:, None, :], axis=0)
n0 = int(np.floor((2 * r.min() / P.c0 + wt0 - (wtend - wt0)) * P.fs)); T = int(np.ceil((2 * r.max() / P.c0 + wtend) * P.fs)) - n0 + 1
t = ev_time(lambda: ultrasound.greens_raw(ps, amp, pn, pn, kern, n0, T, P.fs, P.c0, wt0, 1.0, 2e-4, "cubic"), 2)
print(f"greens 10k scat 64x64 el T={T} K={len(kern)}: {t:.1f} ms  {S*64*64/t/1e6:.2f} G scat-rx-tx/s")

# headline geometry with a dense I x N real apodization (1 GB): the staged kernel with NAP = 1
P = synth.config_c2()
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
g = (dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv))
apod = torch.rand((1024, 1024, 1, 256, 1), device="cuda")
t = ev_time(lambda: qups_b200.das_spec("DAS", *g, x, 0.0, P.fs, P.c0, "interp", "cubic", "apod", apod), 2)
print(f"C2 + dense IxN real apod: {t:.1f} ms {P.I/t/1e3:.2f} Mpix/s ({qups_b200.last_das_kernel()})")
apod2 = torch.rand((1, 1, 1, 256, 256), device="cuda")
t = ev_time(lambda: qups_b200.das_spec("DAS", *g, x, 0.0, P.fs, P.c0, "interp", "cubic", "apod", apod2), 2)
print(f"C2 + N x M real apod: {t:.1f} ms {P.I/t/1e3:.2f} Mpix/s ({qups_b200.last_das_kernel()})")
