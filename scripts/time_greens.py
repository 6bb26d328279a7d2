"""Timing of the greens kernels at BASELINE config 5 scale on ONE GPU: 10 k scatterers, 256 x 256 elements (FSA)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qups_b200 import synth, ultrasound

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--S", type=int, default=10000)
ap.add_argument("--modes", default="conv,binned")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
P = synth.config_c2(64, 64, a.N, a.N, 2048)
rng = np.random.default_rng(1)
ps = np.stack([rng.uniform(-25e-3, 25e-3, a.S), np.zeros(a.S), rng.uniform(1e-3, 51e-3, a.S)], 0)
amp = rng.standard_normal(a.S)
pn = P.Pr
kern, wt0, wtend = synth.greens_kernel(7.5e6, 0.6, P.fs)
r = np.linalg.norm(ps[:, :, None] - pn[:, None, :], axis=0)
n0 = int(np.floor((2 * r.min() / P.c0 + wt0 - (wtend - wt0)) * P.fs))
T = int(np.ceil((2 * r.max() / P.c0 + wtend) * P.fs)) - n0 + 1
print(f"{a.S} scatterers, {a.N} x {a.N} elements, T = {T} output samples, K = {len(kern)} kernel samples", flush=True)
out = {}
for mode in a.modes.split(","):
    os.environ["QUPS_B200_GREENS"] = mode
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x = ultrasound.greens_raw(ps, amp, pn, pn, kern, n0, T, P.fs, P.c0, wt0, 1.0, 2e-4, "cubic")
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    out[mode] = x
    print(f"greens[{mode:6s}] {min(ts):9.2f} ms (median {sorted(ts)[len(ts)//2]:.2f})   {a.S*a.N*a.N/min(ts)/1e6:8.2f} G scatterer-rx-tx/s   out {x.numel()*8/1e9:.2f} GB", flush=True)
if len(out) == 2:
    k = list(out)
    d = (out[k[0]] - out[k[1]]).abs().max() / out[k[1]].abs().max()
    print(f"max |{k[0]} - {k[1]}| / max = {float(d):.2e}")
