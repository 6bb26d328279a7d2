"""Quick device-side timing of one DAS configuration (not the bench contract; see bench.py)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200  # noqa: E402
from qups_b200 import synth, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nz", type=int, default=1024)
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--M", type=int, default=256)
ap.add_argument("--T", type=int, default=2048)
ap.add_argument("--interp", default="cubic")
ap.add_argument("--path", default="auto")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--config", default="c2")
ap.add_argument("--slab", default="", help="k/n: beamform only the k-th of n pixel slabs along x (what rank k of n GPUs does)")
a = ap.parse_args()

if a.config == "c5":
    P = synth.config_c5_das(a.nz, a.nx, a.N, a.T)
    P.interp = a.interp
else:
    P = {"c2": synth.config_c2, "c3": synth.config_c3}[a.config](a.nz, a.nx, a.N, a.M, a.T, a.interp)
if a.slab:
    from qups_b200 import shard
    k, n = (int(v) for v in a.slab.split("/"))
    P.Pi = np.ascontiguousarray(shard.pixel_shard(P.Pi, k, n)[0])
t = time.time()
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
print(f"data gen {time.time()-t:.1f}s", flush=True)
f32 = np.float32
dev = lambda v: torch.from_numpy(np.asarray(v, f32)).cuda()
args = (dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv), x, 0.0, P.fs, P.c0, *[o for o in P.opts if o != "modulation" and not isinstance(o, float)], "interp", P.interp)
pth = {"generic": _lib.PATH_GENERIC, "tiled": _lib.PATH_TILED, "auto": _lib.PATH_AUTO}[a.path]
for it in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = qups_b200.das_spec("DAS", *args, _path=pth)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pairs = P.I * P.N * P.M
    print(f"{a.config} {a.interp} {qups_b200.last_das_kernel()} iter {it}: {ms:.2f} ms  {P.I/ms/1e3:.3f} Mpix/s  "
          f"{pairs/ms/1e6:.1f} Gpair/s  alg {P.bytes_alg()/ms/1e6:.0f} GB/s  |y|max {float(y.abs().max()):.3f}", flush=True)
