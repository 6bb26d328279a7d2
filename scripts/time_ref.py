"""Times the REFERENCE's own DASf kernel (oracle/_ref/bf.ptx = src/bf.cu unmodified, --use_fast_math, compute_100
PTX JIT'd on the box) with the reference's launch geometry (kern/das_spec.m:301-306) on a DAS workload."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qups_b200 import synth  # noqa: E402
from oracle import ref_ptx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nz", type=int, default=1024)
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--M", type=int, default=256)
ap.add_argument("--T", type=int, default=2048)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
P = synth.config_c2(a.nz, a.nx, a.N, a.M, a.T)
x = synth.noise_cube(P.T, P.N, P.M)
k = ref_ptx.RefDASf()
k.prepare(P.Pi, P.Pr, P.Pv, P.Nv, x, 0.0, P.fs, P.c0, interp=2, VS=True, DV=False)
res = []
for it in range(a.iters + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k.launch()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
    print(f"reference DASf iter {it}: {res[-1]:.1f} ms  {P.I/res[-1]/1e3:.3f} Mpix/s  (block {k.block}, grid {k.grid})", flush=True)
print(json.dumps({"kernel": "reference DASf (src/bf.cu, compute_100 PTX, --use_fast_math)", "ms": min(res[1:]),
                  "mpix_s": P.I / min(res[1:]) / 1e3, "block": k.block, "grid": k.grid,
                  "workload": f"{a.nz}x{a.nx} px, N={a.N}, M={a.M}, T={a.T}, cubic fp32"}))
