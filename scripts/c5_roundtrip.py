"""BASELINE.json config 5: greens() point-scatterer simulation (FSA, N x N elements) -> DAS round trip with the
TRANSMIT axis partitioned over the ranks and one NCCL all-reduce of the partial images (SURVEY.md §8e mode 2).

    torchrun --nnodes=1 --nproc-per-node G scripts/c5_roundtrip.py [--scat 10000] [--nel 256] [--npx 1024]

Each rank simulates x(:,:,m in M_g) for its transmit shard (scatterers replicated), beamforms its shard into a
full-size partial image (diverging-wave / FSA delays) and the images are summed with all_reduce."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200  # noqa: E402
from qups_b200 import synth, shard, ultrasound  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scat", type=int, default=10000)
ap.add_argument("--nel", type=int, default=256)
ap.add_argument("--npx", type=int, default=1024)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

P = synth.config_c5_das(a.npx, a.npx, a.nel)
fc, fs, c0 = P.meta["fc"], P.fs, P.c0
rng = np.random.Generator(np.random.PCG64(1))
ps = np.stack([rng.uniform(-25e-3, 25e-3, a.scat), np.zeros(a.scat), rng.uniform(1e-3, 51e-3, a.scat)], 0)
amp = rng.standard_normal(a.scat)
m0, mc = shard.tx_shard(a.nel, rank, world)
pn = P.Pr
pv = P.Pr[:, m0:m0 + mc]
kern, wt0, wtend = synth.greens_kernel(fc, 0.6, fs)
r = np.linalg.norm(ps[:, :, None] - pn[:, None, :], axis=0)
n0 = int(np.floor((2 * r.min() / c0 + wt0 - (wtend - wt0)) * fs))
T = int(np.ceil((2 * r.max() / c0 + wtend) * fs)) - n0 + 1
f32 = np.float32
t = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).to(dev)
Pi, Pr, Pv, Nv = t(P.Pi), t(P.Pr), t(pv), t(np.tile(np.array([[0.0], [0.0], [1.0]]), (1, mc)))


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = []
for it in range(a.steps + 1):
    sync()
    t0 = time.perf_counter()
    x = ultrasound.greens_raw(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, c0 / fc, "cubic", device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    b = qups_b200.das_spec("DAS", Pi, Pr, Pv, Nv, x, n0 / fs, fs, c0, "diverging-waves", "interp", "cubic")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    b = shard.allreduce_image(b.contiguous())
    sync()
    t3 = time.perf_counter()
    res.append((t1 - t0, t2 - t1, t3 - t2))
if rank == 0:
    g, d, r_ = (min(v[i] for v in res[1:]) for i in range(3))
    img = b.abs()
    print(json.dumps({"config": "C5 greens->DAS, tx partition + all-reduce", "n_gpus": world, "scatterers": a.scat,
                      "elements": a.nel, "pixels": a.npx * a.npx, "T": T, "greens_ms": 1e3 * g, "das_ms": 1e3 * d,
                      "allreduce_ms": 1e3 * r_, "greens_triples_per_s": a.scat * a.nel * a.nel / g / world * world,
                      "das_mpix_s": a.npx * a.npx / d / 1e6, "image_max": float(img.max()), "kernel": qups_b200.last_das_kernel()}))
if world > 1:
    dist.destroy_process_group()
