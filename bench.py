#!/usr/bin/env python
"""bench.py — headline benchmark of the DAS hot path (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2-small|c1|c3f32]

A "step" = one delay-and-sum pass over the workload BASELINE.json's metric is quoted on:
C2 = 1024 x 1024 ScanCartesian, 256 focused transmits x 256 receives, T = 2048, fp32 complex, cubic.
Metric = DAS Mpixels/s (whole job).  N > 1: one process per GPU (torchrun), the pixel grid is sharded
along x (strong scaling: the image is fixed), the channel cube is replicated, no data-path collective.

Printed JSON line (rank 0): value = device-timed throughput with inputs resident in HBM; e2e = the same
metric through the C-ABI call with HOST buffers (qups_das_host: H2D + kernel + D2H inside the timed region);
roofline = algorithmic bytes (I*N*M*k*B_s + I*B_s) / CUDA-event kernel time vs the measured
HBM peak; cpu_baseline = the oracle port of kern/das_spec.m's CPU branch (oracle/qups_oracle.c, OpenMP) timed
on this box's host cores on a bounded sample.  --impl reference prints the CPU path as its own line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DAS Mpixels/s (1024^2, 256x256 tx/rx)"
UNIT = "Mpixels/s"


def workload(name):
    from qups_b200 import synth
    if name == "c2":
        return synth.config_c2()
    if name == "c2-small":
        return synth.config_c2(256, 256, 64, 64, 1024)
    if name == "c1":
        return synth.config_c1()
    if name == "c3f32":
        p = synth.config_c3()
        p.opts = ("plane-waves",)
        return p
    raise SystemExit(f"unknown workload {name}")


def workload_label(P, name):
    return (f"{name.upper()}: {P.Isz[0]}x{P.Isz[1]}x{P.Isz[2]} px, N={P.N} rx, M={P.M} tx, T={P.T}, "
            f"{P.interp} fp32 complex, scalar c0, apod=1")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port(P, x_np, budget_s=15.0, threads=None):
    """Time the oracle port of kern/das_spec.m:462-481 on the host cores, bounded sample: all pixels of a
    pixel subset x a transmit subset, scaled to the metric's unit."""
    from oracle import oracle_c
    if threads:
        oracle_c.set_num_threads(threads)
    cores = oracle_c.num_threads()
    kw = dict(VS="plane-waves" not in P.opts, DV="diverging-waves" in P.opts)
    nzs = P.Isz[0]
    msub = int(min(P.M, x_np.shape[2]))
    xs = np.asfortranarray(x_np[:, :, :msub])
    Pv = P.Pv[:, :msub] if P.Pv.shape[1] > 1 else P.Pv
    Nv = P.Nv[:, :msub] if P.Nv.shape[1] > 1 else P.Nv

    def run(nx_):
        Pi = np.ascontiguousarray(P.Pi[:, :, :nx_, :])
        t = time.perf_counter()
        oracle_c.das_spec("DAS", Pi, P.Pr, Pv, Nv, xs, P.t0, P.fs, P.c0, interp=P.interp, **kw)
        return time.perf_counter() - t

    nxs = min(P.Isz[1], 16)
    t1 = run(nxs)  # calibration pass (also warms the pages)
    pairs_per_s = nzs * nxs * P.N * msub / t1
    nxs = int(max(16, min(P.Isz[1], budget_s * pairs_per_s / (nzs * P.N * msub))))
    tt = run(nxs)
    pairs_per_s = nzs * nxs * P.N * msub / tt
    mpix = pairs_per_s / (P.N * P.M) / 1e6  # pixels/s for the full N x M aperture
    sample = (f"{nzs}x{nxs} px x {P.N} rx x {msub}/{P.M} tx of the workload in {tt:.2f} s "
              f"({pairs_per_s/1e6:.1f} M pixel-rx-tx pairs/s), scaled to all {P.N}x{P.M} pairs per pixel")
    return {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}, tt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shard", default="pixels", choices=["tx", "pixels"],
                    help="N>1 decomposition of the device-resident leg: pixel slabs without collective (default) or transmit "
                         "partition + all-reduce; the e2e leg always uses the transmit partition (1/N of the cube per GPU)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    P = workload(a.workload)
    cfg = {"workload": workload_label(P, a.workload), "sharding": "pixels along x (I2), cube replicated, no collective",
           "l2": "inputs (cube %.2f GB) larger than L2; no flush needed" % (P.T * P.N * P.M * 8 / 1e9)}

    # ---------------- reference arm: the reference's CPU path (oracle port), host cores ----------------
    if a.impl == "reference":
        if rank != 0:
            return
        from qups_b200 import synth
        x_np = synth.noise_cube(P.T, P.N, min(P.M, 16))
        vals, tts = [], []
        for _ in range(max(1, a.warmup if a.warmup < 2 else 1)):
            cpu_port(P, x_np, budget_s=2.0, threads=host_threads())
        for _ in range(max(1, a.steps)):
            cb, tt = cpu_port(P, x_np, budget_s=max(2.0, 60.0 / max(1, a.steps)), threads=host_threads())
            vals.append(cb["value"]); tts.append(tt)
        cb["value"] = float(np.mean(vals))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(tts)),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "reference CPU path = oracle port of kern/das_spec.m:462-481 (MATLAB absent); each step is a "
                        "bounded sample scaled to the full aperture"}
        print(json.dumps(line), flush=True)
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    import ctypes as C
    import qups_b200
    from qups_b200 import synth, _lib, shard

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f32 = np.float32

    # N > 1 (strong scaling over the fixed image), two one-step decompositions:
    #   tx (default): rank g holds x(:,:,m in M_g) — 1/N of the cube resident per GPU —, beamforms a full-size partial
    #                 image and the images are summed by ONE NCCL all-reduce of 2*I floats inside the timed step;
    #   pixels      : rank g beamforms a slab of x-columns from a replicated cube, no collective.
    tx_mode = world > 1 and a.shard == "tx"
    x_np = synth.noise_cube(P.T, P.N, P.M, seed=0)
    t = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).to(dev)
    if tx_mode:
        msel = np.arange(rank, P.M, world)  # interleaved transmits: every rank sees the same mix of geometries
        Pi_slab, Isz = P.Pi, P.Pi.shape[1:]
        Pv_l = np.broadcast_to(np.asarray(P.Pv, f32), (3, P.M))[:, msel]
        Nv_l = np.broadcast_to(np.asarray(P.Nv, f32), (3, P.M))[:, msel]
        x_d = torch.from_numpy(np.asfortranarray(x_np[:, :, msel])).to(dev)
        M_loc = len(msel)
        cfg["sharding"] = "transmits: each rank holds 1/N of the cube, full-size partial image, one NCCL all-reduce per step"
    else:
        Pi_slab, axis, s0, cnt = shard.pixel_shard(P.Pi, rank, world)
        Isz = Pi_slab.shape[1:]
        Pv_l, Nv_l, M_loc = P.Pv, P.Nv, P.M
        x_d = torch.from_numpy(x_np).to(dev)
    I_loc, I_tot = int(np.prod(Isz)), P.I
    args = (t(Pi_slab), t(P.Pr), t(Pv_l), t(Nv_l), x_d, float(P.t0), float(P.fs), float(P.c0), *P.opts, "interp", P.interp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Device-resident leg: every input is staged ONCE in the layout the C ABI takes (column-major, Pv4 row 4 = t0) and the
    # timed loop issues nothing but qups_das calls (+ the all-reduce in tx mode) — no per-step Python/torch conversions,
    # allocations or host synchronisation inside the timed region.
    from qups_b200 import kern
    L = _lib.lib()
    cm = lambda v: kern._colmajor(kern._mod_dim(kern._mod_size(v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v, f32)))), torch.float32, dev)
    dPi = kern._colmajor(args[0].reshape(3, *Isz), torch.float32, dev)
    dPr = cm(args[1])
    Pv_t = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(Pv_l, f32), (3, M_loc))))
    dPv4 = kern._colmajor(torch.cat([Pv_t, torch.full((1, M_loc), float(P.t0))], 0), torch.float32, dev)
    dNv = kern._colmajor(torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(Nv_l, f32), (3, M_loc)))), torch.float32, dev)
    dC = torch.tensor([np.float32(1.0) / np.float32(P.c0)], dtype=torch.float32, device=dev)
    dX = kern._cplx_buf(x_d.reshape(P.T, P.N, M_loc), "single", dev)
    yb = torch.empty(I_loc, dtype=torch.complex64, device=dev)
    dp = _lib.DasParams()
    dp.struct_size = C.sizeof(_lib.DasParams)
    dp.dtype = _lib.F32
    dp.I1, dp.I2, dp.I3 = (int(v) for v in Isz)
    dp.N, dp.M, dp.T, dp.F, dp.S = P.N, M_loc, P.T, 1, 0
    dp.flag = _lib.INTERP[P.interp]
    dp.vs, dp.dv = int("plane-waves" not in P.opts), int("diverging-waves" in P.opts)
    dp.fs = float(P.fs)
    acs0 = (C.c_uint64 * 6)(*([0] * 6))
    vpt = lambda tt: C.c_void_p(tt.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run_step():
        _lib.check(L.qups_das(C.byref(dp), vpt(yb), vpt(dPi), vpt(dPr), vpt(dPv4), vpt(dNv), None, vpt(dC), acs0, vpt(dX), stream))
        if tx_mode:  # the path's one exchange step: sum of the partial images (8 MB at 1024^2)
            return shard.allreduce_image(yb)
        return yb

    for _ in range(max(3, a.warmup)):
        y = run_step()
    kern_name = qups_b200.last_das_kernel()
    barrier()
    sampler = ClockSampler(local).start()  # every rank samples ITS GPU: a clock-locked / throttled peer must show up in the line
    time.sleep(0.25)
    _lib.launch_count(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    tw0 = time.time()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record()
    for k in range(a.steps):
        ev[k][0].record()
        y = run_step()
        ev[k][1].record()
    e_all1.record()
    barrier()
    tw1 = time.time()
    launches = _lib.launch_count()
    ms_total = e_all0.elapsed_time(e_all1)
    kern_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev]))
    if tx_mode:  # roofline wants the DAS kernel alone: re-time it without the all-reduce (outside the timed region)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for e0, e1 in kev:
            e0.record(); _lib.check(L.qups_das(C.byref(dp), vpt(yb), vpt(dPi), vpt(dPr), vpt(dPv4), vpt(dNv), None, vpt(dC), acs0, vpt(dX), stream)); e1.record()
        torch.cuda.synchronize()
        kern_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in kev]))
    tmax = torch.tensor([ms_total, kern_ms], device=dev, dtype=torch.float64)
    per_rank_ms = [kern_ms]
    if world > 1:
        gath = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gath, tmax[1:2].clone())
        per_rank_ms = [float(g[0]) for g in gath]   # a single slow GPU (clock lock, throttling) is visible in the line
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax[0]) / a.steps
    clocks = sampler.stop(tw0, tw1)
    if world > 1:  # report the slowest GPU of the job (median SM clock under load) and the union of the throttle reasons
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        meds = [c["sm_mhz"] for c in allc if c and c.get("sm_mhz")]
        clocks = {"sm_mhz": min(meds) if meds else None,
                  "sm_max_mhz": max((c["sm_max_mhz"] for c in allc if c and c.get("sm_max_mhz")), default=None),
                  "reasons": sorted(set(r for c in allc if c for r in c.get("reasons", []))),
                  "samples": sum(c.get("samples", 0) for c in allc if c), "per_gpu_sm_mhz": [c.get("sm_mhz") if c else None for c in allc]}
    ysum = float(torch.view_as_real(y).abs().sum())

    # ---------------- e2e: host buffers through the C ABI (H2D + kernel + D2H timed) ----------------
    e2e = None
    if not a.no_e2e and world > 1:
        # N > 1: transmit partition (SURVEY.md §8e mode 2).  Each rank uploads ONLY its 1/N of the channel cube from
        # pinned host memory, beamforms a full-size partial image and the images are summed with one NCCL all-reduce;
        # rank 0 reads the image back.  (Pixel sharding would make every rank upload the whole 1 GB cube.)
        msel = np.arange(rank, P.M, world)  # interleaved transmit shard (balanced across ranks)
        hX = torch.from_numpy(np.ascontiguousarray(x_np[:, :, msel].transpose(2, 1, 0))).pin_memory()
        Pvb = np.broadcast_to(np.asarray(P.Pv, f32), (3, P.M))[:, msel]
        Nvb = np.broadcast_to(np.asarray(P.Nv, f32), (3, P.M))[:, msel]
        hg = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v in (np.asarray(P.Pi, f32), np.asarray(P.Pr, f32), Pvb, Nvb)]
        hY = torch.empty((P.I,), dtype=torch.complex64).pin_memory()

        def e2e_step():
            xd = hX.to(dev, non_blocking=True).permute(2, 1, 0)
            gd = [h.to(dev, non_blocking=True) for h in hg]
            b = qups_b200.das_spec("DAS", gd[0], gd[1], gd[2], gd[3], xd, float(P.t0), float(P.fs), float(P.c0), *P.opts,
                                   "interp", P.interp)
            b = shard.allreduce_image(b.permute(*reversed(range(b.ndim))).contiguous().reshape(-1))
            if rank == 0:
                hY.copy_(b, non_blocking=True)
            torch.cuda.synchronize()
            return b
        ne = max(2, min(a.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / ne], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = hX.numel() * 8 + sum(h.numel() for h in hg) * 4
        full = float(torch.view_as_real(hY).abs().sum()) if rank == 0 else 0.0
        ysum_all = torch.tensor([ysum], device=dev, dtype=torch.float64)
        dist.all_reduce(ysum_all)
        e2e = {"value": I_tot / float(dt[0]) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(P.I * 8), "ms_per_step": 1e3 * float(dt[0]),
               "api": "qups_b200.das_spec on this rank's transmit shard (pinned host -> device) + NCCL all-reduce of the "
                      "partial images + image read-back on rank 0",
               "sharding": "transmits (each rank uploads 1/N of the cube)", "image_abs_sum": full}
    elif not a.no_e2e:
        L = _lib.lib()
        pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        hPi = pin(np.asarray(Pi_slab, f32).reshape(3, -1, order="F").T)
        hPr = pin(np.asarray(P.Pr, f32).T)
        Pv = np.broadcast_to(np.asarray(P.Pv, f32), (3, P.M))
        hPv = pin(np.concatenate([Pv, np.full((1, P.M), P.t0, f32)], 0).T)
        hNv = pin(np.broadcast_to(np.asarray(P.Nv, f32), (3, P.M)).T)
        hC = pin(np.array([f32(1) / f32(P.c0)], f32))
        hX = torch.from_numpy(x_np.transpose(2, 1, 0)).pin_memory()   # C-contiguous view of the column-major cube
        hY = torch.empty(I_loc, dtype=torch.complex64).pin_memory()
        p = _lib.DasParams()
        p.struct_size = C.sizeof(_lib.DasParams)
        p.dtype = _lib.F32
        p.I1, p.I2, p.I3 = Isz
        p.N, p.M, p.T, p.F, p.S = P.N, P.M, P.T, 1, 0
        p.flag = _lib.INTERP[P.interp]
        p.vs, p.dv = int("plane-waves" not in P.opts), int("diverging-waves" in P.opts)
        p.fs = float(P.fs)
        acs = (C.c_uint64 * 6)(*([0] * 6))
        vp = lambda tt: C.c_void_p(tt.data_ptr())

        def e2e_step():
            _lib.check(L.qups_das_host(C.byref(p), vp(hY), vp(hPi), vp(hPr), vp(hPv), vp(hNv), None, 0, vp(hC), 1, acs,
                                       vp(hX), local))
        ne = max(2, min(a.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / ne], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = hX.numel() * 8 + (hPi.numel() + hPr.numel() + hPv.numel() + hNv.numel() + 1) * 4
        e2e = {"value": I_tot / float(dt[0]) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(I_loc * 8), "ms_per_step": 1e3 * float(dt[0]),
               "api": "qups_das_host (C ABI, pinned host buffers)",
               "checksum_matches_device_path": bool(abs(float(torch.view_as_real(hY).abs().sum()) - ysum) <= 1e-3 * ysum)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    from qups_b200.synth import DasProblem
    loc = DasProblem(P.name, Pi_slab, P.Pr, np.zeros((3, M_loc)), np.zeros((3, M_loc)), P.T, P.fs, P.t0, P.c0, P.opts, P.interp)
    bytes_launch = loc.bytes_alg()
    achieved = bytes_launch / (float(tmax[1]) * 1e-3) / 1e9
    traffic, smem = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get(kern_name)
        # what actually bounds the kernel: shared-memory wavefronts per launch from the committed ncu capture,
        # scaled to this rank's share of the pixels, over the live kernel time, against 1 wavefront / clk / SM
        wf = tj.get(kern_name + "_smem_wavefronts")
        if wf and clocks and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            rate = wf * (I_loc / P.I) / (float(tmax[1]) * 1e-3)
            smem = {"bound": "shared-memory", "achieved": rate * 128 / 1e12, "peak": sms * clocks["sm_mhz"] * 1e6 * 128 / 1e12,
                    "unit": "TB/s", "frac": rate / (sms * clocks["sm_mhz"] * 1e6),
                    "source": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of the ncu capture (profiles/), live kernel time"}
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": I_tot / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kern_name, "kernel_ms": float(tmax[1]),
                     "kernel_ms_per_rank": per_rank_ms, "bytes_per_launch": bytes_launch, "peak_source": peak_src, "limiter": smem,
                     "note": "algorithmic (no-reuse gather) bytes per SURVEY.md §8d; neighbouring pixels share "
                             "samples so this legitimately exceeds 1 — the kernel is issue/LDS bound, see DESIGN.md §6"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "pairs_per_s": I_tot * P.N * P.M / (ms_step * 1e-3), "checksum": ysum,
    }
    if not a.no_cpu and world == 1:
        line["cpu_baseline"], _ = cpu_port(P, x_np[:, :, :min(P.M, 16)], budget_s=15.0, threads=host_threads())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
