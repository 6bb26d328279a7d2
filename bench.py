#!/usr/bin/env python
"""bench.py — benchmark of the DAS / greens hot path (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c1|c3|c3f32|c4|c5|c2-small|bfdas|greens] [--no-cpu] [--no-e2e] [--no-also]

A "step" = one pass of the hot path over the workload.  The default workload is the one BASELINE.json's metric is quoted
on, C2 = 1024 x 1024 ScanCartesian, 256 focused transmits x 256 receives, T = 2048, fp32 complex, cubic; metric =
DAS Mpixels/s (whole job).  The other BASELINE configs are selectable (`--workload`) and the default N = 1 run also reports a
short device-timed measurement of each of them in `also` (outside the timed region of the headline), so that one line
carries every config.

N > 1: one process per GPU (torchrun).  `value` leg: the pixel grid is sharded over the ranks in interleaved groups of
32 image columns (planes for volumes), the channel cube is replicated, no data-path collective (strong scaling: the image
is fixed).  `e2e` leg: the transmit axis is partitioned, each rank uploads 1/N of the cube from pinned host memory
through qups_das_host, and the partial images are summed by ONE NCCL all-reduce.

Printed JSON line (rank 0): value = device-timed throughput with inputs resident in HBM; e2e = the same metric through the
C ABI with HOST buffers (H2D + kernel + D2H inside the timed region); roofline = what bounds the dominant kernel
(shared-memory wavefronts for the staged kernel) with the SURVEY §8d algorithmic-HBM-bytes figure beside it (`hbm_model`);
cpu_baseline = the oracle port of kern/das_spec.m's CPU branch (oracle/qups_oracle.c, OpenMP) timed on this box's host
cores over the FULL workload; ref_kernel = the reference's own DASf kernel (oracle/_ref/bf.ptx, unmodified source, the
reference's flags and launch geometry) timed on the same GPU.  oracle/ is test infrastructure: it is only executed in those
two baseline legs and by --impl reference, never inside our timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
TAPS = {"nearest": 1, "linear": 2, "cubic": 4, "lanczos3": 4}


# ----------------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------------
def workload(name):
    """-> (DasProblem, dict(dtype 'f32'|'f16', fmod, kind 'das'|'bfdas'|'greens'|'c5'))."""
    from qups_b200 import synth
    o = {"dtype": "f32", "fmod": 0.0, "kind": "das"}
    if name == "c2":
        return synth.config_c2(), o
    if name == "c2-small":
        return synth.config_c2(256, 256, 64, 64, 1024), o
    if name == "c1":
        return synth.config_c1(), o
    if name == "c3":      # BASELINE config 3 as SURVEY §8d defines it: half2 IQ, fmod = fc (baseband), cubic
        p = synth.config_c3()
        o.update(dtype="f16", fmod=float(p.meta["fc"]))
        p.opts = ("plane-waves",)
        return p, o
    if name == "c3f32":
        p = synth.config_c3()
        p.opts = ("plane-waves",)
        return p, o
    if name == "c4":      # 256^3 voxels, 32 x 32 matrix array, 64 diverging waves
        return synth.config_c4(), o
    if name == "c5":      # greens (10 k scatterers, FSA 256 x 256) -> DAS round trip
        o["kind"] = "c5"
        return synth.config_c5_das(), o
    if name == "bfdas":   # C2 through the look-up-table path (bfDAS -> bfDASLUT -> sample2sep -> wsinterpd2)
        o["kind"] = "bfdas"
        return synth.config_c2(), o
    if name == "greens":
        o["kind"] = "greens"
        return synth.config_c5_das(), o
    raise SystemExit(f"unknown workload {name}")


def workload_label(P, name, o):
    return (f"{name.upper()}: {P.Isz[0]}x{P.Isz[1]}x{P.Isz[2]} px, N={P.N} rx, M={P.M} tx, T={P.T}, {P.interp} "
            f"{'half2' if o['dtype'] == 'f16' else 'fp32'} complex" + (f", fmod={o['fmod']/1e6:g} MHz" if o["fmod"] else "") +
            ", scalar c0, apod=1")


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of kern/das_spec.m:462-481 on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port(P, o, x_np, threads=None, full_budget_s=100.0, sample_s=15.0, force_sample=False):
    """Times the oracle port on the host cores.  After a short calibration pass the FULL workload (every pixel, receive and
    transmit) is run once when it is predicted to fit `full_budget_s`; otherwise a bounded sample (a column subset x a transmit
    subset of the same workload) is timed and scaled — and labelled as such.  Returns (cpu_baseline dict, seconds per full
    workload pass [measured or scaled])."""
    from oracle import oracle_c
    if threads:
        oracle_c.set_num_threads(threads)
    cores = oracle_c.num_threads()
    kw = dict(VS="plane-waves" not in P.opts, DV="diverging-waves" in P.opts)
    if o["fmod"]:
        kw["fmod"] = o["fmod"]
    nz, nxf = P.Isz[0], P.Isz[1] * P.Isz[2]
    Pi_cols = P.Pi.reshape(3, nz, nxf)
    M = P.M
    Pvb = np.broadcast_to(P.Pv, (3, M)) if P.Pv.shape[1] > 1 else P.Pv
    Nvb = np.broadcast_to(P.Nv, (3, M)) if P.Nv.shape[1] > 1 else P.Nv

    def run(ncols, msub):
        Pi = np.ascontiguousarray(Pi_cols[:, :, :ncols]).reshape(3, nz, ncols, 1)
        xs = x_np if msub == x_np.shape[2] else np.asfortranarray(x_np[:, :, :msub])
        Pv = Pvb[:, :msub] if Pvb.shape[1] > 1 else Pvb
        Nv = Nvb[:, :msub] if Nvb.shape[1] > 1 else Nvb
        t = time.perf_counter()
        oracle_c.das_spec("DAS", Pi, P.Pr, Pv, Nv, xs, P.t0, P.fs, P.c0, interp=P.interp, **kw)
        return time.perf_counter() - t

    mcal = int(min(M, 8))
    ccal = int(min(nxf, 16))
    t1 = run(ccal, mcal)  # calibration pass (also warms the pages)
    rate = nz * ccal * P.N * mcal / t1
    t_full_pred = P.I * P.N * M / rate
    if not force_sample and t_full_pred <= full_budget_s and x_np.shape[2] == M:
        tt = run(nxf, M)
        rate = P.I * P.N * M / tt
        sample = (f"FULL workload, one pass: {P.Isz[0]}x{P.Isz[1]}x{P.Isz[2]} px x {P.N} rx x {M} tx in {tt:.2f} s "
                  f"({rate/1e6:.1f} M pixel-rx-tx pairs/s), not extrapolated")
        t_full = tt
    else:
        msub = int(min(M, x_np.shape[2], 16))
        ncols = int(max(16, min(nxf, sample_s * rate / (nz * P.N * msub))))
        tt = run(ncols, msub)
        rate = nz * ncols * P.N * msub / tt
        t_full = P.I * P.N * M / rate
        sample = (f"bounded sample: {nz}x{ncols} px x {P.N} rx x {msub}/{M} tx of the workload in {tt:.2f} s "
                  f"({rate/1e6:.1f} M pixel-rx-tx pairs/s), SCALED to all {P.N}x{M} pairs per pixel (full pass predicted "
                  f"{t_full_pred:.0f} s)")
    mpix = P.I / t_full / 1e6
    return {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "full_workload_s": t_full}, t_full


def host_cube(P, o, M=None):
    """Synthetic channel cube on the host: complex64 (T,N,M) Fortran order; fp16 workloads round it to half precision."""
    from qups_b200 import synth
    x = synth.noise_cube(P.T, P.N, P.M if M is None else M, seed=0)
    if o["dtype"] == "f16":
        x = (x.real.astype(np.float16).astype(np.float32) + 1j * x.imag.astype(np.float16).astype(np.float32)).astype(np.complex64, order="F")
    return x


# ----------------------------------------------------------------------------------------------------------------------
# our arm: staged C-ABI calls
# ----------------------------------------------------------------------------------------------------------------------
class DasRunner:
    """Every input staged ONCE in the layout the C ABI takes (column-major, Pv4 row 4 = t0); step() issues nothing but the
    qups_das call — no per-step Python/torch conversions, allocations or host synchronisation inside a timed region."""

    def __init__(self, P, o, dev, Pi=None, msel=None, x_dev=None, x_np=None, t0=None):
        import torch
        from qups_b200 import kern, _lib
        f32 = np.float32
        self.P, self.o, self.dev = P, o, dev
        Pi = P.Pi if Pi is None else Pi
        Pi = Pi.reshape(Pi.shape + (1,) * (4 - Pi.ndim))
        self.Isz = tuple(int(v) for v in Pi.shape[1:])
        self.I = int(np.prod(self.Isz))
        M = P.M
        Pv = np.broadcast_to(np.asarray(P.Pv, f32), (3, M))
        Nv = np.broadcast_to(np.asarray(P.Nv, f32), (3, M))
        if msel is not None:
            Pv, Nv = Pv[:, msel], Nv[:, msel]
        self.M = Pv.shape[1]
        t0 = P.t0 if t0 is None else t0
        tt = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32)))
        self.dPi = kern._colmajor(tt(Pi), torch.float32, dev)
        self.dPr = kern._colmajor(tt(P.Pr), torch.float32, dev)
        self.dPv4 = kern._colmajor(torch.cat([tt(Pv), torch.full((1, self.M), float(t0))], 0), torch.float32, dev)
        self.dNv = kern._colmajor(tt(Nv), torch.float32, dev)
        self.dC = torch.tensor([f32(1.0) / f32(P.c0)], dtype=torch.float32, device=dev)
        if x_dev is None:
            xs = x_np if msel is None else np.asfortranarray(x_np[:, :, msel])
            x_dev = torch.from_numpy(np.ascontiguousarray(xs.transpose(2, 1, 0))).to(dev)   # memory = column-major T x N x M
        if o["dtype"] == "f16" and x_dev.dtype != torch.float16:
            x_dev = torch.view_as_real(x_dev).to(torch.float16).contiguous()
        self.dX = x_dev
        self.y = torch.empty(self.I, dtype=torch.complex64, device=dev)
        p = _lib.DasParams()
        p.struct_size = C.sizeof(_lib.DasParams)
        p.dtype = _lib.F16 if o["dtype"] == "f16" else _lib.F32
        p.y_f32 = 1
        p.I1, p.I2, p.I3 = self.Isz
        p.N, p.M, p.T, p.F, p.S = P.N, self.M, P.T, 1, 0
        p.flag = _lib.INTERP[P.interp]
        p.vs, p.dv = int("plane-waves" not in P.opts), int("diverging-waves" in P.opts)
        p.fs, p.fmod = float(P.fs), float(o["fmod"])
        if self.Isz[0] > 1 and self.Isz[1] > 1:   # launcher hints: the kernel never has to read the grid pitch back
            P0 = Pi[:, 0, 0, 0]
            p.pitch_hint[0] = float(np.linalg.norm(Pi[:, 1, 0, 0] - P0))
            p.pitch_hint[1] = float(np.linalg.norm(Pi[:, 0, 1, 0] - P0))
            p.c_hint = float(P.c0)
        self.p = p
        self.acs = (C.c_uint64 * 6)(*([0] * 6))
        self.L = _lib.lib()
        self.stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self._lib = _lib

    def step(self):
        vp = lambda t: C.c_void_p(t.data_ptr())
        self._lib.check(self.L.qups_das(C.byref(self.p), vp(self.y), vp(self.dPi), vp(self.dPr), vp(self.dPv4), vp(self.dNv), None,
                                        vp(self.dC), self.acs, vp(self.dX), self.stream))
        return self.y


def time_device(fn, warm, iters):
    """CUDA-event timing of fn() on torch's current stream (the stream every C-ABI call is issued on)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.min(ts))


def rand_cube(T, N, M, dev, half=False, seed=0):
    """Synthetic cube generated ON the device (for the `also` legs: no 1 GB host generation); memory = column-major T x N x M."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn((M, N, T, 2), generator=g, device=dev, dtype=torch.float32) * float(np.sqrt(0.5))
    x[:, :, :4] = 0
    x[:, :, T - 4:] = 0
    return x.to(torch.float16).contiguous() if half else torch.view_as_complex(x)


def greens_setup(P, nscat=10000, seed=1):
    from qups_b200 import synth
    fc, fs, c0 = P.meta["fc"], P.fs, P.c0
    rng = np.random.Generator(np.random.PCG64(seed))
    ps = np.stack([rng.uniform(-25e-3, 25e-3, nscat), np.zeros(nscat), rng.uniform(1e-3, 51e-3, nscat)], 0)
    amp = rng.standard_normal(nscat)
    kern, wt0, wtend = synth.greens_kernel(fc, 0.6, fs)
    r = np.linalg.norm(ps[:, :, None] - P.Pr[:, None, :], axis=0)
    n0 = int(np.floor((2 * r.min() / c0 + wt0 - (wtend - wt0)) * fs))
    T = int(np.ceil((2 * r.max() / c0 + wtend) * fs)) - n0 + 1
    return dict(ps=ps, amp=amp, kern=kern, wt0=wt0, n0=n0, T=T, fs=fs, c0=c0, R0=c0 / fc)


def run_greens(P, G, dev, pv=None):
    from qups_b200 import ultrasound
    return ultrasound.greens_raw(G["ps"], G["amp"], P.Pr, P.Pr if pv is None else pv, G["kern"], G["n0"], G["T"], G["fs"], G["c0"],
                                 G["wt0"], 1.0, G["R0"], "cubic", device=dev)


class LutRunner:
    """bfDAS at the C-ABI level: tau tables (in samples, as ChannelData.sample2sep passes them, src/ChannelData.m:1428-1445)
    built once on the device, step() = one qups_wsinterpd2 call summing both apertures."""

    def __init__(self, P, dev, x_dev):
        import torch
        from qups_b200 import _lib
        f32 = torch.float32
        I, N, M, T = P.I, P.N, P.M, P.T
        Pi = torch.from_numpy(np.ascontiguousarray(P.Pi.reshape(3, -1, order="F").T.astype(np.float32))).to(dev)   # I x 3
        Pr = torch.from_numpy(np.ascontiguousarray(P.Pr.T.astype(np.float32))).to(dev)
        Pv = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(P.Pv, (3, M)).T.astype(np.float32))).to(dev)
        Nv = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(P.Nv, (3, M)).T.astype(np.float32))).to(dev)
        self.t_rx = torch.empty((N, I), dtype=f32, device=dev)     # memory: I fastest (column-major I x N)
        self.t_tx = torch.empty((M, I), dtype=f32, device=dev)
        for n in range(N):
            self.t_rx[n] = (Pi - Pr[n]).norm(dim=1) * (P.fs / P.c0)
        VS, DV = "plane-waves" not in P.opts, "diverging-waves" in P.opts
        for m in range(M):
            rv = Pi - Pv[m]
            if not VS:
                d = rv @ Nv[m]
            else:
                d = rv.norm(dim=1)
                if not DV:
                    d = d * torch.sign(rv @ Nv[m])
            self.t_tx[m] = (d / P.c0 - P.t0) * P.fs
        self.x = x_dev
        self.y = torch.zeros(I, dtype=torch.complex64, device=dev)
        self.w = torch.ones(1, dtype=f32, device=dev)
        p = _lib.Ws2Params()
        p.struct_size = C.sizeof(_lib.Ws2Params)
        p.dtype, p.T, p.D, p.interp, p.w_real, p.omega = _lib.F32, T, 5, _lib.INTERP[P.interp], 1, 0.0
        # the MATLAB shapes of sample2sep with apdim = [4, 5]: tables I1 x I2 x I3 x N x 1 and I1 x I2 x I3 x 1 x M, x is T x 1 x 1 x N x M
        I1, I2, I3 = P.Isz
        sizes = [I1, I2, I3, N, M]
        #          w  y  t1 t2 x      t1 = transmit table (I x 1 x M), t2 = receive table (I x N)
        strides = [[0, 1, 1, 1, 0], [0, I1, I1, I1, 0], [0, I1 * I2, I1 * I2, I1 * I2, 0], [0, 0, 0, I, 1], [0, 0, I, 0, N]]
        strides = [st if sizes[k] > 1 else [0] * 5 for k, st in enumerate(strides)]
        for k in range(5):
            p.sizes[k] = sizes[k]
            for r in range(5):
                p.dstride[r + 5 * k] = strides[k][r]
        self.p, self.L, self._lib = p, _lib.lib(), _lib
        self.stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.table_bytes = (self.t_rx.numel() + self.t_tx.numel()) * 4

    def step(self):
        vp = lambda t: C.c_void_p(t.data_ptr())
        self._lib.check(self.L.qups_wsinterpd2(C.byref(self.p), vp(self.y), vp(self.w), vp(self.x), vp(self.t_tx), vp(self.t_rx),
                                               self.stream))
        return self.y


def also_legs(dev, peak, log):
    """Short device-timed measurements of the other BASELINE configs + the LUT path + greens + the reference's kernel
    (N = 1, after the headline's timed region).  Each entry: ms per step, the metric, the kernel that ran."""
    import torch
    import qups_b200
    from qups_b200 import _lib
    out = {}

    def das_leg(name, warm=2, iters=3):
        P, o = workload(name)
        x = rand_cube(P.T, P.N, P.M, dev, half=o["dtype"] == "f16")
        r = DasRunner(P, o, dev, x_dev=x)
        _lib.launch_count(reset=True)
        ms, mn = time_device(r.step, warm, iters)
        nl = _lib.launch_count() // (warm + iters)
        Bs = 4 if o["dtype"] == "f16" else 8
        alg = P.bytes_alg(Bs=Bs)
        res = {"workload": workload_label(P, name, o), "ms_per_step": ms, "ms_min": mn, "value": P.I / ms / 1e3, "unit": UNIT,
               "kernel": qups_b200.last_das_kernel(), "launches_per_step": int(nl),
               "pairs_per_s": P.I * P.N * P.M / (ms * 1e-3), "hbm_model": {"bytes_per_launch": alg, "achieved": alg / ms / 1e6,
                                                                           "unit": "GB/s", "frac": alg / ms / 1e6 / peak},
               "checksum": float(torch.view_as_real(r.y).abs().sum())}
        del r, x
        torch.cuda.empty_cache()
        return res

    for name, it in (("c1", 5), ("c3", 3), ("c4", 1)):
        try:
            out[name] = das_leg(name, warm=2 if it > 1 else 1, iters=it)
        except Exception as e:  # a failing side leg must not take the headline down; it is reported
            out[name] = {"error": repr(e)}
        log(f"also {name}: {out[name]}")
    # C5: greens (10 k scatterers, FSA 256 x 256) -> DAS round trip on one GPU
    try:
        P, o = workload("c5")
        G = greens_setup(P)
        gms, _ = time_device(lambda: run_greens(P, G, dev), 1, 2)
        xg = run_greens(P, G, dev)
        P5 = P
        P5.T = G["T"]
        r = DasRunner(P5, o, dev, x_dev=xg.permute(2, 1, 0).contiguous(), t0=G["n0"] / G["fs"])
        dms, _ = time_device(r.step, 1, 2)
        nsc = G["ps"].shape[1]
        out["c5"] = {"workload": f"C5: greens {nsc} scatterers, FSA {P.N}x{P.N} elements, T={G['T']} -> DAS {P.Isz[0]}x{P.Isz[1]} px cubic",
                     "greens_ms": gms, "das_ms": dms, "ms_per_step": gms + dms, "value": P.I / (gms + dms) / 1e3, "unit": UNIT,
                     "greens_G_scat_rx_tx_per_s": nsc * P.N * P.N / gms / 1e6, "das_kernel": qups_b200.last_das_kernel(),
                     "greens_out_GBps": G["T"] * P.N * P.N * 8 / gms / 1e6, "image_abs_max": float(r.y.abs().max())}
        del r, xg
        torch.cuda.empty_cache()
    except Exception as e:
        out["c5"] = {"error": repr(e)}
    log(f"also c5: {out['c5']}")
    # bfDAS at C2: the look-up-table path (2 x 1 GB delay tables + the cube): the HBM-bound DAS variant
    try:
        P, o = workload("bfdas")
        x = rand_cube(P.T, P.N, P.M, dev)
        r = LutRunner(P, dev, x)
        ms, _ = time_device(r.step, 1, 2)
        comp = r.table_bytes + P.T * P.N * P.M * 8 + P.I * 8
        out["bfdas"] = {"workload": workload_label(P, "bfdas", o) + " via qups_wsinterpd2 (tau_rx I x N, tau_tx I x M)",
                        "ms_per_step": ms, "value": P.I / ms / 1e3, "unit": UNIT, "compulsory_bytes": comp,
                        "compulsory_GBps": comp / ms / 1e6, "frac_of_hbm_peak": comp / ms / 1e6 / peak,
                        "kernel": qups_b200.last_ws2_kernel()}
        # same problem through qups_das on the same device data: the two paths must agree to tolerance
        d = DasRunner(P, {"dtype": "f32", "fmod": 0.0, "kind": "das"}, dev, x_dev=x)
        d.step()
        torch.cuda.synchronize()
        out["bfdas"]["rel_linf_vs_das"] = float((r.y - d.y).abs().max() / d.y.abs().max())
        del r, d, x
        torch.cuda.empty_cache()
    except Exception as e:
        out["bfdas"] = {"error": repr(e)}
    log(f"also bfdas: {out['bfdas']}")
    return out


def next_row_legs(dev, peak, log):
    """SURVEY.md §8f rows at the headline cube size, each through its C-ABI entry point on device-resident data (N = 1, after the
    headline's timed region): ChannelData pre-processing, aperture-domain reductions of the keep_rx cube, pwznxcorr, refocus.
    These ARE bound by HBM (one read + one write of the cube): frac = algorithmic bytes / time / measured HBM peak."""
    import torch
    from qups_b200 import _lib
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    out = {}
    T, N, M = 2048, 256, 256
    K = N * M

    def rec(name, fn, gbytes, extra=None, warm=1, iters=3):
        try:
            ms, mn = time_device(fn, warm, iters)
            out[name] = {"ms": mn, "algorithmic_GB": gbytes, "GBps": gbytes / mn * 1e3, "frac_of_hbm_peak": gbytes / mn * 1e3 / peak}
            if extra: out[name].update(extra)
        except Exception as e:
            out[name] = {"error": repr(e)}
        log(f"next-row {name}: {out[name]}")

    # ---- qups_das_cohfac at C2: DAS + coherence factor of the per-receive images in one pass (no 2.15 GB keep_rx cube) ----
    try:
        import qups_b200
        P2, o2 = workload("c2")
        x2 = rand_cube(P2.T, P2.N, P2.M, dev)
        r2 = DasRunner(P2, o2, dev, x_dev=x2)
        cf = torch.empty(r2.I, dtype=torch.float32, device=dev)
        fn = lambda: _lib.check(L.qups_das_cohfac(C.byref(r2.p), vp(r2.y), vp(cf), vp(r2.dPi), vp(r2.dPr), vp(r2.dPv4), vp(r2.dNv), vp(r2.dC),
                                                  vp(r2.dX), st))
        ms, mn = time_device(fn, 1, 3)
        out["das_cohfac_c2"] = {"ms": mn, "value": P2.I / mn / 1e3, "unit": UNIT, "kernel": qups_b200.last_das_kernel(),
                                "cf_mean": float(torch.nan_to_num(cf).mean()),
                                "note": "replaces DAS(keep_rx) -> cohfac + sum (SYN on the staged kernel + aperture pass)"}
        log(f"next-row das_cohfac_c2: {out['das_cohfac_c2']}")
        del r2, x2, cf
        torch.cuda.empty_cache()
    except Exception as e:
        out["das_cohfac_c2"] = {"error": repr(e)}
    # ---- qups_chd_prep (src/ChannelData.m zeropad / hilbert / downmix / cast) ----
    g = torch.Generator(device=dev); g.manual_seed(1)
    rf = torch.randn((K, T), generator=g, device=dev, dtype=torch.float32)
    yc = torch.empty((K, T), dtype=torch.complex64, device=dev)
    def prep(hilbert, fmix, inp, in_dtype):
        p = _lib.PrepParams()
        p.struct_size = C.sizeof(_lib.PrepParams)
        p.in_dtype, p.out_dtype, p.hilbert = in_dtype, _lib.F32, int(hilbert)
        p.T, p.K, p.B, p.A, p.traces_per_t0, p.n_t0, p.fs, p.fmix = T, K, 0, 0, N, 0, 30e6, fmix
        return lambda: _lib.check(L.qups_chd_prep(C.byref(p), vp(yc), vp(inp), None, st))
    rec("prep_hilbert_f32", prep(True, 0.0, rf, _lib.IN_REAL_F32), (4 + 8) * K * T / 1e9)
    rec("prep_hilbert_downmix_f32", prep(True, 7.5e6, rf, _lib.IN_REAL_F32), (4 + 8) * K * T / 1e9)
    rec("prep_cast_f32_to_complex", prep(False, 0.0, rf, _lib.IN_REAL_F32), (4 + 8) * K * T / 1e9)
    del rf
    # ---- qups_pwznxcorr (kern/pwznxcorr.m) on the complex cube: neighbouring channels, 9 lags, window 16 ----
    try:
        lags = (C.c_int32 * 9)(*range(-4, 5))
        w = torch.ones(16, dtype=torch.float32, device=dev)
        F = 32
        xin = yc[: N * F]                                        # T x N x F (time fastest)
        yx = torch.empty((9, F, N - 1, T), dtype=torch.complex64, device=dev)
        px = _lib.XcorrParams()
        px.struct_size = C.sizeof(_lib.XcorrParams)
        px.dtype, px.is_complex, px.ref, px.zero, px.norm, px.pad, px.stride = _lib.F32, 1, _lib.XC_NEIGHBOR, 1, 1, 1, 1
        px.L, px.W, px.T, px.N, px.F, px.x0N, px.x0F = 9, 16, T, N, F, 1, 1
        rec("pwznxcorr_9lags_w16", lambda: _lib.check(L.qups_pwznxcorr(C.byref(px), vp(yx), vp(xin), None, vp(w), lags, st)),
            (xin.numel() + yx.numel()) * 8 / 1e9, {"shape": f"T={T} N={N} F={F} lags=9 W=16"})
        del yx
    except Exception as e:
        out["pwznxcorr_9lags_w16"] = {"error": repr(e)}
    # ---- qups_refocus (REFoCUS decode): T x 128 x 128 data, 128 x 128 x T decoder ----
    try:
        Nr = Er = Vr = 128
        xr = yc[: Nr * Vr]
        Hi = torch.randn((T, Vr, Er, 2), generator=g, device=dev, dtype=torch.float32)
        yr = torch.empty((Er, Nr, T), dtype=torch.complex64, device=dev)
        pr = _lib.RefocusParams()
        pr.struct_size = C.sizeof(_lib.RefocusParams)
        pr.dtype, pr.T, pr.N, pr.V, pr.E, pr.n_t0, pr.fs = _lib.F32, T, Nr, Vr, Er, 1, 30e6
        t0c = (C.c_double * 1)(0.0)
        flop = 8.0 * T * Nr * Vr * Er
        rec("refocus_128x128x128", lambda: _lib.check(L.qups_refocus(C.byref(pr), vp(yr), vp(xr), vp(Hi), t0c, None, st)),
            (xr.numel() + yr.numel() + Hi.numel() // 2) * 8 / 1e9, {"decode_GFLOP": flop / 1e9})
        if "ms" in out["refocus_128x128x128"]:
            out["refocus_128x128x128"]["TFLOPs_fp32"] = flop / out["refocus_128x128x128"]["ms"] / 1e9
        del Hi, yr
    except Exception as e:
        out["refocus_128x128x128"] = {"error": repr(e)}
    del yc
    torch.cuda.empty_cache()
    # ---- qups_aperture on a keep_rx-sized cube: 1024^2 pixels x 256 receives (2.15 GB) ----
    try:
        Cn, A = 1024 * 1024, 256
        b = torch.randn((A, Cn, 2), generator=g, device=dev, dtype=torch.float32)
        o1 = torch.empty(Cn, dtype=torch.complex64, device=dev)
        for name, op, lagv in (("aperture_cohfac", _lib.APD_COHFAC, []), ("aperture_dmas_L16", _lib.APD_DMAS, list(range(1, 17))),
                               ("aperture_slsc_average_L16", _lib.APD_SLSC_AVERAGE, list(range(1, 17)))):
            pa = _lib.ApertureParams()
            pa.struct_size = C.sizeof(_lib.ApertureParams)
            pa.dtype, pa.op, pa.nlags, pa.C, pa.A, pa.S, pa.gamma = _lib.F32, op, len(lagv), Cn, A, 1, 1.0
            lg = (C.c_uint32 * max(1, len(lagv)))(*lagv)
            rec(name, (lambda pa=pa, lg=lg: _lib.check(L.qups_aperture(C.byref(pa), vp(o1), None, vp(b), lg, st))), Cn * A * 8 / 1e9)
        del b, o1
    except Exception as e:
        out["aperture"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def ref_kernel_leg(P, o, dev, x_dev, log):
    """The reference's own GPU kernel on this box (test infrastructure, oracle/ref_ptx.py): DASf from the unmodified
    src/bf.cu, compiled with the reference's flags, launched with the reference's geometry (kern/das_spec.m:301-306)."""
    try:
        import torch
        from oracle import ref_ptx
        if o["dtype"] != "f32" or not ref_ptx.available("bf", "fast"):
            return {"unavailable": "oracle/_ref/bf.ptx or cuda-python missing" if o["dtype"] == "f32" else "fp32 workloads only"}
        k = ref_ptx.RefDASf()
        # x_dev holds the column-major T x N x M cube; prepare() reads the logical shape (T, N, M) and takes the pointer as is
        k.prepare(P.Pi, P.Pr, P.Pv, P.Nv, x_dev.permute(2, 1, 0), P.t0, P.fs, P.c0, interp={"nearest": 0, "linear": 1, "cubic": 2}[P.interp],
                  VS="plane-waves" not in P.opts, DV="diverging-waves" in P.opts)
        ms, mn = time_device(k.launch, 1, 2)
        return {"kernel": "DASf (src/bf.cu unmodified, --use_fast_math, PTX compute_100 JIT)", "ms_per_step": ms, "value": P.I / ms / 1e3,
                "unit": UNIT, "grid": k.grid, "block": k.block,
                "note": "reference GPU semantics differ at trace ends and in cubic (DESIGN.md §3): timing only"}
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / ref_kernel legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short measurements of the other configs (N = 1 default run)")
    ap.add_argument("--cpu-sample", action="store_true", help="cpu_baseline: bounded sample instead of the full workload pass")
    ap.add_argument("--shard", default="pixels", choices=["tx", "pixels"],
                    help="N>1 decomposition of the device-resident leg: interleaved pixel groups without collective (default) or "
                         "transmit partition + all-reduce; the e2e leg always uses the transmit partition (1/N of the cube per GPU)")
    ap.add_argument("--contiguous", action="store_true", help="N>1 pixel sharding with contiguous slabs (round-1 behaviour)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    P, o = workload(a.workload)
    Bs = 4 if o["dtype"] == "f16" else 8
    log = lambda s: print(f"[bench r{rank}] {s}", file=sys.stderr, flush=True)
    cfg = {"workload": workload_label(P, a.workload, o),
           "sharding": "interleaved groups of 32 image columns (planes for volumes), cube replicated, no collective",
           "l2": "inputs (cube %.2f GB) larger than L2; no flush needed" % (P.T * P.N * P.M * Bs / 1e9)}
    metric = METRIC if a.workload in ("c2", "bfdas") else f"DAS Mpixels/s ({a.workload})"

    # ---------------- reference arm: the reference's CPU path (oracle port), host cores ----------------
    if a.impl == "reference":
        if rank != 0:
            return
        if o["kind"] not in ("das", "bfdas"):
            P, o = workload("c2")
        x_np = host_cube(P, o)
        # warm-up = the calibration passes; every timed step is one FULL pass over the workload when that fits the time box
        # (a few minutes in total), otherwise bounded samples scaled to the workload (labelled)
        cb, t_full = cpu_port(P, o, x_np, threads=host_threads(), full_budget_s=150.0, force_sample=a.cpu_sample)
        vals, tts = [cb["value"]], [t_full]
        budget = 150.0 - t_full
        while len(vals) < max(1, a.steps) and budget > t_full and "FULL" in cb["sample"]:
            cb2, t2 = cpu_port(P, o, x_np, threads=host_threads(), full_budget_s=1e9)
            vals.append(cb2["value"]); tts.append(t2); budget -= t2
        cb["value"] = float(np.mean(vals))
        cb["steps_timed"] = len(vals)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "steps_timed": len(vals), "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(tts)),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": o["dtype"], "data": "synthetic",
                "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "reference CPU path = oracle port of kern/das_spec.m:462-481 (MATLAB absent) on all host threads; "
                        "ms_per_step is the time of one pass over the whole workload; as many full passes as fit ~150 s are timed"}
        print(json.dumps(line), flush=True)
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    import qups_b200
    from qups_b200 import _lib, shard

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f32 = np.float32
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if o["kind"] in ("greens", "c5", "bfdas") and world > 1 and o["kind"] != "c5":
        raise SystemExit(f"--workload {a.workload} is a single-GPU leg (use c5 for the multi-GPU round trip)")

    # ------------------------------------------------------------------------------------------------
    # build the step for this workload
    # ------------------------------------------------------------------------------------------------
    x_np = None
    tx_mode = world > 1 and (a.shard == "tx" or o["kind"] == "c5")
    extra = {}
    if o["kind"] == "das":
        x_np = host_cube(P, o)
        if tx_mode:
            msel = np.arange(rank, P.M, world)  # interleaved transmits: every rank sees the same mix of geometries
            run = DasRunner(P, o, dev, msel=msel, x_np=x_np)
            cfg["sharding"] = "transmits: each rank holds 1/N of the cube, full-size partial image, one NCCL all-reduce per step"
        else:
            if world > 1 and not a.contiguous:
                Pi_loc, axis, idx = shard.pixel_shard_interleaved(P.Pi, rank, world)
            elif world > 1:
                Pi_loc, axis, s0, cnt = shard.pixel_shard(P.Pi, rank, world)
                cfg["sharding"] = "contiguous slabs of image columns, cube replicated, no collective"
            else:
                Pi_loc = P.Pi
            run = DasRunner(P, o, dev, Pi=np.ascontiguousarray(Pi_loc), x_np=x_np)
        I_loc, I_tot = run.I, P.I

        def run_step():
            y = run.step()
            return shard.allreduce_image(y) if tx_mode else y
        kern_only = run.step
        bytes_launch = run.I * P.N * run.M * TAPS[P.interp] * Bs + run.I * Bs
        pairs_launch = run.I * P.N * run.M
    elif o["kind"] == "bfdas":
        x = rand_cube(P.T, P.N, P.M, dev)
        run = LutRunner(P, dev, x)
        I_loc = I_tot = P.I
        run_step = kern_only = run.step
        bytes_launch = P.bytes_alg() + run.table_bytes
        pairs_launch = P.I * P.N * P.M
        cfg["data_path"] = "delay tables tau_rx (I x N) + tau_tx (I x M) fp32 = %.2f GB read per step" % (run.table_bytes / 1e9)
    elif o["kind"] in ("greens", "c5"):
        G = greens_setup(P)
        m0, mc = shard.tx_shard(P.N, rank, world)
        pv = P.Pr[:, m0:m0 + mc]
        P.T = G["T"]
        I_loc = I_tot = P.I
        state = {}

        def sim():
            state["x"] = run_greens(P, G, dev, pv)
            return state["x"]
        sim()
        Pd = P
        Pd.Pv, Pd.Nv = pv, np.tile(np.array([[0.0], [0.0], [1.0]]), (1, mc))
        run = DasRunner(Pd, o, dev, x_dev=state["x"].permute(2, 1, 0).contiguous(), t0=G["n0"] / G["fs"])
        if o["kind"] == "greens":
            run_step = kern_only = sim
            metric, cfg["workload"] = "greens G scatterer-rx-tx/s", f"greens: {G['ps'].shape[1]} scatterers, FSA {P.N}x{P.N} elements, T={G['T']}, cubic fp32"
            bytes_launch = G["T"] * P.N * mc * 8
            pairs_launch = G["ps"].shape[1] * P.N * mc
        else:
            def run_step():
                xg = sim()
                run.dX = xg.permute(2, 1, 0)   # greens writes the column-major T x N x M cube: a view, no copy
                y = run.step()
                return shard.allreduce_image(y)
            kern_only = run.step
            cfg["workload"] = (f"C5: greens {G['ps'].shape[1]} scatterers, FSA {P.N}x{P.N} elements, T={G['T']} -> DAS "
                               f"{P.Isz[0]}x{P.Isz[1]} px cubic, transmit partition + NCCL all-reduce")
            cfg["sharding"] = "transmits (each rank simulates and beamforms 1/N of the transmits), one all-reduce of the partial images"
            bytes_launch = P.I * P.N * mc * 4 * 8 + P.I * 8
            pairs_launch = P.I * P.N * mc
    else:
        raise SystemExit("unknown workload kind")

    W = max(3, a.warmup)
    for _ in range(W):
        y = run_step()
    kern_name = qups_b200.last_das_kernel() if o["kind"] in ("das", "c5") else ("greens" if o["kind"] == "greens" else qups_b200.last_ws2_kernel())
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
    if run_step is not kern_only:  # roofline wants the dominant kernel alone: re-time it without the collective / the simulation
        kern_ms, _ = time_device(kern_only, 0, 3)
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
    ysum = float(torch.view_as_real(y).abs().sum()) if y.is_complex() else float(y.abs().sum())

    # ---------------- e2e: host buffers through the C ABI (H2D + kernel + D2H timed) ----------------
    e2e = None
    if not a.no_e2e and o["kind"] == "das":
        # N = 1: the whole problem through qups_das_host.  N > 1: transmit partition (SURVEY.md §8e mode 2) — each rank uploads ONLY
        # its 1/N of the channel cube from pinned host memory through the same call (chunked copy/compute pipeline inside the
        # library), the partial image stays on the device (y_device), the images are summed with one NCCL all-reduce and rank 0
        # reads the sum back.  (Pixel sharding would make every rank upload the whole cube.)
        L = _lib.lib()
        pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        msel = np.arange(rank, P.M, world)
        Pi_e = P.Pi   # the e2e leg always beamforms the whole image (N > 1: from this rank's transmit shard)
        Isz_e = tuple(int(v) for v in (tuple(P.Pi.shape[1:]) + (1,) * (4 - P.Pi.ndim)))
        hPi = pin(np.asarray(Pi_e, f32).reshape(3, -1, order="F").T)
        hPr = pin(np.asarray(P.Pr, f32).T)
        Pvb = np.broadcast_to(np.asarray(P.Pv, f32), (3, P.M))[:, msel]
        Nvb = np.broadcast_to(np.asarray(P.Nv, f32), (3, P.M))[:, msel]
        hPv = pin(np.concatenate([Pvb, np.full((1, len(msel)), P.t0, f32)], 0).T)
        hNv = pin(Nvb.T)
        hC = pin(np.array([f32(1) / f32(P.c0)], f32))
        xs = x_np if world == 1 else x_np[:, :, msel]
        xt = torch.from_numpy(np.ascontiguousarray(xs.transpose(2, 1, 0)))       # memory = column-major T x N x M_loc
        hX = (torch.view_as_real(xt).to(torch.float16).contiguous() if o["dtype"] == "f16" else xt).pin_memory()
        I_e = int(np.prod(Isz_e))
        hY = torch.empty(I_e, dtype=torch.complex64).pin_memory()
        dYp = torch.empty(I_e, dtype=torch.complex64, device=dev)
        p = _lib.DasParams()
        C.memmove(C.byref(p), C.byref(run.p), C.sizeof(p))
        p.I1, p.I2, p.I3 = Isz_e
        p.M = len(msel)
        p.y_device = int(world > 1)
        acs = (C.c_uint64 * 6)(*([0] * 6))
        vp = lambda tt: C.c_void_p(tt.data_ptr())

        def e2e_step():
            _lib.check(L.qups_das_host(C.byref(p), vp(dYp) if world > 1 else vp(hY), vp(hPi), vp(hPr), vp(hPv), vp(hNv), None, 0,
                                       vp(hC), 1, acs, vp(hX), local))
            if world > 1:
                shard.allreduce_image(dYp)
                if rank == 0:
                    hY.copy_(dYp, non_blocking=True)
                torch.cuda.synchronize()
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
        h2d = hX.numel() * hX.element_size() + (hPi.numel() + hPr.numel() + hPv.numel() + hNv.numel() + 1) * 4
        full = float(torch.view_as_real(hY).abs().sum()) if rank == 0 else 0.0
        e2e = {"value": I_tot / float(dt[0]) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(I_e * 8), "ms_per_step": 1e3 * float(dt[0]),
               "api": "qups_das_host (C ABI, pinned host buffers)" + ("" if world == 1 else
                      " on this rank's transmit shard, partial image kept on the device + NCCL all-reduce + read-back on rank 0"),
               "image_abs_sum": full}
        if world == 1:
            e2e["checksum_matches_device_path"] = bool(abs(full - ysum) <= 1e-3 * ysum)
        else:
            e2e["sharding"] = "transmits (each rank uploads 1/N of the cube)"
        del hX, dYp

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline ----------------
    achieved = bytes_launch / (float(tmax[1]) * 1e-3) / 1e9
    hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "bytes_per_launch": int(bytes_launch), "peak_source": peak_src,
           "note": "SURVEY.md §8d algorithmic (no-reuse gather) bytes: I*N*M*k*B_s + I*B_s.  Neighbouring pixels share samples, so "
                   "this figure legitimately exceeds 1 and is NOT what bounds the kernel"}
    traffic, roof = None, None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    clk = (clocks or {}).get("sm_mhz") or 1965.0
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
    except Exception:
        tj = {}
    if kern_name.startswith("das_tiled") or kern_name == "ws2_tiled":
        # what bounds the staged kernel: shared-memory wavefronts.  C2 cubic: the count of the committed ncu capture
        # scaled to this rank's pixels; other workloads: the analytic count (k taps x LDS.64 = 2 wavefronts per 32 pairs), an
        # upper bound because pairs outside the data are skipped.
        key = "das_tiled_smem_wavefronts"
        if a.workload == "c2" and tj.get(key):
            wf, src = tj[key] * (I_loc / P.I), "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of the committed ncu capture (profiles/), scaled to this rank's pixels"
            traffic = tj.get("das_tiled")
        else:
            wf, src = pairs_launch * TAPS[P.interp] / 16.0, "analytic: pairs x taps x 2 wavefronts / 32 lanes (upper bound: skipped pairs not subtracted)"
        rate = wf / (float(tmax[1]) * 1e-3)
        roof = {"bound": "shared-memory", "achieved": rate * 128 / 1e12, "peak": sms * clk * 1e6 * 128 / 1e12, "unit": "TB/s",
                "frac": rate / (sms * clk * 1e6), "traffic": traffic, "wavefronts_per_launch": wf, "source": src,
                "peak_source": f"{sms} SMs x 128 B/clk x {clk:.0f} MHz (SM clock sampled during the timed region)"}
    else:
        roof = dict(hbm)
        roof["traffic"] = tj.get(kern_name)
    roof.update({"kernel": kern_name, "kernel_ms": float(tmax[1]), "kernel_ms_per_rank": per_rank_ms, "hbm_model": hbm})
    if o["kind"] == "greens":
        value = G["ps"].shape[1] * P.N * P.N / (ms_step * 1e-3) / 1e9
        unit = "G scatterer-rx-tx/s"
    else:
        value, unit = I_tot / (ms_step * 1e-3) / 1e6, UNIT
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": o["dtype"], "data": "synthetic", "config": cfg,
        "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "pairs_per_s": pairs_launch * world / (float(tmax[1]) * 1e-3),
        "checksum": ysum,
    }
    if not a.no_cpu and world == 1 and o["kind"] == "das":
        log("cpu baseline (full workload pass on the host cores) ...")
        line["cpu_baseline"], _ = cpu_port(P, o, x_np, threads=host_threads(), force_sample=a.cpu_sample)
        line["ref_kernel"] = ref_kernel_leg(P, o, dev, run.dX, log)
    if world == 1 and not a.no_also and a.workload == "c2":
        del run
        torch.cuda.empty_cache()
        line["also"] = also_legs(dev, peak, log)
        line["next_rows"] = next_row_legs(dev, peak, log)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
