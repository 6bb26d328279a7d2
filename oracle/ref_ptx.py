"""ref_ptx.py — launches the REFERENCE's own CUDA kernels (oracle/_ref/*.ptx) on the GPU box.

TEST INFRASTRUCTURE ONLY.  oracle/_ref/{bf,interpd,greens}.ptx are src/{bf,interpd,greens}.cu of the reference compiled
UNMODIFIED with the reference's flags (`nvcc --ptx -arch=compute_100 --use_fast_math`, oracle/Makefile); the `*_ieee.ptx`
files are the same sources with IEEE flags instead (no --use_fast_math, -fmad=false).  This module does what the reference's
MATLAB launchers do with parallel.gpu.CUDAKernel — load the PTX, set the QUPS_* __constant__ symbols, pick the reference's
launch geometry, call the kernel by its base name:

    DASf          kern/das_spec.m:284-373
    wsinterpd2f   kern/wsinterpd2.m:193-235
    greensf       src/UltrasoundSystem.m:649-718
    convf/convcf  kern/convd.m:135-201

It is used to (i) pin the oracle against the real reference kernels (tests/test_gpu_reference_kernel.py) and (ii) time
"the reference's kernel on this box" beside ours (bench.py `ref_kernel`).  Never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
BF_PTX = os.path.join(_HERE, "_ref", "bf.ptx")


def ptx_path(unit: str, variant: str = "fast") -> str:
    return os.path.join(_HERE, "_ref", unit + ("" if variant == "fast" else "_" + variant) + ".ptx")


def available(unit: str = "bf", variant: str = "fast") -> bool:
    try:
        from cuda.bindings import driver  # noqa: F401
    except Exception:
        return False
    return os.path.exists(ptx_path(unit, variant)) and torch.cuda.is_available()


def _chk(res):
    from cuda.bindings import driver
    err = res[0]
    if err != driver.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _chk_none(res):
    from cuda.bindings import driver
    err = res[0] if isinstance(res, tuple) else res
    if err != driver.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")


class RefModule:
    """One reference PTX module + one entry point resolved by base name (what CUDAKernel(ptx, cu, name) does)."""

    def __init__(self, unit: str, entry: str, variant: str = "fast"):
        from cuda.bindings import driver
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make torch's primary context current
        with open(ptx_path(unit, variant), "rb") as f:
            ptx = f.read() + b"\0"
        self.drv = driver
        self.mod = _chk(driver.cuModuleLoadData(ptx))
        found = None
        for mname, n, rest in re.findall(rb"\.entry\s+(_Z(\d+)([A-Za-z0-9_]+))\s*\(", ptx):
            if int(n) == len(entry) and rest[:int(n)] == entry.encode():
                found = mname
        if found is None:
            raise RuntimeError(f"{entry} entry not found in {ptx_path(unit, variant)}")
        self.fn = _chk(driver.cuModuleGetFunction(self.mod, found))
        self.max_threads = int(_chk(driver.cuFuncGetAttribute(
            driver.CUfunction_attribute.CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, self.fn)))

    def set_const(self, name: str, value, ctype=C.c_uint64):
        dptr, size = _chk(self.drv.cuModuleGetGlobal(self.mod, name.encode()))
        v = ctype(value)
        assert C.sizeof(v) == size, (name, size)
        _chk_none(self.drv.cuMemcpyHtoD(dptr, C.addressof(v), size))

    def launch(self, grid, block, args, stream=None):
        """args: torch tensors (passed as device pointers) or ctypes scalars (passed by value)."""
        holders = [C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a for a in args]
        argv = (C.c_void_p * len(holders))(*[C.addressof(h) for h in holders])
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        g, b = tuple(grid) + (1,) * (3 - len(grid)), tuple(block) + (1,) * (3 - len(block))
        _chk_none(self.drv.cuLaunchKernel(self.fn, int(g[0]), int(g[1]), int(g[2]), int(b[0]), int(b[1]), int(b[2]), 0, st,
                                          C.addressof(argv), 0))


class RefDASf:
    """The reference DASf kernel (src/bf.cu:154-161) behind the argument list of kern/das_spec.m:372."""

    def __init__(self, variant: str = "fast", double: bool = False):
        """double=True launches `DAS` (src/bf.cu:144-151), the fp64 instantiation of the same template."""
        self.k = RefModule("bf", "DAS" if double else "DASf", variant)
        self.max_threads = self.k.max_threads
        self.rdt, self.cdt = (np.float64, np.complex128) if double else (np.float32, np.complex64)

    def prepare(self, Pi, Pr, Pv, Nv, x, t0, fs, c, interp=2, VS=True, DV=False, fmod=0.0):
        """Device-side argument pack; Pi (3,I1,I2,I3), x (T,N,M) complex64 (Fortran order), fp32 everything."""
        dev = "cuda"
        f32, c64 = self.rdt, self.cdt
        tdt, tct = (torch.float64, torch.complex128) if f32 is np.float64 else (torch.float32, torch.complex64)
        Isz = tuple(Pi.shape[1:]) + (1,) * (4 - Pi.ndim)
        I = int(np.prod(Isz))
        T, N, M = x.shape
        col = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dt).T)).to(dev)
        self.I, self.N, self.M, self.T = I, N, M, T
        a = {}
        a["Pi"] = torch.from_numpy(np.ascontiguousarray(np.asarray(Pi, f32).reshape(3, I, order="F").T)).to(dev)
        a["Pr"] = col(Pr, f32)
        Pv4 = np.concatenate([np.broadcast_to(np.asarray(Pv, f32), (3, M)),
                              np.broadcast_to(np.asarray(t0, f32).reshape(1, -1), (1, M))], 0)
        a["Pv"] = col(Pv4, f32)
        a["Nv"] = col(np.broadcast_to(np.asarray(Nv, f32), (3, M)), f32)
        a["apod"] = torch.tensor([[1.0, 0.0]], dtype=tdt, device=dev)   # apod = {1}, complex
        a["cinv"] = torch.tensor([f32(1) / f32(c)], dtype=tdt, device=dev)
        a["strides"] = torch.zeros(12, dtype=torch.int64, device=dev)            # scalar cinv / scalar apod
        if isinstance(x, torch.Tensor):
            a["x"] = x
        else:
            a["x"] = torch.from_numpy(np.ascontiguousarray(np.asarray(x, c64).transpose(2, 1, 0))).to(dev)
        a["fsfc"] = torch.tensor([fs, fmod], dtype=tdt, device=dev)
        a["y"] = torch.zeros(I, dtype=tct, device=dev)
        for nm, v in (("QUPS_I", I), ("QUPS_T", T), ("QUPS_M", M), ("QUPS_N", N), ("QUPS_I1", Isz[0]),
                      ("QUPS_I2", Isz[1]), ("QUPS_I3", Isz[2]), ("QUPS_S", 1)):
            self.k.set_const(nm, v)
        self.k.set_const("QUPS_VS", bool(VS), C.c_bool)
        self.k.set_const("QUPS_DV", bool(DV), C.c_bool)
        self.k.set_const("QUPS_BF_FLAG", int(interp), C.c_int32)
        self.args = a
        # launch geometry of kern/das_spec.m:301-306
        self.block = int(self.max_threads)
        self.grid = int(min(1024, -(-I // self.block)))
        return a

    def launch(self, stream=None):
        a = self.args
        self.k.launch((self.grid,), (self.block,), [a[k] for k in ("y", "Pi", "Pr", "Pv", "Nv", "apod", "cinv", "strides", "x", "fsfc")],
                      stream)

    def result(self, Isz):
        y = self.args["y"].cpu().numpy()
        return y.reshape(tuple(Isz), order="F")


def ref_wsinterpd2f_inm(x, t1, t2, interp=1, omega=0.0, variant="fast", double=False):
    """The reference wsinterpd2f kernel (src/interpd.cu:344-396,460-467) on the canonical bfDAS shapes: x (T,N,M) complex64,
    t1 (I,N,1) receive sample indices, t2 (I,1,M) transmit sample indices (0-based, as ChannelData.sample2sep passes them,
    src/ChannelData.m:1428-1445), w = 1, summed over N and M -> y (I,).  Argument construction follows kern/wsinterpd2.m:
    dims = [I, N, M], iflags = [0, 1, 1] (:221-224), strides = [w; y; t1; t2; x] (:226), grid/block (:213-217)."""
    dev = "cuda"
    k = RefModule("interpd", "wsinterpd2" if double else "wsinterpd2f", variant)
    rdt, cdt = (np.float64, np.complex128) if double else (np.float32, np.complex64)
    tdt, tct = (torch.float64, torch.complex128) if double else (torch.float32, torch.complex64)
    T, N, M = x.shape
    I = t1.shape[0]
    QN = N * M
    for nm, v in (("QUPS_I", I), ("QUPS_T", T), ("QUPS_S", 3), ("QUPS_N", QN), ("QUPS_F", 1)):
        k.set_const(nm, v)
    cm = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dt).reshape(-1, order="F"))).to(dev)
    dx = torch.from_numpy(np.ascontiguousarray(np.asarray(x, cdt).transpose(2, 1, 0))).to(dev)
    d1, d2 = cm(t1, rdt), cm(t2, rdt)
    dw = torch.tensor([[1.0, 0.0]], dtype=tdt, device=dev)
    y = torch.zeros(I, dtype=tct, device=dev)
    sizes = torch.tensor([I, N, M], dtype=torch.int64, device=dev)
    iflags = torch.tensor([0, 1, 1], dtype=torch.uint8, device=dev)
    #                      w          y          t1         t2         x        (per dim: columns of the 5 x S matrix)
    strides = torch.tensor([0, 1, 1, 1, 0,   0, 0, I, 0, 1,   0, 0, 0, I, N], dtype=torch.int64, device=dev)
    K = 65535
    Lb = max(1, min(k.max_threads, -(-QN // K)))
    block = (min(I, k.max_threads // Lb), Lb, 1)
    grid = (max(1, -(-I // block[0])), max(1, -(-min(QN, K * Lb) // block[1])), max(1, -(-QN // (K * Lb))))
    k.launch(grid, block, [y, dw, dx, d1, d2, sizes, iflags, strides, C.c_int32(int(interp)),
                           (C.c_double if double else C.c_float)(float(omega))])
    torch.cuda.synchronize()
    return y.cpu().numpy()


def ref_greensf(ps, amp, pn, pv, kern, n0, S, fs, c0, wv_t0, fsr=1.0, R0=1e-3, interp=2, variant="fast", double=False):
    """The reference greensf kernel (src/greens.cu:8-98) with the argument pack of src/UltrasoundSystem.m:718.  The host-side
    window arrays only skip work (sb = per-scatterer first sample, iblock = scatterer range per time block): they are set wide
    open here so every (sample, scatterer) pair is evaluated.  Output S x N x M complex64, already divided by R0^2*fsr (:84)."""
    dev = "cuda"
    k = RefModule("greens", "greens" if double else "greensf", variant)
    rdt, cdt = (np.float64, np.complex128) if double else (np.float32, np.complex64)
    tdt, tct = (torch.float64, torch.complex128) if double else (torch.float32, torch.complex64)
    I, N, M, T = int(np.asarray(ps).shape[1]), int(np.asarray(pn).shape[1]), int(np.asarray(pv).shape[1]), int(len(kern))
    for nm, v in (("QUPS_S", S), ("QUPS_T", T), ("QUPS_N", N), ("QUPS_M", M), ("QUPS_I", I)):
        k.set_const(nm, v)
    col = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, rdt).T)).to(dev)
    x = torch.zeros(S * N * M, dtype=tct, device=dev)
    das = torch.from_numpy(np.asarray(amp, rdt).astype(cdt)).to(dev)
    dk = torch.from_numpy(np.asarray(kern, cdt)).to(dev)
    bx = min(k.max_threads, 32)
    gx = -(-S // bx)
    sb = torch.zeros(2 * I, dtype=tdt, device=dev)
    sb[0::2] = -1e30
    sb[1::2] = 1e30
    iblock = torch.tensor([0, I - 1] * gx, dtype=torch.int64, device=dev)
    f32 = rdt
    pack = torch.tensor([f32(n0) / f32(fs), wv_t0, fs, fsr, f32(1) / f32(c0), R0], dtype=tdt, device=dev)
    E = torch.tensor([1, 1], dtype=torch.int32, device=dev)
    k.launch((gx, N, M), (bx, 1, 1), [x, col(ps), das, col(pn), col(pv), dk, sb, iblock, pack, E, C.c_int32(int(interp))])
    torch.cuda.synchronize()
    return x.cpu().numpy().reshape((S, N, M), order="F")


class RefConvd:
    """The reference's batched convolution kernels convf / convcf (src/convd.cu:133-156) behind the launcher of
    kern/convd.m:98-201: x is C x M x S, y is C x N x S, z is C x L x S (column-major), start lag L0 as a __constant__."""

    def __init__(self, cplx: bool, variant: str = "ieee"):
        self.cplx = cplx
        self.k = RefModule("convd", "convcf" if cplx else "convf", variant)

    def run(self, x, y, shape="full"):
        """x: (C, M, S), y: (C, N, S) NumPy arrays (logical shapes).  Returns z (C, L, S) and the lags."""
        dt, tdt = (np.complex64, torch.complex64) if self.cplx else (np.float32, torch.float32)
        Cn, M, S = x.shape
        N = y.shape[1]
        if shape == "full":
            lags = np.arange(-(N - 1), M)                       # kern/convd.m:104-105
        elif shape == "same":
            lags = np.arange(0, M) - (N - 1) // 2               # :106-107
        else:
            lags = np.arange(0, M - N + 1)                      # :108-109
        L = len(lags)
        l0 = int(-lags[0]) if L else 0                          # :113
        col = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dt).transpose(2, 1, 0))).cuda()  # memory: C fastest
        dx, dy = col(x), col(y)
        dz = torch.zeros((S, L, Cn), dtype=tdt, device="cuda")
        sizes = torch.tensor([Cn, M, Cn, N, Cn, L, Cn], dtype=torch.int64, device="cuda")   # [xstr, M, ystr, N, zstr, L, C]  (:120-121)
        self.k.set_const("L0", l0, C.c_int32)
        mt = self.k.max_threads
        blk = (1, min(L, mt), 1) if Cn == 1 else (min(Cn, mt), 1, 1)                        # :187-191
        grid = tuple(-(-a // b) for a, b in zip((Cn, L, S), blk))
        self.k.launch(grid, blk, [dx, dy, dz, sizes])
        torch.cuda.synchronize()
        return dz.cpu().numpy().transpose(2, 1, 0), lags
