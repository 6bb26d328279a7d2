"""ref_ptx.py — launches the REFERENCE's own CUDA kernels (oracle/_ref/*.ptx) on the GPU box.

TEST INFRASTRUCTURE ONLY.  oracle/_ref/bf.ptx is src/bf.cu of the reference compiled UNMODIFIED with the
reference's flags (`nvcc --ptx -arch=compute_100 --use_fast_math`, oracle/Makefile); this module does what
kern/das_spec.m:284-373 does with parallel.gpu.CUDAKernel: load the PTX, set the QUPS_* __constant__ symbols,
pick the reference launch geometry and call DASf.  It is used to
  (i) pin the oracle against the real reference kernel on interior samples (tests/test_gpu_reference_kernel.py;
      the kernel's trace-edge behaviour differs from the CPU path), and
 (ii) time "the reference's kernel on this box" beside ours (bench.py extra field).
Never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
BF_PTX = os.path.join(_HERE, "_ref", "bf.ptx")


def available() -> bool:
    try:
        from cuda.bindings import driver  # noqa: F401
    except Exception:
        return False
    return os.path.exists(BF_PTX) and torch.cuda.is_available()


def _chk(res):
    from cuda.bindings import driver
    err = res[0]
    if err != driver.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


class RefDASf:
    """The reference DASf kernel (src/bf.cu:154-161) behind the argument list of kern/das_spec.m:372."""

    def __init__(self):
        from cuda.bindings import driver
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make torch's primary context current
        with open(BF_PTX, "rb") as f:
            ptx = f.read() + b"\0"
        self.drv = driver
        self.mod = _chk(driver.cuModuleLoadData(ptx))
        # the reference's entry points are C++-mangled; CUDAKernel(ptx, cu, 'DASf') resolves them by base name
        import re
        entry = None
        for mname, n, rest in re.findall(rb"\.entry\s+(_Z(\d+)([A-Za-z0-9_]+))\s*\(", ptx):
            if rest[:int(n)] == b"DASf" and int(n) == 4:
                entry = mname
        if entry is None:
            raise RuntimeError("DASf entry not found in oracle/_ref/bf.ptx")
        self.fn = _chk(driver.cuModuleGetFunction(self.mod, entry))
        self.max_threads = _chk(driver.cuFuncGetAttribute(
            driver.CUfunction_attribute.CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, self.fn))

    def _set(self, name: str, value, ctype):
        dptr, size = _chk(self.drv.cuModuleGetGlobal(self.mod, name.encode()))
        v = ctype(value)
        assert C.sizeof(v) == size, (name, size)
        _chk_none(self.drv.cuMemcpyHtoD(dptr, C.addressof(v), size))

    def prepare(self, Pi, Pr, Pv, Nv, x, t0, fs, c, interp=2, VS=True, DV=False, fmod=0.0):
        """Device-side argument pack; Pi (3,I1,I2,I3), x (T,N,M) complex64 (Fortran order), fp32 everything."""
        dev = "cuda"
        f32 = np.float32
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
        a["apod"] = torch.tensor([[1.0, 0.0]], dtype=torch.float32, device=dev)   # apod = {1}, complex
        a["cinv"] = torch.tensor([f32(1) / f32(c)], dtype=torch.float32, device=dev)
        a["strides"] = torch.zeros(12, dtype=torch.int64, device=dev)            # scalar cinv / scalar apod
        if isinstance(x, torch.Tensor):
            a["x"] = x
        else:
            a["x"] = torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.complex64).transpose(2, 1, 0))).to(dev)
        a["fsfc"] = torch.tensor([fs, fmod], dtype=torch.float32, device=dev)
        a["y"] = torch.zeros(I, dtype=torch.complex64, device=dev)
        u64 = C.c_uint64
        for nm, v in (("QUPS_I", I), ("QUPS_T", T), ("QUPS_M", M), ("QUPS_N", N), ("QUPS_I1", Isz[0]),
                      ("QUPS_I2", Isz[1]), ("QUPS_I3", Isz[2]), ("QUPS_S", 1)):
            self._set(nm, v, u64)
        self._set("QUPS_VS", bool(VS), C.c_bool)
        self._set("QUPS_DV", bool(DV), C.c_bool)
        self._set("QUPS_BF_FLAG", int(interp), C.c_int32)
        self.args = a
        # launch geometry of kern/das_spec.m:301-306
        self.block = int(self.max_threads)
        self.grid = int(min(1024, -(-I // self.block)))
        return a

    def launch(self, stream=None):
        a = self.args
        ptrs = [a[k].data_ptr() for k in ("y", "Pi", "Pr", "Pv", "Nv", "apod", "cinv", "strides", "x", "fsfc")]
        holders = [C.c_void_p(p) for p in ptrs]
        argv = (C.c_void_p * len(holders))(*[C.addressof(h) for h in holders])
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _chk_none(self.drv.cuLaunchKernel(self.fn, self.grid, 1, 1, self.block, 1, 1, 0, st, C.addressof(argv), 0))

    def result(self, Isz):
        y = self.args["y"].cpu().numpy()
        return y.reshape(tuple(Isz), order="F")


def _chk_none(res):
    from cuda.bindings import driver
    err = res[0] if isinstance(res, tuple) else res
    if err != driver.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
