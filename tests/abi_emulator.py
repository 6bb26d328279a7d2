"""CPU emulation of the qups_wsinterpd2 C-ABI call (TEST INFRASTRUCTURE ONLY).

Lets the CPU suite run the REAL host-side mirror (qups_b200/kern.py::wsinterpd2: dim moves, broadcast sizes, the 5 x D
stride matrix, column-major buffers) without a GPU: the library call is replaced by a Python interpreter of exactly the
arguments the mirror hands to the C ABI (include/qups_b200.h: qups_ws2_params + device pointers), evaluated with the
NumPy oracle's interp1.  What is under test is the packing, not the kernel."""
import contextlib
import ctypes as C
import itertools

import numpy as np
import torch

from oracle import oracle_np


def _buf(ptr, n, dtype):
    if not ptr or n == 0:
        return None
    return np.ctypeslib.as_array((C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)).view(dtype)


class FakeLib:
    """Stands in for ctypes.CDLL(libqups_b200.so) for the wsinterpd2 / wsinterpd entry points."""

    def __init__(self):
        self.calls = 0

    def qups_last_error(self):
        return b"emulator"

    def qups_wsinterpd2(self, p_ref, y, w, x, t1, t2, stream):
        p = p_ref._obj
        self.calls += 1
        D, T = int(p.D), int(p.T)
        sizes = [int(p.sizes[k]) for k in range(D)]
        st = np.array([[int(p.dstride[r + 5 * k]) for k in range(D)] for r in range(5)])  # rows: w, y, t1, t2, x-trace
        span = lambda r: 1 + int(sum((sizes[k] - 1) * st[r][k] for k in range(D)))
        rdt, cdt = (np.float64, np.complex128) if p.dtype == 2 else (np.float32, np.complex64)
        val = lambda v: v.value if hasattr(v, "value") else v
        W = _buf(val(w), span(0), rdt if p.w_real else cdt)
        Y = _buf(val(y), span(1), cdt)
        T1 = _buf(val(t1), span(2), rdt)
        T2 = _buf(val(t2), span(3), rdt) if val(t2) else None
        X = _buf(val(x), span(4) * T, cdt)
        method = {0: "nearest", 1: "linear", 2: "cubic"}[int(p.interp)]
        Y[:] = 0
        for idx in itertools.product(*[range(s) for s in sizes]):
            off = [int(sum(i * s for i, s in zip(idx, st[r]))) for r in range(5)]
            t = rdt(T1[off[2]]) + (rdt(T2[off[3]]) if T2 is not None else rdt(0))
            v = oracle_np.interp1(X[off[4] * T:(off[4] + 1) * T], np.array([1 + t]), method, 0)[0]
            Y[off[1]] += np.exp(1j * p.omega * t) * W[off[0]] * v
        return 0

    def qups_wsinterpd(self, p_ref, y, w, x, t, stream):
        return self.qups_wsinterpd2(p_ref, y, w, x, t, None, stream)


@contextlib.contextmanager
def emulated(monkeypatch):
    """Route qups_b200.kern's device plumbing to the CPU and its library handle to FakeLib."""
    from qups_b200 import kern, _lib
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "check", lambda rc: None if rc == 0 else (_ for _ in ()).throw(RuntimeError(rc)))
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    real_device = torch.device
    monkeypatch.setattr(kern.torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(kern, "_stream", lambda dev: C.c_void_p(0))
    yield fake
