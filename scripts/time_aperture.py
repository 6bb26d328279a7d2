"""Timing of the keep_rx pipeline at the headline size: SYN (generic kernel, I x N output = 2 GB) + cohfac / dmas / pcf / slsc."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200
from qups_b200 import synth

def ev(fn, n=2):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), r

P = synth.config_c2()
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
dev = lambda v: torch.from_numpy(np.asarray(v, np.float32)).cuda()
g = (dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv))
t, bn = ev(lambda: qups_b200.das_spec("SYN", *g, x, 0.0, P.fs, P.c0, "interp", "cubic"), 1)
print(f"SYN (keep_rx) C2 {qups_b200.last_das_kernel()}: {t:.1f} ms, output {bn.numel()*8/1e9:.2f} GB", flush=True)
from qups_b200 import _lib
t2, _ = ev(lambda: qups_b200.das_spec("SYN", *g, x, 0.0, P.fs, P.c0, "interp", "cubic", _path=_lib.PATH_GENERIC), 1)
print(f"SYN (keep_rx) C2 generic kernel: {t2:.1f} ms", flush=True)
t3, bm = ev(lambda: qups_b200.das_spec("MUL", *g, x, 0.0, P.fs, P.c0, "interp", "cubic"), 1)
print(f"MUL (keep_tx) C2 {qups_b200.last_das_kernel()}: {t3:.1f} ms, output {bm.numel()*8/1e9:.2f} GB", flush=True)
del bm
for name, fn in (("cohfac", lambda: qups_b200.cohfac(bn, 4)), ("pcf", lambda: qups_b200.pcf(bn, 4)),
                 ("dmas L=16", lambda: qups_b200.dmas(bn, 4, 16)), ("slsc L=16 ensemble", lambda: qups_b200.slsc(bn, 4, 16, "ensemble")),
                 ("slsc L=16 average", lambda: qups_b200.slsc(bn, 4, 16, "average"))):
    t, _ = ev(fn)
    print(f"{name:22s} {t:8.2f} ms  ({bn.numel()*8/1e9/t*1e3:.0f} GB/s over the 2.15 GB cube; includes the mirror's layout copy)", flush=True)
