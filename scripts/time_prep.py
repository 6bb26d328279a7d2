"""Timing of the fused ChannelData pre-processing pass at the headline cube size (T = 2048, 256 x 256 traces):
algorithmic bytes = input + output of the cube once; HBM roofline = measured copy bandwidth."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qups_b200 import ultrasound as U

T, N, M, fs = 2048, 256, 256, 30e6
rng = np.random.default_rng(0)
rf16 = torch.from_numpy(rng.integers(-2000, 2000, (T, N, M)).astype(np.int16)).cuda()
rf32 = rf16.float()
iq = torch.complex(rf32, rf32.flip(0))

def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

# note: the mirror's _colmajor permute+copy of the input is inside the timed call (extra pass the C ABI does not have)
for label, x, kw, bin_, bout in (
        ("cast int16 -> complex fp32", rf16, dict(), 2, 8), ("downmix complex fp32", iq, dict(fmix=7.5e6), 8, 8),
        ("hilbert real fp32 (L = 2048)", rf32, dict(hilbert=True), 4, 8),
        ("zeropad+hilbert+downmix int16 -> fp32 (L = 2048+0)", rf16, dict(hilbert=True, fmix=7.5e6), 2, 8),
        ("zeropad(8,24)+hilbert (L = 2080, Bluestein)", rf16, dict(B=8, A=24, hilbert=True), 2, 8),
        ("hilbert+downmix -> half2", rf16, dict(hilbert=True, fmix=7.5e6, out="halfT"), 2, 4)):
    chd = U.ChannelData(x, 0.0, fs)
    t = ev(lambda: chd.prep(**kw))
    L = T + kw.get("B", 0) + kw.get("A", 0)
    gb = (T * bin_ + L * bout) * N * M / 1e9
    print(f"{label:55s} {t:8.2f} ms   {gb:.2f} GB algorithmic  {gb/t*1e3:7.0f} GB/s", flush=True)
