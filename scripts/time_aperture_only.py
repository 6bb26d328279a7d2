"""Timing of dmas / slsc on a random C2-size keep_rx cube (1024 x 1024 x 1 x 256 complex64 = 2.15 GB), no beamforming first."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200

def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

g = torch.Generator(device="cuda").manual_seed(0)
bn = torch.view_as_complex(torch.randn((256, 1, 1024, 1024, 2), device="cuda", generator=g)).permute(3, 2, 1, 0)  # column-major I1 x I2 x I3 x N
for name, fn in (("dmas L=16", lambda: qups_b200.dmas(bn, 4, 16)), ("slsc L=16 ensemble", lambda: qups_b200.slsc(bn, 4, 16, "ensemble")),
                 ("slsc L=16 average", lambda: qups_b200.slsc(bn, 4, 16, "average")), ("dmas all lags", lambda: qups_b200.dmas(bn, 4))):
    print(f"{name:20s} {ev(fn):8.3f} ms", flush=True)
