"""Timing of the generic (strided N-D) wsinterpd2 kernel on the bfDAS table form, forced off the staged path
(QUPS_B200_WS2_GENERIC=1), and of focusTx's call shape."""
import os, sys
os.environ["QUPS_B200_WS2_GENERIC"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200
from qups_b200 import synth, kern

def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

P = synth.config_c2(256, 256, 64, 64, 1024)
x = torch.from_numpy(synth.noise_cube(P.T, P.N, P.M)).cuda()
I = P.I
rng = np.random.default_rng(0)
trx = torch.from_numpy(rng.uniform(100, 400, (I, P.N)).astype(np.float32)).cuda()
ttx = torch.from_numpy(rng.uniform(100, 400, (I, P.M)).astype(np.float32)).cuda()
# y(i) = sum_n sum_m interp1(x(:, n, m), 1 + trx(i, n) + ttx(i, m)):  x T x N x M, t1 I x N x 1, t2 I x 1 x M (column-major)
xv = x.reshape(P.T, P.N, P.M)
fn = lambda: kern.wsinterpd2(xv, trx.reshape(I, P.N, 1), ttx.reshape(I, 1, P.M), 1, 1.0, (2, 3), "cubic")
t = ev_time(fn)
print(f"generic wsinterpd2, table DAS {I} px x {P.N} x {P.M}: {t:.2f} ms  {I*P.N*P.M/t/1e6:.1f} Gterm/s  ({qups_b200.last_ws2_kernel()})")
