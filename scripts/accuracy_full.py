"""Full-size accuracy: errors of the fp32 oracle, the generic kernel and the staged kernel vs the fp64 oracle on a
random pixel subset of the headline configuration (relative to max|b| of the image)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qups_b200
from qups_b200 import synth, _lib
from oracle import oracle_c
P = synth.config_c2(); x = synth.noise_cube(P.T, P.N, P.M, seed=0)
f32 = np.float32
dev = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).cuda()
xd = torch.from_numpy(x).cuda(); g = (dev(P.Pr), dev(P.Pv), dev(P.Nv))
full = qups_b200.das_spec("DAS", dev(P.Pi), *g, xd, 0.0, P.fs, P.c0, "interp", "cubic").cpu().numpy().reshape(1024, 1024, order="F")
scale = np.abs(full).max()
rng = np.random.default_rng(3); n = 256
iz, ix = rng.integers(0, 1024, n), rng.integers(0, 1024, n)
sub = np.ascontiguousarray(P.Pi[:, iz, ix, 0]).reshape(3, -1, 1, 1)
gen = qups_b200.das_spec("DAS", dev(sub), *g, xd, 0.0, P.fs, P.c0, "interp", "cubic", _path=_lib.PATH_GENERIC).cpu().numpy().reshape(-1)
o32 = oracle_c.das_spec("DAS", sub, P.Pr, P.Pv, P.Nv, x, 0.0, P.fs, P.c0, interp="cubic").reshape(-1)
o64 = oracle_c.das_spec("DAS", sub, P.Pr, P.Pv, P.Nv, x, 0.0, P.fs, P.c0, interp="cubic", dtype=np.float64).reshape(-1)
t = full[iz, ix]
e = lambda a, b: float(np.max(np.abs(a - b)) / scale)
print(f"scale {scale:.1f}; vs fp64 oracle: fp32 oracle {e(o32,o64):.2e}, generic {e(gen,o64):.2e}, tiled {e(t,o64):.2e}; tiled vs fp32 oracle {e(t,o32):.2e}; generic==oracle32 {np.array_equal(gen,o32)}")
