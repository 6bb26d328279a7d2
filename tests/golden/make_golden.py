"""Generates tests/golden/*.npz — ORACLE-generated regression pins (NOT reference outputs: the reference's CPU path
is MATLAB and cannot run here; see oracle/qups_oracle.h).  Run from the repo root: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_c, oracle_np  # noqa: E402
from qups_b200 import synth  # noqa: E402
from tests.util import small_problem, oracle_kwargs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, P, interp, fun="DAS", t0=0.0, apod=(), fmod=0.0):
    kw = oracle_kwargs(P["opts"])
    y32 = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], P["c"], interp=interp, apod=apod, fmod=fmod, **kw)
    y64 = oracle_np.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], P["c"], interp=interp, apod=apod, fmod=fmod,
                             dtype=np.float64, **kw)
    np.savez_compressed(os.path.join(HERE, name), Pi=P["Pi"].astype(np.float32), Pr=P["Pr"].astype(np.float32),
                        Pv=P["Pv"].astype(np.float32), Nv=P["Nv"].astype(np.float32), x=P["x"], t0=np.asarray(t0), fs=P["fs"],
                        c=P["c"], opts=np.array(P["opts"], dtype=object), interp=interp, fun=fun, fmod=fmod,
                        apod=np.array(list(apod), dtype=object) if apod else np.zeros(0), y32=y32, y64=y64.astype(np.complex128))


if __name__ == "__main__":
    # C1-like: plane wave, linear (BASELINE.json configs[0], reduced)
    C1 = synth.config_c1(32, 24, 256)
    x = synth.noise_cube(256, 128, 1, seed=1)
    P = dict(Pi=C1.Pi, Pr=C1.Pr, Pv=C1.Pv, Nv=C1.Nv, x=x, fs=C1.fs, c=C1.c0, opts=C1.opts)
    # shrink the depth range so delays stay inside 256 samples
    P["Pi"] = synth.scan_cartesian(np.linspace(-19.05e-3, 19.05e-3, 24), np.linspace(1e-3, 5.5e-3, 32))
    save("c1_pw_linear.npz", P, "linear")
    P = small_problem("FC", nz=20, nx=33, N=17, M=5, T=260, zlim=(2e-3, 12e-3), seed=4)
    save("fc_cubic_t0.npz", P, "cubic", t0=np.linspace(-2e-7, 2e-7, 5))
    P = small_problem("DV", nz=18, nx=12, N=9, M=4, T=260, zlim=(2e-3, 12e-3), seed=5)
    save("dv_nearest_syn.npz", P, "nearest", fun="SYN")
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
