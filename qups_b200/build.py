"""build.py — compiles libqups_b200.so in-tree with nvcc for sm_100a (B200).

    python -m qups_b200.build [--force] [--verbose]

Two translation-unit classes:
  * "exact" units (das_generic.cu, greens.cu, wsinterpd2.cu) are compiled with
    -fmad=false so every floating-point operation is individually rounded and
    the fp32 result is bit-exact against oracle/qups_oracle.c;
  * the hot kernel (das_tiled.cu) keeps FMA contraction for the interpolation
    arithmetic and pins the delay sequence with explicit _rn intrinsics.
No --use_fast_math anywhere (the reference builds with it; we need IEEE sqrt/div).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libqups_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
UNITS = {
    "das_generic.cu": ["-fmad=false"],
    "greens.cu": ["-fmad=false"],
    "wsinterpd2.cu": ["-fmad=false"],
    "convd.cu": ["-fmad=false"],
    "apod_gen.cu": ["-fmad=false"],
    "chd_prep.cu": [],
    "aperture.cu": [],
    "xcorr.cu": [],
    "refocus.cu": [],
    "das_tiled.cu": [],
    "qups_b200.cu": [],
}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """defines: extra -D macros for tuning experiments (QUPS_NT, QUPS_CW, QUPS_STAGES, QUPS_MINBLOCKS, QUPS_WMAX);
    out: alternative .so path (select at run time with QUPS_B200_LIB)."""
    global OBJ
    if defines:
        OBJ = os.path.join(HERE, "csrc", "_obj_" + "_".join(d.replace("=", "") for d in defines))
        force = True
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "qups_b200.h"))
    objs = []
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a gcc wrapper without libgomp; nvcc should use PATH gcc
    env.pop("CXX", None)
    jobs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = ["nvcc", *ARCH, *COMMON, *extra, *["-D" + d for d in defines], "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            jobs.append((unit, cmd))
    if jobs:  # translation units are independent: compile them side by side (das_tiled.cu alone takes about a minute)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            results = list(ex.map(lambda j: (j[0], subprocess.run(j[1], capture_output=True, text=True, env=env)), jobs))
        for unit, r in results:
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {unit}")
    if force or _stale(out, objs):
        cmd = ["nvcc", *ARCH, "-shared", "-o", out, *objs, "-Xlinker", "--exclude-libs,ALL", "-cudart", "shared"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outp = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), OUT)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outp))
