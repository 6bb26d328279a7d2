"""GPU parity at the STATED size of every BASELINE.json config other than the headline (C2 lives in test_gpu_das.py):

  C1  128 x 128 px, L11-5v 128 el, 1 plane wave, T = 2048, linear, fp32   — the whole image against the C oracle
  C3  512 x 512 px, 192 rx x 128 plane waves, T = 2048, half2 IQ, fmod = fc, cubic — full image on the GPU, a pixel subset at the
      full N x M against the oracle on the fp16-rounded inputs (the fp16 parity definition)
  C4  256^3 voxels, 32 x 32 matrix array (N = 1024), 64 diverging waves — one 256 x 256 x 32 slab at full N x M on the GPU, a voxel
      subset against the oracle and against the bit-exact generic kernel
  C5  greens(10 000 scatterers, 256 x 256 FSA) -> DAS 1024^2: a 64-trace subset of the simulated cube against oracle_greens,
      the point-target known answer (test/BFTest.m:230-317) on a single-scatterer round trip at the same array size.
"""
import numpy as np
import pytest

from tests.util import rel_linf

pytestmark = pytest.mark.gpu
f32 = np.float32
TOL = 1e-5   # BASELINE.json north_star: within 1e-5 relative L-inf of das_spec.m (CPU semantics)


def _dev(v):
    import torch
    return torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).cuda()


def test_c1_full_size_whole_image_vs_oracle(oracle_c):
    import qups_b200
    from qups_b200 import synth
    P = synth.config_c1()
    assert P.Isz == (128, 128, 1) and P.N == 128 and P.M == 1 and P.T == 2048 and P.interp == "linear"
    x = synth.noise_cube(P.T, P.N, P.M, seed=1)
    ref = oracle_c.das_spec("DAS", P.Pi, P.Pr, P.Pv, P.Nv, x, P.t0, P.fs, P.c0, interp="linear", VS=False)
    got = qups_b200.das_spec("DAS", *(v.astype(f32) for v in (P.Pi, P.Pr, P.Pv, P.Nv)), x, P.t0, P.fs, P.c0, "plane-waves", "interp", "linear")
    assert qups_b200.last_das_kernel() == "das_tiled"
    assert np.abs(ref).max() > 1
    assert rel_linf(got.reshape(ref.shape), ref) < TOL
    from qups_b200 import _lib
    gen = qups_b200.das_spec("DAS", *(v.astype(f32) for v in (P.Pi, P.Pr, P.Pv, P.Nv)), x, P.t0, P.fs, P.c0, "plane-waves", "interp", "linear",
                             _path=_lib.PATH_GENERIC)
    assert np.array_equal(gen.reshape(ref.shape), ref)   # generic kernel: bit-exact


def test_c3_full_size_half2_modulated_vs_oracle_on_rounded_inputs(oracle_c):
    import torch
    import qups_b200
    from qups_b200 import synth, _lib
    P = synth.config_c3()
    fc = float(P.meta["fc"])
    assert P.Isz == (512, 512, 1) and P.N == 192 and P.M == 128 and P.T == 2048
    x = synth.noise_cube(P.T, P.N, P.M, seed=2)
    xh = (x.real.astype(np.float16).astype(f32) + 1j * x.imag.astype(np.float16).astype(f32)).astype(np.complex64, order="F")
    xd = torch.from_numpy(xh).cuda()
    t0 = 0.0
    g = (_dev(P.Pr), _dev(P.Pv), _dev(P.Nv))
    opts = ("plane-waves", "interp", "cubic", "modulation", fc, "input-precision", "halfT")
    qups_b200.lib()
    _lib.launch_count(reset=True)
    full = qups_b200.das_spec("DAS", _dev(P.Pi), *g, xd, t0, P.fs, P.c0, *opts, _y_f32=True)
    assert qups_b200.last_das_kernel() == "das_tiled"      # fp16 + fmod stays on the staged kernel
    full = full.cpu().numpy().reshape(512, 512, order="F")
    scale = np.abs(full).max()
    rng = np.random.default_rng(5)
    iz, ix = rng.integers(0, 512, 256), rng.integers(0, 512, 256)
    sub = np.ascontiguousarray(P.Pi[:, iz, ix, 0]).reshape(3, -1, 1, 1)
    ref = oracle_c.das_spec("DAS", sub, P.Pr, P.Pv, P.Nv, xh, t0, P.fs, P.c0, interp="cubic", VS=False, fmod=fc).reshape(-1)
    assert np.max(np.abs(ref - full[iz, ix])) / scale < TOL, np.max(np.abs(ref - full[iz, ix])) / scale
    # half2 output (what the reference's DASh writes): the same image rounded to fp16
    out16 = qups_b200.das_spec("DAS", _dev(P.Pi), *g, xd, t0, P.fs, P.c0, *opts).cpu().numpy().reshape(512, 512, order="F")
    assert np.max(np.abs(out16 - full)) / scale < 2e-3


def test_c4_slab_full_aperture_vs_oracle_and_generic(oracle_c):
    import torch
    import qups_b200
    from qups_b200 import synth, _lib
    P = synth.config_c4()
    assert P.Isz == (256, 256, 256) and P.N == 1024 and P.M == 64 and P.T == 2048
    x = synth.noise_cube(P.T, P.N, P.M, seed=3)
    xd = torch.from_numpy(x).cuda()
    k0 = 112
    slab = np.ascontiguousarray(P.Pi[:, :, :, k0:k0 + 32])           # 256 x 256 x 32 voxels (what one of 8 GPUs beamforms)
    g = (_dev(P.Pr), _dev(P.Pv), _dev(P.Nv))
    opts = ("diverging-waves", "interp", "cubic")
    full = qups_b200.das_spec("DAS", _dev(slab), *g, xd, 0.0, P.fs, P.c0, *opts)
    assert qups_b200.last_das_kernel() == "das_tiled"
    full = full.cpu().numpy().reshape(256, 256, 32, order="F")
    scale = np.abs(full).max()
    assert scale > 1
    rng = np.random.default_rng(7)
    iz, ix, iy = rng.integers(0, 256, 1024), rng.integers(0, 256, 1024), rng.integers(0, 32, 1024)
    sub = np.ascontiguousarray(slab[:, iz, ix, iy]).reshape(3, -1, 1, 1)
    gen = qups_b200.das_spec("DAS", _dev(sub), *g, xd, 0.0, P.fs, P.c0, *opts, _path=_lib.PATH_GENERIC).cpu().numpy().reshape(-1)
    assert np.max(np.abs(gen - full[iz, ix, iy])) / scale < TOL
    ref = oracle_c.das_spec("DAS", sub[:, :48], P.Pr, P.Pv, P.Nv, x, 0.0, P.fs, P.c0, interp="cubic", VS=True, DV=True).reshape(-1)
    assert np.array_equal(ref, gen[:48])
    assert np.max(np.abs(ref - full[iz[:48], ix[:48], iy[:48]])) / scale < TOL


def test_c5_greens_at_scale_trace_subset_and_round_trip(oracle_c):
    import torch
    import qups_b200
    from qups_b200 import synth, ultrasound as U
    P = synth.config_c5_das()
    fc, fs, c0 = P.meta["fc"], P.fs, P.c0
    S = 10000
    rng = np.random.Generator(np.random.PCG64(1))
    ps = np.stack([rng.uniform(-25e-3, 25e-3, S), np.zeros(S), rng.uniform(1e-3, 51e-3, S)], 0)
    amp = rng.standard_normal(S)
    kern, wt0, wtend = synth.greens_kernel(fc, 0.6, fs)
    r = np.linalg.norm(ps[:, :, None] - P.Pr[:, None, :], axis=0)
    n0 = int(np.floor((2 * r.min() / c0 + wt0 - (wtend - wt0)) * fs))
    T = int(np.ceil((2 * r.max() / c0 + wtend) * fs)) - n0 + 1
    R0 = c0 / fc
    x = U.greens_raw(ps, amp, P.Pr, P.Pr, kern, n0, T, fs, c0, wt0, 1.0, R0, "cubic")     # T x 256 x 256 on the GPU
    assert tuple(x.shape) == (T, 256, 256)
    # 64 traces (8 receives x 8 transmits spread over the aperture) against oracle_greens: fp64 arbiter on the fp32 inputs at
    # 5e-6 (the kernel computes the arrival times in fp64), fp32 oracle at the reference's own 1e-3 bar
    sel = np.arange(0, 256, 36)[:8]
    f = lambda a: np.asarray(a, f32)
    ref64 = oracle_c.greens(f(ps), f(amp), f(P.Pr[:, sel]), f(P.Pr[:, sel]), kern, n0, T, fs, c0, wt0, 1.0, R0, "cubic", dtype=np.float64)
    ref32 = oracle_c.greens(ps, amp, P.Pr[:, sel], P.Pr[:, sel], kern, n0, T, fs, c0, wt0, 1.0, R0, "cubic")
    got = x[:, torch.from_numpy(sel).cuda()][:, :, torch.from_numpy(sel).cuda()].cpu().numpy()
    assert rel_linf(got, ref64) < 5e-6, rel_linf(got, ref64)
    assert rel_linf(got, ref32) < 1e-3
    # DAS of the simulated cube at the headline grid: staged kernel, finite, and equal to the generic kernel on a pixel subset
    g = (_dev(P.Pr), _dev(P.Pv), _dev(P.Nv))
    img = qups_b200.das_spec("DAS", _dev(P.Pi), *g, x, n0 / fs, fs, c0, "diverging-waves", "interp", "cubic")
    assert qups_b200.last_das_kernel() == "das_tiled"
    img = img.cpu().numpy().reshape(1024, 1024, order="F")
    assert np.isfinite(img).all()
    from qups_b200 import _lib
    iz, ix = rng.integers(0, 1024, 512), rng.integers(0, 1024, 512)
    sub = np.ascontiguousarray(P.Pi[:, iz, ix, 0]).reshape(3, -1, 1, 1)
    gen = qups_b200.das_spec("DAS", _dev(sub), *g, x, n0 / fs, fs, c0, "diverging-waves", "interp", "cubic", _path=_lib.PATH_GENERIC)
    assert np.max(np.abs(gen.cpu().numpy().reshape(-1) - img[iz, ix])) / np.abs(img).max() < TOL
    del x
    torch.cuda.empty_cache()
    # point-target known answer at the same array / grid: one scatterer -> PSF peak within 1.1 mm (test/BFTest.m:230-317)
    p0 = np.array([[3e-3], [0.0], [25e-3]])
    r1 = np.linalg.norm(p0[:, :, None] - P.Pr[:, None, :], axis=0)
    n1 = int(np.floor((2 * r1.min() / c0 + wt0 - (wtend - wt0)) * fs))
    T1 = int(np.ceil((2 * r1.max() / c0 + wtend) * fs)) - n1 + 1
    x1 = U.greens_raw(p0, np.ones(1), P.Pr, P.Pr, kern, n1, T1, fs, c0, wt0, 1.0, R0, "cubic")
    b = qups_b200.das_spec("DAS", _dev(P.Pi), *g, x1, n1 / fs, fs, c0, "diverging-waves", "interp", "cubic").abs().cpu().numpy().reshape(1024, 1024, order="F")
    kz, kx = np.unravel_index(np.argmax(b), b.shape)
    pk = P.Pi[:, kz, kx, 0]
    assert np.linalg.norm(pk - p0[:, 0]) < 1.1e-3, pk
