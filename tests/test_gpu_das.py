"""GPU parity tests (run with -m gpu on the B200 box): CUDA DAS through the C ABI vs the CPU oracle.

Bars (BASELINE.json north_star): bit-exact for nearest-neighbour indexing, <= 1e-5 relative L-inf vs the fp32
oracle (kern/das_spec.m CPU semantics) for linear / cubic; the generic kernel is bit-exact for everything.
"""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu

TOL = 1e-5  # relative L-inf vs max|b|, stated in BASELINE.json


def _gpu(fun, P, interp, path, extra=(), x=None, t0=None, c=None, **kw):
    import qups_b200
    from qups_b200 import _lib
    pth = {"generic": _lib.PATH_GENERIC, "tiled": _lib.PATH_TILED, "auto": _lib.PATH_AUTO}[path]
    f32 = np.float32
    out = qups_b200.das_spec(fun, P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32),
                             P["x"] if x is None else x, P["t0"] if t0 is None else t0, P["fs"],
                             P["c"] if c is None else c, *P["opts"], "interp", interp, *extra, _path=pth, **kw)
    return out


def _ora(oracle_c, fun, P, interp, x=None, t0=None, c=None, **kw):
    okw = oracle_kwargs(P["opts"])
    okw.update(kw)
    xx = P["x"] if x is None else x
    ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], xx,
                            P["t0"] if t0 is None else t0, P["fs"], P["c"] if c is None else c, interp=interp, **okw)
    if fun != "delays" and xx.ndim == 3:
        ref = ref[..., 0]  # MATLAB drops the trailing singleton frame dim
    return ref


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA", "DV"])
@pytest.mark.parametrize("fun", ["DAS", "SYN", "MUL", "BF", "delays"])
def test_generic_bitexact_all_funs(oracle_c, kind, fun):
    P = small_problem(kind, nz=19, nx=13, N=7, M=5, T=150)
    for interp in ("nearest", "linear", "cubic"):
        ref = _ora(oracle_c, fun, P, interp)
        got = _gpu(fun, P, interp, "generic")
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), (fun, kind, interp, rel_linf(got, ref))


def test_generic_lanczos3_close(oracle_c):
    P = small_problem("FC", nz=19, nx=13, N=7, M=5, T=150)
    ref = _ora(oracle_c, "DAS", P, "lanczos3")
    got = _gpu("DAS", P, "lanczos3", "generic")
    assert rel_linf(got, ref) < 1e-5


def test_generic_apod_cinv_t0_frames_transpose(oracle_c):
    P = small_problem("FC", nz=11, nx=9, ny=2, N=6, M=4, T=150, F=3)
    rng = np.random.default_rng(3)
    Isz = P["Pi"].shape[1:]
    apods = [rng.uniform(0, 1, Isz + (6, 1)).astype(np.float32), rng.uniform(0, 1, (1, 1, 1, 1, 4)).astype(np.float32),
             (rng.uniform(0, 1, (Isz[0], 1, 1, 6, 4)) > 0.3).astype(np.float32)]
    c = rng.uniform(1500, 1580, Isz).astype(np.float32)
    t0 = rng.uniform(-2e-7, 2e-7, 4)
    extra = sum((("apod", a) for a in apods), ())
    for fun in ("DAS", "SYN", "MUL", "BF"):
        ref = _ora(oracle_c, fun, P, "cubic", t0=t0, c=c, apod=apods)
        got = _gpu(fun, P, "cubic", "generic", extra, t0=t0, c=c)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), fun
    xt = np.asfortranarray(np.swapaxes(P["x"], 1, 2))
    ref = _ora(oracle_c, "DAS", P, "linear", t0=t0)
    got = _gpu("DAS", P, "linear", "generic", ("transpose", True), x=xt, t0=t0)
    assert np.array_equal(got, ref)
    # complex apodization (the reference GPU path forces complex weights: kern/das_spec.m:237-243)
    ac = [(apods[0] * np.exp(1j * 0.3)).astype(np.complex64)]
    ref = _ora(oracle_c, "DAS", P, "linear", apod=ac)
    got = _gpu("DAS", P, "linear", "generic", ("apod", ac[0]))
    assert np.array_equal(got, ref)


def test_generic_fp64_and_fp16(oracle_c):
    P = small_problem("PW", nz=13, nx=9, N=6, M=4, T=150)
    import qups_b200
    ref = _ora(oracle_c, "DAS", P, "cubic", dtype=np.float64)
    got = qups_b200.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"].astype(np.complex128), P["t0"], P["fs"],
                             P["c"], *P["opts"], "interp", "cubic", "input-precision", "double")
    assert got.dtype == np.complex128
    assert rel_linf(got, ref) < 1e-12
    # fp16 parity definition: oracle on fp16-rounded inputs, fp32 math
    xh = (P["x"].real.astype(np.float16).astype(np.float32) + 1j * P["x"].imag.astype(np.float16).astype(np.float32)).astype(np.complex64)
    ref = _ora(oracle_c, "DAS", P, "cubic", x=xh)
    got = _gpu("DAS", P, "cubic", "generic", ("input-precision", "halfT"), _y_f32=True)
    assert np.array_equal(got, ref)
    import qups_b200 as qb
    got = _gpu("DAS", P, "cubic", "auto", ("input-precision", "halfT"), _y_f32=True)   # widened + staged kernel
    assert qb.last_das_kernel() == "das_tiled" and rel_linf(got, ref) < TOL
    for path in ("auto", "generic"):
        got16 = _gpu("DAS", P, "cubic", path, ("input-precision", "halfT"))
        assert rel_linf(got16, ref) < 2e-3  # half2 output rounding (reference DASh writes half2)
    rsyn = _ora(oracle_c, "SYN", P, "cubic", x=xh)
    gsyn = _gpu("SYN", P, "cubic", "generic", ("input-precision", "halfT"), _y_f32=True)        # generic mixed types
    assert np.array_equal(gsyn, rsyn)
    gsyn = _gpu("SYN", P, "cubic", "auto", ("input-precision", "halfT"), _y_f32=True)           # fp16 data on the staged kernel
    assert qb.last_das_kernel() == "das_tiled" and rel_linf(gsyn, rsyn) < TOL


def test_generic_modulation(oracle_c):
    P = small_problem("PW", nz=13, nx=9, N=6, M=4, T=150)
    t0 = np.array([1e-7, -1e-7, 2e-7, 0.0])
    ref = _ora(oracle_c, "DAS", P, "cubic", t0=t0, fmod=5e6)
    for path in ("generic", "tiled"):
        got = _gpu("DAS", P, "cubic", path, ("modulation", 5e6), t0=t0)
        assert rel_linf(got, ref) < 5e-6, path


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA", "DV"])
@pytest.mark.parametrize("shape", [(40, 70, 1), (33, 37, 1), (16, 32, 3), (5, 3, 2), (1, 100, 1), (100, 1, 1)])
def test_tiled_nearest_bitexact_integer_data(oracle_c, kind, shape):
    """Integer-valued samples make every partial sum exact in fp32, so any tap-index mismatch shows as a
    bit difference: the tiled kernel must pick exactly the oracle's sample for every (pixel, rx, tx)."""
    import qups_b200
    nz, nx, ny = shape
    P = small_problem(kind, nz=nz, nx=nx, ny=ny, N=21, M=6, T=300, int_data=True, zlim=(2e-3, 14e-3))
    ref = _ora(oracle_c, "DAS", P, "nearest")
    got = _gpu("DAS", P, "nearest", "tiled")
    assert qups_b200.last_das_kernel() == "das_tiled"
    assert np.abs(ref).max() > 0
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA", "DV"])
@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_tiled_vs_oracle_tolerance(oracle_c, kind, interp):
    P = small_problem(kind, nz=70, nx=45, N=37, M=9, T=400, zlim=(2e-3, 14e-3))
    t0 = np.linspace(-3e-7, 3e-7, 9)
    ref = _ora(oracle_c, "DAS", P, interp, t0=t0)
    got = _gpu("DAS", P, interp, "tiled", t0=t0)
    gen = _gpu("DAS", P, interp, "generic", t0=t0)
    assert np.array_equal(gen, ref)
    assert rel_linf(got, ref) < TOL, rel_linf(got, ref)


@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_tiled_edges_out_of_range_and_small_windows(oracle_c, interp, monkeypatch):
    """Pixels whose delays fall before sample 1 / after sample T (extrapval 0), traces only partly in range,
    the cubic end-padding zone, and windows larger than the smem slot (slow path) must all match."""
    P = small_problem("FSA", nz=64, nx=40, N=20, M=5, T=96, zlim=(0.2e-3, 9e-3), pad=0, int_data=(interp == "nearest"))
    ref = _ora(oracle_c, "DAS", P, interp, t0=2e-6)  # t0 > 0 pushes shallow pixels before the first sample
    got = _gpu("DAS", P, interp, "tiled", t0=2e-6)
    if interp == "nearest":
        assert np.array_equal(got, ref)
    else:
        assert rel_linf(got, ref) < TOL
    monkeypatch.setenv("QUPS_B200_WMAX", "8")  # force most traces through the slow path
    got = _gpu("DAS", P, interp, "tiled", t0=2e-6)
    assert (np.array_equal(got, ref) if interp == "nearest" else rel_linf(got, ref) < TOL)
    monkeypatch.setenv("QUPS_B200_LANE_AXIS", "1")
    got = _gpu("DAS", P, interp, "tiled", t0=2e-6)
    assert (np.array_equal(got, ref) if interp == "nearest" else rel_linf(got, ref) < TOL)


def test_tiled_nan_pixel_and_frames(oracle_c):
    P = small_problem("DV", nz=40, nx=33, N=17, M=4, T=300, F=2, zlim=(2e-3, 14e-3))
    P["Pi"] = P["Pi"].copy()
    P["Pi"][:, 7, 5, 0] = np.nan
    ref = _ora(oracle_c, "DAS", P, "cubic")
    got = _gpu("DAS", P, "cubic", "tiled")
    assert got.shape == ref.shape
    assert ref[7, 5, 0].max() == 0 and np.all(got[7, 5, 0] == 0)
    assert rel_linf(got, ref) < TOL


def test_auto_dispatch_and_errors():
    import qups_b200
    from qups_b200 import QupsError
    P = small_problem("FC", nz=40, nx=33, N=17, M=4, T=300)
    _gpu("DAS", P, "cubic", "auto")
    assert qups_b200.last_das_kernel() == "das_tiled"
    _gpu("SYN", P, "cubic", "auto")
    assert qups_b200.last_das_kernel() == "das_tiled"      # one kept aperture: staged kernel (roles swapped for keep_rx)
    _gpu("BF", P, "cubic", "auto")
    assert qups_b200.last_das_kernel() == "das_generic"    # both kept: one output element per pair, nothing to stage for
    with pytest.raises(QupsError):
        _gpu("BF", P, "cubic", "tiled")
    with pytest.raises(ValueError):
        _gpu("DAS", P, "spline", "auto")
    with pytest.raises(AssertionError):
        _gpu("DAS", P, "cubic", "auto", ("apod", np.ones((3, 3), np.float32)))


def test_host_entry_point_matches_device_entry_point(oracle_c):
    """qups_das_host (host buffers, copies inside) == qups_das (device buffers)."""
    import ctypes as C
    from qups_b200 import _lib
    P = small_problem("FC", nz=40, nx=33, N=17, M=4, T=300)
    ref = _ora(oracle_c, "DAS", P, "cubic")
    f32 = np.float32
    Pi = np.asfortranarray(P["Pi"].reshape(3, -1, order="F").astype(f32))
    Pr = np.asfortranarray(P["Pr"].astype(f32))
    Pv4 = np.asfortranarray(np.concatenate([P["Pv"], np.zeros((1, 4))], 0).astype(f32))
    Nv = np.asfortranarray(P["Nv"].astype(f32))
    cinv = np.array([1.0 / f32(1540.0)], dtype=f32)
    x = np.asfortranarray(P["x"])
    p = _lib.DasParams()
    p.struct_size = C.sizeof(_lib.DasParams)
    p.dtype = _lib.F32
    p.I1, p.I2, p.I3 = P["Pi"].shape[1:]
    p.N, p.M, p.T, p.F, p.S = 17, 4, 300, 1, 0
    p.flag, p.vs, p.dv = _lib.CUBIC, 1, 0
    p.fs = P["fs"]
    y = np.empty(P["Pi"].shape[1:], dtype=np.complex64, order="F")
    acs = (C.c_uint64 * 6)(*([0] * 6))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for chunks in (0, 3):   # single shot, then the transmit-chunked copy/compute pipeline (accumulating launches)
        p.host_chunks = chunks
        y[:] = 0
        rc = _lib.lib().qups_das_host(C.byref(p), vp(y), vp(Pi), vp(Pr), vp(Pv4), vp(Nv), None, 0, vp(cinv), 1, acs, vp(x), 0)
        assert rc == 0, _lib.lib().qups_last_error()
        assert rel_linf(y.reshape(ref.shape, order="F"), ref) < TOL
    _lib.lib().qups_host_release()


def test_tiled_volume_matrix_array(oracle_c):
    """3-D ScanCartesian voxels + TransducerMatrix (BASELINE config 4, reduced): tiles over (I1, I2) per I3 slice."""
    import qups_b200
    from qups_b200 import synth
    Pr = synth.matrix_array(6, 5, 0.3e-3)
    g = np.linspace(-1e-3, 1e-3, 3)
    X, Y = np.meshgrid(g, g, indexing="ij")
    Pv = np.stack([X.reshape(-1), Y.reshape(-1), np.full(9, -4e-3)], 0)
    Nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, 9))
    ax = np.linspace(-1.5e-3, 1.5e-3, 19)
    Pi = synth.scan_cartesian(ax, np.linspace(3e-3, 9e-3, 37), ax[:7])
    rng = np.random.default_rng(9)
    x = np.asfortranarray((rng.standard_normal((400, 30, 9)) + 1j * rng.standard_normal((400, 30, 9))).astype(np.complex64))
    P = dict(Pi=Pi, Pr=Pr, Pv=Pv, Nv=Nv, x=x, t0=0.0, fs=25e6, c=1540.0, opts=("diverging-waves",))
    ref = _ora(oracle_c, "DAS", P, "cubic")
    got = _gpu("DAS", P, "cubic", "tiled")
    assert got.shape == ref.shape == (37, 19, 7, 1, 1)
    assert np.abs(ref).max() > 0 and rel_linf(got, ref) < TOL


@pytest.mark.parametrize("kind", ["FC", "PW"])
def test_tiled_real_apodization(oracle_c, kind):
    """Real apodization arrays (1 or 2, any broadcast shape) ride the staged kernel: bit-exact for nearest with
    integer-valued data and weights (masks), tolerance for cubic; complex weights fall back to the generic kernel."""
    import qups_b200
    rng = np.random.default_rng(12)
    P = small_problem(kind, nz=37, nx=45, N=21, M=6, T=300, int_data=True, zlim=(2e-3, 14e-3))
    Isz = P["Pi"].shape[1:]
    shapes = [Isz + (21, 1), (1, 1, 1, 21, 6), (Isz[0], Isz[1], 1, 1, 6), (1, 1, 1, 1, 6), Isz + (21, 6)]
    for shp in shapes:
        a1 = rng.integers(0, 3, shp).astype(np.float32)        # 0 / 1 / 2 : masks and integer weights
        ref = _ora(oracle_c, "DAS", P, "nearest", apod=[a1])
        got = _gpu("DAS", P, "nearest", "tiled", ("apod", a1))
        assert qups_b200.last_das_kernel() == "das_tiled"
        assert np.array_equal(got, ref), shp
    a1 = rng.uniform(0, 1, Isz + (21, 1)).astype(np.float32)
    a2 = (rng.uniform(0, 1, (1, 1, 1, 21, 6)) > 0.3).astype(np.float32)
    P2 = small_problem(kind, nz=37, nx=45, N=21, M=6, T=300, zlim=(2e-3, 14e-3))
    for interp in ("linear", "cubic"):
        ref = _ora(oracle_c, "DAS", P2, interp, apod=[a1, a2])
        got = _gpu("DAS", P2, interp, "tiled", ("apod", a1, "apod", a2))
        assert rel_linf(got, ref) < TOL, interp
    ac = (a1 * np.exp(0.2j)).astype(np.complex64)
    _gpu("DAS", P2, "cubic", "auto", ("apod", ac))
    assert qups_b200.last_das_kernel() == "das_generic"
    _gpu("DAS", P2, "cubic", "auto", ("apod", a1, "apod", a2, "apod", a2))
    assert qups_b200.last_das_kernel() == "das_generic"


def test_full_size_headline_parity_on_pixel_subsets(oracle_c):
    """BASELINE.json headline size (1024^2 px, 256 x 256, T = 2048, cubic): the full staged-kernel image is checked
    (a) against the C oracle on 96 random pixels (all 65 536 pairs each), (b) against the bit-exact generic kernel on
    4096 random pixels, (c) for linearity and a checksum of checksums across pixel slabs."""
    import torch
    import qups_b200
    from qups_b200 import synth, _lib
    P = synth.config_c2()
    x = synth.noise_cube(P.T, P.N, P.M, seed=0)
    f32 = np.float32
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).cuda()
    xd = torch.from_numpy(x).cuda()
    g = (dev(P.Pr), dev(P.Pv), dev(P.Nv))
    full = qups_b200.das_spec("DAS", dev(P.Pi), *g, xd, 0.0, P.fs, P.c0, "interp", "cubic")
    assert qups_b200.last_das_kernel() == "das_tiled"
    full = full.cpu().numpy().reshape(1024, 1024, order="F")
    scale = np.abs(full).max()
    rng = np.random.default_rng(3)
    iz, ix = rng.integers(0, 1024, 4096), rng.integers(0, 1024, 4096)
    sub = np.ascontiguousarray(P.Pi[:, iz, ix, 0]).reshape(3, -1, 1, 1)
    gen = qups_b200.das_spec("DAS", dev(sub), *g, xd, 0.0, P.fs, P.c0, "interp", "cubic", _path=_lib.PATH_GENERIC)
    gen = gen.cpu().numpy().reshape(-1)
    assert np.max(np.abs(gen - full[iz, ix])) / scale < TOL
    ref = oracle_c.das_spec("DAS", sub[:, :96], P.Pr, P.Pv, P.Nv, x, 0.0, P.fs, P.c0, interp="cubic").reshape(-1)
    assert np.array_equal(ref, gen[:96])                      # generic kernel == oracle, bit for bit, at full N x M
    assert np.max(np.abs(ref - full[iz[:96], ix[:96]])) / scale < TOL
    # linearity + slab additivity (checksum of checksums): two x-slabs beamformed separately == the full image
    left = qups_b200.das_spec("DAS", dev(P.Pi[:, :, :512]), *g, xd, 0.0, P.fs, P.c0, "interp", "cubic").cpu().numpy()
    assert np.max(np.abs(left.reshape(1024, 512, order="F") - full[:, :512])) / scale < TOL   # receive-split changes the sum order
    twice = qups_b200.das_spec("DAS", dev(P.Pi[:, :, 512:]), *g, 2 * xd, 0.0, P.fs, P.c0, "interp", "cubic").cpu().numpy()
    assert np.max(np.abs(twice.reshape(1024, 512, order="F") - 2 * full[:, 512:])) / scale < 2 * TOL


def test_full_size_kept_aperture_and_fused_mask_properties():
    """Headline size, size-independent properties: (a) summing the staged SYN / MUL outputs over the kept dimension gives the
    staged DAS image; (b) the in-kernel f-number mask equals the same mask passed as a dense array, bit for bit in the mask
    and to tolerance in the image; (c) on 2048 random pixels both equal the bit-exact generic kernel."""
    import torch
    import qups_b200
    from qups_b200 import synth, _lib, ultrasound as U
    P = synth.config_c2()
    x = synth.noise_cube(P.T, P.N, P.M, seed=0)
    f32 = np.float32
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(np.asarray(v, f32))).cuda()
    xd = torch.from_numpy(x).cuda()
    g = (dev(P.Pi), dev(P.Pr), dev(P.Pv), dev(P.Nv))
    full = qups_b200.das_spec("DAS", *g, xd, 0.0, P.fs, P.c0, "interp", "cubic").reshape(1024, 1024)
    scale = float(full.abs().max())
    rng = np.random.default_rng(11)
    iz, ix = rng.integers(0, 1024, 2048), rng.integers(0, 1024, 2048)
    sub = dev(np.ascontiguousarray(P.Pi[:, iz, ix, 0]).reshape(3, -1, 1, 1))
    for fun, axis in (("SYN", 3), ("MUL", 4)):
        b = qups_b200.das_spec(fun, *g, xd, 0.0, P.fs, P.c0, "interp", "cubic")
        assert qups_b200.last_das_kernel() == "das_tiled"
        tot = b.sum(dim=(3, 4)).reshape(1024, 1024)
        assert float((tot - full).abs().max()) / scale < 2 * TOL, fun
        gen = qups_b200.das_spec(fun, sub, *g[1:], xd, 0.0, P.fs, P.c0, "interp", "cubic", _path=_lib.PATH_GENERIC)
        pick = b.reshape(1024, 1024, b.shape[3], b.shape[4])[torch.from_numpy(iz).cuda(), torch.from_numpy(ix).cuda()]
        assert float((pick.reshape(gen.shape[0], -1) - gen.reshape(gen.shape[0], -1)).abs().max()) / float(gen.abs().max()) < TOL, fun
        del b, tot, gen, pick
        torch.cuda.empty_cache()
    us = U.UltrasoundSystem(tx=P.Pr, rx=P.Pr, seq=U.Sequence("FC", P.Pv), scan=P.Pi, fs=P.fs)
    spec = us.apApertureGrowth(1.5)
    fused = qups_b200.das_spec("DAS", *g, xd, 0.0, P.fs, P.c0, "interp", "cubic", "apod", spec)
    dense = spec.dense(g[0], g[1], which="rx")
    viaarr = qups_b200.das_spec("DAS", *g, xd, 0.0, P.fs, P.c0, "interp", "cubic", "apod", dense)
    assert float((fused - viaarr).abs().max()) / float(viaarr.abs().max()) < TOL
    genm = qups_b200.das_spec("DAS", sub, *g[1:], xd, 0.0, P.fs, P.c0, "interp", "cubic", "apod", spec, _path=_lib.PATH_GENERIC)
    assert qups_b200.last_das_kernel() == "das_generic+apod_generate"
    pick = fused.reshape(1024, 1024)[torch.from_numpy(iz).cuda(), torch.from_numpy(ix).cuda()]
    assert float((pick - genm.reshape(-1)).abs().max()) / float(genm.abs().max()) < TOL


def test_empty_and_degenerate_shapes_through_the_c_abi(oracle_c):
    """Empty and degenerate inputs at the boundary (the reference's sum over an empty aperture is zeros(size of the image);
    an empty image is a no-op), and the smallest ragged grids: one pixel, one row, one column, one receive, one transmit."""
    import ctypes as C
    from qups_b200 import _lib
    f32 = np.float32
    acs = (C.c_uint64 * 6)(*([0] * 6))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def host(P, N, M, interp=_lib.CUBIC):
        Pi = np.asfortranarray(P["Pi"].reshape(3, -1, order="F").astype(f32))
        Pr = np.asfortranarray(P["Pr"][:, :max(N, 1)].astype(f32))
        Mf = P["Pv"].shape[1]
        Pv4 = np.asfortranarray(np.concatenate([P["Pv"], np.broadcast_to(np.asarray(P["t0"], float).reshape(1, -1), (1, Mf))], 0)[:, :max(M, 1)].astype(f32))
        Nv = np.asfortranarray(P["Nv"][:, :max(M, 1)].astype(f32))
        cinv = np.array([1.0 / f32(P["c"])], dtype=f32)
        x = np.asfortranarray(P["x"][:, :max(N, 1), :max(M, 1)])
        p = _lib.DasParams()
        p.struct_size = C.sizeof(_lib.DasParams)
        p.dtype = _lib.F32
        p.I1, p.I2, p.I3 = P["Pi"].shape[1:]
        p.N, p.M, p.T, p.F, p.S = N, M, P["x"].shape[0], 1, 0
        p.flag, p.vs, p.dv = interp, 1, 0
        p.fs = P["fs"]
        y = np.full(P["Pi"].shape[1:], 7 + 7j, dtype=np.complex64, order="F")
        rc = _lib.lib().qups_das_host(C.byref(p), vp(y), vp(Pi), vp(Pr), vp(Pv4), vp(Nv), None, 0, vp(cinv), 1, acs, vp(x), 0)
        assert rc == 0, _lib.lib().qups_last_error()
        return y

    P = small_problem("FC", nz=9, nx=5, N=6, M=3, T=200)
    assert np.all(host(P, 0, 3) == 0)      # no receives: empty sum
    assert np.all(host(P, 6, 0) == 0)      # no transmits
    Pe = dict(P)
    Pe["Pi"] = P["Pi"][:, :0]              # empty image: nothing is written, the call succeeds
    assert host(Pe, 6, 3).size == 0
    for nz, nx, N, M in ((1, 1, 1, 1), (1, 7, 6, 3), (5, 1, 6, 3), (9, 5, 1, 3), (9, 5, 6, 1)):
        Q = small_problem("FC", nz=nz, nx=nx, N=N, M=M, T=200)
        ref = _ora(oracle_c, "DAS", Q, "cubic")
        got = host(Q, N, M)
        assert rel_linf(got.reshape(ref.shape, order="F"), ref) < TOL, (nz, nx, N, M)
    _lib.lib().qups_host_release()
