"""GPU parity of the staged DAS kernel across its run-time geometry choices (tile shape, lane patch, ring depth / slot
length — csrc/das_tiled.cu picks them from the pixel spacing) and on the grids that exercise them: anisotropic / coarse
pixel grids, volumes, focal-plane sign flips (dual-cluster windows), per-transmit t0.  Oracle = CPU restatement of
kern/das_spec.m:393-482; nearest is bit-exact (integer data), linear / cubic <= 1e-5 relative L-inf."""
import os

import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu
f32 = np.float32


def _tiled(P, interp, **env):
    import qups_b200
    from qups_b200 import _lib
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        out = qups_b200.das_spec("DAS", P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"],
                                 P["t0"], P["fs"], P["c"], *P["opts"], "interp", interp, _path=_lib.PATH_TILED)
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    assert qups_b200.last_das_kernel() == "das_tiled"
    return out


def _oracle(oracle_c, P, interp):
    return oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp,
                             **oracle_kwargs(P["opts"]))[..., 0]


@pytest.mark.parametrize("tile", ["32,8", "8,2", "256,32", "1,1", "16,4", "64,8", "4,4"])
@pytest.mark.parametrize("kind", ["FC", "PW"])
def test_forced_tile_shapes(oracle_c, tile, kind):
    P = small_problem(kind, nz=83, nx=71, N=20, M=6, T=420, zlim=(2e-3, 15e-3))
    ref = _oracle(oracle_c, P, "cubic")
    got = _tiled(P, "cubic", QUPS_B200_TILE=tile)
    assert rel_linf(got, ref) < 1e-5
    Pn = small_problem(kind, nz=83, nx=71, N=20, M=6, T=420, zlim=(2e-3, 15e-3), int_data=True)
    assert np.array_equal(_tiled(Pn, "nearest", QUPS_B200_TILE=tile), _oracle(oracle_c, Pn, "nearest"))


@pytest.mark.parametrize("ring", [(2, 256), (3, 170), (4, 128), (2, 64), (4, 32), (2, 512)])
def test_forced_ring_geometry(oracle_c, ring):
    """Short slots push traces onto the EDGE / SLOW / dual-window paths; long ones onto one CTA per SM."""
    P = small_problem("FC", nz=90, nx=64, N=18, M=7, T=500, zlim=(2e-3, 16e-3), t0=np.linspace(-3e-7, 4e-7, 7))
    ref = _oracle(oracle_c, P, "cubic")
    got = _tiled(P, "cubic", QUPS_B200_STAGES=ring[0], QUPS_B200_WMAX=ring[1])
    assert rel_linf(got, ref) < 1e-5
    for interp in ("linear", "nearest"):
        assert rel_linf(_tiled(P, interp, QUPS_B200_STAGES=ring[0], QUPS_B200_WMAX=ring[1]), _oracle(oracle_c, P, interp)) < 1e-5


@pytest.mark.parametrize("aniso", [(8.0, 1.0), (1.0, 6.0), (16.0, 1.0)])
def test_anisotropic_grids_pick_their_own_shape(oracle_c, aniso):
    """Coarse lateral (or axial) sampling: the automatic tile / ring choice must stay exact (it only affects speed)."""
    from qups_b200 import synth
    dz = 1540.0 / 20e6 / 4
    nx, nz = 48, 96
    xs = (np.arange(nx) - nx / 2) * dz * aniso[0]
    zs = 2e-3 + np.arange(nz) * dz * aniso[1]
    P = small_problem("FC", N=16, M=5, T=900)
    P["Pi"] = synth.scan_cartesian(xs, zs)
    ref = _oracle(oracle_c, P, "cubic")
    assert np.any(ref != 0)
    assert rel_linf(_tiled(P, "cubic"), ref) < 1e-5


def test_volume_and_focal_plane_inside_every_tile(oracle_c):
    """3-D grid (slices along I3) with the foci in the middle of the depth range: dv flips sign inside the tiles."""
    P = small_problem("FC", nz=40, nx=36, ny=3, N=12, M=5, T=300, zlim=(3e-3, 7e-3))
    ref = _oracle(oracle_c, P, "cubic")
    assert rel_linf(_tiled(P, "cubic"), ref) < 1e-5
    assert rel_linf(_tiled(P, "cubic", QUPS_B200_WMAX=48, QUPS_B200_STAGES=4), ref) < 1e-5  # forces split windows
    Pn = small_problem("FC", nz=40, nx=36, ny=3, N=12, M=5, T=300, zlim=(3e-3, 7e-3), int_data=True)
    assert np.array_equal(_tiled(Pn, "nearest", QUPS_B200_WMAX=48), _oracle(oracle_c, Pn, "nearest"))


# ---- kept apertures on the staged kernel: MUL (keep_tx) and SYN (keep_rx, roles of the apertures swapped) --------------
def _tiled_fun(fun, P, interp, **env):
    import qups_b200
    from qups_b200 import _lib
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        out = qups_b200.das_spec(fun, P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"],
                                 P["t0"], P["fs"], P["c"], *P["opts"], "interp", interp, _path=_lib.PATH_TILED)
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    assert qups_b200.last_das_kernel() == "das_tiled"
    return out


@pytest.mark.parametrize("fun", ["SYN", "MUL"])
@pytest.mark.parametrize("kind", ["FC", "PW", "DV", "FSA"])
def test_kept_aperture_on_the_staged_kernel(oracle_c, fun, kind):
    M = 21 if kind != "FSA" else 19
    P = small_problem(kind, nz=70, nx=45, N=19, M=M, T=400, zlim=(2e-3, 14e-3), t0=np.linspace(-2e-7, 3e-7, M))
    for interp in ("cubic", "linear"):
        ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp,
                                **oracle_kwargs(P["opts"]))[..., 0]
        got = _tiled_fun(fun, P, interp)
        assert got.shape == ref.shape
        assert rel_linf(got, ref) < 1e-5, (fun, kind, interp)
    Pn = small_problem(kind, nz=70, nx=45, N=19, M=M, T=400, zlim=(2e-3, 14e-3), int_data=True)
    refn = oracle_c.das_spec(fun, Pn["Pi"], Pn["Pr"], Pn["Pv"], Pn["Nv"], Pn["x"], Pn["t0"], Pn["fs"], Pn["c"], interp="nearest",
                             **oracle_kwargs(Pn["opts"]))[..., 0]
    assert np.array_equal(_tiled_fun(fun, Pn, "nearest"), refn)


@pytest.mark.parametrize("fun", ["SYN", "MUL"])
def test_kept_aperture_short_slots_and_sum_consistency(oracle_c, fun):
    """Short slots force the EDGE / dual-window / global-memory paths; summing the kept dimension gives back DAS."""
    P = small_problem("FC", nz=64, nx=40, ny=2, N=12, M=5, T=300, zlim=(3e-3, 8e-3))
    ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="cubic",
                            **oracle_kwargs(P["opts"]))[..., 0]
    for env in (dict(), dict(QUPS_B200_WMAX=48, QUPS_B200_STAGES=4), dict(QUPS_B200_WMAX=24, QUPS_B200_STAGES=2), dict(QUPS_B200_TILE="8,2")):
        got = _tiled_fun(fun, P, "cubic", **env)
        assert rel_linf(got, ref) < 1e-5, env
    das = _tiled(P, "cubic")
    assert rel_linf(got.sum(axis=(3, 4)), das[..., 0, 0] if das.ndim == 5 else das) < 1e-5


@pytest.mark.parametrize("kind", ["FC", "PW", "DV"])
@pytest.mark.parametrize("interp", ["nearest", "cubic"])
def test_receive_split_with_bounds_prepass(oracle_c, kind, interp):
    """M >= 16 transmits and several receive tiles on a small image: the launcher splits the receive axis down to one tile per
    CTA and takes the per-tile path-length bounds from das_bounds_kernel (csrc/das_tiled.cu) instead of the in-kernel phase 0.
    Same result as with the pre-pass disabled (the bounds only select windows; the receive split — hence the summation order —
    differs between the two launch policies, so bit-equality holds for integer data only) and as the oracle; also with a real
    apodization array, per-transmit t0 and two frames."""
    import qups_b200
    from qups_b200 import _lib
    P = small_problem(kind, nz=70, nx=45, N=40, M=19, T=520, zlim=(2e-3, 14e-3), int_data=(interp == "nearest"), F=2,
                      t0=np.linspace(0.0, 0.4e-6, 19))
    got = _tiled(P, interp)
    n0 = _lib.launch_count()
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp,
                            **oracle_kwargs(P["opts"]))
    nob = _tiled(P, interp, QUPS_B200_NOBOUNDS=1)
    assert np.array_equal(got, nob) if interp == "nearest" else rel_linf(got, nob) <= 2e-6
    got, ref = np.squeeze(got), np.squeeze(ref)
    assert got.shape == ref.shape
    if interp == "nearest":
        assert np.array_equal(got, ref)
    else:
        assert rel_linf(got, ref) <= 1e-5
    # a real apodization array over (pixels x receives) rides along (NAP = 1)
    rng = np.random.default_rng(5)
    A = rng.uniform(0.2, 1.0, P["Pi"].shape[1:] + (40, 1)).astype(f32)
    ga = qups_b200.das_spec("DAS", P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"], P["t0"],
                            P["fs"], P["c"], *P["opts"], "interp", interp, "apod", A, _path=_lib.PATH_TILED)
    assert qups_b200.last_das_kernel() == "das_tiled"
    ra = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp, apod=[A],
                           **oracle_kwargs(P["opts"]))
    assert rel_linf(np.squeeze(ga), np.squeeze(ra)) <= 1e-5
    assert n0 > 0
