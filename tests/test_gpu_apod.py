"""GPU parity tests of the closed-form apodization path (src/UltrasoundSystem.m:4892-5429):
qups_apod_generate vs the canonical-fp32 oracle (bit-exact for masks), and DAS with the generator FUSED into the
kernel vs DAS fed the dense array through the oracle (kern/das_spec.m:473-478 semantics: a .* interp1)."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu
f32 = np.float32


def _us(P, kind="FC", ae=None):
    from qups_b200 import ultrasound as U
    N = P["Pr"].shape[1]
    nn = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, N))
    if ae is not None:
        nn = np.stack([np.sin(np.deg2rad(ae)), 0 * ae, np.cos(np.deg2rad(ae))])
    return U.UltrasoundSystem(tx=P["Pr"], rx=P["Pr"], seq=U.Sequence(kind, P["Pv"]), scan=P["Pi"], fs=P["fs"], rx_normal=nn, rx_angle=ae)


def _dense_oracle(name, P, us, **kw):
    from oracle import apod_np as ap
    Pi, Pn = P["Pi"], P["Pr"]
    xv = P["Pv"][0]
    if name == "acc": return ap.apAcceptanceAngle(Pi, Pn, us._rx_normals(), kw["theta"], literal=False)
    if name == "cos": return ap.apCosineAngle(Pi, Pn, us._rx_normals(), kw["theta"], literal=False)
    if name == "grow": return ap.apApertureGrowth(Pi, Pn, ae=us.rx_angle, f=kw["f"], Dmax=kw.get("Dmax", np.inf), literal=False)
    if name == "scan": return ap.apScanline(Pi, xv, kw["tol"], literal=False)
    if name == "trans": return ap.apTranslatingAperture(Pi, xv, Pn[0], kw["tol"], literal=False)
    if name == "para": return ap.apTxParallelogram(Pi, kw["theta"], kw["phi"], kw["bounds"], literal=False)
    raise KeyError(name)


def _spec(name, us, **kw):
    if name == "acc": return us.apAcceptanceAngle(kw["theta"])
    if name == "cos": return us.apCosineAngle(kw["theta"])
    if name == "grow": return us.apApertureGrowth(kw["f"], kw.get("Dmax", np.inf))
    if name == "scan": return us.apScanline(kw["tol"])
    if name == "trans": return us.apTranslatingAperture(kw["tol"])
    if name == "para": return us.apTxParallelogram(kw["theta"], kw["phi"], kw["bounds"])
    raise KeyError(name)


CASES = [("acc", dict(theta=25.0)), ("cos", dict(theta=35.0)), ("grow", dict(f=1.2, Dmax=2.5e-3)),
         ("scan", dict(tol=0.26e-3)), ("trans", dict(tol=(0.3e-3, 1.1e-3))),
         ("para", dict(theta=np.linspace(-12, 12, 5), phi=(-4.0, 4.0), bounds=(-1.5e-3, 1.5e-3)))]


@pytest.mark.parametrize("name,kw", CASES)
def test_generate_matches_oracle(name, kw):
    P = small_problem("FC", nz=37, nx=41, N=12, M=5, T=200)
    us = _us(P)
    spec = _spec(name, us, **kw)
    ref = _dense_oracle(name, P, us, **kw)
    Pi = P["Pi"].astype(f32)
    got = []
    if spec.rx_kind: got.append(spec.dense(Pi, P["Pr"].astype(f32), which="rx")[..., None])
    if spec.tx_kind: got.append(spec.dense(Pi, M=P["Pv"].shape[1], which="tx"))
    g = got[0] if len(got) == 1 else got[0] * got[1]
    g = g.reshape(ref.shape) if g.size == ref.size else g
    assert g.shape == ref.shape
    if name == "cos":
        assert np.max(np.abs(g - ref)) < 2e-6
    else:
        assert np.array_equal(g, ref.astype(f32)), float(np.mean(g != ref))
    assert 0.02 < float(np.mean(g != 0)) < 0.98  # the case exercises both sides of the mask


def test_generate_nonplanar_growth_and_complex_output():
    P = small_problem("FC", nz=33, nx=29, N=10, M=4, T=200)
    ae = np.linspace(-25, 25, 10)
    us = _us(P, ae=ae)
    spec = us.apApertureGrowth(1.0)
    ref = _dense_oracle("grow", P, us, f=1.0)
    got = spec.dense(P["Pi"].astype(f32), P["Pr"].astype(f32), which="rx")
    assert np.array_equal(got, ref.astype(f32))
    gc = spec.dense(P["Pi"].astype(f32), P["Pr"].astype(f32), which="rx", complex_=True)
    assert gc.dtype == np.complex64 and np.array_equal(gc.real, got) and not np.any(gc.imag)


def _das(P, interp, *apod, path="auto", fun="DAS"):
    import qups_b200
    from qups_b200 import _lib
    pth = {"generic": _lib.PATH_GENERIC, "tiled": _lib.PATH_TILED, "auto": _lib.PATH_AUTO}[path]
    extra = sum((("apod", a) for a in apod), ())
    out = qups_b200.das_spec(fun, P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"],
                             P["t0"], P["fs"], P["c"], *P["opts"], "interp", interp, *extra, _path=pth)
    return out, qups_b200.last_das_kernel()


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("interp", ["nearest", "cubic"])
def test_fused_das_matches_dense_oracle(oracle_c, name, kw, interp):
    """DAS with the generator evaluated in-kernel == oracle DAS with the dense array as 'apod'."""
    P = small_problem("FC", nz=70, nx=66, N=20, M=5, T=400, zlim=(2e-3, 14e-3), int_data=(interp == "nearest"))
    us = _us(P)
    spec = _spec(name, us, **kw)
    A = _dense_oracle(name, P, us, **kw).astype(f32)
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp,
                            apod=[A], **oracle_kwargs(P["opts"]))[..., 0]
    got, k = _das(P, interp, spec, path="tiled")
    assert k == "das_tiled"
    if interp == "nearest" and name != "cos":
        assert np.array_equal(got, ref)  # integer data, 0/1 weights: bit-exact, a flipped mask entry would show
    else:
        assert rel_linf(got, ref) < 1e-5
    assert np.any(got != 0)


def test_fused_plus_array_and_rx_tx_product(oracle_c):
    P = small_problem("FC", nz=48, nx=40, N=16, M=6, T=300, zlim=(2e-3, 12e-3))
    us = _us(P)
    rng = np.random.default_rng(5)
    W = rng.uniform(0.2, 1.0, (1, P["Pi"].shape[2], 1, 1, 6)).astype(f32)  # e.g. apMultiline-shaped
    acc, scan = us.apAcceptanceAngle(30.0), us.apScanline(0.9e-3)
    A1 = _dense_oracle("acc", P, us, theta=30.0).astype(f32)
    A2 = _dense_oracle("scan", P, us, tol=0.9e-3).astype(f32)
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                            apod=[A1, A2, W], **oracle_kwargs(P["opts"]))[..., 0]
    got, k = _das(P, "linear", acc, scan, W, path="tiled")
    assert k == "das_tiled" and rel_linf(got, ref) < 1e-5


@pytest.mark.parametrize("fun", ["SYN", "MUL", "BF"])
def test_fused_falls_back_to_dense_for_kept_apertures(oracle_c, fun):
    """Outside the staged kernel's envelope the library materialises the dense weights itself (still on the GPU)."""
    P = small_problem("FC", nz=21, nx=17, N=8, M=4, T=200)
    us = _us(P)
    spec = us.apTranslatingAperture((0.6e-3, 1.0e-3))
    A = _dense_oracle("trans", P, us, tol=(0.6e-3, 1.0e-3)).astype(f32)
    ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="cubic",
                            apod=[A], **oracle_kwargs(P["opts"]))[..., 0]
    got, k = _das(P, "cubic", spec, fun=fun)
    assert k == "das_generic+apod_generate"
    assert rel_linf(got, ref) < 1e-6


def test_fused_errors():
    import qups_b200
    from qups_b200 import kern
    P = small_problem("FC", nz=9, nx=9, N=4, M=3, T=100)
    us = _us(P)
    with pytest.raises(qups_b200.QupsError):
        us.apAcceptanceAngle(30.0).merged(us.apCosineAngle(30.0))
    bad = kern.FusedApod(rx_kind=1, rx_p=(0.5,), rx_aux=None)
    with pytest.raises(qups_b200.QupsError):
        _das(P, "linear", bad)
    with pytest.raises(qups_b200.QupsError):  # fp64 geometry is outside the closed-form path
        qups_b200.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"].astype(np.complex128), P["t0"], P["fs"], P["c"],
                           "apod", us.apAcceptanceAngle(30.0))


def test_fused_scanline_skips_work_and_matches(oracle_c):
    """Scanline imaging at a realistic shape: every tile uses a handful of transmits; result identical to the dense mask."""
    from qups_b200 import synth
    Pc = synth.config_c2(128, 128, 32, 32, 768)
    x = synth.noise_cube(Pc.T, Pc.N, Pc.M, seed=2)
    P = dict(Pi=Pc.Pi, Pr=Pc.Pr, Pv=Pc.Pv, Nv=Pc.Nv, x=x, t0=Pc.t0, fs=Pc.fs, c=Pc.c0, opts=Pc.opts)
    us = _us(P)
    tol = float(abs(Pc.Pv[0, 1] - Pc.Pv[0, 0])) * 0.6
    spec = us.apScanline(tol).merged(us.apApertureGrowth(1.5))
    A1 = _dense_oracle("scan", P, us, tol=tol).astype(f32)
    A2 = _dense_oracle("grow", P, us, f=1.5).astype(f32)
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], x, P["t0"], P["fs"], P["c"], interp="cubic",
                            apod=[A1, A2], **oracle_kwargs(P["opts"]))[..., 0]
    got, k = _das(P, "cubic", spec, path="tiled")
    assert k == "das_tiled" and rel_linf(got, ref) < 1e-5
