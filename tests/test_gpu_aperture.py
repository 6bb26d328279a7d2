"""GPU parity of the aperture-domain post-processing kernels vs NumPy float64 restatements of
kern/cohfac.m, kern/dmas.m, kern/pcf.m, kern/slsc.m, in the reference tests' style (test/KernTest.m:220-242 only
smoke-tests them; here values are pinned).  fp32 tolerance 2e-5 relative, fp64 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(shape, dtype=np.complex64, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)


def _close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))  # 0/0 (an all-zero aperture) is NaN in the reference too
    a, b = np.nan_to_num(a), np.nan_to_num(b)
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("dtype,tol", [(np.complex64, 2e-5), (np.complex128, 1e-12)])
def test_cohfac_dmas_pcf_slsc(dim, dtype, tol):
    import qups_b200
    from oracle import aperture_np as ap
    b = _data((9, 12, 7), dtype, seed=dim)
    _close(qups_b200.cohfac(b, dim), ap.cohfac(b, dim), tol)
    _close(qups_b200.dmas(b, dim), ap.dmas(b, dim), tol * 10)
    _close(qups_b200.dmas(b, dim, 3), ap.dmas(b, dim, 3), tol * 10)
    _close(qups_b200.dmas(b, dim, [2, 5, 40]), ap.dmas(b, dim, [2, 5, 40]), tol * 10)
    w, sf = qups_b200.pcf(b, dim, 0.7)
    wr, sfr = ap.pcf(b, dim, 0.7)
    _close(w, wr, tol * 5)
    _close(sf, sfr, tol * 5)
    for method in ("average", "ensemble"):
        _close(qups_b200.slsc(b, dim, 3, method), ap.slsc(b, dim, 3, method), tol * 10)
        _close(qups_b200.slsc(b, dim, [1, 2, 6], method), ap.slsc(b, dim, [1, 2, 6], method), tol * 10)
        _close(qups_b200.slsc(b, dim, None, method), ap.slsc(b, dim, None, method), tol * 10)


def test_defaults_and_known_answers():
    import qups_b200
    b = np.ones((5, 8), np.complex64) * (2 - 1j)
    assert np.allclose(qups_b200.cohfac(b), 1.0)                      # perfectly coherent aperture
    assert np.allclose(qups_b200.slsc(b, 2, 2, "average").real, 1.0, atol=1e-6)
    assert np.allclose(qups_b200.slsc(b, 2, 2, "ensemble").real, 1.0, atol=1e-6)
    w, sf = qups_b200.pcf(b)
    assert np.allclose(sf, 0, atol=1e-6) and np.allclose(w, 1.0, atol=1e-6)
    z = np.zeros((4, 6), np.complex64)
    assert np.all(qups_b200.slsc(z, 2, 2, "ensemble") == 0)           # nan2zero
    with pytest.raises(ValueError):
        qups_b200.pcf(np.ones((3, 3), np.float32))


def test_coherence_factor_on_beamformed_receive_cube(oracle_c):
    """The intended pipeline: b = DAS(..., keep_rx) (fun 'SYN') -> cohfac over the receive dimension."""
    import qups_b200
    from oracle import aperture_np as ap
    from tests.util import small_problem, oracle_kwargs
    f32 = np.float32
    P = small_problem("FC", nz=24, nx=20, N=10, M=4, T=200)
    from qups_b200 import _lib
    bn = qups_b200.das_spec("SYN", P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"], P["t0"],
                            P["fs"], P["c"], *P["opts"], "interp", "cubic", _path=_lib.PATH_GENERIC)
    ref = oracle_c.das_spec("SYN", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="cubic",
                            **oracle_kwargs(P["opts"]))[..., 0]
    assert np.array_equal(bn, ref)
    cf = qups_b200.cohfac(bn, 4)
    assert cf.shape == bn.shape[:3] + (1,) + bn.shape[4:]
    _close(cf, ap.cohfac(ref.astype(np.complex128), 4), 2e-5)
    assert np.all((cf >= 0) & (cf <= 1 + 1e-5) | np.isnan(cf))
