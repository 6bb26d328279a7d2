"""GPU parity of qups_das_cohfac (coherence mode of the staged kernel, csrc/das_tiled.cu KEEP = 3): the DAS image and the
coherence factor of the per-receive images in one pass, against the oracle's SYN cube (kern/das_spec.m:483-521) reduced by the
NumPy restatement of kern/cohfac.m.  Image: <= 1e-5 relative L-inf (nearest: bit-exact on integer data); factor: <= 1e-4
absolute (a ratio of two sums of ~N terms in fp32)."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu
f32 = np.float32


def _run(P, interp, **kw):
    import qups_b200
    return qups_b200.das_spec("DAS+cohfac", P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32), P["x"],
                              P["t0"], P["fs"], P["c"], *P["opts"], "interp", interp, **kw)


def _oracle(oracle_c, P, interp):
    from oracle import aperture_np
    b = oracle_c.das_spec("SYN", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp,
                          **oracle_kwargs(P["opts"]))
    b = np.squeeze(np.asarray(b, np.complex128))                 # I1 x I2 x N
    with np.errstate(invalid="ignore", divide="ignore"):
        return b.sum(-1), np.squeeze(aperture_np.cohfac(b, b.ndim))


@pytest.mark.parametrize("kind", ["FC", "PW", "DV", "FSA"])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_das_cohfac_matches_oracle(oracle_c, kind, interp):
    import qups_b200
    # N = 21 receives (three groups of 8, the last one short), M = 7 transmits (odd: the last transmit pair is half empty),
    # focal plane inside the image for FC, per-transmit t0
    P = small_problem(kind, nz=70, nx=45, N=21, M=7, T=520, zlim=(2e-3, 14e-3), int_data=(interp == "nearest"),
                      t0=np.linspace(0.0, 0.3e-6, 7))
    y, cf = _run(P, interp)
    assert qups_b200.last_das_kernel() == "das_tiled"
    yr, cfr = _oracle(oracle_c, P, interp)
    y, cf = np.squeeze(y), np.squeeze(cf)
    assert y.shape == yr.shape and cf.shape == cfr.shape
    if interp == "nearest":
        assert np.array_equal(y, yr.astype(np.complex64))
    else:
        assert rel_linf(y, yr) <= 1e-5
    assert np.array_equal(np.isnan(cf), np.isnan(cfr))
    assert np.nanmax(np.abs(cf - cfr)) <= 1e-4
    assert np.nanmax(cf) <= 1.0 + 1e-5 and np.nanmin(cf) >= 0.0


def test_das_cohfac_receive_split_and_fallback(oracle_c):
    """M >= 16 transmits: receive split with the bounds pre-pass (partials per split summed by das_cf_reduce_kernel); and the
    same call forced off the staged kernel (QUPS_PATH_GENERIC): keep_rx cube + cohfac reduction, same numbers."""
    from qups_b200 import _lib
    P = small_problem("FC", nz=60, nx=40, N=40, M=18, T=520, zlim=(2e-3, 14e-3))
    y, cf = _run(P, "cubic")
    yr, cfr = _oracle(oracle_c, P, "cubic")
    assert rel_linf(np.squeeze(y), yr) <= 1e-5 and np.nanmax(np.abs(np.squeeze(cf) - cfr)) <= 1e-4
    yg, cfg = _run(P, "cubic", _path=_lib.PATH_GENERIC)
    assert rel_linf(np.squeeze(yg), yr) <= 1e-5 and np.nanmax(np.abs(np.squeeze(cfg) - cfr)) <= 1e-4


def test_das_cohfac_identical_receives_give_one():
    """All receives at the same position see the same trace: b_n identical, cohfac == 1 wherever the image is non-zero."""
    P = small_problem("PW", nz=40, nx=30, N=16, M=4, T=400, zlim=(2e-3, 10e-3))
    P["Pr"] = np.repeat(P["Pr"][:, :1], 16, 1)
    P["x"] = np.asfortranarray(np.repeat(P["x"][:, :1, :], 16, 1))
    y, cf = _run(P, "linear")
    m = np.abs(np.squeeze(y)) > 1e-3 * np.abs(y).max()
    assert np.max(np.abs(np.squeeze(cf)[m] - 1.0)) < 1e-5
