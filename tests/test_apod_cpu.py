"""CPU tests of the apodization-generator oracle (oracle/apod_np.py): the literal float64 restatement of
src/UltrasoundSystem.m:4892-5429 against analytic answers, and against the canonical fp32 sequence the device
functions follow (they may differ only for pixels within rounding distance of a mask boundary)."""
import numpy as np
import pytest

from oracle import apod_np as ap
from qups_b200 import synth


def _geom(N=16, nz=20, nx=24, pitch=0.3e-3):
    Pn = synth.linear_array(N, pitch)
    nn = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, N))
    Pi = synth.scan_cartesian(np.linspace(-3e-3, 3e-3, nx), np.linspace(1e-3, 12e-3, nz))
    return Pi, Pn, nn


def test_acceptance_angle_is_a_cone():
    Pi, Pn, nn = _geom()
    A = ap.apAcceptanceAngle(Pi, Pn, nn, theta=30.0)
    assert A.shape == Pi.shape[1:] + (Pn.shape[1],)
    # analytic: |x - xn| <= z tan(theta) for a +z normal
    X, Z = Pi[0][..., None], Pi[2][..., None]
    want = np.abs(X - Pn[0].reshape(1, 1, 1, -1)) <= Z * np.tan(np.deg2rad(30.0)) * (1 + 1e-12)
    assert np.mean(A != want) < 1e-3


def test_cosine_angle_limits():
    Pi, Pn, nn = _geom()
    W = ap.apCosineAngle(Pi, Pn, nn, theta=45.0)
    assert W.min() >= 0 and W.max() <= 1
    # directly under an element the weight is 1, beyond theta it is 0
    P1 = np.array([Pn[0, 3], 0, 5e-3]).reshape(3, 1, 1, 1)
    assert ap.apCosineAngle(P1, Pn[:, 3:4], nn[:, 3:4], 45.0)[0, 0, 0, 0] == pytest.approx(1.0)
    P2 = np.array([Pn[0, 3] + 6e-3, 0, 5e-3]).reshape(3, 1, 1, 1)
    assert ap.apCosineAngle(P2, Pn[:, 3:4], nn[:, 3:4], 45.0)[0, 0, 0, 0] == pytest.approx(0.0, abs=1e-12)


def test_aperture_growth_fnumber_and_nonplanar_rotation():
    Pi, Pn, _ = _geom()
    A = ap.apApertureGrowth(Pi, Pn, f=1.5, Dmax=3e-3)
    X, Z = Pi[0][..., None], Pi[2][..., None]
    d2 = np.abs(2 * (Pn[0].reshape(1, 1, 1, -1) - X))
    assert np.array_equal(A, ((Z > 1.5 * d2) & (d2 < 3e-3)).astype(float))
    # the algebraic rotation the kernel uses equals the reference's atan2d/sind/cosd form away from mask boundaries
    ae = np.linspace(-20, 20, Pn.shape[1])
    L = ap.apApertureGrowth(Pi, Pn, ae=ae, f=1.0, literal=True)
    C = ap.apApertureGrowth(Pi, Pn, ae=ae, f=1.0, literal=False)
    assert np.mean(L != C) < 2e-3


@pytest.mark.parametrize("gen", ["acc", "scan", "trans", "para"])
def test_fp32_canonical_matches_literal_away_from_boundaries(gen):
    Pi, Pn, nn = _geom()
    xv = np.linspace(-2e-3, 2e-3, 9)
    if gen == "acc":
        L, C = (ap.apAcceptanceAngle(Pi, Pn, nn, 40.0, literal=l) for l in (True, False))
    elif gen == "scan":
        L, C = (ap.apScanline(Pi, xv, 0.3e-3, literal=l) for l in (True, False))
    elif gen == "trans":
        L, C = (ap.apTranslatingAperture(Pi, xv, Pn[0], (0.3e-3, 1e-3), literal=l) for l in (True, False))
    else:
        th = np.linspace(-15, 15, 9)
        L, C = (ap.apTxParallelogram(Pi, th, (-5.0, 5.0), (Pn[0].min(), Pn[0].max()), literal=l) for l in (True, False))
    assert L.shape == C.shape
    assert np.mean(L != C) < 2e-3


def test_multiline_rows_sum_to_one_inside_the_transmit_span():
    x = np.linspace(-3e-3, 3e-3, 31)
    xv = np.linspace(-2e-3, 2e-3, 5)
    A = ap.apMultiline(x, xv)
    inside = (x >= xv.min()) & (x <= xv.max())
    assert np.allclose(A[inside].sum(1), 1.0)
    assert np.all(A[~inside] == 0)
    # a scan line on a transmit gets that transmit only
    k = np.argmin(np.abs(x - xv[2]))
    if x[k] == xv[2]:
        assert A[k, 2] == 1.0 and A[k].sum() == 1.0


def test_ultrasound_mirror_multiline_matches_oracle():
    from qups_b200 import ultrasound as U
    Pi, Pn, _ = _geom()
    xv = np.linspace(-2e-3, 2e-3, 5)
    fo = np.stack([xv, 0 * xv, np.full(5, 8e-3)])
    us = U.UltrasoundSystem(tx=Pn, rx=Pn, seq=U.Sequence("FC", fo), scan=Pi, fs=20e6)
    A = us.apMultiline()
    assert A.shape == (1, Pi.shape[2], 1, 1, 5)
    assert np.array_equal(A[0, :, 0, 0, :], ap.apMultiline(Pi[0, 0, :, 0], xv))


def test_prep_oracle_hilbert_equals_scipy_and_weights():
    """oracle/prep_np.py: the reference's hilbert weights (src/ChannelData.m:961-962) == scipy.signal.hilbert (MATLAB's definition)."""
    import scipy.signal as ss
    from oracle import prep_np
    rng = np.random.default_rng(0)
    for T in (64, 101, 100, 7):
        x = rng.standard_normal((T, 2, 2))
        y, _ = prep_np.prep(x, 0.0, 1.0, hilbert=True)
        assert np.max(np.abs(y - ss.hilbert(x, axis=0))) < 1e-6
    assert list(prep_np.hilbert_weights(5)) == [1, 2, 2, 0, 0] and list(prep_np.hilbert_weights(6)) == [1, 2, 2, 1, 0, 0]
    y, t0 = prep_np.prep(np.ones((4, 1, 1)), 1e-6, 1e6, B=2, A=1)
    assert y.shape == (7, 1, 1) and t0 == pytest.approx(-1e-6) and list(y[:, 0, 0].real) == [0, 0, 1, 1, 1, 1, 0]
