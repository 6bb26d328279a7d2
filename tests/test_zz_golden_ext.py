"""Golden fixtures for the SURVEY §8f rows and greens (tests/golden/ext/*.npz, oracle-generated regression pins — see
tests/golden/make_golden_ext.py).  CPU: the oracles still reproduce them; GPU: the CUDA paths match them with the same
tolerances as the live-oracle parity tests."""
import os

import numpy as np
import pytest

from tests.util import rel_linf

EXT = os.path.join(os.path.dirname(__file__), "golden", "ext")
f32 = np.float32


def _ld(name):
    return np.load(os.path.join(EXT, name), allow_pickle=True)


# ------------------------------------------------------------------ CPU: oracle == fixtures ---------------------------
def test_oracle_reproduces_apod_golden():
    from oracle import apod_np as ap
    d = _ld("apod.npz")
    Pi, Pn, nn = d["Pi"].astype(np.float64), d["Pn"].astype(np.float64), d["nn"].astype(np.float64)
    assert np.array_equal(ap.apAcceptanceAngle(Pi, Pn, nn, 30.0, literal=False), d["acc"])
    assert np.allclose(ap.apCosineAngle(Pi, Pn, nn, 40.0, literal=False), d["cos"], atol=1e-7)
    assert np.array_equal(ap.apApertureGrowth(Pi, Pn, f=1.3, Dmax=2e-3, literal=False), d["grow"])
    assert np.array_equal(ap.apScanline(Pi, d["xv"], 0.41e-3, literal=False), d["scan"])
    assert np.array_equal(ap.apTranslatingAperture(Pi, d["xv"], Pn[0], (0.41e-3, 0.9e-3), literal=False), d["trans"])
    assert np.array_equal(ap.apTxParallelogram(Pi, d["th"], (-3.0, 3.0), (-1.2e-3, 1.2e-3), literal=False), d["para"])


def test_oracle_reproduces_prep_aperture_greens_golden(oracle_c):
    from oracle import prep_np, aperture_np as apd
    d = _ld("prep.npz")
    y, t0p = prep_np.prep(d["x"].astype(np.float64), d["t0"], float(d["fs"]), B=int(d["B"]), A=int(d["A"]), hilbert=True, fmix=float(d["fmix"]))
    assert np.allclose(y, d["y"], atol=1e-6 * np.abs(d["y"]).max()) and np.allclose(t0p, d["t0p"])
    a = _ld("aperture.npz")
    assert np.allclose(apd.cohfac(a["b"], 2), a["cohfac"]) and np.allclose(apd.dmas(a["b"], 2, 3), a["dmas3"])
    assert np.allclose(apd.slsc(a["b"], 2, 3, "average"), a["slsc_avg"]) and np.allclose(apd.slsc(a["b"], 2, 3, "ensemble"), a["slsc_ens"])
    g = _ld("greens.npz")
    y32 = oracle_c.greens(g["ps"], g["amp"], g["pn"], g["pv"], g["kern"], int(g["n0"]), int(g["T"]), float(g["fs"]), float(g["c0"]),
                          float(g["wt0"]), 1.0, float(g["R0"]), "cubic")
    assert np.array_equal(y32, g["y32"])
    assert rel_linf(g["y32"], g["y64"]) < 1e-3       # the fp32 oracle's own delay rounding vs the fp64 arbiter


def test_oracle_reproduces_xcorr_refocus_golden():
    from oracle import xcorr_np, refocus_np
    d = _ld("xcorr.npz")
    x, w = d["x"], d["w"]
    assert np.allclose(xcorr_np.pwznxcorr(x, [-3, 0, 2], w), d["y_neighbor"], atol=1e-6, equal_nan=True)
    assert np.allclose(xcorr_np.pwznxcorr(x, [-3, 0, 2], 6, ref="center", norm=False), d["y_center_nonorm"], atol=1e-6 * np.abs(d["y_center_nonorm"]).max())
    assert np.allclose(xcorr_np.pwznxcorr(x, 2, 8, ref="x0", x0=x[:, 1:2], zero=False), d["y_x0"], atol=1e-6, equal_nan=True)
    r = _ld("refocus.npz")
    for m in ("tikhonov", "adjoint"):
        y, t0m, Hi = refocus_np.refocus(r["x"], r["t0"], float(r["fs"]), r["tau"], r["apd"], m)
        assert np.allclose(y, r["y_" + m], atol=1e-9 * np.abs(r["y_" + m]).max()) and np.allclose(Hi, r["Hi_" + m], atol=1e-9 * np.abs(r["Hi_" + m]).max())
        assert t0m == float(r["t0_out"])


# ------------------------------------------------------------------ GPU: CUDA == fixtures -----------------------------
@pytest.mark.gpu
def test_cuda_xcorr_refocus_match_golden():
    import qups_b200
    from qups_b200 import synth, ultrasound as U
    d = _ld("xcorr.npz")
    x, w = d["x"], d["w"]
    assert rel_linf(qups_b200.pwznxcorr(x, [-3, 0, 2], w), d["y_neighbor"]) < 1e-5
    assert rel_linf(qups_b200.pwznxcorr(x, [-3, 0, 2], 6, ref="center", norm=False), d["y_center_nonorm"]) < 2e-4
    assert rel_linf(qups_b200.pwznxcorr(x, 2, 8, ref="x0", x0=x[:, 1:2], zero=False), d["y_x0"]) < 1e-5
    r = _ld("refocus.npz")
    th = np.deg2rad(r["angles"])
    seq = U.Sequence("PW", np.stack([np.sin(th), 0 * th, np.cos(th)]), float(r["c0"]))
    tx = synth.linear_array(8, 0.3e-3)
    assert np.allclose(U.seq_delays(seq, tx), r["tau"], atol=1e-15)
    for m in ("tikhonov", "adjoint"):
        chd, Hi = U.refocus(U.ChannelData(r["x"], r["t0"], float(r["fs"])), seq, tx, m)
        assert rel_linf(Hi, r["Hi_" + m]) < 1e-9 and abs(chd.t0 - float(r["t0_out"])) < 1e-15
        assert rel_linf(np.asarray(chd.data), r["y_" + m]) < 1e-5


@pytest.mark.gpu
def test_cuda_apod_matches_golden():
    from qups_b200 import ultrasound as U
    d = _ld("apod.npz")
    Pi, Pn = d["Pi"], d["Pn"]
    fo = np.stack([d["xv"], 0 * d["xv"], np.full(5, 8e-3)])
    us = U.UltrasoundSystem(tx=Pn, rx=Pn, seq=U.Sequence("FC", fo), scan=Pi, fs=20e6, rx_normal=d["nn"])
    rx = lambda s: s.dense(Pi, Pn, which="rx")
    tx = lambda s: s.dense(Pi, M=5, which="tx")
    assert np.array_equal(rx(us.apAcceptanceAngle(30.0)), d["acc"].astype(f32))
    assert np.max(np.abs(rx(us.apCosineAngle(40.0)) - d["cos"])) < 2e-6
    assert np.array_equal(rx(us.apApertureGrowth(1.3, 2e-3)), d["grow"].astype(f32))
    assert np.array_equal(tx(us.apScanline(0.41e-3)), d["scan"].astype(f32))
    s = us.apTranslatingAperture((0.41e-3, 0.9e-3))
    assert np.array_equal(rx(s)[..., None] * tx(s), d["trans"].astype(f32))
    assert np.array_equal(tx(us.apTxParallelogram(d["th"], (-3.0, 3.0), (-1.2e-3, 1.2e-3))), d["para"].astype(f32))


@pytest.mark.gpu
def test_cuda_prep_aperture_greens_match_golden(monkeypatch):
    import qups_b200
    from qups_b200 import ultrasound as U
    from qups_b200.ultrasound import greens_raw
    d = _ld("prep.npz")
    chd = U.ChannelData(d["x"], d["t0"], float(d["fs"]))
    y = chd.prep(B=int(d["B"]), A=int(d["A"]), hilbert=True, fmix=float(d["fmix"]))
    assert rel_linf(y.data.cpu().numpy(), d["y"]) < 1e-5 and np.allclose(y.t0, d["t0p"])
    assert rel_linf(chd.hilbert().data.cpu().numpy(), d["y_hilbert100"]) < 1e-5       # L = 100: Bluestein
    a = _ld("aperture.npz")
    b = a["b"]
    assert rel_linf(qups_b200.cohfac(b, 2), a["cohfac"]) < 2e-5
    assert rel_linf(qups_b200.dmas(b, 2, 3), a["dmas3"]) < 2e-4
    w, sf = qups_b200.pcf(b, 2, 0.8)
    assert rel_linf(w, a["pcf_w"]) < 1e-4 and rel_linf(sf, a["pcf_sf"]) < 1e-4
    assert rel_linf(qups_b200.slsc(b, 2, 3, "average"), a["slsc_avg"]) < 2e-4
    assert rel_linf(qups_b200.slsc(b, 2, 3, "ensemble"), a["slsc_ens"]) < 2e-4
    g = _ld("greens.npz")
    args = (g["ps"], g["amp"], g["pn"], g["pv"], g["kern"], int(g["n0"]), int(g["T"]), float(g["fs"]), float(g["c0"]), float(g["wt0"]), 1.0,
            float(g["R0"]), "cubic")
    assert rel_linf(greens_raw(*args).cpu().numpy(), g["y64"]) < 5e-6                  # convolution kernel vs the fp64 arbiter
    monkeypatch.setenv("QUPS_B200_GREENS", "binned")
    assert rel_linf(greens_raw(*args).cpu().numpy(), g["y32"]) < 2e-6
    monkeypatch.setenv("QUPS_B200_GREENS", "simple")
    assert np.array_equal(greens_raw(*args).cpu().numpy(), g["y32"])
