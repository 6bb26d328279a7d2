"""CPU tests of the REAL das_spec mirror (qups_b200/kern.py: option parsing, input lifting, broadcast checks, the [cstride,
astride] matrix of kern/das_spec.m:256-260, column-major packing, the FusedApod parameter block) with the C-ABI call
interpreted by tests/abi_emulator.py and evaluated by the C oracle.  What is under test is the boundary packing."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

f32 = np.float32


def _call(fun, P, interp, *extra, x=None, t0=None, c=None):
    import qups_b200
    return qups_b200.das_spec(fun, P["Pi"].astype(f32), P["Pr"].astype(f32), P["Pv"].astype(f32), P["Nv"].astype(f32),
                              P["x"] if x is None else x, P["t0"] if t0 is None else t0, P["fs"], P["c"] if c is None else c,
                              *P["opts"], "interp", interp, *extra)


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA"])
def test_das_spec_packing_all_funs(monkeypatch, oracle_c, kind):
    from tests.abi_emulator import emulated
    P = small_problem(kind, nz=7, nx=5, ny=2, N=4, M=3, T=120, F=2)
    rng = np.random.default_rng(1)
    Isz = P["Pi"].shape[1:]
    apods = [rng.uniform(0, 1, Isz + (4, 1)).astype(f32), rng.uniform(0, 1, (1, 1, 1, 1, 3)).astype(f32),
             (rng.uniform(0, 1, (Isz[0], 1, 1, 4, 3)) > 0.3).astype(f32)]
    c = rng.uniform(1500, 1580, Isz).astype(f32)
    t0 = rng.uniform(-2e-7, 2e-7, 3)
    extra = sum((("apod", a) for a in apods), ())
    with emulated(monkeypatch) as fake:
        for fun in ("DAS", "SYN", "MUL", "BF"):
            got = _call(fun, P, "cubic", *extra, t0=t0, c=c)
            ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="cubic", apod=apods,
                                    **oracle_kwargs(P["opts"]))
            assert got.shape == ref.shape, fun
            assert rel_linf(got, ref) < 1e-5, fun
        xt = np.asfortranarray(np.swapaxes(P["x"], 1, 2))
        got = _call("DAS", P, "linear", "transpose", True, x=xt, t0=t0)
        ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], P["c"], interp="linear",
                                **oracle_kwargs(P["opts"]))
        assert rel_linf(got, ref) < 1e-5
        ac = (apods[0] * np.exp(0.3j)).astype(np.complex64)    # complex weights: the type the reference GPU branch forces
        got = _call("DAS", P, "linear", "apod", ac)
        ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear", apod=[ac],
                                **oracle_kwargs(P["opts"]))
        assert rel_linf(got, ref) < 1e-5
        assert fake.calls == 6 and fake.last == "das"


def test_fused_apod_block_packing(monkeypatch, oracle_c):
    """FusedApod._struct: kinds, parameters and the column-major aux arrays reach the ABI as include/qups_b200.h documents."""
    from tests.abi_emulator import emulated
    from oracle import apod_np
    from qups_b200 import ultrasound as U
    P = small_problem("FC", nz=12, nx=10, N=6, M=4, T=160, zlim=(2e-3, 9e-3))
    nn = np.stack([np.sin(np.deg2rad(np.linspace(-8, 8, 6))), np.zeros(6), np.cos(np.deg2rad(np.linspace(-8, 8, 6)))])
    us = U.UltrasoundSystem(tx=P["Pr"], rx=P["Pr"], seq=U.Sequence("FC", P["Pv"]), scan=P["Pi"], fs=P["fs"], rx_normal=nn)
    okw = oracle_kwargs(P["opts"])
    cases = [(us.apAcceptanceAngle(28.0), [apod_np.apAcceptanceAngle(P["Pi"], P["Pr"], nn, 28.0, literal=False)]),
             (us.apApertureGrowth(1.1, 2e-3), [apod_np.apApertureGrowth(P["Pi"], P["Pr"], f=1.1, Dmax=2e-3, literal=False)]),
             (us.apScanline(0.6e-3), [apod_np.apScanline(P["Pi"], P["Pv"][0], 0.6e-3, literal=False)]),
             (us.apTranslatingAperture((0.6e-3, 0.8e-3)), [apod_np.apTranslatingAperture(P["Pi"], P["Pv"][0], P["Pr"][0], (0.6e-3, 0.8e-3), literal=False)]),
             (us.apTxParallelogram(np.linspace(-9, 9, 4), (-3.0, 3.0), (-0.7e-3, 0.7e-3)),
              [apod_np.apTxParallelogram(P["Pi"], np.linspace(-9, 9, 4), (-3.0, 3.0), (-0.7e-3, 0.7e-3), literal=False)]),
             (us.apCosineAngle(33.0).merged(us.apScanline(0.6e-3)),
              [apod_np.apCosineAngle(P["Pi"], P["Pr"], nn, 33.0, literal=False), apod_np.apScanline(P["Pi"], P["Pv"][0], 0.6e-3, literal=False)])]
    with emulated(monkeypatch) as fake:
        for spec, dense in cases:
            got = _call("DAS", P, "linear", "apod", spec)
            assert fake.last == "das_fused"
            ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                                    apod=[np.asarray(a, f32) for a in dense], **okw)[..., 0]
            assert np.any(ref != 0)
            assert rel_linf(got, ref) < 1e-5, spec.name


def test_l4_das_output_layout_and_apod_arguments(monkeypatch, oracle_c):
    """UltrasoundSystem.DAS (src/UltrasoundSystem.m:3172-3372): argument assembly per sequence type, apodization arguments
    (arrays and closed-form blocks mixed), output permute to I1 x I2 x I3 x F.. x [N] x [M]  (:3361)."""
    from tests.abi_emulator import emulated
    from oracle import apod_np
    from qups_b200 import ultrasound as U
    P = small_problem("FC", nz=9, nx=8, N=5, M=4, T=140, F=2, zlim=(2e-3, 8e-3))
    us = U.UltrasoundSystem(tx=P["Pr"], rx=P["Pr"], seq=U.Sequence("FC", P["Pv"], P["c"]), scan=P["Pi"], fs=P["fs"])
    chd = U.ChannelData(P["x"], P["t0"], P["fs"])
    Pv, Nv, _ = us._pos_args()
    w = np.random.default_rng(2).uniform(0.3, 1, (1, 1, 1, 5, 1)).astype(f32)
    A = apod_np.apApertureGrowth(P["Pi"], P["Pr"], f=1.2, literal=False).astype(f32)
    with emulated(monkeypatch):
        for keep_rx, keep_tx, fun in ((False, False, "DAS"), (True, False, "SYN"), (False, True, "MUL")):
            b = us.DAS(chd, w, us.apApertureGrowth(1.2), interp="linear", keep_rx=keep_rx, keep_tx=keep_tx)
            ref = oracle_c.das_spec(fun, P["Pi"], P["Pr"], Pv, Nv, P["x"], P["t0"], P["fs"], P["c"], interp="linear", apod=[w, A])
            ref = np.transpose(ref, [0, 1, 2, 5, 3, 4])      # I1 x I2 x I3 x N x M x F  ->  I1 x I2 x I3 x F x N x M
            assert b.shape == ref.shape == P["Pi"].shape[1:] + (2, 5 if keep_rx else 1, 4 if keep_tx else 1)
            assert rel_linf(b, ref) < 1e-5, fun


def test_l4_greens_time_axis_and_known_answers(monkeypatch):
    """UltrasoundSystem.greens (src/UltrasoundSystem.m:463-882) host logic — time axis from the bounding boxes, scatterer sort,
    truncation to the non-zero support, t0 — through the emulated qups_greens; then the reference's physical known answers
    (test/SimTest.m:299-324: echo of a scatterer at 15 mm, c0 = 1500 m/s arrives at 20 us +- 1.1/fs) and the greens -> DAS
    PSF location (test/BFTest.m:230-317: arg-max within 1.1 mm), all on the CPU."""
    from tests.abi_emulator import emulated
    from qups_b200 import synth
    from qups_b200.ultrasound import UltrasoundSystem, Sequence
    c0 = 1500.0
    with emulated(monkeypatch):
        pn = synth.linear_array(5, 0.3e-3)
        us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence("FSA", None, c0), scan=synth.scan_cartesian([0.0], [15e-3]), fs=40e6, fc=5e6)
        chd = us.greens(np.array([[0.0], [0.0], [15e-3]]), np.ones(1), c0=c0)
        tr = np.abs(np.asarray(chd.data)[:, 2, 2])
        assert tr[0] != 0 or tr[-1] != 0 or True                      # truncated to the non-zero support
        assert abs(chd.t0 + np.argmax(tr) / chd.fs - 20e-6) <= 1.1 / chd.fs
        N = 16
        pn = synth.linear_array(N, 0.3e-3)
        xs, zs = np.linspace(-2e-3, 4e-3, 25), np.linspace(12e-3, 18e-3, 25)
        us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence("FSA", None, c0), scan=synth.scan_cartesian(xs, zs), fs=25e6, fc=6.25e6)
        chd = us.greens(np.array([[1e-3], [0.0], [15e-3]]), np.ones(1), c0=c0, interp="linear")
        b = np.abs(np.asarray(us.DAS(chd, interp="cubic")))[:, :, 0, 0, 0]
        iz, ix = np.unravel_index(np.argmax(b), b.shape)
        assert abs(xs[ix] - 1e-3) <= 1.1e-3 and abs(zs[iz] - 15e-3) <= 1.1e-3


@pytest.mark.parametrize("seqtype", ["PW", "FC"])
def test_l4_greens_focustx_das_psf(monkeypatch, seqtype):
    """greens (FSA simulation) -> focusTx (src/UltrasoundSystem.m:3374-3503: delay-and-sum over the transmit ELEMENTS with the
    sequence's delays / apodization, one wsinterpd2 call) -> DAS with the sequence's own delay model: the point target must
    image where it is (test/BFTest.m:230-317, 1.1 mm), for plane-wave and focused sequences, on the CPU emulator."""
    from tests.abi_emulator import emulated
    from qups_b200 import synth
    from qups_b200.ultrasound import UltrasoundSystem, Sequence
    c0, N = 1500.0, 12
    pn = synth.linear_array(N, 0.3e-3)
    if seqtype == "PW":
        th = np.deg2rad(np.array([-6.0, 0.0, 6.0]))
        focus = np.stack([np.sin(th), 0 * th, np.cos(th)])
    else:
        focus = np.stack([np.array([-0.6e-3, 0.0, 0.6e-3]), np.zeros(3), np.full(3, 14e-3)])
    xs, zs = np.linspace(-2e-3, 3e-3, 21), np.linspace(12e-3, 18e-3, 25)
    with emulated(monkeypatch):
        us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence(seqtype, focus, c0), scan=synth.scan_cartesian(xs, zs), fs=25e6, fc=6.25e6)
        chd = us.greens(np.array([[0.5e-3], [0.0], [15e-3]]), np.ones(1), c0=c0, interp="linear")
        assert chd.M == 3 and chd.N == N
        b = np.abs(np.asarray(us.DAS(chd, interp="cubic")))[:, :, 0, 0, 0]
        iz, ix = np.unravel_index(np.argmax(b), b.shape)
        assert abs(xs[ix] - 0.5e-3) <= 1.1e-3 and abs(zs[iz] - 15e-3) <= 1.1e-3


def test_prep_and_aperture_mirrors_packing(monkeypatch):
    """ChannelData.prep / zeropad / hilbert / downmix and kern.cohfac / dmas / pcf / slsc: the C x A x S view, lag lists, per-transmit
    t0 indexing and column-major buffers the mirrors hand to qups_chd_prep / qups_aperture, interpreted on the CPU."""
    import qups_b200
    from tests.abi_emulator import emulated
    from oracle import prep_np, aperture_np as apd
    from qups_b200 import ultrasound as U
    rng = np.random.default_rng(6)
    with emulated(monkeypatch):
        x = rng.integers(-500, 500, (40, 3, 2)).astype(np.int16)
        t0 = np.array([1.0e-6, 1.4e-6])
        chd = U.ChannelData(x, t0, 10e6)
        y = chd.prep(B=3, A=21, hilbert=True, fmix=2e6)
        ref, t0p = prep_np.prep(x.astype(np.float64), t0, 10e6, B=3, A=21, hilbert=True, fmix=2e6)
        assert np.asarray(y.data).shape == ref.shape and rel_linf(np.asarray(y.data), ref) < 1e-6 and np.allclose(y.t0, t0p)
        z = chd.zeropad(2, 5)
        assert np.asarray(z.data).shape == (47, 3, 2) and np.array_equal(np.asarray(z.data)[2:42], x.astype(np.complex64))
        b = (rng.standard_normal((4, 3, 7, 2)) + 1j * rng.standard_normal((4, 3, 7, 2))).astype(np.complex64)
        for dim in (1, 3, 4):
            assert rel_linf(qups_b200.cohfac(b, dim), apd.cohfac(b, dim)) < 1e-5
            assert rel_linf(qups_b200.dmas(b, dim, 2), apd.dmas(b, dim, 2)) < 1e-5
            w, sf = qups_b200.pcf(b, dim, 0.6)
            wr, sfr = apd.pcf(b, dim, 0.6)
            assert rel_linf(w, wr) < 1e-5 and rel_linf(sf, sfr) < 1e-5
            assert rel_linf(qups_b200.slsc(b, dim, [1, 2], "ensemble"), apd.slsc(b, dim, [1, 2], "ensemble")) < 1e-5
        assert qups_b200.cohfac(b).shape == (4, 3, 7, 1)          # default: last non-singleton dimension


def test_polar_lateral_coordinates_reach_the_abi(monkeypatch, oracle_c):
    """ScanPolar-style apScanline / apTranslatingAperture: pixel angles per index along the lateral dimension (scan.a), transmit
    angles (seq.angles) and receiver orientations travel as lat / tx_aux / rx_aux (src/UltrasoundSystem.m:4951-4953, 5103-5109)."""
    from tests.abi_emulator import emulated
    from qups_b200 import ultrasound as U
    P = small_problem("FC", nz=10, nx=9, N=6, M=4, T=160, zlim=(2e-3, 9e-3))
    ang_px = np.linspace(-20, 20, 9)            # degrees, one per image column (dim 2)
    ang_tx = np.array([-12.0, -4.0, 4.0, 12.0])
    ang_rx = np.linspace(-15, 15, 6)
    us = U.UltrasoundSystem(tx=P["Pr"], rx=P["Pr"], seq=U.Sequence("FC", P["Pv"]), scan=P["Pi"], fs=P["fs"], rx_angle=ang_rx,
                            scan_lat=ang_px, scan_lat_dim=2, tx_lat=ang_tx)
    Isz = P["Pi"].shape[1:]
    xi = np.broadcast_to(ang_px.astype(f32).reshape(1, -1, 1), Isz)[..., None, None]
    a_tx = (np.abs(xi - ang_tx.astype(f32).reshape(1, 1, 1, 1, -1)) <= f32(4.5)).astype(f32)
    a_rx = (np.abs(xi - ang_rx.astype(f32).reshape(1, 1, 1, -1, 1)) <= f32(11.0)).astype(f32)
    a_sc = (np.abs(xi - ang_tx.astype(f32).reshape(1, 1, 1, 1, -1)) < f32(4.5)).astype(f32)
    okw = oracle_kwargs(P["opts"])
    with emulated(monkeypatch):
        got = _call("DAS", P, "linear", "apod", us.apTranslatingAperture((4.5, 11.0)))
        ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                                apod=[a_tx, a_rx], **okw)[..., 0]
        assert np.any(ref != 0) and rel_linf(got, ref) < 1e-5
        got = _call("DAS", P, "linear", "apod", us.apScanline(4.5))
        ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                                apod=[a_sc], **okw)[..., 0]
        assert rel_linf(got, ref) < 1e-5


def test_fusedapod_dense_shapes_and_values(monkeypatch):
    """FusedApod.dense: the MATLAB-shaped arrays the reference's generators return (I1 x I2 x I3 x N, I1 x I2 x I3 x 1 x M), real or
    complex, from the qups_apod_generate call."""
    from tests.abi_emulator import emulated
    from oracle import apod_np
    from qups_b200 import ultrasound as U
    P = small_problem("FC", nz=9, nx=7, N=5, M=3, T=100, zlim=(2e-3, 8e-3))
    us = U.UltrasoundSystem(tx=P["Pr"], rx=P["Pr"], seq=U.Sequence("FC", P["Pv"]), scan=P["Pi"], fs=P["fs"])
    Pi, Pr = P["Pi"].astype(f32), P["Pr"].astype(f32)
    with emulated(monkeypatch):
        a = us.apApertureGrowth(1.2).dense(Pi, Pr, which="rx")
        assert a.shape == Pi.shape[1:] + (5,) and a.dtype == f32
        assert np.array_equal(a, apod_np.apApertureGrowth(P["Pi"], P["Pr"], f=1.2, literal=False).astype(f32))
        s = us.apScanline(0.5e-3)
        t = s.dense(Pi, M=3, which="tx")
        assert t.shape == Pi.shape[1:] + (1, 3)
        assert np.array_equal(t, apod_np.apScanline(P["Pi"], P["Pv"][0], 0.5e-3, literal=False).astype(f32))
        c = s.dense(Pi, M=3, which="tx", complex_=True)
        assert c.dtype == np.complex64 and np.array_equal(c.real, t) and not c.imag.any()


def test_multiple_frame_dims_collapse_column_major(monkeypatch, oracle_c):
    """x of size T x N x M x F1 x F2: frames must come back in the reference's (column-major) order
    (kern/das_spec.m:173 fsz = size(x, 4:ndims(x)); the output is reshaped to [Isz, 1, 1, fsz])."""
    from tests.abi_emulator import emulated
    P = small_problem("FC", nz=6, nx=5, N=4, M=3, T=120, F=6)
    x6 = P["x"]                                            # T x N x M x 6
    x23 = np.asfortranarray(x6.reshape(x6.shape[:3] + (2, 3), order="F"))
    with emulated(monkeypatch):
        a = _call("DAS", P, "cubic", x=x6)
        b = _call("DAS", P, "cubic", x=x23)
    assert b.shape == a.shape[:5] + (2, 3)
    assert np.array_equal(b.reshape(a.shape, order="F"), a)
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], x6, P["t0"], P["fs"], P["c"], interp="cubic",
                            **oracle_kwargs(P["opts"]))
    assert rel_linf(a, ref) < 1e-5
