"""CPU test of the host logic of ultrasound.bfDAS / bfDASLUT (mirror of src/UltrasoundSystem.m:4334-4673): delay tables,
validation (the reference's error identifiers, test/USTest.m:260-279), transmit blocking (bsize) and apodization reduction.
The GPU call (kern.wsinterpd2) is replaced by the NumPy oracle's wsinterpd2, so only the mirror's own code is under test;
the result must equal the oracle's das_spec (kern/das_spec.m CPU branch) up to the table rounding (tau = d/c0 vs cinv*d)."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf


@pytest.fixture()
def cpu_ws2(monkeypatch, oracle_np):
    from qups_b200 import kern
    monkeypatch.setattr(kern, "wsinterpd2", lambda *a, **k: oracle_np.wsinterpd2(*a, **k))


def _us(P, kind):
    from qups_b200 import ultrasound as U
    foc = P["Nv"] if kind == "PW" else P["Pv"]
    return U.UltrasoundSystem(tx=P["Pv"] if kind == "FSA" else P["Pr"], rx=P["Pr"], seq=U.Sequence(kind, foc, c0=P["c"]), scan=P["Pi"], fs=P["fs"])


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA"])
@pytest.mark.parametrize("keep", [(False, False), (True, False), (False, True)])
def test_bfdas_equals_das_spec_oracle(cpu_ws2, oracle_np, kind, keep):
    from qups_b200 import ultrasound as U
    M = 5
    P = small_problem(kind, nz=9, nx=7, N=6, M=M, T=160, t0=np.linspace(-1e-7, 2e-7, M))
    us = _us(P, kind)
    chd = U.ChannelData(P["x"], P["t0"], P["fs"])
    keep_rx, keep_tx = keep
    fun = {(False, False): "DAS", (True, False): "SYN", (False, True): "MUL"}[keep]
    ref = oracle_np.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="cubic",
                             **oracle_kwargs(P["opts"]))
    ref = ref.reshape(ref.shape[:5])
    for bsize in (None, 2, 1):
        b = us.bfDASLUT(chd, *_tables(us, chd), interp="cubic", keep_rx=keep_rx, keep_tx=keep_tx, bsize=bsize)
        assert b.shape == ref.shape
        assert rel_linf(b, ref) < 2e-4, (kind, keep, bsize)   # fp32 tables: tau = d / c0 rounds differently from cinv * (dv + dr)
    b2 = us.bfDAS(chd, interp="cubic", keep_rx=keep_rx, keep_tx=keep_tx)
    assert rel_linf(b2, ref) < 2e-4


def _tables(us, chd):
    """What bfDAS builds (src/UltrasoundSystem.m:4431-4463), in float64 for the test."""
    Pv, Nv, _ = us._pos_args()
    Isz = tuple(us.scan.shape[1:])
    Pi = np.asarray(us.scan, np.float64).reshape(3, -1, order="F")
    Pv = np.broadcast_to(np.asarray(Pv, np.float64), (3, chd.M))
    rv = Pi[:, :, None] - Pv[:, None, :]
    t = us.seq.type
    if t in ("DV", "FSA"): dv = np.linalg.norm(rv, axis=0)
    elif t == "PW": dv = (rv * np.asarray(Nv, np.float64)[:, None, :]).sum(0)
    else:
        nf = np.asarray(us.seq.focus, np.float64) - us.tx_offset
        nf = nf / np.linalg.norm(nf, axis=0, keepdims=True)
        dv = np.linalg.norm(rv, axis=0) * np.sign((rv * nf[:, None, :]).sum(0))
    dr = np.linalg.norm(Pi[:, :, None] - np.asarray(us.rx, np.float64)[:, None, :], axis=0)
    return (dr / us.seq.c0).reshape(Isz + (chd.N,), order="F"), (dv / us.seq.c0).reshape(Isz + (1, chd.M), order="F")


def test_bfdaslut_apodization_blocks_and_errors(cpu_ws2, oracle_np):
    import qups_b200
    from qups_b200 import ultrasound as U
    P = small_problem("FC", nz=8, nx=6, N=5, M=4, T=160)
    us = _us(P, "FC")
    chd = U.ChannelData(P["x"], P["t0"], P["fs"])
    trx, ttx = _tables(us, chd)
    rng = np.random.default_rng(0)
    Isz = P["Pi"].shape[1:]
    a_rx = rng.uniform(0, 1, Isz + (5, 1))            # common to all transmits: reduced once
    a_tx = rng.uniform(0, 1, (1, 1, 1, 1, 4))         # per transmit: sliced per block
    ref = oracle_np.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                             apod=[a_rx, a_tx], **oracle_kwargs(P["opts"]))
    ref = ref.reshape(ref.shape[:5])
    for bsize in (None, 3, 1):
        b = us.bfDASLUT(chd, trx, ttx, a_rx, a_tx, interp="linear", bsize=bsize)
        assert rel_linf(b, ref) < 2e-4
    # flattened tables are accepted when the element counts match (:4588-4627) ...
    b = us.bfDASLUT(chd, trx.reshape(-1, 5, order="F"), ttx.reshape(-1, 4, order="F"), a_rx, a_tx, interp="linear")
    assert rel_linf(b, ref) < 2e-4
    # ... and rejected with the reference's identifiers otherwise
    with pytest.raises(qups_b200.QupsError, match="incompatibleReceiveDelayTable"):
        us.bfDASLUT(chd, trx[:-1], ttx)
    with pytest.raises(qups_b200.QupsError, match="incompatibleTransmitDelayTable"):
        us.bfDASLUT(chd, trx, ttx[..., :-1])


@pytest.fixture()
def cpu_ws(monkeypatch, oracle_np):
    from qups_b200 import kern
    monkeypatch.setattr(kern, "wsinterpd2", lambda *a, **k: oracle_np.wsinterpd2(*a, **k))
    monkeypatch.setattr(kern, "wsinterpd", lambda *a, **k: oracle_np.wsinterpd(*a, **k))


def test_channeldata_sample_sample2sep_rectifyt0(cpu_ws, oracle_np):
    """Mirrors of ChannelData.sample / sample2sep / rectifyt0 (src/ChannelData.m:1205-1447) against interp1 written out."""
    from qups_b200 import ultrasound as U
    rng = np.random.default_rng(4)
    T, N, M, fs = 64, 4, 3, 10e6
    x = (rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M))).astype(np.complex64)
    t0 = np.array([1.03e-6, 1.31e-6, 0.8e-6])   # off the sample grid: no query lands exactly on the first / last sample
    chd = U.ChannelData(x, t0, fs)
    tau = (2e-6 + np.arange(9)[:, None, None] * 0.37e-6) + np.zeros((1, N, 1))     # I x N x 1, broadcast over M
    y = chd.sample(tau, "linear")
    assert y.shape == (9, N, M)
    for n in range(N):
        for m in range(M):
            ref = oracle_np.interp1(x[:, n, m], 1 + (tau[:, n, 0] - t0[m]) * fs, "linear", 0)
            assert np.max(np.abs(y[:, n, m] - ref)) < 1e-4
    # separable: tau1 over (I, N), tau2 over (I, 1, M); sum over receives and transmits with weights
    t1 = 1.5e-6 + rng.uniform(0, 2e-6, (9, N, 1))
    t2 = rng.uniform(0, 1e-6, (9, 1, M))
    w = rng.uniform(0.5, 1, (1, N, M))
    y2 = chd.sample2sep(t1, t2, "cubic", w, (2, 3))
    ref = np.zeros(9, np.complex128)
    for n in range(N):
        for m in range(M):
            ref += w[0, n, m] * oracle_np.interp1(x[:, n, m], 1 + (t1[:, n, 0] + t2[:, 0, m] - t0[m]) * fs, "cubic", 0)
    assert np.max(np.abs(np.asarray(y2).reshape(-1) - ref)) < 2e-4 * np.max(np.abs(ref))
    # lifted to the beamforming layout (apdim = [4 5]): I1 x I2 x I3 x N x M, as bfDASLUT calls it (:4651)
    t1b = t1.reshape(9, 1, 1, N, 1)[:, :, :, :, :]
    y3 = chd.sample2sep(t1.reshape(9, 1, 1, N, 1), t2.reshape(9, 1, 1, 1, M), "cubic", w.reshape(1, 1, 1, N, M), (4, 5), 0.0, (4, 5))
    assert np.max(np.abs(np.asarray(y3).reshape(-1) - ref)) < 2e-4 * np.max(np.abs(ref))
    # rectifyt0: one scalar t0, every trace resampled onto the common axis
    r = chd.rectifyt0("linear")
    assert r.t0 == pytest.approx(t0.min())
    tt = r.t0 + np.arange(r.data.shape[0]) / fs
    for m in range(M):
        ref = oracle_np.interp1(x[:, 1, m], 1 + (tt - t0[m]) * fs, "linear", 0)
        assert np.max(np.abs(np.asarray(r.data)[:, 1, m] - ref)) < 1e-4
    with pytest.raises(AssertionError):
        chd.sample(np.zeros((5, N + 1, 1)))


def test_real_mirror_packing_through_the_abi_emulator(monkeypatch, oracle_np):
    """kern.wsinterpd2's own host code (dim moves, sizes, 5 x D strides, column-major buffers) driven end to end on the CPU:
    the C-ABI call is interpreted by tests/abi_emulator.py.  Covers the lifted 5-D layout bfDASLUT / sample2sep use."""
    from tests.abi_emulator import emulated
    from qups_b200 import kern, ultrasound as U
    rng = np.random.default_rng(8)
    with emulated(monkeypatch) as fake:
        # plain N-D call, time along dim 2, sum over one dim, complex weights and a phasor
        x = (rng.standard_normal((3, 20, 2)) + 1j * rng.standard_normal((3, 20, 2))).astype(np.complex64)
        t1 = rng.uniform(2, 15, (3, 5, 1)).astype(np.float32)
        t2 = rng.uniform(-1, 1, (1, 1, 2)).astype(np.float32)
        w = (rng.uniform(0.5, 1, (3, 1, 2)) * np.exp(0.3j)).astype(np.complex64)
        got = kern.wsinterpd2(x, t1, t2, 2, w, (3,), "cubic", 0, 0.4j)
        ref = oracle_np.wsinterpd2(x, t1, t2, 2, w, (3,), "cubic", 0, 0.4j)
        assert got.shape == ref.shape and rel_linf(got, ref) < 1e-5
        # the beamforming layout: bfDAS -> bfDASLUT -> ChannelData.sample2sep(apdim = [4 5]) -> wsinterpd2
        P = small_problem("FC", nz=5, nx=4, N=4, M=3, T=120, t0=np.array([0.0, 1e-7, -1e-7]))
        us = _us(P, "FC")
        chd = U.ChannelData(P["x"], P["t0"], P["fs"])
        for keep_rx, keep_tx, fun in ((False, False, "DAS"), (True, False, "SYN")):
            b = us.bfDAS(chd, interp="linear", keep_rx=keep_rx, keep_tx=keep_tx)
            ref = oracle_np.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp="linear",
                                     **oracle_kwargs(P["opts"]))
            assert rel_linf(np.asarray(b), ref.reshape(np.asarray(b).shape)) < 2e-4
        assert fake.calls >= 3


def test_sample_and_rectifyt0_through_the_abi_emulator(monkeypatch, oracle_np):
    """ChannelData.sample / rectifyt0 -> the real kern.wsinterpd (single-table entry point) on the CPU emulator."""
    from tests.abi_emulator import emulated
    from qups_b200 import ultrasound as U
    rng = np.random.default_rng(12)
    T, N, M, fs = 48, 3, 2, 10e6
    x = (rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M))).astype(np.complex64)
    t0 = np.array([0.93e-6, 1.27e-6])
    chd = U.ChannelData(x, t0, fs)
    with emulated(monkeypatch):
        tau = (1.5e-6 + np.arange(7)[:, None, None] * 0.41e-6) + np.zeros((1, N, 1))
        y = np.asarray(chd.sample(tau, "cubic"))
        for n in range(N):
            for m in range(M):
                ref = oracle_np.interp1(x[:, n, m], 1 + (tau[:, n, 0] - t0[m]) * fs, "cubic", 0)
                assert np.max(np.abs(y[:, n, m] - ref)) < 2e-4
        r = chd.rectifyt0("linear")
        tt = r.t0 + np.arange(np.asarray(r.data).shape[0]) / fs
        for m in range(M):
            # ntau in the data's precision (single), as ChannelData.sample computes it: the last sample of the earliest
            # transmit lands exactly on xq == T there, while a float64 evaluation overshoots T by one ulp
            xq = 1 + ((tt - t0[m]) * fs).astype(np.float32).astype(np.float64)
            ref = oracle_np.interp1(x[:, 2, m], xq, "linear", 0)
            assert np.max(np.abs(np.asarray(r.data)[:, 2, m] - ref)) < 1e-4
