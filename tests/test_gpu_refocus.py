"""GPU parity of qups_refocus (csrc/refocus.cu) behind the refocus mirror vs the float64 restatement of
src/UltrasoundSystem.m:3690-3767 (oracle/refocus_np.py).  The decoder is host code in both (float64); the data path
(fft, phase, per-frequency decode, ifft) runs in fp32 on the device: tolerance 1e-5 of max|y|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(kind, T, N, V, seed=0):
    from qups_b200 import synth, ultrasound as U
    rng = np.random.default_rng(seed)
    tx = synth.linear_array(N, 0.3e-3)
    if kind == "PW":
        th = np.deg2rad(np.linspace(-12, 12, V))
        seq = U.Sequence("PW", np.stack([np.sin(th), 0 * th, np.cos(th)]), 1540.0)
    elif kind == "FC":
        seq = U.Sequence("FC", np.stack([np.linspace(-2e-3, 2e-3, V), np.zeros(V), np.full(V, 20e-3)]), 1540.0)
    else:  # Hadamard-encoded FSA
        H = np.array([[1.0]])
        while H.shape[0] < N:
            H = np.block([[H, H], [H, -H]])
        seq = U.Sequence("FSA", None, 1540.0, apd=H[:, :V])
    x = (rng.standard_normal((T, N, V)) + 1j * rng.standard_normal((T, N, V))).astype(np.complex64)
    return tx, seq, x


@pytest.mark.parametrize("kind,T,N,V", [("PW", 256, 16, 11), ("FC", 512, 16, 16), ("HD", 128, 16, 16), ("PW", 8, 5, 3), ("FC", 2048, 70, 9)])
@pytest.mark.parametrize("method", ["tikhonov", "adjoint", "pinv"])
def test_refocus_matches_oracle(kind, T, N, V, method):
    from qups_b200 import ultrasound as U
    from oracle import refocus_np as R
    tx, seq, x = _setup(kind, T, N, V, seed=T + V)
    fs = 25e6
    t0 = 2e-6 if kind != "FC" else np.linspace(1e-6, 3.3e-6, V)  # focused sequences: a start time per transmit
    chd, Hi = U.refocus(U.ChannelData(x, t0, fs), seq, tx, method)
    tau, apd = U.seq_delays(seq, tx), U.seq_apodization(seq, tx)
    want, t0m, Hiw = R.refocus(x, t0, fs, tau, apd, method)
    assert Hi.shape == (N, V, T)
    assert np.max(np.abs(Hi - Hiw)) <= 1e-9 * max(1e-300, np.max(np.abs(Hiw)))
    got = np.asarray(chd.data)
    assert got.shape == want.shape == (T, N, N)
    assert abs(chd.t0 - t0m) < 1e-15
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want)), np.max(np.abs(got - want)) / np.max(np.abs(want))


def test_refocus_decodes_hadamard_exactly_and_rejects_odd_lengths():
    import qups_b200
    from qups_b200 import ultrasound as U
    tx, seq, _ = _setup("HD", 64, 8, 8)
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((64, 8, 8)) + 1j * rng.standard_normal((64, 8, 8))).astype(np.complex64)
    enc = np.einsum("tne,ev->tnv", x, seq.apd).astype(np.complex64)
    chd, _ = U.refocus(U.ChannelData(enc, 0.0, 20e6), seq, tx, "tikhonov", gamma=0.0)
    assert np.max(np.abs(np.asarray(chd.data) - x)) < 1e-5 * np.max(np.abs(x))
    with pytest.raises(qups_b200.QupsError):
        U.refocus(U.ChannelData(enc[:60], 0.0, 20e6), seq, tx, "adjoint")
