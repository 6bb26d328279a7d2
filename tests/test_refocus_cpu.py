"""CPU checks of the REFoCUS restatement (oracle/refocus_np.py, src/UltrasoundSystem.m:3690-3767): a Hadamard-encoded
full-synthetic-aperture acquisition decodes exactly (the reference's own example, :3601-3607), and the decoder follows the
code's `A \\ H.'` (plain transpose) literally."""
import numpy as np

from oracle import refocus_np as R


def _hadamard(n):
    H = np.array([[1.0]])
    while H.shape[0] < n:
        H = np.block([[H, H], [H, -H]])
    return H


def test_hadamard_encoding_decodes_exactly():
    rng = np.random.default_rng(0)
    T, N, E = 64, 8, 8
    x = rng.standard_normal((T, N, E)) + 1j * rng.standard_normal((T, N, E))
    apd = _hadamard(E)                                  # elements x pulses
    enc = np.einsum("tne,ev->tnv", x, apd)              # chd * apd: every pulse fires all elements with +-1 weights
    for method in ("tikhonov", "adjoint", "pinv"):
        y, t0m, Hi = R.refocus(enc, 1e-6, 20e6, np.zeros((E, E)), apd, method, gamma=0.0)
        assert Hi.shape == (E, E, T) and t0m == 1e-6
        # 'pinv' carries the weight w = 1/sigma_max^2 = 1/E on top of the exact inverse (:3719: `w .* pinv(H)`)
        want = x / E if method == "pinv" else x
        assert np.max(np.abs(y - want)) < 1e-10 * np.max(np.abs(x))


def test_decoder_follows_the_plain_transpose_of_the_code():
    T, E, V, fs = 16, 4, 4, 10e6
    rng = np.random.default_rng(1)
    tau = rng.uniform(0, 1e-6, (E, V))
    Hi = R.decoder(tau, np.ones((E, V)), T, fs, E, "tikhonov", 3.0)
    f = np.arange(T) * fs / T
    k = 5
    H = np.exp(-2j * np.pi * f[k] * tau.T)              # V x E
    w = np.linalg.norm(H, 2) ** -2
    want = np.linalg.solve(H.conj().T @ H + 3.0 * w * np.eye(E), H.T)
    assert np.allclose(Hi[:, :, k], want)
    adj = R.decoder(tau, np.ones((E, V)), T, fs, E, "adjoint")
    assert np.allclose(adj[:, :, k], H.T * w)


def test_per_transmit_start_times_realign():
    """Delaying one transmit's record by an integer number of samples and declaring it in t0 leaves the decoded data unchanged."""
    rng = np.random.default_rng(2)
    T, N, E, fs = 64, 3, 4, 10e6
    x = np.zeros((T, N, E), complex)
    x[8:40] = rng.standard_normal((32, N, E))
    apd = _hadamard(E)
    y0, t00, _ = R.refocus(x, np.zeros(E), fs, np.zeros((E, E)), apd, "adjoint")
    xs = x.copy()
    xs[:, :, 2] = np.roll(x[:, :, 2], -3, 0)            # record of pulse 3 starts 3 samples later ...
    t0 = np.zeros(E); t0[2] = 3 / fs                    # ... and says so
    y1, t01, _ = R.refocus(xs, t0, fs, np.zeros((E, E)), apd, "adjoint")
    assert t01 == 0.0 and np.max(np.abs(y1 - y0)) < 1e-10 * np.max(np.abs(y0))
