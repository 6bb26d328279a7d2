"""refocus_np.py — CPU restatement of UltrasoundSystem.refocus (TEST INFRASTRUCTURE ONLY; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).

REFoCUS: frequency-domain decoding of a transmit sequence back to full-synthetic-aperture data, restated line by line
from src/UltrasoundSystem.m:3690-3767 in float64 NumPy:

    :3690-3691  tau = seq.delays(tx), a = seq.apodization(tx)            (elements E x pulses V)
    :3694       f = chd.fftaxis = (0:T-1) * fs / T                        (src/ChannelData.m:1491 — not wrapped)
    :3697       H = a.' .* exp(-2j*pi*f.*tau.')                           (V x E x T)
    :3702       w = pagenorm(H, 2).^-2                                    (spectral norm per frequency)
    :3708-3712  tikhonov: A = H'*H + gamma*w*eye(N);  Hi = A \\ H.'        (NOTE: the code divides by the plain transpose
                `pagetranspose(H)`, not the conjugate transpose its comment names; the restatement follows the code)
    :3714-3715  adjoint:  Hi = H.' .* w
    :3717-3719  pinv:     Hi = w .* pinv(H)
    :3727       Hi(isnan(Hi)) = 0
    :3730-3734  x = fft(x, T, tdim) .* exp(-2i*pi*f.*t0)
    :3746-3750  y(:,:,e) = sum_v Hi(e,v,:) .* x(:,:,v)
    :3755-3757  y = ifft(y .* exp(+2i*pi*f.*min(t0)))
    :3762       t0 = min(t0)
The reference evaluates this in the precision of the data (single), where the Tikhonov system (condition number ~1e5 at
256 elements) loses most of its digits; float64 is the arbiter here.  Parity unpinned (no MATLAB; the reference has no
numeric test for refocus)."""
import numpy as np


def decoder(tau, apd, T, fs, N, method="tikhonov", gamma=None):
    """Hi (E x V x T complex128) from the E x V delays / apodization."""
    tau = np.asarray(tau, np.float64)
    apd = np.asarray(apd, np.complex128 if np.iscomplexobj(apd) else np.float64) * np.ones(tau.shape)
    E, V = tau.shape
    if gamma is None:
        gamma = 10.0 * (N / 10.0) ** 2                                  # arguments block :3683
    f = np.arange(T) * fs / T
    H = apd.T[None] * np.exp(-2j * np.pi * f[:, None, None] * tau.T[None])      # T x V x E (pages first)
    with np.errstate(divide="ignore"):
        w = np.linalg.norm(H, 2, axis=(1, 2)) ** -2.0                   # T
    if method == "tikhonov":
        if N != E:
            raise ValueError("refocus: eye(chd.N) must match the number of transmit elements")
        A = np.conj(np.swapaxes(H, 1, 2)) @ H + (gamma * w)[:, None, None] * np.eye(E)[None]
        Hi = np.empty((T, E, V), np.complex128)
        for k in range(T):
            try:
                Hi[k] = np.linalg.solve(A[k], H[k].T)
            except np.linalg.LinAlgError:
                Hi[k] = np.nan
    elif method == "adjoint":
        Hi = np.swapaxes(H, 1, 2) * w[:, None, None]
    elif method == "pinv":
        Hi = w[:, None, None] * np.linalg.pinv(H)
    else:
        raise ValueError(method)
    Hi = np.where(np.isnan(Hi), 0, Hi)
    return np.ascontiguousarray(np.transpose(Hi, (1, 2, 0)))             # E x V x T


def apply(x, Hi, t0, fs):
    """x: T x N x V data; Hi: E x V x T; t0: scalar or V start times.  Returns (y T x N x E, t0_out)."""
    x = np.asarray(x, np.complex128)
    T, N, V = x.shape
    f = np.arange(T) * fs / T
    t0 = np.broadcast_to(np.asarray(t0, np.float64).reshape(-1), (V,))
    X = np.fft.fft(x, T, 0) * np.exp(-2j * np.pi * f[:, None, None] * t0[None, None, :])
    Y = np.einsum("evt,tnv->tne", Hi, X)
    t0m = t0.min()
    y = np.fft.ifft(Y * np.exp(2j * np.pi * f * t0m)[:, None, None], T, 0)
    return y, t0m


def refocus(x, t0, fs, tau, apd, method="tikhonov", gamma=None):
    T, N, V = np.shape(x)
    Hi = decoder(tau, apd, T, fs, N, method, gamma)
    y, t0m = apply(x, Hi, t0, fs)
    return y, t0m, Hi
