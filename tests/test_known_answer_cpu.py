"""Physical known-answer tests of the reference, restated against the ORACLE (CPU, no GPU needed).

test/SimTest.m:299-324  a scatterer at 15 mm, c0 = 1500 m/s, fs = 40 MHz, 5-element array: the centre rx/tx
                         trace peaks at 20 us within 1.1/fs (FSA).
test/BFTest.m:230-317    one scatterer at (2, 0, 15) mm simulated with greens(..., 'interp','linear'), beamformed
                         with DAS: image non-zero and arg-max within 1.1 mm of the scatterer in x and z.
These pin the oracle's delay model, time axis and sign conventions to the reference's own expectations."""
import numpy as np

from qups_b200 import synth


def _greens_oracle(oracle_c, ps, amp, pn, fs, c0, fc, interp):
    kern, wt0, wtend = synth.greens_kernel(fc, 0.6, fs)
    r = np.linalg.norm(ps[:, :, None] - pn[:, None, :], axis=0)
    tmin, tmax = 2 * r.min() / c0 + wt0 - (wtend - wt0), 2 * r.max() / c0 + wtend
    n0, ne = int(np.floor(tmin * fs)), int(np.ceil(tmax * fs))
    x = oracle_c.greens(ps, amp, pn, pn, kern, n0, ne - n0 + 1, fs, c0, wt0, 1.0, c0 / fc, interp, dtype=np.float64)
    return x, n0 / fs


def test_simtest_echo_arrival_time(oracle_c):
    fs, c0, fc = 40e6, 1500.0, 5e6
    pn = synth.linear_array(5, 0.3e-3)
    ps = np.array([[0.0], [0.0], [15e-3]])
    x, t0 = _greens_oracle(oracle_c, ps, np.ones(1), pn, fs, c0, fc, "cubic")
    tr = np.abs(x[:, 2, 2])
    t_peak = t0 + np.argmax(tr) / fs
    assert abs(t_peak - 20e-6) <= 1.1 / fs


def test_bftest_psf_location_fsa(oracle_c):
    fs, c0, fc = 25e6, 1500.0, 6.25e6
    N = 32
    pn = synth.linear_array(N, 0.3e-3)
    ps = np.array([[2e-3], [0.0], [15e-3]])
    x, t0 = _greens_oracle(oracle_c, ps, np.ones(1), pn, fs, c0, fc, "linear")
    xs, zs = np.linspace(-4e-3, 6e-3, 41), np.linspace(11e-3, 19e-3, 33)
    Pi = synth.scan_cartesian(xs, zs)
    nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, N))
    b = oracle_c.das_spec("DAS", Pi, pn, pn, nv, np.asfortranarray(x), t0, fs, c0, interp="cubic", VS=True, DV=True,
                          dtype=np.float64)[:, :, 0, 0, 0, 0]
    assert np.abs(b).max() > 0
    iz, ix = np.unravel_index(np.argmax(np.abs(b)), b.shape)
    assert abs(zs[iz] - 15e-3) <= 1.1e-3 and abs(xs[ix] - 2e-3) <= 1.1e-3


def test_focustx_host_logic_with_oracle_sampler(oracle_np):
    """focusTx mirror (src/UltrasoundSystem.m:3374-3503) == brute-force delayed sum, with the ORACLE as the sampler
    (the product path samples with the CUDA wsinterpd2 kernel; this pins the host logic on CPU)."""
    from qups_b200.ultrasound import focusTx, Sequence, ChannelData, seq_delays
    pn = synth.linear_array(8, 0.3e-3)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((64, 8, 8)) + 1j * rng.standard_normal((64, 8, 8))).astype(np.complex128)
    th = np.deg2rad([-5, 0, 5])
    seq = Sequence("PW", np.stack([np.sin(th), 0 * th, np.cos(th)], 0), 1540.0)
    fs = 20e6
    out = focusTx(ChannelData(x, 1e-6, fs), seq, pn, "linear", ws2=oracle_np.wsinterpd2)
    tau = -seq_delays(seq, pn)
    nmin = np.floor(tau.min() * fs)
    Tn = out.data.shape[0]
    ref = np.zeros_like(out.data)
    for mp in range(3):
        for m in range(8):
            tq = np.arange(Tn) - (tau[m, mp] * fs - nmin)
            xp = np.concatenate([x[:, :, m], np.zeros((Tn - 64, 8))], 0)
            for n in range(8):
                ref[:, n, mp] += oracle_np.interp1(xp[:, n], 1 + tq, "linear", 0)
    assert np.abs(ref - out.data).max() < 1e-12
    assert abs(out.t0 - (1e-6 + nmin / fs)) < 1e-15
