"""GPU parity: greens, wsinterpd2 / bfDAS through the C ABI vs the CPU oracle, plus the reference's physical
known-answer checks (test/BFTest.m:230-317, test/SimTest.m:299-357) run end to end on the GPU path."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
@pytest.mark.parametrize("R0", [0.0, 3e-4])
def test_greens_bitexact_vs_oracle(oracle_c, interp, R0, monkeypatch):
    from qups_b200 import synth
    from qups_b200.ultrasound import greens_raw
    fs, fc, c0 = 20e6, 5e6, 1500.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    pn = synth.linear_array(7, 0.3e-3)
    pv = synth.linear_array(5, 0.4e-3)
    rng = np.random.default_rng(2)
    S = 300  # more than one smem chunk is exercised in the large test below; this one pins the arithmetic
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), rng.uniform(-1e-3, 1e-3, S), rng.uniform(3e-3, 12e-3, S)], 0)
    amp = rng.standard_normal(S)
    n0, T = 40, 2600
    ref = oracle_c.greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, interp)
    f32 = np.float32  # the fp64 oracle on the SAME fp32-rounded inputs: the arbiter for the convolution kernel's fp64 delays
    ref64 = oracle_c.greens(ps.astype(f32), amp.astype(f32), pn.astype(f32), pv.astype(f32), kern, n0, T, fs, c0, wt0, 1.0, R0, interp,
                            dtype=np.float64)
    assert np.abs(ref).max() > 0
    got = greens_raw(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, interp).cpu().numpy()   # default: convolution kernel
    if interp != "nearest":  # nearest flips a tap where an fp32 delay lands on the other side of .5: compare where it cannot
        assert rel_linf(got, ref64) < 5e-6, rel_linf(got, ref64)
        assert rel_linf(got, ref) < 1e-3                          # vs the fp32 oracle: its own delay rounding (SimTest bar)
    monkeypatch.setenv("QUPS_B200_GREENS", "binned")             # the oracle's fp32 sequence, bucket order
    got = greens_raw(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, interp).cpu().numpy()
    assert rel_linf(got, ref) < 2e-6, rel_linf(got, ref)          # same terms, bucket order instead of scatterer order
    monkeypatch.setenv("QUPS_B200_GREENS", "simple")             # exact scatterer-order variant: bit-exact
    got = greens_raw(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, interp).cpu().numpy()
    assert np.array_equal(got, ref), rel_linf(got, ref)


def test_greens_many_scatterers_fsr_and_fp64(oracle_c):
    from qups_b200 import synth
    from qups_b200.ultrasound import greens_raw
    fs, fc, c0 = 20e6, 5e6, 1540.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, 2 * fs)   # kernel sampled at 2x the output rate: fsr = 2
    pn = synth.linear_array(4, 0.3e-3)
    rng = np.random.default_rng(5)
    S = 5000
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), np.zeros(S), rng.uniform(3e-3, 12e-3, S)], 0)
    amp = rng.standard_normal(S)
    ref = oracle_c.greens(ps, amp, pn, pn, kern, 60, 400, fs, c0, wt0, 2.0, 2e-4, "cubic")
    got = greens_raw(ps, amp, pn, pn, kern, 60, 400, fs, c0, wt0, 2.0, 2e-4, "cubic").cpu().numpy()
    assert rel_linf(got, ref) < 5e-6
    ref64 = oracle_c.greens(ps, amp, pn, pn, kern, 60, 400, fs, c0, wt0, 2.0, 2e-4, "cubic", dtype=np.float64)
    got64 = greens_raw(ps, amp, pn, pn, kern, 60, 400, fs, c0, wt0, 2.0, 2e-4, "cubic", dtype=np.float64).cpu().numpy()
    assert rel_linf(got64, ref64) < 1e-12
    assert rel_linf(got, ref64) < 1e-3   # the reference's own CPU-vs-GPU greens tolerance (test/SimTest.m:327-357)


def test_wsinterpd_interptest_shapes(oracle_np):
    """test/interpTest.m:28-47 generator, :96-143 check, on the GPU kernel."""
    import qups_b200
    I, T, N, M, F = 16, 32, 4, 3, 2
    t = np.arange(T)[:, None, None]
    n = np.arange(N)[None, :, None]
    f = np.arange(F)[None, None, :]
    x = np.exp(2j * np.pi * (0.5 + f / 2 * n / 4) * t / T).astype(np.complex64)[:, :, None, :]   # T x N x 1 x F
    rng = np.random.default_rng(11)
    tau = rng.uniform(-2, T + 1, (I, N, M, 1)).astype(np.float32)
    w = rng.uniform(0, 1, (I, N, M, 1)).astype(np.float32)
    for terp in ("cubic", "nearest", "linear"):
        for dsum in ((), (2,), (3,), (2, 3), (4,)):
            ref = oracle_np.wsinterpd(x, tau, 1, w, dsum, terp, 0)
            got = qups_b200.wsinterpd(x, tau, 1, w, dsum, terp, 0)
            assert got.shape == ref.shape, (terp, dsum)
            assert rel_linf(got, ref) < 1e4 * np.finfo(np.float32).eps, (terp, dsum)   # tolerance of test/interpTest.m:126
    # two-table form with a phasor and time along dim 2
    x2 = np.ascontiguousarray(np.swapaxes(x, 0, 1))                                               # N x T x 1 x F
    t1 = np.swapaxes(tau, 0, 1)
    t2 = rng.uniform(-1, 1, (1, 1, M, 1)).astype(np.float32)
    ref = oracle_np.wsinterpd2(x2, t1, t2, 2, np.swapaxes(w, 0, 1), (3,), "linear", 0, 0.7j)
    got = qups_b200.wsinterpd2(x2, t1, t2, 2, np.swapaxes(w, 0, 1), (3,), "linear", 0, 0.7j)
    assert got.shape == ref.shape and rel_linf(got, ref) < 2e-6


@pytest.mark.parametrize("kind,seq", [("FC", "FC"), ("PW", "PW"), ("FSA", "FSA"), ("DV", "DV")])
def test_bfdas_matches_das_and_oracle(oracle_c, kind, seq):
    from qups_b200.ultrasound import UltrasoundSystem, Sequence, ChannelData
    P = small_problem(kind, nz=21, nx=15, N=8, M=5, T=200)
    focus = P["Nv"] if seq == "PW" else P["Pv"]
    us = UltrasoundSystem(tx=P["Pv"] if seq == "FSA" else P["Pr"], rx=P["Pr"], seq=Sequence(seq, focus, P["c"]),
                          scan=P["Pi"], fs=P["fs"])
    chd = ChannelData(P["x"], 1e-7, P["fs"])
    b1 = us.DAS(chd, interp="cubic")
    b2 = us.bfDAS(chd, interp="cubic")
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], 1e-7, P["fs"], P["c"], interp="cubic",
                            dtype=np.float64, **oracle_kwargs(P["opts"]))[..., 0]
    ref32 = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], 1e-7, P["fs"], P["c"], interp="cubic",
                              **oracle_kwargs(P["opts"]))[..., 0]
    assert rel_linf(np.asarray(b1).reshape(ref.shape), ref32) < 1e-5
    if kind != "FC":  # focused waves: sign(rv . Nv) flips between fp32 and fp64 for pixels at the focal depth
        assert rel_linf(np.asarray(b1).reshape(ref.shape), ref) < 2e-3  # fp32 vs the fp64 arbiter on white noise
    assert rel_linf(np.asarray(b2).reshape(ref.shape), ref) < 5e-3     # table path rounds tau twice (fp32 tables)


def test_physical_known_answers_on_gpu():
    """greens -> DAS round trip: SimTest echo time and BFTest PSF location, all on the GPU path."""
    from qups_b200 import synth
    from qups_b200.ultrasound import UltrasoundSystem, Sequence
    c0 = 1500.0
    pn = synth.linear_array(5, 0.3e-3)
    us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence("FSA", None, c0), scan=synth.scan_cartesian([0.0], [15e-3]),
                          fs=40e6, fc=5e6)
    chd = us.greens(np.array([[0.0], [0.0], [15e-3]]), np.ones(1), c0=c0)
    tr = chd.data[:, 2, 2].abs().cpu().numpy()
    assert abs(chd.t0 + np.argmax(tr) / chd.fs - 20e-6) <= 1.1 / chd.fs
    N = 32
    pn = synth.linear_array(N, 0.3e-3)
    xs, zs = np.linspace(-4e-3, 6e-3, 41), np.linspace(11e-3, 19e-3, 33)
    us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence("FSA", None, c0), scan=synth.scan_cartesian(xs, zs), fs=25e6,
                          fc=6.25e6)
    chd = us.greens(np.array([[2e-3], [0.0], [15e-3]]), np.ones(1), c0=c0, interp="linear")
    b = us.DAS(chd, interp="cubic")
    b = b.cpu().numpy() if hasattr(b, "cpu") else np.asarray(b)
    b = np.abs(b).reshape(len(zs), len(xs), order="F")
    assert b.max() > 0
    iz, ix = np.unravel_index(np.argmax(b), b.shape)
    assert abs(zs[iz] - 15e-3) <= 1.1e-3 and abs(xs[ix] - 2e-3) <= 1.1e-3


def test_convd_matches_conv_and_xcorr():
    """test/KernTest.m:115-161: convd vs conv for full/same/valid along each dim; convd(A) == xcorr(A)."""
    import qups_b200
    rng = np.random.default_rng(4)
    A = (rng.standard_normal((5, 13, 3)) + 1j * rng.standard_normal((5, 13, 3))).astype(np.complex64)
    B = (rng.standard_normal((5, 4, 3)) + 1j * rng.standard_normal((5, 4, 3))).astype(np.complex64)
    for shape in ("full", "same", "valid"):
        z, lags = qups_b200.convd(A, B, 2, shape)
        ref = np.stack([np.stack([np.convolve(A[c, :, s].astype(np.complex128), B[c, :, s].astype(np.complex128), "full")
                                  for s in range(3)], -1) for c in range(5)], 0)
        k0 = {"full": 0, "same": 2, "valid": 3}[shape]
        L = {"full": 16, "same": 13, "valid": 10}[shape]
        assert z.shape == (5, L, 3) and lags.size == L
        assert rel_linf(z, ref[:, k0:k0 + L, :]) < 1e-6, shape
    # broadcast kernel (singleton before and after dim), real double, dim 1
    a = rng.standard_normal((17, 4)); b = rng.standard_normal((5, 1))
    z, _ = qups_b200.convd(a, b, 1, "same")
    ref = np.stack([np.convolve(a[:, j], b[:, 0], "same") for j in range(4)], 1)
    assert z.dtype == np.float64 and rel_linf(z, ref) < 1e-13
    # auto-correlation: convd(A) == xcorr(A)
    v = np.array([1.0, -2, 3, -4, 5])
    z, lags = qups_b200.convd(v)
    assert np.allclose(z, np.correlate(v, v, "full")) and list(lags.ravel()) == list(range(-4, 5))


@pytest.mark.parametrize("seqtype", ["PW", "FC"])
def test_greens_focustx_das_psf(seqtype):
    """test/BFTest.m:230-317 for synthesised transmits: greens (FSA) -> focusTx -> DAS, PSF within 1.1 mm."""
    from qups_b200 import synth
    from qups_b200.ultrasound import UltrasoundSystem, Sequence
    c0, N = 1500.0, 32
    pn = synth.linear_array(N, 0.3e-3)
    xs, zs = np.linspace(-4e-3, 6e-3, 41), np.linspace(11e-3, 19e-3, 33)
    if seqtype == "PW":
        th = np.deg2rad(np.linspace(-10, 10, 5))
        focus = np.stack([np.sin(th), 0 * th, np.cos(th)], 0)
    else:
        focus = np.stack([np.linspace(-3e-3, 3e-3, 5), np.zeros(5), np.full(5, 15e-3)], 0)
    us = UltrasoundSystem(tx=pn, rx=pn, seq=Sequence(seqtype, focus, c0), scan=synth.scan_cartesian(xs, zs), fs=25e6, fc=6.25e6)
    chd = us.greens(np.array([[2e-3], [0.0], [15e-3]]), np.ones(1), c0=c0, interp="linear")
    assert chd.data.shape[1:] == (N, 5)
    b = us.DAS(chd, interp="cubic")
    b = np.abs(b.cpu().numpy() if hasattr(b, "cpu") else np.asarray(b)).reshape(len(zs), len(xs), order="F")
    assert b.max() > 0
    iz, ix = np.unravel_index(np.argmax(b), b.shape)
    assert abs(zs[iz] - 15e-3) <= 1.1e-3 and abs(xs[ix] - 2e-3) <= 1.1e-3


@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_greens_convolution_kernel_exact_cases(oracle_c, interp, monkeypatch):
    """The convolution form on delays that are exact in fp32 AND fp64 (dyadic geometry): integer and half-integer arrivals hit
    the closed ends of interp1's support (xq == 1, xq == K), the cubic end padding and the nearest tie; must equal the
    fp32 oracle to summation order, for every interpolator, across block boundaries and many chunks."""
    from qups_b200.ultrasound import greens_raw
    fs, c0 = 2.0 ** 24, 2.0 ** 10                     # fs / c0 = 2^14 samples per metre: distances k/2^15 m are exact half-samples
    rng = np.random.default_rng(7)
    K = 37
    kern = (rng.standard_normal(K) + 1j * rng.standard_normal(K)).astype(np.complex64)
    pn = np.array([[0.0], [0.0], [0.0]])
    S = 2500
    zs = rng.integers(40, 9000, S) / 2.0 ** 16         # one-way path z => arrival 2 z 2^14 = k / 2 samples exactly
    ps = np.stack([np.zeros(S), np.zeros(S), zs], 0)
    amp = rng.integers(-4, 5, S).astype(np.float64)
    n0, T = 10, 5000                                   # two train windows
    ref = oracle_c.greens(ps, amp, pn, pn, kern, n0, T, fs, c0, 0.0, 1.0, 0.0, interp)
    got = greens_raw(ps, amp, pn, pn, kern, n0, T, fs, c0, 0.0, 1.0, 0.0, interp).cpu().numpy()
    assert np.abs(ref).max() > 0
    assert rel_linf(got, ref) < 3e-6, rel_linf(got, ref)
    z = np.abs(ref) == 0
    assert np.all(got[z] == 0)                         # exact zeros outside every scatterer's support


@pytest.mark.parametrize("mode", ["conv", "binned", "simple"])
def test_greens_sub_elements(oracle_c, mode, monkeypatch):
    """Element sub-divisions (E sub-elements per element, entry order scatterer -> em -> en, src/UltrasoundSystem.m:785-790):
    every kernel variant against the oracle, and against the sum of the E x E single-sub-element simulations."""
    from qups_b200 import synth
    from qups_b200.ultrasound import greens_raw
    fs, fc, c0 = 20e6, 5e6, 1500.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    N, M, E = 6, 4, 2
    base_n, base_v = synth.linear_array(N, 0.3e-3), synth.linear_array(M, 0.4e-3)
    off = np.array([[-0.07e-3, 0.07e-3], [0.0, 0.0], [0.0, 0.0]])           # sub-element offsets along x
    pn = np.concatenate([base_n + off[:, [e]] for e in range(E)], axis=1)   # column n + N*en
    pv = np.concatenate([base_v + off[:, [e]] for e in range(E)], axis=1)
    rng = np.random.default_rng(11)
    S = 700                                                                  # 2800 entries: three chunks of the staged kernels
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), rng.uniform(-1e-3, 1e-3, S), rng.uniform(3e-3, 12e-3, S)], 0)
    amp = rng.standard_normal(S)
    n0, T, R0 = 40, 700, 3e-4
    monkeypatch.setenv("QUPS_B200_GREENS", mode)
    got = greens_raw(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, "cubic", E=E).cpu().numpy()
    assert got.shape == (T, N, M)
    ref = oracle_c.greens(ps, amp, pn, pv, kern, n0, T, fs, c0, wt0, 1.0, R0, "cubic", E=E)
    assert np.abs(ref).max() > 0
    if mode == "simple":
        assert np.array_equal(got, ref)
    elif mode == "binned":
        assert rel_linf(got, ref) < 2e-6, rel_linf(got, ref)
    else:   # fp64 arrivals: the fp64 oracle on the fp32-rounded inputs is the arbiter (see test_greens_bitexact_vs_oracle)
        f32 = np.float32
        ref64 = oracle_c.greens(ps.astype(f32), amp.astype(f32), pn.astype(f32), pv.astype(f32), kern, n0, T, fs, c0, wt0, 1.0, R0,
                                "cubic", dtype=np.float64, E=E)
        assert rel_linf(got, ref64) < 5e-6, rel_linf(got, ref64)
        assert rel_linf(got, ref) < 1e-3
    parts = sum(greens_raw(ps, amp, pn[:, en * N:(en + 1) * N], pv[:, em * M:(em + 1) * M], kern, n0, T, fs, c0, wt0, 1.0, R0,
                           "cubic").cpu().numpy().astype(np.complex128) for em in range(E) for en in range(E))
    assert rel_linf(got, parts) < 5e-6, rel_linf(got, parts)
