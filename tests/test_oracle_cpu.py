"""CPU suite: pins the oracle (C restatement vs independent NumPy restatement vs analytic properties).

Reference semantics under test: kern/das_spec.m:391-561 (CPU branch), MATLAB interp1(v, xq, method, 0),
kern/wsinterpd2.m:240-308, src/UltrasoundSystem.m:778-851.
"""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf


def test_interp1_linear_matches_numpy_interp(oracle_np):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(37) + 1j * rng.standard_normal(37)
    xq = rng.uniform(-3, 42, 500)
    got = oracle_np.interp1(v, xq, "linear", 0)
    grid = np.arange(1, 38)
    ref = np.interp(xq, grid, v.real, left=0, right=0) + 1j * np.interp(xq, grid, v.imag, left=0, right=0)
    assert np.allclose(got, ref, atol=1e-12)


def test_interp1_grid_points_and_extrapolation(oracle_np, oracle_c):
    v = (np.arange(1, 12) ** 2 + 1j * np.arange(11)).astype(np.complex128)
    xq = np.array([1.0, 2.0, 5.0, 11.0, 0.999, 11.001, np.nan, -np.inf, np.inf])
    for m in ("nearest", "linear", "cubic"):
        for impl in (oracle_np.interp1(v, xq, m, 0), oracle_c.interp1(v, xq, m, np.float64)):
            assert np.allclose(impl[:4], v[[0, 1, 4, 10]])
            assert np.all(impl[4:] == 0)


def test_interp1_nearest_rounds_half_away(oracle_np, oracle_c):
    v = np.arange(10, 20).astype(np.complex64)
    xq = np.array([1.5, 2.5, 3.49, 9.5, 1.4999999], dtype=np.float32)
    exp = np.array([11, 12, 12, 19, 10])
    assert np.array_equal(oracle_np.interp1(v, xq, "nearest", 0).real, exp)
    assert np.array_equal(oracle_c.interp1(v, xq, "nearest").real, exp)


def test_interp1_cubic_reproduces_quadratics_everywhere(oracle_np):
    # Keys a=-1/2 with the 3v1-3v2+v3 end padding reproduces polynomials up to degree 2 on [1,T]
    t = np.arange(1, 21, dtype=np.float64)
    v = 0.3 * t * t - 2 * t + 1
    xq = np.linspace(1, 20, 401)
    got = oracle_np.interp1(v, xq, "cubic", 0)
    assert np.allclose(got, 0.3 * xq * xq - 2 * xq + 1, atol=1e-10)


def test_interp1_cubic_interior_equals_catmull_rom(oracle_np):
    # interior weights identical to the reference GPU sampler (src/interpd.cu:103-112)
    rng = np.random.default_rng(3)
    v = rng.standard_normal(30)
    xq = rng.uniform(3, 28, 200)
    k = np.floor(xq).astype(int)
    u = xq - k
    s0, s1, s2, s3 = v[k - 2], v[k - 1], v[k], v[k + 1]
    a0 = u * (-1 + u * (2 * u - 1) * -1) if False else 0 + u * (-1 + u * (2 - u))
    a1 = 2 + u * (0 + u * (-5 + 3 * u))
    a2 = 0 + u * (1 + u * (4 - 3 * u))
    a3 = 0 + u * (0 + u * (-1 + u))
    ref = 0.5 * (s0 * a0 + s1 * a1 + s2 * a2 + s3 * a3)
    assert np.allclose(oracle_np.interp1(v, xq, "cubic", 0), ref, atol=1e-12)


@pytest.mark.parametrize("method", ["nearest", "linear", "cubic", "lanczos3"])
def test_interp1_c_equals_numpy_fp32_bitexact(oracle_np, oracle_c, method):
    rng = np.random.default_rng(5)
    v = (rng.standard_normal(40) + 1j * rng.standard_normal(40)).astype(np.complex64)
    xq = rng.uniform(-2, 44, 600).astype(np.float32)
    xq[:8] = [1, 1.5, 2, 39, 39.5, 40, 40.0001, 0.9999]
    a = oracle_np.interp1(v, xq, method, 0).astype(np.complex64)
    b = oracle_c.interp1(v, xq, method)
    if method == "lanczos3":
        assert np.allclose(a, b, atol=2e-6)
    else:
        assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA", "DV"])
@pytest.mark.parametrize("fun", ["DAS", "SYN", "MUL", "BF", "delays"])
def test_das_c_equals_numpy_bitexact(oracle_np, oracle_c, kind, fun):
    P = small_problem(kind, nz=9, nx=7, N=6, M=4, T=120)
    kw = oracle_kwargs(P["opts"])
    for interp in ("nearest", "linear", "cubic"):
        a = oracle_np.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp, **kw)
        b = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], P["t0"], P["fs"], P["c"], interp=interp, **kw)
        assert a.shape == b.shape
        assert np.array_equal(a, b), (fun, kind, interp, rel_linf(a, b))
        if fun != "delays":
            assert np.abs(b).max() > 0


def test_das_apod_cinv_tpose_frames_t0(oracle_np, oracle_c):
    P = small_problem("FC", nz=8, nx=6, N=5, M=4, T=120, F=2)
    rng = np.random.default_rng(7)
    Isz = P["Pi"].shape[1:]
    apods = [rng.uniform(0, 1, Isz + (5, 1)).astype(np.float32), rng.uniform(0, 1, (1, 1, 1, 1, 4)).astype(np.float32),
             (rng.uniform(0, 1, (Isz[0], 1, 1, 5, 4)) > 0.3).astype(np.float32)]
    c = rng.uniform(1500, 1580, Isz).astype(np.float32)
    t0 = rng.uniform(-2e-7, 2e-7, 4)
    kw = oracle_kwargs(P["opts"])
    for fun in ("DAS", "SYN", "MUL", "BF"):
        a = oracle_np.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="cubic", apod=apods, **kw)
        b = oracle_c.das_spec(fun, P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="cubic", apod=apods, **kw)
        assert np.array_equal(a, b), fun
    # transposed data layout gives the same image
    xt = np.asfortranarray(np.swapaxes(P["x"], 1, 2))
    a = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="linear", **kw)
    b = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], xt, t0, P["fs"], c, interp="linear", tpose=True, **kw)
    assert np.array_equal(a, b)
    # complex apodization
    ac = [(apods[0] * np.exp(1j * 0.3)).astype(np.complex64)]
    a = oracle_np.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="linear", apod=ac, **kw)
    b = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], t0, P["fs"], c, interp="linear", apod=ac, **kw)
    assert rel_linf(a, b) < 1e-6


def test_das_modulation_and_fp64_arbiter(oracle_np, oracle_c):
    P = small_problem("PW", nz=10, nx=6, N=6, M=3, T=140)
    kw = oracle_kwargs(P["opts"])
    args = (P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], 1e-7, P["fs"], P["c"])
    a = oracle_np.das_spec("DAS", *args, interp="cubic", fmod=5e6, **kw)
    b = oracle_c.das_spec("DAS", *args, interp="cubic", fmod=5e6, **kw)
    assert rel_linf(a, b) < 2e-6  # cos/sin differ by an ulp between libm and NumPy
    d = oracle_c.das_spec("DAS", *args, interp="cubic", fmod=5e6, dtype=np.float64, **kw)
    assert rel_linf(b, d) < 2e-3  # fp32 delay rounding on white noise vs the fp64 arbiter
    d2 = oracle_np.das_spec("DAS", *args, interp="cubic", fmod=5e6, dtype=np.float64, **kw)
    assert rel_linf(d, d2) < 1e-12


def test_das_linearity_and_sum_consistency(oracle_c):
    P = small_problem("FC", nz=8, nx=5, N=6, M=4, T=120)
    kw = oracle_kwargs(P["opts"])
    args = lambda x: (P["Pi"], P["Pr"], P["Pv"], P["Nv"], x, 0.0, P["fs"], P["c"])
    bf = oracle_c.das_spec("BF", *args(P["x"]), interp="cubic", dtype=np.float64, **kw)
    das = oracle_c.das_spec("DAS", *args(P["x"]), interp="cubic", dtype=np.float64, **kw)
    syn = oracle_c.das_spec("SYN", *args(P["x"]), interp="cubic", dtype=np.float64, **kw)
    mul = oracle_c.das_spec("MUL", *args(P["x"]), interp="cubic", dtype=np.float64, **kw)
    assert np.allclose(bf.sum(axis=(3, 4), keepdims=True), das, atol=1e-10)
    assert np.allclose(bf.sum(axis=4, keepdims=True), syn, atol=1e-10)
    assert np.allclose(bf.sum(axis=3, keepdims=True), mul, atol=1e-10)
    das2 = oracle_c.das_spec("DAS", *args(2.5 * P["x"]), interp="cubic", dtype=np.float64, **kw)
    assert np.allclose(das2, 2.5 * das, atol=1e-9)


def test_wsinterpd2_inm_equals_general_nd_and_bfdas_equals_das(oracle_np, oracle_c):
    # bfDAS: tau_rx = dr./c0, tau_tx = dv./c0 (src/UltrasoundSystem.m:4460-4463) -> sample2sep -> wsinterpd2
    P = small_problem("FC", nz=7, nx=5, N=5, M=3, T=120)
    kw = oracle_kwargs(P["opts"])
    Pi = P["Pi"].reshape(3, -1, order="F")
    dv, dr = oracle_np.tx_rx_distances(Pi, P["Pr"], P["Pv"], P["Nv"], **kw)
    fs, c0 = P["fs"], P["c"]
    t_rx = (dr / c0 * fs).astype(np.float64)[:, :, None]     # I x N x 1
    t_tx = (dv / c0 * fs).astype(np.float64)[:, None, :]     # I x 1 x M
    y_c = oracle_c.wsinterpd2_inm(P["x"], t_rx, t_tx, interp="cubic", dtype=np.float64)
    # general N-D form as sample2sep lifts it: x -> T x 1(I) ... use dims (T|I, N, M)
    y_np = oracle_np.wsinterpd2(P["x"].astype(np.complex128), t_rx, t_tx, 1, 1, (2, 3), "cubic", 0, 0)
    assert np.allclose(y_c.reshape(-1, order="F"), y_np.reshape(-1, order="F"), atol=1e-10)
    das = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], 0.0, fs, c0, interp="cubic", dtype=np.float64, **kw)
    assert np.allclose(das.reshape(-1, order="F"), y_c.reshape(-1, order="F"), atol=1e-7)


def test_interptest_generator_wsinterpd(oracle_np):
    # test/interpTest.m:28-47 data generator and :96-143 check: wsinterpd vs a loop of interp1 then sum(w.*y0, dsum)
    I, T, N, M, F = 16, 32, 4, 3, 2
    t = np.arange(T)[:, None, None]
    n = np.arange(N)[None, :, None]
    f = np.arange(F)[None, None, :]
    x = np.exp(2j * np.pi * (0.5 + f / 2 * n / 4) * t / T)            # T x N x F
    rng = np.random.default_rng(11)
    tau = rng.uniform(-2, T + 1, (I, N, M))                               # I x N x M
    w = rng.uniform(0, 1, (I, N, M))
    x4 = x[:, :, None, :]                                                 # T x N x 1 x F
    for terp in ("cubic", "nearest", "linear"):
        y0 = np.zeros((I, N, M, F), complex)
        for ff in range(F):
            for m in range(M):
                for nn in range(N):
                    y0[:, nn, m, ff] = oracle_np.interp1(x[:, nn, ff], 1 + tau[:, nn, m], terp, 0)
        for dsum in ((), (2,), (3,), (2, 3)):
            ref = (w[..., None] * y0)
            if dsum:
                ref = ref.sum(axis=tuple(d - 1 for d in dsum), keepdims=True)
            got = oracle_np.wsinterpd(x4, tau[..., None], 1, w[..., None], dsum, terp, 0)
            assert np.allclose(got, ref, atol=1e-12), (terp, dsum)


def test_greens_c_equals_numpy(oracle_np, oracle_c):
    from qups_b200 import synth
    fs, fc = 20e6, 5e6
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    pn = synth.linear_array(5, 0.3e-3)
    rng = np.random.default_rng(2)
    ps = np.stack([rng.uniform(-2e-3, 2e-3, 7), np.zeros(7), rng.uniform(4e-3, 8e-3, 7)], 0)
    amp = rng.standard_normal(7)
    n0, T = 60, 200
    for interp in ("linear", "cubic"):
        a = oracle_np.greens(ps, amp, pn, pn, kern, n0, T, fs, 1500.0, wt0, 1.0, 3e-4, interp)
        b = oracle_c.greens(ps, amp, pn, pn, kern, n0, T, fs, 1500.0, wt0, 1.0, 3e-4, interp)
        assert np.abs(a).max() > 0
        assert rel_linf(a, b) < 1e-6
    d = oracle_c.greens(ps, amp, pn, pn, kern, n0, T, fs, 1500.0, wt0, 1.0, 3e-4, "cubic", dtype=np.float64)
    assert rel_linf(b, d) < 1e-3


def test_interp1_cubic_equals_independent_keys_convolution_incl_ends(oracle_c):
    """Second, independent statement of interp1(..., 'cubic') (R2020b+: cubic convolution, Keys 1981, a = -1/2): the sample
    sequence extended by the Keys boundary condition v(0) = 3v(1) - 3v(2) + v(3), v(T+1) = 3v(T) - 3v(T-1) + v(T-2),
    convolved with the piecewise-cubic KERNEL W(s) (no per-interval polynomial, no floor) in fp64.  Covers the first and last
    intervals, the grid points and out-of-range queries.  Also checks the polynomials the reference GPU kernel states in its
    comment (src/interpd.cu:108-111), evaluated literally, against the same kernel form."""
    rng = np.random.default_rng(11)
    T = 37
    v = rng.standard_normal(T) + 1j * rng.standard_normal(T)
    vp = np.concatenate([[3 * v[0] - 3 * v[1] + v[2]], v, [3 * v[-1] - 3 * v[-2] + v[-3]]])   # positions 0 .. T+1
    a = -0.5

    def W(s):
        s = np.abs(s)
        return np.where(s <= 1, (a + 2) * s**3 - (a + 3) * s**2 + 1, np.where(s < 2, a * s**3 - 5 * a * s**2 + 8 * a * s - 4 * a, 0.0))

    xq = np.concatenate([np.linspace(1, T, 1441), np.arange(1, T + 1), [0.999, T + 0.001, -3.0, np.nan]])
    pos = np.arange(0, T + 2)
    with np.errstate(invalid="ignore"):
        indep = (vp[None, :] * W(xq[:, None] - pos[None, :])).sum(1)
    indep = np.where((xq >= 1) & (xq <= T), indep, 0)
    got = oracle_c.interp1(v, xq, "cubic", dtype=np.float64)
    assert np.max(np.abs(got - indep)) < 1e-12 * np.max(np.abs(v))
    # the commented Catmull-Rom polynomials of the reference GPU kernel == the Keys kernel on an interior interval
    u = np.linspace(0, 1, 33)
    cr = np.stack([-u**3 + 2 * u**2 - u, 3 * u**3 - 5 * u**2 + 2, -3 * u**3 + 4 * u**2 + u, u**3 - u**2]) * 0.5
    kw = np.stack([W(u + 1), W(u), W(u - 1), W(u - 2)])
    assert np.max(np.abs(cr - kw)) < 1e-14


def test_greens_sub_elements_sum_of_single_sub_element_simulations(oracle_c):
    """Element sub-divisions (entry order scatterer -> em -> en, src/UltrasoundSystem.m:785-790): the E = 2 simulation is the sum
    of the four (em, en) simulations with one sub-element each, to summation order."""
    from qups_b200 import synth
    fs, fc, c0 = 20e6, 5e6, 1500.0
    kern, wt0, _ = synth.greens_kernel(fc, 0.6, fs)
    N, M, E = 5, 3, 2
    bn, bv = synth.linear_array(N, 0.3e-3), synth.linear_array(M, 0.4e-3)
    off = np.array([[-0.07e-3, 0.07e-3], [0.0, 0.0], [0.0, 0.0]])
    pn = np.concatenate([bn + off[:, [e]] for e in range(E)], axis=1)   # column n + N*en
    pv = np.concatenate([bv + off[:, [e]] for e in range(E)], axis=1)
    rng = np.random.default_rng(11)
    S = 200
    ps = np.stack([rng.uniform(-3e-3, 3e-3, S), rng.uniform(-1e-3, 1e-3, S), rng.uniform(3e-3, 12e-3, S)], 0)
    amp = rng.standard_normal(S)
    ref = oracle_c.greens(ps, amp, pn, pv, kern, 40, 600, fs, c0, wt0, 1.0, 3e-4, "cubic", dtype=np.float64, E=E)
    assert ref.shape == (600, N, M) and np.abs(ref).max() > 0
    parts = sum(oracle_c.greens(ps, amp, pn[:, en * N:(en + 1) * N], pv[:, em * M:(em + 1) * M], kern, 40, 600, fs, c0, wt0, 1.0,
                                3e-4, "cubic", dtype=np.float64) for em in range(E) for en in range(E))
    assert np.abs(ref - parts).max() <= 1e-12 * np.abs(parts).max()
