"""GPU parity of the staged look-up-table delay-and-sum (qups_wsinterpd2 canonical form -> das_tiled LUT mode): what
UltrasoundSystem.bfDAS / bfDASLUT hand to ChannelData.sample2sep -> wsinterpd2 (src/UltrasoundSystem.m:4640-4656,
src/ChannelData.m:1428-1445, kern/wsinterpd2.m:226-235; reference kernel src/interpd.cu:344-396).  Checked against the C
oracle (oracle_wsinterpd2), against the generic strided kernel, and bit-exactly for nearest on integer-valued data."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu
f32 = np.float32


def _tables(P, kind):
    """tau tables in samples as sample2sep builds them: ntau_tx = (tau_tx - t0) fs (I x 1 x M), ntau_rx = tau_rx fs (I x N)."""
    Pi = P["Pi"].reshape(3, -1, order="F")
    c, fs = P["c"], P["fs"]
    M = max(P["Pv"].shape[1], P["Nv"].shape[1])
    Pv = np.broadcast_to(P["Pv"], (3, M))
    Nv = np.broadcast_to(P["Nv"], (3, M))
    rv = Pi[:, :, None] - Pv[:, None, :]
    if kind == "PW":
        dv = (rv * Nv[:, None, :]).sum(0)
    elif kind in ("DV", "FSA"):
        dv = np.linalg.norm(rv, axis=0)
    else:
        dv = np.linalg.norm(rv, axis=0) * np.sign((rv * Nv[:, None, :]).sum(0))
    dr = np.linalg.norm(Pi[:, :, None] - P["Pr"][:, None, :], axis=0)
    return (dr / c * fs).astype(f32), (dv / c * fs).astype(f32)   # I x N, I x M


@pytest.mark.parametrize("kind", ["FC", "PW", "DV"])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_ws2_tiled_matches_oracle_and_generic(oracle_c, kind, interp, monkeypatch):
    import qups_b200
    P = small_problem(kind, nz=40, nx=70, N=21, M=6, T=300, zlim=(2e-3, 14e-3), int_data=(interp == "nearest"))
    trx, ttx = _tables(P, kind)
    I, N, M = trx.shape[0], trx.shape[1], ttx.shape[1]
    Isz = P["Pi"].shape[1:]
    x = P["x"]
    ref = oracle_c.wsinterpd2_inm(x, trx.reshape(I, N, 1), ttx.reshape(I, 1, M), interp=interp).reshape(-1)
    # MATLAB shapes of sample2sep with apdim = [4, 5]: x is T x 1 x 1 x N x M, tables I1 x I2 x I3 x N x 1 and I1 x I2 x I3 x 1 x M
    x5 = x.reshape(x.shape[0], 1, 1, N, M)
    t_rx = trx.reshape(Isz + (N, 1), order="F")
    t_tx = ttx.reshape(Isz + (1, M), order="F")
    got = np.asarray(qups_b200.wsinterpd2(x5, t_tx, t_rx, 1, 1, (4, 5), interp)).reshape(-1, order="F")
    assert qups_b200.last_ws2_kernel() == "ws2_tiled"
    if interp == "nearest":
        assert np.array_equal(got, ref)
    else:
        assert rel_linf(got, ref) < 1e-5, rel_linf(got, ref)
    monkeypatch.setenv("QUPS_B200_WS2_GENERIC", "1")
    gen = np.asarray(qups_b200.wsinterpd2(x5, t_tx, t_rx, 1, 1, (4, 5), interp)).reshape(-1, order="F")
    assert qups_b200.last_ws2_kernel() == "wsinterpd2"
    assert rel_linf(got, gen) < 1e-5


@pytest.mark.parametrize("order", ["t1_inner", "t2_inner"])
@pytest.mark.parametrize("w", [1.0, 0.25, 0.5 - 2j])
def test_ws2_tiled_argument_orders_weights_and_pixel_ranks(oracle_c, order, w):
    """Either table may index the aperture whose traces are adjacent in x; scalar real / complex weight; 1-D pixel list."""
    import qups_b200
    rng = np.random.default_rng(3)
    T, N, M, I = 200, 18, 5, 333
    x = (rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M))).astype(np.complex64)
    x[:4] = 0
    x[-4:] = 0
    base = np.linspace(20, 90, I)[:, None]
    tn = (base + rng.uniform(0, 25, (1, N))).astype(f32)      # smooth over the pixels (windows fit), arbitrary over the aperture
    tm = (0.5 * base + rng.uniform(0, 25, (1, M))).astype(f32)
    ref = oracle_c.wsinterpd2_inm(x, tn.reshape(I, N, 1), tm.reshape(I, 1, M), interp="cubic").reshape(-1) * np.complex64(w)
    if order == "t2_inner":
        got = qups_b200.wsinterpd2(x.reshape(T, N, M), tm.reshape(I, 1, M), tn.reshape(I, N, 1), 1, w, (2, 3), "cubic")
    else:
        got = qups_b200.wsinterpd2(x.reshape(T, N, M), tn.reshape(I, N, 1), tm.reshape(I, 1, M), 1, w, (2, 3), "cubic")
    assert qups_b200.last_ws2_kernel() == "ws2_tiled"
    assert rel_linf(np.asarray(got).reshape(-1), ref) < 1e-5


def test_ws2_tiled_rough_tables_nan_inf_and_out_of_range(oracle_c):
    """Tables that jump between neighbouring pixels (windows larger than a slot -> per-pair path), NaN / Inf / far out-of-range
    entries (contribute 0, kern/wsinterpd2.m: interp1 extrapval 0 then sum 'omitnan')."""
    import qups_b200
    rng = np.random.default_rng(9)
    T, N, M = 180, 17, 4
    I1, I2 = 37, 41
    I = I1 * I2
    x = (rng.standard_normal((T, N, M)) + 1j * rng.standard_normal((T, N, M))).astype(np.complex64)
    tn = rng.uniform(-20, 120, (I, N)).astype(f32)
    tm = rng.uniform(-20, 120, (I, M)).astype(f32)
    tn[5, 3] = np.inf
    tn[77, 0] = -np.inf
    tm[100, 2] = np.nan
    tm[101, 1] = 1e30
    ref = oracle_c.wsinterpd2_inm(x, tn.reshape(I, N, 1), tm.reshape(I, 1, M), interp="linear").reshape(-1)
    got = qups_b200.wsinterpd2(x.reshape(T, 1, N, M), tn.reshape((I1, I2, N, 1), order="F"), tm.reshape((I1, I2, 1, M), order="F"), 1, 1,
                               (3, 4), "linear")
    assert qups_b200.last_ws2_kernel() == "ws2_tiled"
    got = np.asarray(got).reshape(-1, order="F")
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), ok)
    assert rel_linf(got[ok], ref[ok]) < 1e-5


def test_bfdas_lut_path_equals_das_on_the_staged_kernels(oracle_c):
    """UltrasoundSystem.bfDAS (tables from the geometry) vs UltrasoundSystem.DAS on the same data: the two staged paths agree to
    the rounding of the fp32 tables (the reference's own bfDAS-vs-DAS check, test/BFTest.m)."""
    import qups_b200
    from qups_b200 import ultrasound as U
    P = small_problem("FSA", nz=48, nx=64, N=16, M=16, T=400, zlim=(3e-3, 14e-3))
    us = U.UltrasoundSystem(tx=P["Pv"], rx=P["Pr"], seq=U.Sequence("FSA", None, c0=P["c"]), scan=P["Pi"], fs=P["fs"])
    chd = U.ChannelData(P["x"], 0.0, P["fs"])
    b1 = np.asarray(us.DAS(chd, interp="cubic")).reshape(-1)
    b2 = np.asarray(us.bfDAS(chd, interp="cubic")).reshape(-1)
    assert qups_b200.last_ws2_kernel() == "ws2_tiled"
    assert rel_linf(b2, b1) < 2e-3
