"""Pins the oracle against the REFERENCE's own DASf kernel (src/bf.cu compiled unmodified -> oracle/_ref/bf.ptx).

Interior samples only: the reference GPU sampler has different trace-edge behaviour than the CPU path
 and is built with --use_fast_math, so the bar is a tolerance, not bit equality."""
import numpy as np
import pytest

from tests.util import small_problem, oracle_kwargs, rel_linf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["FC", "PW", "FSA", "DV"])
@pytest.mark.parametrize("interp", [("nearest", 0), ("linear", 1), ("cubic", 2)])
def test_reference_dasf_matches_oracle_in_the_interior(oracle_c, kind, interp):
    from oracle import ref_ptx
    if not ref_ptx.available():
        pytest.skip("oracle/_ref/bf.ptx or cuda-python not available")
    name, flag = interp
    P = small_problem(kind, nz=48, nx=40, N=16, M=6, T=400, zlim=(3e-3, 12e-3))
    # smooth band-limited traces so fast-math delay differences stay small
    T, N, M = P["x"].shape
    t = np.arange(T)[:, None, None]
    ph = np.random.default_rng(0).uniform(0, 2 * np.pi, (1, N, M))
    x = np.asfortranarray((np.exp(1j * (2 * np.pi * 0.04 * t + ph)) * np.hanning(T)[:, None, None]).astype(np.complex64))
    kw = oracle_kwargs(P["opts"])
    ref = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], x, 0.0, P["fs"], P["c"], interp=name, **kw)
    k = ref_ptx.RefDASf()
    k.prepare(P["Pi"], P["Pr"], P["Pv"], P["Nv"], x, 0.0, P["fs"], P["c"], interp=flag, VS=kw["VS"], DV=kw["DV"])
    k.launch()
    got = k.result(P["Pi"].shape[1:])
    import qups_b200
    ours = qups_b200.das_spec("DAS", P["Pi"].astype(np.float32), P["Pr"].astype(np.float32), P["Pv"].astype(np.float32),
                              P["Nv"].astype(np.float32), x, 0.0, P["fs"], P["c"], *P["opts"], "interp", name)
    assert np.abs(ref).max() > 1
    # nearest flips indices under --use_fast_math; the reference GPU cubic is NOT Catmull-Rom: its Horner
    # nesting (src/interpd.cu:103-106) evaluates 2u^3-u^2-u, -5u^3+3u^2+2, 4u^3-3u^2+u, -u^3+u^2 instead of the
    # commented Catmull-Rom polynomials (:108-111) -- it interpolates the nodes but differs in between.
    tol = {"nearest": 8e-2, "linear": 2e-3, "cubic": 3e-2}[name]
    assert rel_linf(got.reshape(ref.shape), ref) < tol, rel_linf(got.reshape(ref.shape), ref)
    assert rel_linf(ours.reshape(ref.shape), ref) < 1e-5
