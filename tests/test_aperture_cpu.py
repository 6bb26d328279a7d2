"""CPU checks of the aperture-domain restatement (oracle/aperture_np.py) against a SECOND, independent restatement: scalar
loops over one aperture vector written from the reference's code, statement by statement (kern/cohfac.m:45, kern/dmas.m:46-54,
kern/pcf.m:62-76, kern/slsc.m:113-118 + 170-195), and against the properties the reference's help texts state.  The reference's
own tests only smoke-test these functions (test/KernTest.m:220-242: they run and return something), so this is what pins
the array-style restatement that the GPU kernels are checked against."""
import cmath
import math

import numpy as np
import pytest

from oracle import aperture_np as ap


def _vec(seed, A, zeros=()):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(A) + 1j * rng.standard_normal(A)
    for k in zeros:
        v[k] = 0
    return v


def loop_cohfac(v):   # r = abs(sum(b)).^2 ./ sum(abs(b).^2) / numel   (kern/cohfac.m:45)
    s, p = 0j, 0.0
    for x in v:
        s += x
        p += abs(x) ** 2
    return abs(s) ** 2 / p / len(v)


def loop_dmas(v, L=None):   # kern/dmas.m:46-54
    N = len(v)
    if L is None:
        lags = list(range(1, N))
    elif np.ndim(L) == 0:
        lags = list(range(1, int(L) + 1))
    else:
        lags = [i for i in range(1, N) if i in set(int(q) for q in L)]   # intersect(1:N-1, L)
    b = 0j
    for i in lags:
        for n in range(0, N - i):   # sub(bn, 1:N-i) .* sub(bn, 1+i:N)
            b += v[n] * v[n + i]
    return cmath.exp(1j * cmath.phase(b)) * math.sqrt(abs(b))


def _std1(ph):   # std(phi, 1, dim, "omitnan"): population standard deviation of the non-NaN entries
    ph = [p for p in ph if p == p]
    m = sum(ph) / len(ph)
    return math.sqrt(sum((p - m) ** 2 for p in ph) / len(ph))


def loop_pcf(v, gamma=1.0):   # kern/pcf.m:62-76
    phi = [cmath.phase(x) for x in v]
    s0 = _std1(phi)
    aux = [p - math.pi * ((p > 0) - (p < 0)) for p in phi]
    sa = _std1(aux)
    sf = min(s0, sa)
    return max(0.0, 1 - (gamma / math.sqrt(math.pi / 3)) * sf), sf


def loop_slsc(v, L=None, method="average"):   # kern/slsc.m:113-118, 170-195 (kdim singleton)
    A = len(v)
    L = max(1, A // 4) if L is None else L
    lags = list(range(1, int(L) + 1)) if np.ndim(L) == 0 else [int(q) for q in L]
    nl = len(lags)
    if method == "average":
        x = [(q / abs(q)) if abs(q) > 0 else 0j for q in v]   # nan2zero(x ./ vecnorm(x, 2, kdim))
        z = 0j
        for i in range(A):
            for j in range(A):
                if abs(i - j) in lags:
                    z += np.conj(x[i]) * x[j] / (A - abs(i - j)) / 2 / nl
        return z
    z = a = b = 0j
    for i in range(A):
        for j in range(A):
            if abs(i - j) in lags:
                z += np.conj(v[i]) * v[j]
                a += np.conj(v[j]) * v[j]
                b += np.conj(v[i]) * v[i]
    sc = 1 / math.sqrt(a.real) / math.sqrt(b.real) if a.real > 0 and b.real > 0 else 0.0   # nan2zero(rsqrt(a) .* rsqrt(b))
    return z * sc


@pytest.mark.parametrize("A", [1, 2, 7, 33])
def test_cohfac_and_pcf_match_the_loops(A):
    v = _vec(A, A)
    assert np.allclose(ap.cohfac(v, 1)[0], loop_cohfac(v), rtol=1e-12)
    if A > 1:
        w, sf = ap.pcf(v, 1, 0.7)
        lw, lsf = loop_pcf(v, 0.7)
        assert np.allclose(w[0], lw, rtol=1e-12, atol=1e-14) and np.allclose(sf[0], lsf, rtol=1e-12)


@pytest.mark.parametrize("L", [None, 1, 3, 40, [2, 5], [0, 1, 50], []])
def test_dmas_matches_the_loops(L):
    v = _vec(3, 12)
    got = ap.dmas(v, 1, L)[0]
    ref = loop_dmas(v, L)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14), (L, got, ref)


@pytest.mark.parametrize("method", ["average", "ensemble"])
@pytest.mark.parametrize("L", [None, 1, 3, [1, 2, 6], [0, 2], [0]])
def test_slsc_matches_the_loops(method, L):
    v = _vec(5, 13, zeros=(4,))     # one dead element: x ./ |x| is NaN -> 0 in the average estimator
    got = ap.slsc(v, 1, L, method)[0]
    ref = loop_slsc(v, L, method)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14), (method, L, got, ref)


def test_reduced_dimension_and_broadcast_over_the_others():
    rng = np.random.default_rng(9)
    b = rng.standard_normal((3, 9, 4)) + 1j * rng.standard_normal((3, 9, 4))
    for i in range(3):
        for k in range(4):
            v = b[i, :, k]
            assert np.allclose(ap.cohfac(b, 2)[i, 0, k], loop_cohfac(v))
            assert np.allclose(ap.dmas(b, 2, 4)[i, 0, k], loop_dmas(v, 4))
            assert np.allclose(ap.slsc(b, 2, 2, "ensemble")[i, 0, k], loop_slsc(v, 2, "ensemble"))
            assert np.allclose(ap.pcf(b, 2)[0][i, 0, k], loop_pcf(v)[0])


def test_documented_properties():
    A = 16
    coh = np.full(A, 2.0 - 1.0j)                        # identical signals on every element
    assert np.allclose(ap.cohfac(coh, 1), 1.0)          # coherence factor in [0, 1], 1 = fully coherent
    rnd = _vec(1, A)
    assert 0.0 <= ap.cohfac(rnd, 1)[0] <= 1.0
    assert np.allclose(ap.slsc(coh, 1, 4, "average"), 1.0) and np.allclose(ap.slsc(coh, 1, 4, "ensemble"), 1.0)
    w, sf = ap.pcf(coh, 1)
    assert np.allclose(sf, 0.0) and np.allclose(w, 1.0)  # no phase diversity: weight 1
    # dmas of a constant aperture c: sum over lags of (A - lag) c^2, then sign-preserving square root
    z = sum(A - lag for lag in range(1, A)) * coh[0] ** 2
    assert np.allclose(ap.dmas(coh, 1), np.exp(1j * np.angle(z)) * np.sqrt(abs(z)))
    # the ensemble estimator is invariant to a common scale (the reference rescales by a power of two first, kern/slsc.m:182)
    assert np.allclose(ap.slsc(rnd, 1, 3, "ensemble"), ap.slsc(1e-12 * rnd, 1, 3, "ensemble"))
