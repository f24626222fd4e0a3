"""CPU tests of the extrema oracle (oracle/extrema_oracle.c).

The oracle's real-root finder is not the reference's Jenkins-Traub; it is pinned here
  * against the reference's OWN root finder, src/rpoly.cpp compiled where it lies into
    oracle/_ref/librpoly_ref.so (skipped only where neither that file nor /root/reference exists),
  * against dense sampling, the way the reference's tests check extrema
    (test/test_polynomial_optimization.cpp:396-507: candidate times within 0.01 of the sampled
    ones, maxima within 0.01 of the sampled maximum),
  * against known answers (polynomials built from their roots).
"""
import numpy as np
import pytest

from oracle.oracle_py import ReferenceRpoly

from helpers import random_batch


def solved(oracle, K, D, seed, box=None, lo=None, hi=None):
    if lo is not None:
        pos = oracle.create_random_positions(K, np.asarray(lo, float), np.asarray(hi, float), seed)
        times = oracle.estimate_segment_times(pos, 3.0, 5.0)
    else:
        p, t = random_batch(oracle, 1, K, D, seed=seed, box=box)
        pos, times = p[0], t[0]
    coeffs = oracle.solve_batch_standard(pos[None], times[None])[0][0]
    return np.asarray(coeffs, np.float64), np.asarray(times, np.float64)


def poly_from_roots(roots):
    c = np.array([1.0])
    for r in roots:
        c = np.convolve(c, np.array([-r, 1.0]))   # increasing powers
    return c


# ---- known answers ------------------------------------------------------------------------
def test_real_roots_known_answers(oracle):
    g = poly_from_roots([1.0, 2.0, 3.0])
    np.testing.assert_allclose(oracle.real_roots_in_range(g, 0.0, 4.0), [1.0, 2.0, 3.0], rtol=1e-13)
    np.testing.assert_allclose(oracle.real_roots_in_range(g, 1.5, 2.5), [2.0], rtol=1e-13)
    assert len(oracle.real_roots_in_range(g, 3.5, 9.0)) == 0
    # roots exactly on the interval ends are in range (ref: t < t_start || t > t_stop are dropped)
    np.testing.assert_allclose(oracle.real_roots_in_range(g, 1.0, 3.0), [1.0, 2.0, 3.0], rtol=1e-13)
    # complex pair: (t^2 + 1)(t - 0.5)
    g = np.convolve(np.array([1.0, 0.0, 1.0]), np.array([-0.5, 1.0]))
    np.testing.assert_allclose(oracle.real_roots_in_range(g, -5.0, 5.0), [0.5], rtol=1e-13)
    # degree 15 with clustered but distinct roots
    roots = np.linspace(0.1, 1.5, 15)
    got = oracle.real_roots_in_range(poly_from_roots(roots), 0.0, 2.0)
    assert len(got) == 15
    np.testing.assert_allclose(got, roots, rtol=1e-6)
    # a root of multiplicity: t^3 (t - 2) has one distinct root at 0 and one at 2
    g = np.array([0.0, 0.0, 0.0, -2.0, 1.0])
    np.testing.assert_allclose(oracle.real_roots_in_range(g, 0.0, 3.0), [0.0, 2.0], atol=1e-15)
    # constants and the zero polynomial have no roots (ref src/rpoly.cpp:61-75)
    assert len(oracle.real_roots_in_range(np.array([3.0]), 0.0, 1.0)) == 0


def test_trailing_coefficient_removal(oracle):
    # ref src/rpoly.cpp:44-55: |c| >= machine epsilon counts as non-zero, absolute
    eps = np.finfo(np.float64).eps
    assert oracle.last_nonzero_coefficient(np.array([1.0, 2.0, eps / 2, 0.0])) == 1
    assert oracle.last_nonzero_coefficient(np.array([1.0, 2.0, eps, 0.0])) == 2
    assert oracle.last_nonzero_coefficient(np.zeros(4)) == -1


def test_convolution_matches_numpy(oracle):
    # the candidate polynomial of several dimensions is sum_dim conv(p^(k), p^(k+1))
    rng = np.random.default_rng(3)
    c = rng.uniform(-1, 1, (3, 10))
    for k in range(0, 5):
        g = oracle.candidate_polynomial(c, k)
        want = np.zeros(2 * (10 - k) - 2)
        for d in range(3):
            pk = np.polynomial.polynomial.polyder(c[d], k)
            pk1 = np.polynomial.polynomial.polyder(c[d], k + 1)
            want += np.convolve(pk, pk1)
        np.testing.assert_allclose(g, want, rtol=1e-12, atol=1e-12)
    # one dimension: the next derivative itself (ref LIN.i:412-416)
    g = oracle.candidate_polynomial(c, 1, dims=[2])
    np.testing.assert_allclose(g, np.polynomial.polynomial.polyder(c[2], 2), rtol=1e-14)


# ---- pinned against the reference's Jenkins-Traub ------------------------------------------
CASES = [
    # (K, D, seed, box lo, box hi): the reference's own extrema tests
    (100, 1, 1234, [-10.0], [10.0]),                               # test_polynomial_optimization.cpp:418-421
    (100, 3, 978, [-10.0, -9.0, -8.0], [8.0, 9.0, 10.0]),          # :509-516
    (10, 3, 12345, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]),     # bench configuration
]


@pytest.mark.parametrize("K,D,seed,lo,hi", CASES)
def test_roots_match_reference_rpoly(oracle, K, D, seed, lo, hi):
    if not ReferenceRpoly.available():
        pytest.skip("oracle/_ref/librpoly_ref.so absent and /root/reference not mounted")
    ref = ReferenceRpoly()
    coeffs, times = solved(oracle, K, D, seed, lo=lo, hi=hi)
    n_checked = 0
    for k in (1, 2):
        for s in range(K):
            T = times[s]
            g = oracle.candidate_polynomial(coeffs[s], k)
            theirs = ref.real_roots_in_range(g, 0.0, T)
            assert theirs is not None
            mine = oracle.segment_candidate_roots(coeffs[s], k, 0.0, T)
            # Simple roots away from the rest-to-rest ends must agree one to one.  (At t = 0 of
            # the first and t = T of the last segment g has a root of multiplicity >= 3, which
            # Jenkins-Traub returns as a cluster of perturbed copies: not comparable, and
            # irrelevant to the extrema because the ends are candidates anyway.)
            inner = lambda r: [t for t in r if 0.03 * T < t < 0.97 * T]  # noqa: E731
            a, b = inner(theirs), inner(mine)
            assert len(a) == len(b), (k, s, theirs, mine)
            if a:
                np.testing.assert_allclose(b, a, rtol=0, atol=2e-6 * T)
                n_checked += len(a)
            # the value the reference would report for this segment
            cands_ref = [0.0, T] + list(theirs)
            cands_mine = [0.0, T] + list(mine)
            v_ref = max(oracle.segment_magnitude(coeffs[s], k, t) for t in cands_ref)
            v_mine = max(oracle.segment_magnitude(coeffs[s], k, t) for t in cands_mine)
            assert abs(v_ref - v_mine) <= 1e-8 * max(v_ref, 1e-12), (k, s)
    assert n_checked > K


def test_random_polynomials_match_reference_rpoly(oracle):
    """test/test_polynomial.cpp:79-128 draws coefficients in [-100, 100] and a random range."""
    if not ReferenceRpoly.available():
        pytest.skip("oracle/_ref/librpoly_ref.so absent and /root/reference not mounted")
    ref = ReferenceRpoly()
    rng = np.random.default_rng(1234567)
    n_roots = 0
    for _ in range(300):
        n = int(rng.integers(3, 12))
        c = rng.uniform(-100.0, 100.0, n)
        t0 = rng.uniform(-3.0, 0.0)
        t1 = rng.uniform(t0, 3.0)
        theirs = ref.real_roots_in_range(c, t0, t1)
        mine = oracle.real_roots_in_range(c[: oracle.last_nonzero_coefficient(c) + 1], t0, t1)
        assert len(theirs) == len(mine), (c, t0, t1, theirs, mine)
        if len(mine):
            np.testing.assert_allclose(mine, theirs, rtol=1e-9, atol=1e-9)
        n_roots += len(mine)
    assert n_roots > 100


# ---- the reference's own property: analytic extrema vs dense sampling --------------------------
def sampled_candidates(oracle, seg, k, T, dt=0.001):
    """ref computeSegmentMaximumMagnitudeCandidatesBySampling (LIN.i:439-468)."""
    ts = np.arange(-1, int(np.ceil((T + dt) / dt)) + 2) * dt
    v = np.array([oracle.segment_magnitude(seg, k, t) ** 2 for t in ts])
    direction = np.diff(v)
    out = []
    for i in range(1, len(direction)):
        if np.sign(direction[i - 1]) != np.sign(direction[i]):
            out.append(ts[i])
    return np.array(out)


@pytest.mark.parametrize("K,D,seed,lo,hi", [(20, 1, 1234, [-10.0], [10.0]),
                                            (20, 3, 978, [-10.0, -9.0, -8.0], [8.0, 9.0, 10.0])])
def test_candidates_against_sampling(oracle, K, D, seed, lo, hi):
    coeffs, times = solved(oracle, K, D, seed, lo=lo, hi=hi)
    for s in range(K):
        T = times[s]
        mine = oracle.segment_candidate_roots(coeffs[s], 1, 0.0, T)
        sampled = sampled_candidates(oracle, coeffs[s], 1, T)
        for t in mine:   # ref checkExtrema(testee = analytic, reference = sampling, tol 0.01)
            # At the rest ends g has a root of multiplicity 7 ((t - T)^4 (t - T)^3): in floating point it
            # splits into a cluster of radius ~eps^(1/7) T (the reference's root finder returns the same
            # kind of cluster, see test_roots_match_reference_rpoly); the magnitude there is ~0.
            if (s == 0 and t < 0.02 * T) or (s == K - 1 and t > 0.98 * T):
                continue
            assert np.min(np.abs(sampled - t)) < 0.01, (s, t, sampled)


@pytest.mark.parametrize("K,D,seed,lo,hi", CASES)
def test_maximum_against_sampling(oracle, K, D, seed, lo, hi):
    coeffs, times = solved(oracle, K, D, seed, lo=lo, hi=hi)
    for k in (1, 2):
        # ref getMaximumMagnitude (test_polynomial_optimization.cpp:48-59), dt = 0.01
        best = -1e9
        for s in range(K):
            for t in np.arange(0.0, times[s], 0.01):
                best = max(best, oracle.segment_magnitude(coeffs[s], k, t))
        for mode in (0, 1):
            # keep_small: without the reference's coefficient truncation (see the next test)
            r = oracle.minmax_magnitude(coeffs, times, k, mode, keep_small=True)
            t, v, seg = r["max"]
            assert abs(v - best) < 0.01          # EXPECT_NEAR(v_max_ref, v_max.value, 0.01)
            assert v >= best - 1e-9              # an analytic maximum is never below a sample
            assert 0 <= seg < K and 0.0 <= t <= times[seg]
            assert abs(oracle.segment_magnitude(coeffs[seg], k, t) - v) <= 1e-12 * max(v, 1.0)
            # the reference-compatible result can only lose candidates on truncated segments
            t_c, v_c, seg_c = oracle.minmax_magnitude(coeffs, times, k, mode)["max"]
            assert v_c <= v * (1 + 1e-12)
            assert abs(oracle.segment_magnitude(coeffs[seg_c], k, t_c) - v_c) <= 1e-12 * max(v_c, 1.0)
            if times.max() < 10.0:
                assert abs(v_c - v) <= 1e-9 * v
        t, v, seg = oracle.minmax_magnitude(coeffs, times, k, 1)["min"]
        assert v <= 1e-9                         # rest-to-rest: the magnitude is zero at the start
        assert abs(oracle.segment_magnitude(coeffs[seg], k, t) - v) <= 1e-12


def test_reference_truncation_misses_extrema(oracle):
    """Documents a reference defect that the parity mode reproduces on purpose: src/rpoly.cpp:44-55
    drops trailing coefficients below 2.2e-16 ABSOLUTE.  For the bench trajectory seed 12345 the
    8th segment lasts 17.2 s; its candidate polynomial loses its two highest coefficients, the
    root at t = 16.74 disappears and with it the true maximum speed of the whole trajectory."""
    coeffs, times = solved(oracle, 10, 3, 12345, lo=[-10.0, -20.0, -10.0], hi=[10.0, 20.0, 10.0])
    s = 7
    g = oracle.candidate_polynomial(coeffs[s], 1)
    assert oracle.last_nonzero_coefficient(g) < len(g) - 1          # truncated
    compat = oracle.segment_candidate_roots(coeffs[s], 1, 0.0, times[s])
    exact = oracle.segment_candidate_roots(coeffs[s], 1, 0.0, times[s], keep_small=True)
    assert len(exact) == len(compat) + 1
    sampled = max(oracle.segment_magnitude(coeffs[s], 1, t) for t in np.arange(0.0, times[s], 0.001))
    v_exact = oracle.minmax_magnitude(coeffs, times, 1, 0, keep_small=True)["max"][1]
    v_compat = oracle.minmax_magnitude(coeffs, times, 1, 0)["max"][1]
    assert abs(v_exact - sampled) < 1e-5 and v_exact >= sampled
    assert v_compat < sampled - 0.02
    if ReferenceRpoly.available():     # the reference's own root finder agrees with the parity mode
        theirs = ReferenceRpoly().real_roots_in_range(g, 0.0, times[s])
        np.testing.assert_allclose(theirs, compat, rtol=1e-6)


def test_modes_and_dimension_subsets(oracle, oracle_ld):
    coeffs, times = solved(oracle, 10, 3, 12345 + 7)
    for dims in ([0, 1, 2], [0, 2], [1]):
        for k in (0, 1, 2, 3):
            a = oracle.minmax_magnitude(coeffs, times, k, 1, dims=dims, want_candidates=True)
            b = oracle_ld.minmax_magnitude(coeffs, times, k, 1, dims=dims)
            assert abs(a["max"][1] - float(b["max"][1])) <= 1e-10 * max(a["max"][1], 1.0)
            # the maximum over the listed dimensions is bounded by the one over all of them
            full = oracle.minmax_magnitude(coeffs, times, k, 1)
            assert a["max"][1] <= full["max"][1] * (1 + 1e-12)
            for s, roots in enumerate(a["candidates"]):
                assert np.all(np.diff(roots) >= 0) and np.all(roots >= 0) and np.all(roots <= times[s])
    # mode 0 never looks at interior segment ends; both modes agree on the maximum when it is
    # attained at a root (always the case for rest-to-rest velocity)
    m0 = oracle.minmax_magnitude(coeffs, times, 1, 0)["max"]
    m1 = oracle.minmax_magnitude(coeffs, times, 1, 1)["max"]
    assert abs(m0[1] - m1[1]) <= 1e-12 * m1[1]
