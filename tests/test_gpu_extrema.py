"""GPU parity: extrema of the magnitude of a derivative through the C ABI
(minsnap_extrema / minsnap_extrema_host) against the CPU oracle (oracle/extrema_oracle.c), which is
itself pinned against the reference's own root finder in tests/test_extrema_oracle.py.

Parity bars (written here because north_star is silent on this row):
  * extremum VALUE                    <= 1e-8 relative  (well conditioned: the derivative of the
                                         magnitude vanishes at an interior extremum)
  * reported (segment, time)          the magnitude re-evaluated there reproduces the value to 1e-9
                                         (time and segment themselves are compared only when the
                                         runner-up is clearly lower: ties at shared vertices are
                                         decided by rounding in the reference too)
  * candidate roots                   simple roots away from the rest ends: <= 1e-7 T
"""
import numpy as np
import pytest

from helpers import random_batch

pytestmark = pytest.mark.gpu

KEEP = 16


def solve_batch(ms, torch, pos, times):
    p = torch.from_numpy(pos).cuda()
    t = torch.from_numpy(times).cuda()
    r = ms.solve_standard(p, t, want_status=False)
    return r["coeffs"], t


def check_against_oracle(oracle, coeffs_h, times_h, got, k, mode, dims=None, keep_small=False, check_roots=True):
    B, K = times_h.shape
    worst = 0.0
    for b in range(B):
        want = oracle.minmax_magnitude(coeffs_h[b], times_h[b], k, mode & 1, dims=dims, want_candidates=True,
                                       keep_small=keep_small)
        for which in (("max", "min") if mode & 1 else ("max",)):
            t_w, v_w, s_w = want[which]
            t_g = float(got[which + "_time"][b])
            v_g = float(got[which + "_value"][b])
            s_g = int(got[which + "_segment"][b])
            scale = max(abs(v_w), 1e-9)
            assert abs(v_g - v_w) <= 1e-8 * scale + 1e-12, (b, which, v_g, v_w)
            worst = max(worst, abs(v_g - v_w) / scale)
            assert 0 <= s_g < K and 0.0 <= t_g <= times_h[b, s_g]
            re = oracle.segment_magnitude(coeffs_h[b, s_g], k, t_g, dims=dims)
            assert abs(re - v_g) <= 1e-9 * max(abs(v_g), 1.0), (b, which, re, v_g)
        if check_roots and "root_count" in got:
            for s in range(K):
                T = times_h[b, s]
                mine = got["root_times"][b, s, : got["root_count"][b, s]]
                theirs = want["candidates"][s]
                assert np.all(np.diff(mine) >= 0)
                # candidate list: start, end, roots -- each with the magnitude there
                n_c = 2 + got["root_count"][b, s]
                ct, cv = got["cand_times"][b, s, :n_c], got["cand_values"][b, s, :n_c]
                assert ct[0] == 0.0 and ct[1] == T
                for t_c, v_c in zip(ct, cv):
                    re = oracle.segment_magnitude(coeffs_h[b, s], k, t_c, dims=dims)
                    assert abs(re - v_c) <= 1e-9 * max(abs(re), 1.0)
                inner = lambda r: np.array([t for t in r if 0.03 * T < t < 0.97 * T])  # noqa: E731
                a, c = inner(theirs), inner(mine)
                if len(a) == len(c):
                    if len(a):
                        np.testing.assert_allclose(c, a, rtol=0, atol=1e-7 * T)
                else:
                    # a pair of nearly coincident roots may exist in one precision and not in the other;
                    # it cannot carry an extremum that matters (the value test above covers it)
                    assert abs(len(a) - len(c)) % 2 == 0, (b, s, a, c)
    return worst


def to_host(r):
    return {k: v.cpu().numpy() for k, v in r.items()}


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("mode", [0, 1])
def test_extrema_batch_matches_oracle(ms, oracle, torch_cuda, k, mode):
    torch = torch_cuda
    pos, times = random_batch(oracle, 96, 10)
    coeffs, t = solve_batch(ms, torch, pos, times)
    got = to_host(ms.extrema(coeffs, t, k, mode=mode, want_roots=True))
    check_against_oracle(oracle, coeffs.cpu().numpy(), times, got, k, mode)


@pytest.mark.parametrize("K,D,seed,lo,hi", [
    (100, 1, 1234, [-10.0], [10.0]),                        # ref test_polynomial_optimization.cpp:418-421
    (100, 3, 978, [-10.0, -9.0, -8.0], [8.0, 9.0, 10.0]),   # ref :509-516
])
def test_extrema_reference_test_cases(ms, oracle, torch_cuda, K, D, seed, lo, hi):
    torch = torch_cuda
    pos = oracle.create_random_positions(K, np.asarray(lo), np.asarray(hi), seed)[None]
    times = oracle.estimate_segment_times(pos[0], 3.0, 5.0)[None].astype(np.float64)
    coeffs, t = solve_batch(ms, torch, np.ascontiguousarray(pos), times)
    ch = coeffs.cpu().numpy()
    for k in (1, 2):                                        # VELOCITY, ACCELERATION as in the reference test
        for mode in (0, 1):
            got = to_host(ms.extrema(coeffs, t, k, mode=mode, want_roots=True))
            check_against_oracle(oracle, ch, times, got, k, mode)
        # ref: EXPECT_NEAR(v_max_ref, v_max.value, 0.01) with v_max_ref from 0.01 s sampling
        best = max(oracle.segment_magnitude(ch[0, s], k, tt) for s in range(K)
                   for tt in np.arange(0.0, times[0, s], 0.01))
        v = float(ms.extrema(coeffs, t, k, mode=0 | KEEP)["max_value"][0])
        assert abs(v - best) < 0.01 and v >= best - 1e-9


def test_extrema_keep_small_coefficients(ms, oracle, torch_cuda):
    """The reference truncation (absolute 2.2e-16) against the exact candidate polynomial."""
    torch = torch_cuda
    pos, times = random_batch(oracle, 64, 10)
    coeffs, t = solve_batch(ms, torch, pos, times)
    ch = coeffs.cpu().numpy()
    for mode in (0, 1):
        compat = to_host(ms.extrema(coeffs, t, 1, mode=mode, want_roots=True))
        exact = to_host(ms.extrema(coeffs, t, 1, mode=mode | KEEP, want_roots=True))
        check_against_oracle(oracle, ch, times, compat, 1, mode)
        check_against_oracle(oracle, ch, times, exact, 1, mode, keep_small=True)
        assert np.all(exact["max_value"] >= compat["max_value"] * (1 - 1e-12))
    # seed 12345 (b = 0): the truncated polynomial of its 8th segment loses the root that carries the maximum
    assert exact["max_value"][0] > compat["max_value"][0] + 0.02
    for b in range(8):
        best = max(oracle.segment_magnitude(ch[b, s], 1, tt) for s in range(10)
                   for tt in np.arange(0.0, times[b, s], 0.01))
        assert abs(exact["max_value"][b] - best) < 1e-3 and exact["max_value"][b] >= best - 1e-9


@pytest.mark.parametrize("dims", [[0, 2], [1], [0, 1]])
def test_extrema_dimension_subsets(ms, oracle, torch_cuda, dims):
    torch = torch_cuda
    pos, times = random_batch(oracle, 32, 6)
    coeffs, t = solve_batch(ms, torch, pos, times)
    for k in (1, 2):
        got = to_host(ms.extrema(coeffs, t, k, mode=1, dimensions=dims, want_roots=True))
        check_against_oracle(oracle, coeffs.cpu().numpy(), times, got, k, 1, dims=dims)


@pytest.mark.parametrize("D", [1, 2])
def test_extrema_other_dimensions(ms, oracle, torch_cuda, D):
    torch = torch_cuda
    pos, times = random_batch(oracle, 32, 8, D=D)
    coeffs, t = solve_batch(ms, torch, pos, times)
    for mode in (0, 1):
        got = to_host(ms.extrema(coeffs, t, 1, mode=mode, want_roots=True))
        check_against_oracle(oracle, coeffs.cpu().numpy(), times, got, 1, mode)


@pytest.mark.parametrize("N", [6, 8, 12])
def test_extrema_random_polynomials(ms, oracle, torch_cuda, N):
    """Not solved trajectories: arbitrary coefficients (ref test/test_polynomial.cpp:79-128 style)."""
    torch = torch_cuda
    rng = np.random.default_rng(100 + N)
    B, K, D = 64, 3, 3
    c = rng.uniform(-100.0, 100.0, (B, K, D, N))
    times = rng.uniform(0.2, 1.5, (B, K))
    cd, td = torch.from_numpy(c).cuda(), torch.from_numpy(times).cuda()
    for k in (0, 1, 2):
        got = to_host(ms.extrema(cd, td, k, mode=1, want_roots=True))
        check_against_oracle(oracle, c, times, got, k, 1)


def test_extrema_host_entry_point(ms, oracle, torch_cuda):
    pos, times = random_batch(oracle, 5, 10)
    coeffs = ms.solve_standard_host(pos, times)["coeffs"]
    # single trajectory (the call a C++ PolynomialOptimization makes) and a small batch
    for sl in (slice(0, 1), slice(0, 5)):
        for mode in (0, 1):
            got = ms.extrema_host(coeffs[sl], times[sl], 2, mode=mode, want_roots=True)
            check_against_oracle(oracle, coeffs[sl], times[sl], got, 2, mode)
    # empty batch
    e = ms.extrema_host(np.zeros((0, 10, 3, 10)), np.zeros((0, 10)), 1, mode=1)
    assert e["max_value"].shape == (0,)


def test_extrema_argument_errors(ms, torch_cuda):
    torch = torch_cuda
    c = torch.zeros((2, 3, 3, 10), dtype=torch.float64, device="cuda")
    t = torch.ones((2, 3), dtype=torch.float64, device="cuda")
    with pytest.raises(ms.MinsnapError):
        ms.extrema(c, t, 9)             # N - derivative - 1 must stay positive (ref static_assert LIN.i:384)
    with pytest.raises(ms.MinsnapError):
        ms.extrema(c, t, 1, mode=3)
    with pytest.raises(ValueError):
        ms.extrema(c, t, 1, dimensions=[3])   # ref src/segment.cpp:97-102: out of bounds dimension
    # all-zero polynomials: no roots, value 0 at (segment 0, time 0) like Extremum()
    r = ms.extrema(c, t, 1, mode=0)
    assert float(r["max_value"][0]) == 0.0 and int(r["max_segment"][0]) == 0 and float(r["max_time"][0]) == 0.0


def test_extrema_full_size_properties(ms, oracle, torch_cuda):
    """65,536 x K=10 (BASELINE configs[1]): size-independent properties."""
    torch = torch_cuda
    B, K = 65536, 10
    pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
    pos = torch.from_numpy(pos_h).cuda()
    times = ms.estimate_segment_times(pos, 3.0, 5.0)
    coeffs = ms.solve_standard(pos, times, want_status=False)["coeffs"]
    for k in (1, 2):
        e = ms.extrema(coeffs, times, k, mode=1 | KEEP)
        # (1) the reported maximum is the magnitude at the reported place (GPU sampler, a17-a19)
        seg = e["max_segment"].long()
        start = torch.cumsum(times, 1) - times
        t_abs = start.gather(1, seg[:, None])[:, 0] + e["max_time"]
        # an extremum at the very end of a segment belongs to the next one for the sampler: nudge inside
        t_abs = torch.minimum(t_abs, times.sum(1) * (1 - 1e-12))
        s = ms.sample_at(coeffs, times, t_abs[:, None].contiguous(), n_deriv=k + 1)
        mag = s[:, 0, k, :].norm(dim=1)
        assert float(((mag - e["max_value"]).abs() / e["max_value"].clamp_min(1e-9)).max()) < 1e-6
        # (2) no sample of a uniform grid exceeds it
        M = 256
        u = ms.sample_uniform(coeffs, times, M, k + 1)
        grid_max = u[:, :, k, :].norm(dim=2).max(dim=1).values
        assert bool((grid_max <= e["max_value"] * (1 + 1e-9) + 1e-12).all())
        assert float(((e["max_value"] - grid_max) / e["max_value"]).max()) < 0.05
        # (3) rest to rest: the minimum is zero
        assert float(e["min_value"].max()) < 1e-6
        # (4) the reference-compatible mode never reports more than the exact one
        c = ms.extrema(coeffs, times, k, mode=1)
        assert bool((c["max_value"] <= e["max_value"] * (1 + 1e-12)).all())
    # spot parity on a slice
    sl = slice(1000, 1032)
    got = to_host({kk: v[sl] for kk, v in ms.extrema(coeffs, times, 1, mode=0).items()})
    check_against_oracle(oracle, coeffs[sl].cpu().numpy(), times[sl].cpu().numpy(), got, 1, 0)


@pytest.mark.parametrize("N,K,D", [(4, 1, 1), (4, 3, 2), (6, 1, 3), (10, 1, 3), (12, 2, 1)])
def test_extrema_small_shapes_and_high_derivatives(ms, oracle, torch_cuda, N, K, D):
    """Single segments, one dimension, the lowest polynomial order, and every derivative up to N - 2
    (where the candidate polynomial is linear or constant)."""
    torch = torch_cuda
    rng = np.random.default_rng(7 * N + K + D)
    B = 16
    c = rng.uniform(-5.0, 5.0, (B, K, D, N))
    c[0] = 0.0                                   # the zero trajectory
    c[1, :, :, 1:] = 0.0                         # constants: no roots at any derivative
    times = rng.uniform(0.5, 2.0, (B, K))
    cd, td = torch.from_numpy(c).cuda(), torch.from_numpy(times).cuda()
    for k in range(0, N - 1):
        for mode in (0, 1):
            got = to_host(ms.extrema(cd, td, k, mode=mode, want_roots=True))
            check_against_oracle(oracle, c, times, got, k, mode)
    assert ms.extrema_max_roots(N, N - 2, D) == (1 if D > 1 else 0)
