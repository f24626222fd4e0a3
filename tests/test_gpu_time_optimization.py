"""GPU parity for SURVEY.md section 8 (f) 2: the time-only objective (ref objectiveFunctionTime,
NL.i:765-832) and the numeric time gradient (ref getCostAndGradientTime, NL.i:2155-2243), plus
the additive batched descent driver built on them.

The oracle side follows the reference literally with the oracle's own primitives:
  objective  = computeCost(solveLinear(times)) + time_penalty * total_time^2
  gradient_n = w_d (J_d(T_n+) - J_d(T_n-)) / (2 increment) + w_t,   J_d(T') = sum_dim d^T R(T') d
with d = [d_f; d_p] of the solve at the unperturbed times and R(T') rebuilt for the moved time
(the reference calls updateSegmentTimes + getR, not solveLinear, inside the loop).
Tolerances: objective 1e-8 relative (COST_TOL); gradient 1e-6 of its largest component (the oracle
forms R through A^-T Q A^-1, accurate to ~1e-10, and the central difference divides that by 0.2).
"""
import numpy as np
import pytest

from helpers import (COST_TOL, oracle_gradient, oracle_objective, random_batch, standard_mask,
                     vertex_values)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [4, 10])
def test_time_objective_matches_oracle(ms, oracle, torch_cuda, K):
    torch = torch_cuda
    B, S, penalty = 12, 5, 500.0
    pos, times = random_batch(oracle, B, K)
    rng = np.random.default_rng(7)
    cand = times[:, None, :] * rng.uniform(0.6, 1.6, (B, S, K))
    obj, cost = ms.time_objective(torch.from_numpy(pos).cuda(), torch.from_numpy(cand).cuda(), penalty, want_cost=True)
    obj, cost = obj.cpu().numpy(), cost.cpu().numpy()
    for b in range(B):
        for s in range(S):
            want, want_cost = oracle_objective(oracle, pos[b], cand[b, s], penalty)
            assert abs(cost[b, s] - want_cost) <= COST_TOL * want_cost
            assert abs(obj[b, s] - want) <= COST_TOL * want
    # without the separate cost buffer the objective is completed in place
    obj2 = ms.time_objective(torch.from_numpy(pos).cuda(), torch.from_numpy(cand).cuda(), penalty).cpu().numpy()
    np.testing.assert_array_equal(obj2, obj)


@pytest.mark.parametrize("K,increment", [(4, 0.1), (10, 0.1), (10, 0.01)])
def test_time_gradient_matches_oracle(ms, oracle, torch_cuda, K, increment):
    torch = torch_cuda
    B = 6
    pos, times = random_batch(oracle, B, K)
    p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    coeffs = ms.solve_standard(p, t, want_status=False)["coeffs"]
    grad, seg = ms.time_gradient(coeffs, t, increment=increment, w_d=0.1, w_t=1.0, want_segment_cost=True)
    grad, seg = grad.cpu().numpy(), seg.cpu().numpy()
    for b in range(B):
        want, j_d = oracle_gradient(oracle, pos[b], times[b], increment, 0.1, 1.0)
        assert np.max(np.abs(grad[b] - want)) <= 1e-6 * np.max(np.abs(want)), (b, grad[b], want)
        # J_d = d^T R d = 2 computeCost (note C)
        assert abs(seg[b].sum() - j_d) <= 1e-8 * j_d
        cost = float(oracle.solve(10, K, 3, 4, standard_mask(K), vertex_values(pos[b]), times[b])["cost"])
        assert abs(seg[b].sum() - 2.0 * cost) <= 1e-8 * cost


def test_time_gradient_clamps_short_segments(ms, oracle, torch_cuda):
    """ref NL.i:2186-2187, 2207-2208: a segment time <= 0.1 is set to 0.1 on both sides."""
    torch = torch_cuda
    # a small-scale problem (5 cm box, segment times of 0.06-0.3 s) so that short segments do not make
    # the QP ill conditioned next to long ones
    # (not createRandomVertices: it redraws any vertex closer than 0.2 m to its predecessor)
    rng = np.random.default_rng(5)
    pos = rng.uniform(-0.05, 0.05, (4, 7, 3))
    times = rng.uniform(0.06, 0.3, (4, 6))
    times[:, 1] = 0.08
    times[:, 4] = 0.1
    p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    coeffs = ms.solve_standard(p, t, want_status=False)["coeffs"]
    grad = ms.time_gradient(coeffs, t, increment=0.01, w_d=0.1, w_t=1.0).cpu().numpy()
    short = times <= 0.1
    assert short[:, 1].all() and short[:, 4].all()
    np.testing.assert_array_equal(grad[short], 1.0)        # both sides equal: only w_t is left
    for b in range(4):
        want, _ = oracle_gradient(oracle, pos[b], times[b], 0.01, 0.1, 1.0)
        assert np.max(np.abs(grad[b] - want)) <= 1e-6 * np.max(np.abs(want)), (grad[b], want)


def test_gradient_agrees_with_resolved_finite_difference(ms, oracle, torch_cuda):
    """Envelope theorem: at the optimum of the QP the partial derivative with d held fixed equals
    the total derivative of the re-solved cost (the two central differences differ at O(h^2);
    measured 6.5e-5 at h = 1e-4, 6.5e-3 at h = 1e-3); checked with the cost sweep."""
    torch = torch_cuda
    B, K, h = 64, 10, 1e-4
    pos, times = random_batch(oracle, B, K)
    p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    coeffs = ms.solve_standard(p, t, want_status=False)["coeffs"]
    grad = ms.time_gradient(coeffs, t, increment=h, w_d=0.5, w_t=0.0)          # d computeCost / dT
    eye = torch.eye(K, dtype=torch.float64, device="cuda")
    sweep = torch.cat([t[:, None, :] + h * eye[None], t[:, None, :] - h * eye[None]], 1).contiguous()
    c = ms.cost_sweep(p, sweep)
    fd = (c[:, :K] - c[:, K:]) / (2 * h)
    rel = ((grad - fd).abs().max(1).values / fd.abs().max(1).values).max()
    assert float(rel) < 2e-4


def test_batched_time_descent(ms, oracle, torch_cuda):
    torch = torch_cuda
    B, K, penalty = 256, 10, 0.05
    pos, times = random_batch(oracle, B, K)
    p, t0 = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    t1, hist = ms.optimize_segment_times(p, t0, iterations=15, time_penalty=penalty)
    hist = hist.cpu().numpy()
    assert np.all(np.diff(hist, axis=0) <= 0)                      # never accepts a worse allocation
    assert np.mean(hist[-1] / hist[0]) < 0.9                        # and the batch improves on average
    assert float(t1.min()) >= 0.1
    # the reported objective is the reference objective at the returned times
    t1h = t1.cpu().numpy()
    for b in (0, 17, 255):
        want, _ = oracle_objective(oracle, pos[b], t1h[b], penalty)
        assert abs(hist[-1, b] - want) <= COST_TOL * want


def test_device_resident_descent_matches_the_glue_version(ms, oracle, torch_cuda):
    """minsnap_optimize_segment_times (the whole loop enqueued by one C-ABI call, glue in two kernels) against the
    same descent with its glue in elementwise torch operations: same accepted steps, same objective history; with
    non-zero end derivatives and a graph capture of the whole optimisation."""
    torch = torch_cuda
    B, K, penalty = 200, 10, 0.05
    pos, times = random_batch(oracle, B, K)
    p, t0 = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    end = torch.from_numpy(np.random.default_rng(3).uniform(-0.5, 0.5, (B, 2, 4, 3))).cuda()
    for e in (None, end):
        # One iteration: identical up to rounding (the torch glue sums the segment times pairwise, the kernel left to
        # right like the reference).
        t_dev, h_dev = ms.optimize_segment_times(p, t0, iterations=1, time_penalty=penalty, end_derivatives=e)
        t_ref, h_ref = ms.api.optimize_segment_times_reference_glue(p, t0, iterations=1, time_penalty=penalty, end_derivatives=e)
        assert h_dev.shape == h_ref.shape == (2, B)
        assert float(((h_dev - h_ref).abs() / h_ref.abs()).max()) <= 1e-12
        assert float((t_dev - t_ref).abs().max()) <= 1e-14 * float(t_ref.abs().max())
        # Later iterations: the reference's central differences (h = 1e-3 on a quadratic form of size 1e3) turn a
        # last-bit difference of the times into a 1e-7 relative difference of the gradient, so the two descents agree
        # to that level only (measured 7e-8 after the second iteration); both never accept a worse allocation.
        t_dev, h_dev = ms.optimize_segment_times(p, t0, iterations=10, time_penalty=penalty, end_derivatives=e)
        t_ref, h_ref = ms.api.optimize_segment_times_reference_glue(p, t0, iterations=10, time_penalty=penalty, end_derivatives=e)
        assert float(((h_dev - h_ref).abs() / h_ref.abs()).max()) <= 1e-4
        assert bool((h_dev[1:] <= h_dev[:-1]).all())
    # the loop holds no host-side decision: it can be captured into a CUDA graph and replayed
    t_buf = t0.clone()
    hist = torch.empty((5, B), dtype=torch.float64, device="cuda")
    lib = ms.capi.load()

    def enqueue():
        ms.capi.check(lib.minsnap_optimize_segment_times(B, K, 3, 10, 4, ms.api._dptr(p), None, ms.api._dptr(t_buf), 4, penalty,
                                                         16, 0.5, 0.1, 1e-3, ms.api._dptr(hist), ms.api._stream()),
                      "minsnap_optimize_segment_times")
    enqueue()
    torch.cuda.synchronize()
    first = hist.clone()
    t_buf.copy_(t0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        enqueue()
    t_buf.copy_(t0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(hist, first)


def test_time_objective_with_soft_constraints(ms, oracle, torch_cuda):
    """ref objectiveFunctionTime with use_soft_constraints (NL.i:765-832, 2346-2426): the soft term is
    min(max_cost, exp(relative violation * weight)) per limited derivative, with the maximum of
    computeMaximumOfMagnitude."""
    torch = torch_cuda
    B, S, K, penalty, weight = 6, 4, 10, 500.0, 100.0
    pos, times = random_batch(oracle, B, K)
    rng = np.random.default_rng(9)
    cand = times[:, None, :] * rng.uniform(0.7, 1.5, (B, S, K))
    limits = [(1, 4.0), (2, 2.0)]
    obj, parts = ms.time_objective_with_soft_constraints(torch.from_numpy(pos).cuda(), torch.from_numpy(cand).cuda(),
                                                         penalty, limits, soft_constraint_weight=weight,
                                                         want_terms=True)
    obj = obj.cpu().numpy()
    for b in range(B):
        for s in range(S):
            r = oracle.solve(10, K, 3, 4, standard_mask(K), vertex_values(pos[b]), cand[b, s])
            base, _ = oracle_objective(oracle, pos[b], cand[b, s], penalty)
            want = base
            for j, (k, limit) in enumerate(limits):
                peak = float(oracle.minmax_magnitude(np.asarray(r["coeffs"], np.float64), cand[b, s], k, 0)["max"][1])
                term = min(1.0e12, np.exp((peak - limit) / limit * weight))
                got_term = float(parts["cost_constraints"][j][b, s])
                assert abs(got_term - term) <= 1e-6 * term, (b, s, k, got_term, term)
                want += term
            assert abs(obj[b, s] - want) <= 1e-6 * want
    # saturation at maximum_cost
    sat = ms.time_objective_with_soft_constraints(torch.from_numpy(pos).cuda(), torch.from_numpy(cand).cuda(),
                                                  penalty, [(1, 0.01)], soft_constraint_weight=weight, maximum_cost=7.0,
                                                  want_terms=True)[1]["cost_constraints"][0]
    assert float(sat.max()) == 7.0 and float(sat.min()) == 7.0
