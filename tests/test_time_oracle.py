"""CPU tests of the oracle-side time objective / numeric time gradient (tests/helpers.py), the
checker of minsnap_time_objective / minsnap_time_gradient.

The reference evaluates the gradient with the end-point derivatives of the last solve held fixed
(getCostAndGradientTime calls updateSegmentTimes + getR, not solveLinear; NL.i:2155-2243).  At the
optimum of the QP that partial derivative equals the total derivative of the re-solved cost
(envelope theorem), which pins the helper against an independent computation."""
import numpy as np

from helpers import oracle_gradient, oracle_objective, random_batch, standard_mask, vertex_values


def test_cost_conventions(oracle):
    # SURVEY 8 note C: getCostAndGradientDerivative returns d^T R d = 2 computeCost
    pos, times = random_batch(oracle, 3, 6)
    for b in range(3):
        _, j_d = oracle_gradient(oracle, pos[b], times[b], 0.1, 0.1, 1.0)
        cost = float(oracle.solve(10, 6, 3, 4, standard_mask(6), vertex_values(pos[b]), times[b])["cost"])
        assert abs(j_d - 2.0 * cost) <= 1e-8 * cost
        obj, c = oracle_objective(oracle, pos[b], times[b], 500.0)
        assert abs(c - cost) <= 1e-12 * cost
        assert abs(obj - (cost + 500.0 * times[b].sum() ** 2)) <= 1e-9 * obj


def test_fixed_d_gradient_is_the_total_derivative(oracle):
    pos, times = random_batch(oracle, 3, 6)
    h = 1e-4
    for b in range(3):
        grad, _ = oracle_gradient(oracle, pos[b], times[b], h, 0.5, 0.0)     # d computeCost / dT, d fixed
        fd = np.zeros(6)
        for n in range(6):
            up, dn = times[b].copy(), times[b].copy()
            up[n] += h
            dn[n] -= h
            fd[n] = (oracle_objective(oracle, pos[b], up, 0.0)[1] - oracle_objective(oracle, pos[b], dn, 0.0)[1]) / (2 * h)
        assert np.max(np.abs(grad - fd)) <= 1e-3 * np.max(np.abs(fd)), (grad, fd)


def test_clamp_rule(oracle):
    # ref NL.i:2186-2187, 2207-2208: a segment time <= 0.1 is set to 0.1 on both sides -> only w_t is left
    rng = np.random.default_rng(5)
    pos = rng.uniform(-0.05, 0.05, (5, 3))
    times = np.array([0.2, 0.08, 0.15, 0.1])
    grad, _ = oracle_gradient(oracle, pos, times, 0.01, 0.1, 1.0)
    assert grad[1] == 1.0 and grad[3] == 1.0
    assert grad[0] != 1.0 and grad[2] != 1.0
