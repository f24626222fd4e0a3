"""Collision cost against a signed-distance grid (ref getCostAndGradientCollision,
NL.i:1523-1709).  CPU: the oracle against closed forms.  GPU: the kernel against the oracle on three
synthetic fields (sphere, box, smooth random), both distance modes, to 1e-8 relative; collision flags and
the number of charged samples exactly."""
import numpy as np
import pytest

from oracle import oracle_py
from oracle.oracle_py import Oracle, standard_mask, vertex_values_from_positions

BOX_LO, BOX_HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]
ORIGIN, RES = np.array([-12.0, -22.0, -12.0]), 0.5
DIMS = (48, 88, 48)          # covers the box +-(10, 20, 10) with a margin of 2


def centres():
    ax = [ORIGIN[k] + (np.arange(DIMS[k]) + 0.5) * RES for k in range(3)]
    return np.meshgrid(*ax, indexing="ij")


def field(name):
    X, Y, Z = centres()
    if name == "sphere":      # distance to a sphere of radius 3 at (1, -2, 0.5), negative inside
        return np.sqrt((X - 1.0) ** 2 + (Y + 2.0) ** 2 + (Z - 0.5) ** 2) - 3.0
    if name == "box":         # signed distance to the box |x| <= 2, |y - 5| <= 4, |z| <= 3
        q = np.stack([np.abs(X) - 2.0, np.abs(Y - 5.0) - 4.0, np.abs(Z) - 3.0])
        return np.linalg.norm(np.maximum(q, 0.0), axis=0) + np.minimum(q.max(axis=0), 0.0)
    rng = np.random.default_rng(5)   # smooth random: a few low-frequency waves around 0.6
    f = 0.6 + 0.0 * X
    for _ in range(6):
        k = rng.normal(size=3) * 0.35
        f = f + 0.35 * np.sin(k[0] * X + k[1] * Y + k[2] * Z + rng.uniform(0, 6.28))
    return f


def trajectories(orc, B, K=10, seed=31):
    pos = np.stack([orc.create_random_positions(K, BOX_LO, BOX_HI, seed + b) for b in range(B)])
    times = np.stack([orc.estimate_segment_times(p, 3.0, 5.0, 6.5) for p in pos])
    coeffs = np.stack([orc.solve(10, K, 3, 4, standard_mask(K), vertex_values_from_positions(pos[b]), times[b])["coeffs"]
                       for b in range(B)]).astype(np.float64)
    return coeffs, times


def test_oracle_potential_closed_forms():
    """getCostPotential (NL.i:2319-2345) and the two distance modes on a field that is linear in x: the
    trilinear blend of a linear field reproduces it, the discrete lookup returns the cell-centre value."""
    X, _, _ = centres()
    g = 0.25 * X + 1.0
    kw = dict(sdf=g, origin=ORIGIN, resolution=RES, min_bound=[-10, -20, -10], max_bound=[10, 20, 10], epsilon=0.5,
              robot_radius=0.5, coll_pot_multiplier=2.0)
    for x in (-7.3, -1.9, 0.1, 3.33):
        p = [x, 1.1, -2.2]
        d_true = 0.25 * x + 1.0
        want = (2.0 * -(d_true - 0.5) + 0.25) if d_true - 0.5 <= 0 else (0.5 / 0.5 * (d_true - 0.5 - 0.5) ** 2 if d_true - 0.5 <= 0.5 else 0.0)
        cost, grad, hit = oracle_py.collision_potential(p, use_continuous_distance=True, **kw)
        assert abs(cost - want) <= 1e-12 and hit == (d_true - 0.5 <= 0)
        centre = ORIGIN[0] + (np.floor((x - ORIGIN[0]) / RES) + 0.5) * RES
        d_cell = 0.25 * centre + 1.0
        want_cell = (2.0 * -(d_cell - 0.5) + 0.25) if d_cell - 0.5 <= 0 else (1.0 * (d_cell - 1.0) ** 2 if d_cell - 0.5 <= 0.5 else 0.0)
        cost_d, _, _ = oracle_py.collision_potential(p, use_continuous_distance=False, **kw)
        assert abs(cost_d - want_cell) <= 1e-12
    # outside [min_bound + res, max_bound - res] the discrete lookup is used even in continuous mode; outside the grid: oob
    cost, _, hit = oracle_py.collision_potential([9.8, 0.0, 0.0], use_continuous_distance=True, **kw)
    centre = ORIGIN[0] + (np.floor((9.8 - ORIGIN[0]) / RES) + 0.5) * RES
    assert abs(cost - 0.0) <= 1e-12 and not hit and 0.25 * centre + 1.0 - 0.5 > 0.5
    cost, _, hit = oracle_py.collision_potential([100.0, 0.0, 0.0], oob_value=-1.0, **kw)
    assert hit and abs(cost - (2.0 * 1.5 + 0.25)) <= 1e-12


def test_oracle_cost_straight_line_known_answer():
    """A straight line x(t) = v t through a constant potential: every charged sample adds c |v| time_sum, and the
    charged intervals tile the walk, so the cost is c |v| times the walked time up to the last charged sample."""
    g = np.full(DIMS, 0.75)            # d - r = 0.25 <= epsilon: c = (0.25 - 0.5)^2 / (2 * 0.5) = 0.0625
    v = np.array([0.8, -0.3, 0.2])
    K, N = 3, 10
    times = np.array([2.0, 3.3, 1.7])
    coeffs = np.zeros((K, 3, N))
    start = np.array([-3.0, 2.0, 1.0])
    for i in range(K):
        coeffs[i, :, 0] = start + v * times[:i].sum()
        coeffs[i, :, 1] = v
    cost, hit, charged = oracle_py.collision_cost(coeffs, times, g, ORIGIN, RES, [-10, -20, -10], [10, 20, 10], dt=0.1,
                                                  map_resolution=0.25)
    speed = np.linalg.norm(v)
    assert not hit and charged > 10
    # samples are charged every ceil(0.25 / (speed * 0.1)) steps; the sum telescopes to c * speed * (time of last charge)
    per = int(np.ceil(0.25 / (speed * 0.1) - 1e-12))
    assert abs(cost - 0.0625 * speed * charged * per * 0.1) <= 0.0625 * speed * 0.25   # segment joins shift single steps


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere", "box", "smooth"])
@pytest.mark.parametrize("continuous", [True, False])
def test_collision_cost_against_oracle(name, continuous):
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    orc = Oracle("f64")
    B = 48
    coeffs, times = trajectories(orc, B)
    g = field(name)
    kw = dict(origin=ORIGIN, resolution=RES, min_bound=[-10.0, -20.0, -10.0], max_bound=[10.0, 20.0, 10.0], dt=0.1,
              map_resolution=RES, epsilon=0.5, robot_radius=0.5, coll_pot_multiplier=1.5,
              use_continuous_distance=continuous, oob_value=0.0)
    out = ms.collision_cost(torch.from_numpy(coeffs).cuda(), torch.from_numpy(times).cuda(), torch.from_numpy(g).cuda(),
                            want_charged=True, **kw)
    cost, hit, charged = out["cost"].cpu().numpy(), out["is_collision"].cpu().numpy(), out["charged"].cpu().numpy()
    n_hit = 0
    for b in range(B):
        want, want_hit, want_charged = oracle_py.collision_cost(coeffs[b], times[b], g, **kw)
        assert charged[b] == want_charged
        assert hit[b] == want_hit
        assert abs(cost[b] - want) <= 1e-8 * max(abs(want), 1e-3), (b, cost[b], want)
        n_hit += want_hit
    if name != "smooth":
        assert 0 < n_hit < B      # the obstacle is met by some trajectories and missed by others


@pytest.mark.gpu
def test_collision_cost_edge_cases():
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    orc = Oracle("f64")
    coeffs, times = trajectories(orc, 5, K=4, seed=77)
    g = field("sphere")
    kw = dict(origin=ORIGIN, resolution=RES, min_bound=[-10.0, -20.0, -10.0], max_bound=[10.0, 20.0, 10.0])
    # a coarse increment (fewer than 32 samples per segment), a fine one (several chunks per segment), a large threshold
    for dt, mr in ((0.9, 0.5), (0.013, 0.5), (0.1, 7.0)):
        out = ms.collision_cost(torch.from_numpy(coeffs).cuda(), torch.from_numpy(times).cuda(), torch.from_numpy(g).cuda(),
                                dt=dt, map_resolution=mr, want_charged=True, **kw)
        for b in range(5):
            want, want_hit, want_charged = oracle_py.collision_cost(coeffs[b], times[b], g, dt=dt, map_resolution=mr, **kw)
            assert out["charged"][b].item() == want_charged and out["is_collision"][b].item() == want_hit
            assert abs(out["cost"][b].item() - want) <= 1e-8 * max(abs(want), 1e-3)
    # empty batch, wrong dimension
    empty = ms.collision_cost(torch.zeros((0, 4, 3, 10), dtype=torch.float64, device="cuda"),
                              torch.zeros((0, 4), dtype=torch.float64, device="cuda"), torch.from_numpy(g).cuda(), **kw)
    assert empty["cost"].numel() == 0
    with pytest.raises(ms.MinsnapError):
        ms.collision_cost(torch.zeros((2, 4, 2, 10), dtype=torch.float64, device="cuda"),
                          torch.ones((2, 4), dtype=torch.float64, device="cuda"), torch.from_numpy(g).cuda(), **kw)
