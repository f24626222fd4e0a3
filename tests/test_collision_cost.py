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


# ---- gradient w.r.t. the free derivatives (ref NL.i:1666-1686) -----------------------------------------------

def solved(orc, K, seed, mask=None, shrink=1.0):
    """One trajectory with its constraint bookkeeping: coeffs, times, index map, n_fixed, n_free, d_all."""
    N = 10
    pos = orc.create_random_positions(K, [shrink * v for v in BOX_LO], [shrink * v for v in BOX_HI], seed)
    times = orc.estimate_segment_times(pos, 3.0, 5.0, 6.5)
    vv = vertex_values_from_positions(pos)
    if mask is None:
        mask = standard_mask(K)
    else:   # fixed derivatives of a non-standard mask get non-zero values
        rng = np.random.default_rng(seed)
        vv = vv + (np.asarray(mask).reshape(K + 1, N // 2, 1) * rng.normal(size=vv.shape) * 0.3) * (np.arange(N // 2) > 0)[None, :, None]
    sol = orc.solve(N, K, 3, 4, mask, vv, times)
    col, n_fixed, n_free = orc.reorder(N, K, mask)
    return np.asarray(sol["coeffs"], np.float64), np.asarray(times, np.float64), col, n_fixed, n_free, sol


def test_oracle_gradient_is_the_frozen_walk_derivative():
    """The reference's gradient differentiates c(pos)|v| time_sum of every charged sample with the walk (which
    samples are charged, and their time_sum) held fixed.  On a potential that is LINEAR where the trajectory goes
    (distance field linear in x, inside the collision branch: c = m (r - d) + eps/2) the central difference of the
    potential is exact, so the oracle's gradient must equal a finite difference of that frozen sum in d_p."""
    orc = Oracle("f64")
    K, N = 4, 10
    X, Y, Z = centres()
    g = -2.0 + 0.02 * X + 0.0 * Y     # always in collision: potential = 1.5 (0.5 - d) + 0.25, linear.  (Linear in x only: the
    # reference's triLerp blends (y0,z0) with (y0,z1) by the Y weight, so it does not reproduce a field that varies in y or z.)
    coeffs, times, col, n_fixed, n_free, sol = solved(orc, K, 5, shrink=0.5)   # stays where the blend is used
    kw = dict(sdf=g, origin=ORIGIN, resolution=RES, min_bound=[-10.0, -20.0, -10.0], max_bound=[10.0, 20.0, 10.0], dt=0.05,
              map_resolution=RES, epsilon=0.5, robot_radius=0.5, coll_pot_multiplier=1.5, use_continuous_distance=True)
    cost, grad, hit, charged = oracle_py.collision_cost_gradient(coeffs, times, col, n_fixed, n_free, **kw)
    cost0, hit0, charged0 = oracle_py.collision_cost(coeffs, times, g, **{k: v for k, v in kw.items() if k != "sdf"})
    assert cost == cost0 and charged == charged0 and hit == hit0 == 1 and charged > 20
    assert np.abs(grad).max() > 0

    # frozen walk: the charged samples (segment, t, time_sum) from a python replay of the loop
    def walk(c):
        out, time_sum, dist_sum, prev = [], -1.0, 0.0, None
        for i in range(K):
            t = 0.0
            while t < times[i]:
                p = np.array([np.polyval(c[i, k, ::-1], t) for k in range(3)])
                if time_sum < 0:
                    time_sum, prev = 0.0, p
                    t += 0.05
                    continue
                time_sum += 0.05
                dist_sum += np.linalg.norm(p - prev)
                prev = p
                if dist_sum >= RES:
                    out.append((i, t, time_sum))
                    dist_sum = time_sum = 0.0
                t += 0.05
            time_sum += -0.05 + (times[i] - t)
        return out
    samples = walk(coeffs)
    assert len(samples) == charged

    def frozen(c):
        J = 0.0
        for i, t, ts in samples:
            p = np.array([np.polyval(c[i, k, ::-1], t) for k in range(3)])
            v = np.array([np.polyval(np.polyder(c[i, k, ::-1]), t) for k in range(3)])
            d = -2.0 + 0.02 * p[0]   # the blend reproduces a field linear in x
            J += (1.5 * (0.5 - d) + 0.25) * np.linalg.norm(v) * ts
        return J
    assert abs(frozen(coeffs) - cost) <= 1e-9 * cost
    d_all = np.concatenate([np.asarray(sol["d_fixed"], np.float64), np.asarray(sol["d_free"], np.float64)], axis=1)   # [3][n_all]
    assert np.abs(np.asarray(orc.coeffs_from_constraints(N, K, 3, col, d_all, times), np.float64) - coeffs).max() <= 1e-9
    h = 1e-5
    for colj in (0, 3, n_free // 2, n_free - 1):
        for k in range(3):
            dp, dm = d_all.copy(), d_all.copy()
            dp[k, n_fixed + colj] += h
            dm[k, n_fixed + colj] -= h
            cp = np.asarray(orc.coeffs_from_constraints(N, K, 3, col, dp, times), np.float64)
            cm = np.asarray(orc.coeffs_from_constraints(N, K, 3, col, dm, times), np.float64)
            fd = (frozen(cp) - frozen(cm)) / (2 * h)
            assert abs(fd - grad[colj, k]) <= 1e-6 * max(np.abs(grad).max(), 1e-3), (colj, k, fd, grad[colj, k])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere", "box", "smooth"])
@pytest.mark.parametrize("continuous", [True, False])
def test_collision_gradient_against_oracle(name, continuous):
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    orc = Oracle("f64")
    B, K, N = 24, 10, 10
    coeffs, times = trajectories(orc, B)
    col, n_fixed, n_free = orc.reorder(N, K, standard_mask(K))
    g = field(name)
    kw = dict(origin=ORIGIN, resolution=RES, min_bound=[-10.0, -20.0, -10.0], max_bound=[10.0, 20.0, 10.0], dt=0.1,
              map_resolution=RES, epsilon=0.5, robot_radius=0.5, coll_pot_multiplier=1.5,
              use_continuous_distance=continuous, oob_value=0.0)
    cd, td, gd = torch.from_numpy(coeffs).cuda(), torch.from_numpy(times).cuda(), torch.from_numpy(g).cuda()
    out = ms.collision_gradient(cd, td, gd, **kw)                      # standard mask in closed form
    out_map = ms.collision_gradient(cd, td, gd, col_of_row=torch.from_numpy(col).cuda(), n_fixed=n_fixed, n_free=n_free, **kw)
    plain = ms.collision_cost(cd, td, gd, want_charged=True, **kw)
    assert torch.equal(out["cost"], plain["cost"]) and torch.equal(out["charged"], plain["charged"])
    assert torch.equal(out["gradient"], out_map["gradient"])
    grad = out["gradient"].cpu().numpy()
    assert grad.shape == (B, n_free, 3)
    n_nonzero = 0
    for b in range(B):
        want_cost, want, want_hit, want_charged = oracle_py.collision_cost_gradient(coeffs[b], times[b], col, n_fixed, n_free,
                                                                                    sdf=g, **kw)
        assert out["charged"][b].item() == want_charged and out["is_collision"][b].item() == want_hit
        assert abs(out["cost"][b].item() - want_cost) <= 1e-8 * max(abs(want_cost), 1e-3)
        scale = max(np.abs(want).max(), 1e-6)
        assert np.abs(grad[b] - want).max() <= 1e-8 * scale, (b, np.abs(grad[b] - want).max(), scale)
        n_nonzero += np.abs(want).max() > 0
    assert n_nonzero > 0


@pytest.mark.gpu
def test_collision_gradient_general_mask_and_host_entry():
    """A non-standard mask (some interior derivatives fixed, an end derivative free) through the index map, short and
    long increments, and the host-buffer entry point."""
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    orc = Oracle("f64")
    K, N = 5, 10
    mask = np.array(standard_mask(K)).reshape(K + 1, N // 2).copy()
    mask[2, 1] = 1      # velocity fixed at vertex 2
    mask[3, 2] = 1      # acceleration fixed at vertex 3
    mask[K, 4] = 0      # snap free at the end
    mask[0, 3] = 0      # jerk free at the start
    g = field("sphere")
    kw = dict(origin=ORIGIN, resolution=RES, min_bound=[-10.0, -20.0, -10.0], max_bound=[10.0, 20.0, 10.0], epsilon=0.5,
              robot_radius=0.5, coll_pot_multiplier=1.5, use_continuous_distance=True)
    rows = [solved(orc, K, 300 + b, mask=mask) for b in range(6)]
    coeffs = np.stack([r[0] for r in rows])
    times = np.stack([r[1] for r in rows])
    col, n_fixed, n_free = rows[0][2], rows[0][3], rows[0][4]
    for dt, mr in ((0.1, 0.5), (0.7, 0.5), (0.02, 1.5)):
        out = ms.collision_gradient(torch.from_numpy(coeffs).cuda(), torch.from_numpy(times).cuda(), torch.from_numpy(g).cuda(),
                                    col_of_row=torch.from_numpy(col).cuda(), n_fixed=n_fixed, n_free=n_free, dt=dt,
                                    map_resolution=mr, **kw)
        host = ms.collision_gradient_host(coeffs, times, g, col_of_row=col, n_fixed=n_fixed, n_free=n_free, dt=dt,
                                          map_resolution=mr, **kw)
        assert np.array_equal(out["gradient"].cpu().numpy(), host["gradient"])
        assert np.array_equal(out["cost"].cpu().numpy(), host["cost"])
        for b in range(len(rows)):
            want_cost, want, want_hit, want_charged = oracle_py.collision_cost_gradient(coeffs[b], times[b], col, n_fixed,
                                                                                        n_free, sdf=g, dt=dt,
                                                                                        map_resolution=mr, **kw)
            assert host["charged"][b] == want_charged
            scale = max(np.abs(want).max(), 1e-6)
            assert np.abs(host["gradient"][b] - want).max() <= 1e-8 * scale
    empty = ms.collision_gradient(torch.zeros((0, K, 3, 10), dtype=torch.float64, device="cuda"),
                                  torch.zeros((0, K), dtype=torch.float64, device="cuda"), torch.from_numpy(g).cuda(), **kw)
    assert empty["gradient"].shape == (0, (K - 1) * 4, 3)
