"""Every *_host entry point of the C ABI against its device-pointer twin (GPU): the staged small path (one
pinned block, one copy each way), the large path (stream-ordered allocations) and the chunked pipeline of
minsnap_solve_standard_host give bit-identical results to the device calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOX_LO, BOX_HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]


@pytest.fixture(scope="module")
def env():
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    return ms, torch


@pytest.mark.parametrize("B", [3, 6000])      # 3: staged through the arena; 6000: > 4 MB, the large paths
def test_host_twins_bit_identical(env, B):
    ms, torch = env
    K = 10
    pos = ms.random_positions_host(B, K, BOX_LO, BOX_HI, 555)
    pos_d = torch.from_numpy(pos).cuda()
    times_d = ms.estimate_segment_times(pos_d, 3.0, 5.0)
    times = ms.estimate_segment_times_host(pos, 3.0, 5.0)
    assert np.array_equal(times, times_d.cpu().numpy())
    dev = ms.solve_standard(pos_d, times_d, want_cost=True, want_free=True)
    host = ms.solve_standard_host(pos, times, want_cost=True, want_free=True, want_status=True)
    assert np.array_equal(host["coeffs"], dev["coeffs"].cpu().numpy())
    assert np.array_equal(host["cost"], dev["cost"].cpu().numpy())
    assert np.array_equal(host["free_values"], dev["free_values"].cpu().numpy())
    assert int((host["status"] != 0).sum()) == 0
    coeffs = host["coeffs"]
    assert np.array_equal(ms.cost_host(coeffs, times), ms.cost(dev["coeffs"], times_d).cpu().numpy())
    n = min(B, 500)
    # sampling
    s_host, t_host = ms.sample_uniform_host(coeffs[:n], times[:n], 40, 5, want_times=True)
    s_dev, t_dev = ms.sample_uniform(dev["coeffs"][:n], times_d[:n], 40, 5, want_times=True)
    assert np.array_equal(s_host, s_dev.cpu().numpy()) and np.array_equal(t_host, t_dev.cpu().numpy())
    at_host, seg_host = ms.sample_at_host(coeffs[:n], times[:n], t_host, 3)
    at_dev, seg_dev = ms.sample_at(dev["coeffs"][:n], times_d[:n], t_dev, 3, want_segment=True)
    assert np.array_equal(at_host, at_dev.cpu().numpy()) and np.array_equal(seg_host, seg_dev.cpu().numpy())
    # configs[4]: the time sweep, the time objective and the time gradient from host buffers
    S = 8
    rng = np.random.default_rng(B)
    sweep_t = times[:n, None, :] * rng.uniform(0.8, 1.25, size=(n, S, K))
    sweep_d = torch.from_numpy(sweep_t).cuda()
    assert np.array_equal(ms.cost_sweep_host(pos[:n], sweep_t), ms.cost_sweep(pos_d[:n], sweep_d).cpu().numpy())
    obj_h, cost_h = ms.time_objective_host(pos[:n], sweep_t, 500.0, want_cost=True)
    obj_d, cost_dv = ms.time_objective(pos_d[:n], sweep_d, 500.0, want_cost=True)
    assert np.array_equal(obj_h, obj_d.cpu().numpy()) and np.array_equal(cost_h, cost_dv.cpu().numpy())
    g_h, sc_h = ms.time_gradient_host(coeffs[:n], times[:n], want_segment_cost=True)
    g_d, sc_d = ms.time_gradient(dev["coeffs"][:n], times_d[:n], want_segment_cost=True)
    assert np.array_equal(g_h, g_d.cpu().numpy()) and np.array_equal(sc_h, sc_d.cpu().numpy())
    # extrema: in optimisation mode the minimum outputs are not produced (and must not be written)
    e_h = ms.extrema_host(coeffs[:n], times[:n], 1)
    e_d = ms.extrema(dev["coeffs"][:n], times_d[:n], 1)
    assert np.array_equal(e_h["max_value"], e_d["max_value"].cpu().numpy())
    # collision cost
    X, Y, Z = np.meshgrid(*[-12.0 + (np.arange(m) + 0.5) * 0.5 for m in (48, 88, 48)], indexing="ij")
    sdf = np.sqrt(X ** 2 + (Y - 3.0) ** 2 + Z ** 2) - 4.0
    kw = dict(origin=[-12.0, -22.0, -12.0], resolution=0.5, min_bound=BOX_LO, max_bound=BOX_HI)
    c_h = ms.collision_cost_host(coeffs[:n], times[:n], sdf, **kw)
    c_d = ms.collision_cost(dev["coeffs"][:n], times_d[:n], torch.from_numpy(sdf).cuda(), want_charged=True, **kw)
    assert np.array_equal(c_h["cost"], c_d["cost"].cpu().numpy())
    assert np.array_equal(c_h["is_collision"], c_d["is_collision"].cpu().numpy())
    assert np.array_equal(c_h["charged"], c_d["charged"].cpu().numpy())


def test_extrema_host_leaves_minimum_untouched_in_optimisation_mode(env):
    ms, torch = env
    import ctypes as C
    K, B = 4, 2
    pos = ms.random_positions_host(B, K, BOX_LO, BOX_HI, 9)
    times = ms.estimate_segment_times_host(pos, 3.0, 5.0)
    coeffs = ms.solve_standard_host(pos, times)["coeffs"]
    lib = ms.capi.load()
    sentinel = -123.25
    mx_t, mx_v, mn_t, mn_v = (np.full(B, sentinel) for _ in range(4))
    mx_s, mn_s = np.full(B, -7, np.int32), np.full(B, -7, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ms.capi.check(lib.minsnap_extrema_host(B, K, 3, 10, p(coeffs), p(times), 1, ms.EXTREMA_OPTIMIZATION, 0, p(mx_t), p(mx_v),
                                           p(mx_s), p(mn_t), p(mn_v), p(mn_s), None, None, None), "minsnap_extrema_host")
    assert (mx_v > 0).all() and (mn_t == sentinel).all() and (mn_v == sentinel).all() and (mn_s == -7).all()
    ms.capi.check(lib.minsnap_extrema_host(B, K, 3, 10, p(coeffs), p(times), 1, ms.EXTREMA_TRAJECTORY, 0, p(mx_t), p(mx_v),
                                           p(mx_s), p(mn_t), p(mn_v), p(mn_s), None, None, None), "minsnap_extrema_host")
    assert (mn_v != sentinel).all() and (mn_s >= 0).all()
