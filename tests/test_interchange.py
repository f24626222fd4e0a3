"""Interchange formats: .npy through the C ABI against numpy itself (CPU), and the
reference's sampled-trajectory table (printMatlabSampledTrajectory, NL.i:2567-2662) against a literal
restatement (GPU)."""
import numpy as np
import pytest

import mav_trajectory_generation_cmake_b200 as ms


@pytest.mark.parametrize("shape", [(3, 10, 3, 10), (7,), (2, 0, 3), (1, 1), (4, 1000, 5, 3)])
def test_npy_round_trips_with_numpy(tmp_path, shape):
    rng = np.random.default_rng(len(shape))
    a = rng.normal(size=shape) * 10.0 ** rng.integers(-300, 300)
    ms.save_npy(tmp_path / "ours.npy", a)
    b = np.load(tmp_path / "ours.npy")                 # numpy reads what the C ABI wrote
    assert b.dtype == np.float64 and b.shape == a.shape and np.array_equal(a, b) and not np.isfortran(b)
    np.save(tmp_path / "theirs.npy", a)
    c = ms.load_npy(tmp_path / "theirs.npy")           # the C ABI reads what numpy wrote
    assert c.shape == a.shape and np.array_equal(a, c)
    assert (tmp_path / "ours.npy").read_bytes()[:6] == b"\x93NUMPY"
    assert (tmp_path / "ours.npy").stat().st_size % 64 == (a.size * 8) % 64   # header padded to 64 bytes


def test_npy_rejects_other_dtypes(tmp_path):
    np.save(tmp_path / "f32.npy", np.zeros(4, np.float32))
    with pytest.raises(ms.MinsnapError) as err:
        ms.load_npy(tmp_path / "f32.npy")
    assert err.value.code == ms.capi.ERR_UNSUPPORTED
    np.save(tmp_path / "fortran.npy", np.asfortranarray(np.zeros((3, 4))))
    with pytest.raises(ms.MinsnapError):
        ms.load_npy(tmp_path / "fortran.npy")
    with pytest.raises(ms.MinsnapError):
        ms.load_npy(tmp_path / "missing.npy")


def reference_table(coeffs, times, dt):
    """ref printMatlabSampledTrajectory (NL.i:2576-2655), restated literally: pow-based evaluation."""
    K, D, N = coeffs.shape
    rows = sum(int(np.ceil(T / dt)) + 1 for T in times)
    out = np.zeros((rows, 5 * D + 2))
    j, current = 0, 0.0
    for i in range(K):
        t = 0.0
        while t < times[i]:
            if j < rows:
                out[j, 0] = t + current
                for k in range(D):
                    c = coeffs[i, k]
                    for d in range(5):
                        acc = 0.0
                        for n in range(d, N):
                            fall = 1.0
                            for q in range(d):
                                fall *= (n - q)
                            acc += fall * c[n] * t ** (n - d)
                        out[j, 1 + d * D + k] = acc
                j += 1
            t += dt
        current += times[i]
        out[i, 1 + 5 * D] = current
    return out


@pytest.mark.gpu
def test_sampled_table_against_reference_layout(tmp_path):
    import torch
    K = 6
    pos = ms.random_positions_host(1, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 4242)
    pos_d = torch.from_numpy(pos).cuda()
    times_d = ms.estimate_segment_times(pos_d, 3.0, 5.0)
    coeffs = ms.solve_standard(pos_d, times_d)["coeffs"][0].cpu().numpy()
    times = times_d[0].cpu().numpy()
    for dt in (0.01, 0.37):
        table = ms.sampled_table_host(coeffs, times, dt=dt, path=tmp_path / "table.txt")
        want = reference_table(coeffs, times, dt)
        assert table.shape == want.shape
        assert np.array_equal(table[:, 0], want[:, 0])                  # row times: bit-exact
        assert np.array_equal(table[:, -1], want[:, -1])                # vertex times
        scale = np.abs(want[:, 1:-1]).max(axis=0)
        assert (np.abs(table[:, 1:-1] - want[:, 1:-1]).max(axis=0) <= 1e-11 * np.maximum(scale, 1.0)).all()
        back = np.loadtxt(tmp_path / "table.txt", ndmin=2)
        assert np.array_equal(back, table)                              # 17 digits: the text round-trips exactly
