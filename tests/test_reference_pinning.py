"""The oracle against the REFERENCE ITSELF (CPU).  oracle/_ref/libmav_ref_core.so is the reference's own
src/vertex.cpp, polynomial.cpp, segment.cpp, trajectory.cpp, motion_defines.cpp and rpoly.cpp compiled from
/root/reference against the Eigen / glog stand-ins of oracle/ref_shim/ (oracle/Makefile target `ref`; the
prebuilt file travels to machines without /root/reference).  Everything whose arithmetic order is the same
is required to agree BIT FOR BIT: the input generator, the segment-time heuristic, the base-coefficient
table, Horner evaluation, the segment search of Trajectory::evaluate, the time accumulation of
evaluateRange, the candidate polynomial of the extrema search.  Extrema values depend on the root finder
(Jenkins-Traub in the reference, interval isolation in the oracle) and agree to 1e-8."""
import numpy as np
import pytest

from oracle.oracle_py import Oracle, ReferenceCore, standard_mask, vertex_values_from_positions

pytestmark = pytest.mark.skipif(not ReferenceCore.available(), reason="neither oracle/_ref nor /root/reference present")

BOX_LO, BOX_HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]


@pytest.fixture(scope="module")
def ref():
    return ReferenceCore()


@pytest.fixture(scope="module")
def orc():
    return Oracle("f64")


@pytest.fixture(scope="module")
def solved(orc):
    """A few solved trajectories (oracle coefficients) to evaluate with both sides."""
    out = []
    for K, seed in ((4, 12345), (10, 978), (10, 12346), (25, 4711)):
        pos = orc.create_random_positions(K, BOX_LO, BOX_HI, seed)
        times = orc.estimate_segment_times(pos, 3.0, 5.0, 6.5)
        sol = orc.solve(10, K, 3, 4, standard_mask(K), vertex_values_from_positions(pos), times)
        out.append((np.asarray(sol["coeffs"], np.float64), np.asarray(times, np.float64)))
    return out


@pytest.mark.parametrize("seed", [12, 123, 1234, 12345, 978, 12345 + 65535])
@pytest.mark.parametrize("K", [1, 4, 10, 100])
def test_create_random_vertices_bit_exact(ref, orc, seed, K):
    """ref createRandomVertices (src/vertex.cpp:27-79): draw order, rejection rule, constraint counts."""
    pos, n_constraints = ref.create_random_positions(4, K, BOX_LO, BOX_HI, seed)
    assert np.array_equal(pos, orc.create_random_positions(K, BOX_LO, BOX_HI, seed))
    want = np.ones(K + 1, np.int32)
    want[0] = want[K] = 5          # makeStartOrEnd(position, 4): derivatives 0..4
    assert np.array_equal(n_constraints, want)
    # 1-D variant of the reference's tests (createRandomVertices1D)
    pos1, _ = ref.create_random_positions(4, K, [-50.0], [50.0], seed)
    assert np.array_equal(pos1, orc.create_random_positions(K, [-50.0], [50.0], seed))


@pytest.mark.parametrize("v_a", [(3.0, 5.0), (2.0, 2.0), (1.0, 10.0)])
def test_estimate_segment_times_bit_exact(ref, orc, v_a):
    """ref estimateSegmentTimes (src/vertex.cpp:162-178)."""
    for seed in (12345, 978, 5):
        pos = orc.create_random_positions(50, BOX_LO, BOX_HI, seed)
        assert np.array_equal(ref.estimate_segment_times(pos, v_a[0], v_a[1], 6.5),
                              orc.estimate_segment_times(pos, v_a[0], v_a[1], 6.5))


def test_base_coefficients_bit_exact(ref, orc):
    """ref computeBaseCoefficients (src/polynomial.cpp:140-155), the 22 x 22 table and smaller ones."""
    for n in (4, 10, 12, 22):
        assert np.array_equal(ref.base_coefficients(n), orc.base_coefficients(n))


def test_base_coeffs_with_time_match_mapping_matrix(ref, orc):
    """ref baseCoeffsWithTime (polynomial.h:215-233) builds the rows of A (LIN.i:101-111)."""
    for N in (6, 10, 12):
        for T in (0.0, 0.7, 3.0, 25.0):
            A = orc.mapping_matrix(N, T)
            for d in range(N // 2):
                assert np.array_equal(A[d], ref.base_coeffs_with_time(N, d, 0.0))
                assert np.array_equal(A[N // 2 + d], ref.base_coeffs_with_time(N, d, T))


def test_polynomial_evaluate_bit_exact(ref, orc):
    """ref Polynomial::evaluate(t, derivative) and evaluate(t, VectorXd*) (polynomial.h:120-151)."""
    rng = np.random.default_rng(3)
    for N in (4, 10, 12):
        for _ in range(40):
            c = rng.normal(size=N) * 10.0 ** rng.integers(-6, 3, size=N)
            t = float(rng.uniform(-2.0, 30.0))
            every = ref.polynomial_evaluate_all(c, t, min(N, 6))
            for d in range(N + 2):
                got = orc.polynomial_evaluate(c, t, d)
                assert got == ref.polynomial_evaluate(c, t, d)
                if d < len(every):
                    assert got == every[d]


def test_candidate_polynomial_bit_exact(ref, orc, solved):
    """ref Segment::computeMinMaxMagnitudeCandidateTimes (src/segment.cpp:93-116): getCoefficients, head, convolve
    and the order in which the dimensions are added."""
    for coeffs, _ in solved:
        for seg in coeffs[:: max(1, len(coeffs) // 4)]:
            for k in (0, 1, 2, 3):
                N = seg.shape[1]
                want = np.zeros((N - k) + (N - k - 1) - 1)
                for d in range(seg.shape[0]):
                    a = ref.polynomial_get_coefficients(seg[d], k)[: N - k]
                    b = ref.polynomial_get_coefficients(seg[d], k + 1)[: N - k - 1]
                    want = want + ref.convolve(a, b)
                got = orc.candidate_polynomial(seg, k)
                assert np.array_equal(got[: len(want)], want)


def test_trajectory_evaluate_bit_exact(ref, orc, solved):
    """ref Segment::evaluate (src/segment.cpp:51-58) and Trajectory::evaluate (src/trajectory.cpp:41-66):
    the strict '>' segment choice (a vertex instant belongs to the segment on its right), local time."""
    rng = np.random.default_rng(11)
    for coeffs, times in solved:
        total = ref.trajectory_max_time(coeffs, times)
        acc = np.cumsum(times)
        instants = list(rng.uniform(0.0, total, size=60)) + list(acc[:-1]) + list(np.nextafter(acc[:-1], 0.0)) + \
            list(np.nextafter(acc[:-1], np.inf)) + [0.0]
        for t in instants:
            if not t < total:
                continue
            for d in range(5):
                want = ref.trajectory_evaluate(coeffs, times, float(t), d)
                got, _ = orc.trajectory_evaluate(coeffs, times, float(t), d)
                assert np.array_equal(got, want), (t, d)
        for k, T in enumerate(times[:3]):
            for t in (0.0, 0.3 * T, T):
                for d in range(5):
                    want = ref.segment_evaluate(coeffs[k], T, t, d)
                    got = np.array([orc.polynomial_evaluate(coeffs[k][dim], t, d) for dim in range(3)])
                    assert np.array_equal(got, want)


@pytest.mark.parametrize("dt", [0.01, 0.1, 0.37])
def test_evaluate_range_bit_exact(ref, orc, solved, dt):
    """ref Trajectory::evaluateRange (src/trajectory.cpp:68-128): sample count, sampling times (accumulated, not
    multiplied) and values."""
    for coeffs, times in solved:
        total = float(np.sum(times))
        for t0, t1 in ((0.0, total), (0.0, 0.5 * total), (0.31 * total, 0.9 * total), (times[0], total)):
            for d in (0, 2):
                want, want_t = ref.trajectory_evaluate_range(coeffs, times, t0, t1, dt, d)
                got, got_t = orc.trajectory_evaluate_range(coeffs, times, t0, t1, dt, d)
                assert len(got) == len(want)
                assert np.array_equal(got_t, want_t)
                assert np.array_equal(got, want)


def test_minmax_magnitude_against_reference(ref, orc, solved):
    """ref Trajectory::computeMinMaxMagnitude (src/trajectory.cpp:181-217) over the reference's own Jenkins-Traub
    roots, against the oracle's interval isolation: extremal values to 1e-8 relative, same segment."""
    for coeffs, times in solved:
        for k in (1, 2):
            ok, mn, mx = ref.trajectory_minmax_magnitude(coeffs, times, k)
            assert ok
            r = orc.minmax_magnitude(coeffs, times, k, 1)
            assert abs(r["max"][1] - mx[1]) <= 1e-8 * max(1.0, abs(mx[1]))
            assert abs(r["min"][1] - mn[1]) <= 1e-8 * max(1.0, abs(mn[1]))
            assert r["max"][2] == mx[2]
            # one dimension: Polynomial::computeMinMax route (src/polynomial.cpp:95-108)
            ok, mn1, mx1 = ref.trajectory_minmax_magnitude(coeffs, times, k, dims=[1])
            r1 = orc.minmax_magnitude(coeffs, times, k, 1, dims=[1])
            assert ok and abs(r1["max"][1] - mx1[1]) <= 1e-8 * max(1.0, abs(mx1[1]))


def test_segment_candidates_against_reference(ref, orc, solved):
    """ref Segment::computeMinMaxMagnitudeCandidates (src/segment.cpp:132-156): the largest candidate value of an
    inner segment agrees with the oracle's segment maximum."""
    coeffs, times = solved[1]
    for s in (2, 5, 7):
        got = ref.segment_minmax_candidates(coeffs[s], times[s], 1, 0.0, times[s])
        assert got is not None
        _, values = got
        cand = orc.segment_candidate_roots(coeffs[s], 1, 0.0, times[s])
        best = max([orc.segment_magnitude(coeffs[s], 1, t) for t in list(cand) + [0.0, times[s]]])
        ref_best = max(list(values) + [orc.segment_magnitude(coeffs[s], 1, 0.0), orc.segment_magnitude(coeffs[s], 1, times[s])])
        assert abs(best - ref_best) <= 1e-8 * max(1.0, ref_best)


def test_product_input_generator_against_reference(ref):
    """The product's host-side createRandomVertices (csrc/minsnap_host_inputs.cpp, no GPU involved) against the
    reference's own function, bit for bit."""
    import mav_trajectory_generation_cmake_b200 as ms
    got = ms.random_positions_host(64, 10, BOX_LO, BOX_HI, 12345)
    for b in (0, 1, 17, 63):
        want, _ = ref.create_random_positions(4, 10, BOX_LO, BOX_HI, 12345 + b)
        assert np.array_equal(got[b], want)


@pytest.mark.gpu
def test_gpu_path_against_reference_itself(ref):
    """Solve on the GPU, then sample and evaluate ranges on the GPU, and compare with the REFERENCE's own
    Trajectory::evaluate / evaluateRange / estimateSegmentTimes on the GPU's coefficients (no oracle in between)."""
    import torch
    import mav_trajectory_generation_cmake_b200 as ms
    B, K = 32, 10
    pos = ms.random_positions_host(B, K, BOX_LO, BOX_HI, 978)
    pos_d = torch.from_numpy(pos).cuda()
    times_d = ms.estimate_segment_times(pos_d, 3.0, 5.0)
    times = times_d.cpu().numpy()
    for b in range(B):
        want = ref.estimate_segment_times(pos[b], 3.0, 5.0, 6.5)
        assert np.abs(times[b] - want).max() <= 4e-16 * np.abs(want).max()
    out = ms.solve_standard(pos_d, times_d)
    coeffs = out["coeffs"].cpu().numpy()
    samples, ts = ms.sample_uniform(out["coeffs"], times_d, 97, 5, want_times=True)
    samples, ts = samples.cpu().numpy(), ts.cpu().numpy()
    for b in range(0, B, 5):
        for m in range(0, 97, 3):
            for d in range(5):
                want = ref.trajectory_evaluate(coeffs[b], times[b], float(ts[b, m]), d)
                assert np.abs(samples[b, m, d] - want).max() <= 1e-6
    vals, t_out, count = ms.evaluate_range(out["coeffs"][:4], times_d[:4], 0.0, float(times[:4].sum(1).min()), 0.05, 1, 4096)
    vals, t_out, count = vals.cpu().numpy(), t_out.cpu().numpy(), count.cpu().numpy()
    for b in range(4):
        want, want_t = ref.trajectory_evaluate_range(coeffs[b], times[b], 0.0, float(times[:4].sum(1).min()), 0.05, 1)
        assert count[b] == len(want)
        assert np.array_equal(t_out[b, : count[b]], want_t)
        assert np.abs(vals[b, : count[b]] - want).max() <= 1e-6
