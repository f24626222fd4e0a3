// C++ drop-in tests: the reference's own test cases for the hot path
// (test/test_polynomial_optimization.cpp of magrimm/mav_trajectory_generation_cmake), re-expressed
// against include/mav_trajectory_generation/*.h, which reach the GPU only through the C ABI.
// Needs a CUDA device (run by tests/test_cpp_dropin.py under the gpu marker).
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "mav_trajectory_generation/polynomial_optimization_linear.h"

using namespace mav_trajectory_generation;

static int g_failures = 0;
static int g_checks = 0;
#define EXPECT_TRUE(cond)                                                        \
  do {                                                                           \
    ++g_checks;                                                                  \
    if (!(cond)) {                                                               \
      ++g_failures;                                                              \
      std::printf("  FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);            \
    }                                                                            \
  } while (0)
#define EXPECT_LT(a, b) EXPECT_TRUE((a) < (b))
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

const int N = 10;
const int max_derivative = derivative_order::SNAP;
const int derivative_to_optimize = derivative_order::SNAP;

static double maxAbsDiff(const Eigen::VectorXd& a, const Eigen::VectorXd& b) {
  double m = 0.0;
  for (long i = 0; i < a.size(); ++i) m = std::max(m, std::fabs(a[i] - b[i]));
  return m;
}

// T:73-131 checkPath: fixed constraints met at both ends of every segment, derivatives 0..N/2-1
// continuous across vertices, tolerance 1e-6.  One batched GPU evaluation per segment end.
static void checkPath(const Vertex::Vector& vertices, const std::vector<Segment>& segments) {
  const double tol = 1e-6;
  EXPECT_EQ(segments.size(), vertices.size() - 1);
  const int h = N / 2;
  std::vector<std::vector<double> > at_start(segments.size()), at_end(segments.size());
  for (size_t i = 0; i < segments.size(); ++i) {
    std::vector<double> ts = {0.0, segments[i].getTime()};
    std::vector<double> out = segments[i].evaluateBatch(ts, h);   // [2][h][D]
    const size_t per = static_cast<size_t>(h) * segments[i].D();
    at_start[i].assign(out.begin(), out.begin() + per);
    at_end[i].assign(out.begin() + per, out.end());
  }
  for (size_t i = 0; i < segments.size(); ++i) {
    const int D = segments[i].D();
    for (int end = 0; end < 2; ++end) {
      const Vertex& v = vertices[i + end];
      const std::vector<double>& val = end == 0 ? at_start[i] : at_end[i];
      for (Vertex::Constraints::const_iterator it = v.cBegin(); it != v.cEnd(); ++it)
        for (int d = 0; d < D; ++d) EXPECT_LT(std::fabs(it->second[d] - val[it->first * D + d]), tol);
    }
    if (i > 0)
      for (int k = 0; k < h; ++k)
        for (int d = 0; d < D; ++d) EXPECT_LT(std::fabs(at_end[i - 1][k * D + d] - at_start[i][k * D + d]), tol);
  }
}

// T:61-71 computeCostNumeric (Riemann sum with dt = 1e-3) and T:133-152 checkCost (10 %).
static bool checkCost(double cost_to_check, const std::vector<Segment>& segments, int derivative, double rel_tol) {
  const double dt = 0.001;
  double cost_numeric = 0.0;
  for (const Segment& s : segments) {
    std::vector<double> ts;
    for (double t = 0; t < s.getTime(); t += dt) ts.push_back(t);
    std::vector<double> out = s.evaluateBatch(ts, derivative + 1);
    const int D = s.D();
    for (size_t m = 0; m < ts.size(); ++m) {
      double sq = 0.0;
      for (int d = 0; d < D; ++d) {
        const double v = out[(m * (derivative + 1) + derivative) * D + d];
        sq += v * v;
      }
      cost_numeric += sq * dt;
    }
  }
  return std::fabs(cost_numeric - cost_to_check) <= cost_numeric * rel_tol;
}

static double maximumMagnitude(const std::vector<Segment>& segments, int derivative, double dt = 0.01) {
  double maximum = -1e9;
  for (const Segment& s : segments) {
    std::vector<double> ts;
    for (double t = 0; t < s.getTime(); t += dt) ts.push_back(t);
    std::vector<double> out = s.evaluateBatch(ts, derivative + 1);
    for (size_t m = 0; m < ts.size(); ++m) {
      double sq = 0.0;
      for (int d = 0; d < s.D(); ++d) {
        const double v = out[(m * (derivative + 1) + derivative) * s.D() + d];
        sq += v * v;
      }
      maximum = std::max(maximum, std::sqrt(sq));
    }
  }
  return maximum;
}

// T:154-192
static void testVertexGeneration() {
  Vertex::Vector vertices = createRandomVertices1D(max_derivative, 100, -50, 50, 0);
  EXPECT_EQ(vertices.front().getNumberOfConstraints(), static_cast<size_t>(N / 2));
  EXPECT_EQ(vertices.back().getNumberOfConstraints(), static_cast<size_t>(N / 2));
  for (const Vertex& v : vertices) {
    EXPECT_TRUE(v.hasConstraint(derivative_order::POSITION));
    Eigen::VectorXd c;
    v.getConstraint(derivative_order::POSITION, &c);
    EXPECT_TRUE(c[0] <= 50 && c[0] >= -50);
  }
  Eigen::VectorXd pos_min(3), pos_max(3);
  pos_min << -10.0, -20.0, -10.0;
  pos_max << 10.0, 20.0, 10.0;
  vertices = createRandomVertices(max_derivative, 100, pos_min, pos_max, 12345);
  EXPECT_EQ(vertices.size(), static_cast<size_t>(101));
  for (const Vertex& v : vertices) {
    Eigen::VectorXd c;
    EXPECT_TRUE(v.getConstraint(derivative_order::POSITION, &c));
    for (int i = 0; i < 3; ++i) EXPECT_TRUE(c[i] <= pos_max[i] && c[i] >= pos_min[i]);
  }
}

// T:194-204: A * invertMappingMatrix(A) == I, and the inverse maps back (1e-10 scale)
static void testMappingMatrixInversion() {
  for (double t = 1; t <= 60; t += 1) {
    PolynomialOptimization<N>::SquareMatrix A, Ai;
    PolynomialOptimization<N>::setupMappingMatrix(t, &A);
    PolynomialOptimization<N>::invertMappingMatrix(A, &Ai);
    // residual of Ai * A against the identity, row-scaled like the inverse itself
    double worst = 0.0;
    for (int r = 0; r < N; ++r) {
      double row_scale = 0.0;
      for (int c = 0; c < N; ++c) row_scale = std::max(row_scale, std::fabs(Ai(r, c)));
      for (int c = 0; c < N; ++c) {
        double acc = 0.0, mag = 0.0;
        for (int k = 0; k < N; ++k) {
          acc += Ai(r, k) * A(k, c);
          mag = std::max(mag, std::fabs(Ai(r, k) * A(k, c)));
        }
        worst = std::max(worst, std::fabs(acc - (r == c ? 1.0 : 0.0)) / std::max(1.0, mag));
      }
    }
    EXPECT_LT(worst, 1e-10);
  }
}

// T:206-394: random 1-D / 3-D paths, checkPath + checkCost + loose magnitude bounds
static void testUnconstrained(int D, int n_segments, size_t seed, double v_factor, double box_1d = 10.0) {
  Vertex::Vector vertices;
  if (D == 1) {
    vertices = createRandomVertices1D(max_derivative, n_segments, -box_1d, box_1d, seed);
  } else {
    Eigen::VectorXd pos_min(3), pos_max(3);
    pos_min << -10.0, -20.0, -10.0;
    pos_max << 10.0, 20.0, 10.0;
    vertices = createRandomVertices(max_derivative, n_segments, pos_min, pos_max, seed);
  }
  const double approximate_v_max = 3.0, approximate_a_max = 5.0;
  std::vector<double> segment_times = estimateSegmentTimes(vertices, approximate_v_max, approximate_a_max);
  PolynomialOptimization<N> opt(D);
  opt.setupFromVertices(vertices, segment_times, derivative_to_optimize);
  opt.solveLinear();
  Segment::Vector segments;
  opt.getSegments(&segments);
  checkPath(vertices, segments);
  const double v_max = maximumMagnitude(segments, derivative_order::VELOCITY);
  const double a_max = maximumMagnitude(segments, derivative_order::ACCELERATION);
  EXPECT_LT(v_max, approximate_v_max * v_factor);
  EXPECT_LT(a_max, approximate_a_max * 2.0);
  EXPECT_TRUE(checkCost(opt.computeCost(), segments, derivative_to_optimize, 0.1));
  EXPECT_EQ(opt.getNumberFixedConstraints(), static_cast<size_t>(n_segments + 9));
  EXPECT_EQ(opt.getNumberFreeConstraints(), static_cast<size_t>(4 * (n_segments - 1)));
  // Trajectory::evaluate: vertex instants belong to the segment on their right, past the end -> zeros
  Trajectory trajectory;
  opt.getTrajectory(&trajectory);
  Eigen::VectorXd p1;
  vertices[1].getConstraint(derivative_order::POSITION, &p1);
  EXPECT_LT(maxAbsDiff(trajectory.evaluate(segment_times[0], derivative_order::POSITION), p1), 1e-6);
  EXPECT_TRUE(trajectory.evaluate(trajectory.getMaxTime() + 1.0, 0).isZero(0.0));
}

// T:700-744: the known-answer vector
static void testTwoVerticesSetup() {
  Vertex start_vertex(1);
  start_vertex.addConstraint(derivative_order::POSITION, 0.0);
  start_vertex.addConstraint(derivative_order::VELOCITY, 0.0);
  start_vertex.addConstraint(derivative_order::ACCELERATION, 0.0);
  start_vertex.addConstraint(derivative_order::JERK, 0.0);
  start_vertex.addConstraint(derivative_order::SNAP, 0.0);
  Vertex goal_vertex = start_vertex;
  goal_vertex.addConstraint(derivative_order::POSITION, 5.0);
  const double kSegmentTime = std::fabs(5.0 - 0.0) * 2.0 / 2.0;
  PolynomialOptimization<10> opt(1);
  Vertex::Vector vertices{start_vertex, goal_vertex};
  std::vector<double> segment_times{kSegmentTime};
  opt.setupFromVertices(vertices, segment_times, derivative_order::SNAP);
  opt.solveLinear();
  Segment::Vector segments;
  opt.getSegments(&segments);
  checkPath(vertices, segments);
  Eigen::VectorXd matlab_coeffs(10);
  matlab_coeffs << -0.000000000000004, 0.000000000000004, -0.000000000000006, 0.000000000000003, -0.000000000000001,
      0.201600000000015, -0.134400000000012, 0.034560000000004, -0.004032000000000, 0.000179200000000;
  Eigen::VectorXd coeffs = segments[0].getPolynomialsRef()[0].getCoefficients();
  EXPECT_LT(maxAbsDiff(matlab_coeffs, coeffs), 2e-14);
  EXPECT_EQ(opt.getNumberFreeConstraints(), static_cast<size_t>(0));
}

// T:747-774
static void testTwoVerticesRand() {
  Eigen::VectorXd min_pos = Eigen::VectorXd::Constant(3, -50.0), max_pos = Eigen::VectorXd::Constant(3, 50.0);
  for (size_t i = 0; i < 100; i++) {
    Vertex::Vector vertices = createRandomVertices(derivative_order::ACCELERATION, 1, min_pos, max_pos, 12345 + i);
    std::vector<double> segment_times = estimateSegmentTimes(vertices, 3.0, 5.0);
    PolynomialOptimization<N> opt(3);
    opt.setupFromVertices(vertices, segment_times);
    opt.solveLinear();
    Segment::Vector segments;
    opt.getSegments(&segments);
    checkPath(vertices, segments);
  }
}

// T:777-836: [d_f; d_p] -> p -> [d_f; d_p]
static void testConstraintPacking() {
  Eigen::VectorXd min_pos = Eigen::VectorXd::Constant(3, -50.0), max_pos = Eigen::VectorXd::Constant(3, 50.0);
  for (size_t s = 0; s < 100; s++) {
    Vertex::Vector vertices = createRandomVertices(derivative_order::JERK, 5, min_pos, max_pos, 12345 + s);
    std::vector<double> segment_times = estimateSegmentTimes(vertices, 3.0, 5.0);
    PolynomialOptimization<N> opt(3);
    opt.setupFromVertices(vertices, segment_times);
    opt.solveLinear();
    Segment::Vector segments;
    opt.getSegments(&segments);
    std::vector<Eigen::VectorXd> fixed_constraints, free_constraints;
    opt.getFixedConstraints(&fixed_constraints);
    opt.getFreeConstraints(&free_constraints);
    Eigen::MatrixXd M, A_inv, A, M_pinv;
    opt.getM(&M);
    opt.getAInverse(&A_inv);
    opt.getA(&A);
    opt.getMpinv(&M_pinv);
    EXPECT_EQ(fixed_constraints.size(), static_cast<size_t>(3));
    EXPECT_EQ(free_constraints.size(), static_cast<size_t>(3));
    for (int i = 0; i < 3; ++i) {
      const long nf = fixed_constraints[i].size(), np = free_constraints[i].size();
      Eigen::VectorXd d_all_ordered(nf + np);
      for (long c = 0; c < nf; ++c) d_all_ordered[c] = fixed_constraints[i][c];
      for (long c = 0; c < np; ++c) d_all_ordered[nf + c] = free_constraints[i][c];
      Eigen::VectorXd p = A_inv * (M * d_all_ordered);
      Eigen::VectorXd d_unordered = A * p;
      Eigen::VectorXd d_reordered = M_pinv * d_unordered;
      EXPECT_LT(maxAbsDiff(d_all_ordered, d_reordered), 1e-6);
      for (size_t j = 0; j < segments.size(); ++j) {
        Eigen::VectorXd p_seg = segments[j][i].getCoefficients(0);
        for (int n = 0; n < N; ++n) EXPECT_LT(std::fabs(p_seg[n] - p[static_cast<long>(j * N + n)]), 1e-6);
      }
    }
    if (s == 0) {
      // getR: symmetric, and its free/free block is positive on the optimum direction
      Eigen::MatrixXd R;
      opt.getR(&R);
      EXPECT_EQ(R.rows(), static_cast<long>(opt.getNumberFixedConstraints() + opt.getNumberFreeConstraints()));
      EXPECT_LT((R - R.transpose()).maxAbs(), 1e-9 * R.maxAbs());
      // setFreeConstraints(getFreeConstraints) reproduces the solved segments
      Segment::Vector before = segments;
      opt.setFreeConstraints(free_constraints);
      Segment::Vector after;
      opt.getSegments(&after);
      for (size_t j = 0; j < after.size(); ++j)
        for (int d = 0; d < 3; ++d)
          EXPECT_LT(maxAbsDiff(before[j][d].getCoefficients(0), after[j][d].getCoefficients(0)),
                    1e-9 * (1.0 + before[j][d].getCoefficients(0).maxAbs()));
    }
  }
}

// evaluateRange (ref src/trajectory.cpp:68-128) against single evaluations; batch API
static void testRangeAndBatch() {
  Eigen::VectorXd pos_min(3), pos_max(3);
  pos_min << -10.0, -20.0, -10.0;
  pos_max << 10.0, 20.0, 10.0;
  Vertex::Vector vertices = createRandomVertices(max_derivative, 10, pos_min, pos_max, 978);
  std::vector<double> segment_times = estimateSegmentTimes(vertices, 3.0, 5.0);
  PolynomialOptimization<N> opt(3);
  opt.setupFromVertices(vertices, segment_times, derivative_to_optimize);
  opt.solveLinear();
  Trajectory trajectory;
  opt.getTrajectory(&trajectory);
  std::vector<Eigen::VectorXd> result;
  std::vector<double> sampling_times;
  trajectory.evaluateRange(trajectory.getMinTime(), trajectory.getMaxTime(), 0.1, derivative_order::VELOCITY, &result,
                           &sampling_times);
  EXPECT_TRUE(result.size() == sampling_times.size() && result.size() > 10);
  std::vector<double> singles = trajectory.evaluateBatch(sampling_times, 2);
  for (size_t m = 0; m < result.size(); ++m)
    for (int d = 0; d < 3; ++d) EXPECT_LT(std::fabs(result[m][d] - singles[(m * 2 + 1) * 3 + d]), 1e-6);

  // A range that starts INSIDE a segment and ends before the trajectory does: the reference's loop runs its
  // accumulated time from the start of that segment (src/trajectory.cpp:104-127) and so emits more samples than
  // (t_end - t_start) / dt; the mirror must return all of them (it repeats the call with the reported count).
  {
    const double t_start = segment_times[0] + segment_times[1] + 0.9 * segment_times[2];
    const double t_end = t_start + 0.5, dt = 0.01;
    size_t want = 0;   // the reference's loop, counted
    {
      const double seg_start = segment_times[0] + segment_times[1];
      double accumulated = seg_start, in_segment = t_start - seg_start;
      size_t i = 2;
      while (accumulated < t_end) {
        if (in_segment > segment_times[i]) {
          in_segment = in_segment - segment_times[i];
          if (++i >= segment_times.size()) break;
          continue;
        }
        ++want;
        in_segment += dt;
        accumulated += dt;
      }
    }
    trajectory.evaluateRange(t_start, t_end, dt, derivative_order::POSITION, &result, &sampling_times);
    EXPECT_TRUE(want > static_cast<size_t>((t_end - t_start) / dt) + 4);
    EXPECT_EQ(result.size(), want);
    EXPECT_EQ(sampling_times.size(), want);
  }

  // the additive batched optimizer agrees with the per-problem class
  const int B = 64, K = 10;
  std::vector<double> positions, times;
  std::vector<Vertex::Vector> all;
  for (int b = 0; b < B; ++b) {
    Vertex::Vector vs = createRandomVertices(max_derivative, K, pos_min, pos_max, 12345 + b);
    std::vector<double> ts = estimateSegmentTimes(vs, 3.0, 5.0);
    for (const Vertex& v : vs) {
      Eigen::VectorXd p;
      v.getConstraint(derivative_order::POSITION, &p);
      for (int d = 0; d < 3; ++d) positions.push_back(p[d]);
    }
    times.insert(times.end(), ts.begin(), ts.end());
    all.push_back(vs);
  }
  PolynomialOptimizationBatch<N> batch(3, K);
  batch.solve(positions, times);
  for (int b = 0; b < B; b += 9) {
    EXPECT_EQ(batch.status(b), 0);
    PolynomialOptimization<N> single(3);
    single.setupFromVertices(all[b], std::vector<double>(times.begin() + b * K, times.begin() + (b + 1) * K));
    single.solveLinear();
    Segment::Vector segs;
    single.getSegments(&segs);
    Trajectory tb;
    batch.getTrajectory(b, &tb);
    for (int i = 0; i < K; ++i)
      for (int d = 0; d < 3; ++d) {
        const Eigen::VectorXd a = segs[i][d].getCoefficients(0), c = tb.segments()[i][d].getCoefficients(0);
        EXPECT_LT(maxAbsDiff(a, c), 1e-9 * a.maxAbs());
      }
    EXPECT_LT(std::fabs(batch.cost(b) / single.computeCost() - 1.0), 1e-8);
  }
}

// T:396-416 checkExtrema: every analytic candidate has a sampled one within tol
static bool checkExtrema(const std::vector<double>& testee, const std::vector<double>& reference, double tol = 0.01) {
  for (double t : testee) {
    bool found_match = false;
    for (double r : reference)
      if (std::fabs(t - r) < tol) {
        found_match = true;
        break;
      }
    if (!found_match) {
      std::printf("  no sampled extremum near t = %.6f\n", t);
      return false;
    }
  }
  return true;
}

// T:418-507 (1-D) and T:509-612 (3-D): segment extrema of magnitude, analytic against sampling
static void testExtremaOfMagnitude(int D, int n_segments, size_t seed) {
  Vertex::Vector vertices;
  if (D == 1) {
    vertices = createRandomVertices1D(max_derivative, n_segments, -10, 10, seed);
  } else {
    Eigen::VectorXd pos_min(3), pos_max(3);
    pos_min << -10.0, -9.0, -8.0;
    pos_max << 8.0, 9.0, 10.0;
    vertices = createRandomVertices(max_derivative, n_segments, pos_min, pos_max, seed);
  }
  std::vector<double> segment_times = estimateSegmentTimes(vertices, 3.0, 5.0);
  PolynomialOptimization<N> opt(D);
  opt.setupFromVertices(vertices, segment_times, derivative_to_optimize);
  opt.solveLinear();
  Segment::Vector segments;
  opt.getSegments(&segments);

  std::vector<int> dimensions;
  for (int d = 0; d < D; ++d) dimensions.push_back(d);
  int segment_idx = 0;
  for (const Segment& s : segments) {
    if (segment_idx % 5 == 0) {   // every fifth segment: the sampled comparison costs 2 launches each
      std::vector<double> res, res_template_free, res_sampling;
      opt.computeSegmentMaximumMagnitudeCandidates<1>(s, 0, s.getTime(), &res);
      s.computeMinMaxMagnitudeCandidateTimes(1, 0.0, s.getTime(), dimensions, &res_template_free);
      opt.computeSegmentMaximumMagnitudeCandidatesBySampling<1>(s, 0, s.getTime(), 0.001, &res_sampling);
      // the rest ends carry a root of multiplicity 7 that floating point splits into a cluster
      std::vector<double> simple;
      for (double t : res) {
        const bool rest_end = (segment_idx == 0 && t < 0.02 * s.getTime()) ||
                              (segment_idx == n_segments - 1 && t > 0.98 * s.getTime());
        if (!rest_end) simple.push_back(t);
      }
      EXPECT_TRUE(checkExtrema(simple, res_sampling, 0.01));
      EXPECT_EQ(res.size(), res_template_free.size() - 2);
      for (size_t i = 0; i < res.size() && i + 2 < res_template_free.size(); i++)
        EXPECT_EQ(res[i], res_template_free[i + 2]);
    }
    ++segment_idx;
  }

  // the exact candidate polynomial for the comparison with sampling (the reference's coefficient
  // threshold loses extrema on the few segments longer than ~12 s)
  gpu::keepSmallCoefficients(true);
  const double v_max_ref = maximumMagnitude(segments, derivative_order::VELOCITY);
  const double a_max_ref = maximumMagnitude(segments, derivative_order::ACCELERATION);
  std::vector<Extremum> candidates;
  const Extremum v_max = opt.computeMaximumOfMagnitude<derivative_order::VELOCITY>(&candidates);
  const Extremum a_max = opt.computeMaximumOfMagnitude<derivative_order::ACCELERATION>(nullptr);
  EXPECT_LT(std::fabs(v_max_ref - v_max.value), 0.01);
  EXPECT_LT(std::fabs(a_max_ref - a_max.value), 0.01);
  // the candidate list holds the reported maximum, and nothing larger
  bool found = false;
  for (const Extremum& c : candidates) {
    EXPECT_TRUE(c.value <= v_max.value);
    found = found || (c.value == v_max.value && c.time == v_max.time && c.segment_idx == v_max.segment_idx);
  }
  EXPECT_TRUE(found);
  EXPECT_TRUE(candidates.size() >= static_cast<size_t>(n_segments + 1));
  // Trajectory::computeMinMaxMagnitude agrees with the optimisation's maximum
  Trajectory trajectory;
  opt.getTrajectory(&trajectory);
  Extremum t_min, t_max;
  EXPECT_TRUE(trajectory.computeMinMaxMagnitude(derivative_order::VELOCITY, dimensions, &t_min, &t_max));
  EXPECT_LT(std::fabs(t_max.value - v_max.value), 1e-9);
  EXPECT_LT(t_min.value, 1e-9);   // rest to rest
  // and with the per-segment route
  Extremum s_min, s_max;
  std::vector<Extremum> seg_candidates;
  const Segment& seg = segments[static_cast<size_t>(t_max.segment_idx)];
  EXPECT_TRUE(seg.computeMinMaxMagnitudeCandidates(derivative_order::VELOCITY, 0.0, seg.getTime(), dimensions,
                                                   &seg_candidates));
  EXPECT_TRUE(seg.selectMinMaxMagnitudeFromCandidates(0.0, seg.getTime(), derivative_order::VELOCITY, dimensions,
                                                      seg_candidates, &s_min, &s_max));
  EXPECT_LT(std::fabs(s_max.value - t_max.value), 1e-9);
  gpu::keepSmallCoefficients(false);
  // the additive batch class reports the same maximum for the same problem
  if (D == 3) {
    std::vector<double> flat;
    for (const Vertex& v : vertices) {
      Eigen::VectorXd pos;
      v.getConstraint(derivative_order::POSITION, &pos);
      for (int d = 0; d < D; ++d) flat.push_back(pos[d]);
    }
    PolynomialOptimizationBatch<N> batch(D, n_segments);
    batch.solve(flat, segment_times);
    const std::vector<Extremum> maxima = batch.computeMaximumOfMagnitude<derivative_order::VELOCITY>();
    EXPECT_EQ(maxima.size(), static_cast<size_t>(1));
    EXPECT_LT(std::fabs(maxima[0].value - v_max.value), 1e-9 * v_max.value);
    EXPECT_EQ(maxima[0].segment_idx, v_max.segment_idx);
  }
  // reference-compatible mode: never above the exact maximum
  const Extremum v_compat = opt.computeMaximumOfMagnitude<derivative_order::VELOCITY>(nullptr);
  EXPECT_TRUE(v_compat.value <= v_max.value * (1 + 1e-12));
  // bad arguments keep the reference's conventions (src/segment.cpp:89-102)
  std::vector<double> none;
  EXPECT_TRUE(!segments[0].computeMinMaxMagnitudeCandidateTimes(1, 0.0, 1.0, std::vector<int>(), &none));
  EXPECT_TRUE(!segments[0].computeMinMaxMagnitudeCandidateTimes(1, 0.0, 1.0, std::vector<int>{0, D}, &none));
}

// test/test_polynomial.cpp:62-128: Convolution and FindMinMax (random polynomials, extrema against sampling)
static void testPolynomialConvolutionAndMinMax() {
  Eigen::VectorXd coeffs_1(2), coeffs_2(2);
  coeffs_1 << 1.0, 2.0;
  coeffs_2 << -1.0, 3.0;
  const Polynomial product = Polynomial(coeffs_1) * Polynomial(coeffs_2);
  EXPECT_EQ(product.N(), 3);
  EXPECT_TRUE(product.getCoefficients()[0] == -1.0 && product.getCoefficients()[1] == 1.0 &&
              product.getCoefficients()[2] == 6.0);

  std::srand(1234567);
  auto uniform = [](double lo, double hi) { return lo + (hi - lo) * (std::rand() / static_cast<double>(RAND_MAX)); };
  for (int trial = 0; trial < 40; ++trial) {
    const int n = std::rand() % (Polynomial::kMaxN - 1) + 1;
    Eigen::VectorXd c(n);
    for (int i = 0; i < n; ++i) c[i] = uniform(-100.0, 100.0);
    const Polynomial p(c);
    const double t_start = uniform(0.0, 2.0), t_end = uniform(t_start, 4.0);
    std::pair<double, double> lo, hi;
    EXPECT_TRUE(p.computeMinMax(t_start, t_end, derivative_order::POSITION, &lo, &hi));
    // sampling (ref findMinMaxBySampling), evaluated in one launch
    std::vector<double> ts;
    for (double t = t_start; t <= t_end; t += 1.0e-3) ts.push_back(t);
    int n_padded = 4;                       // the batched sampler is built for N = 4, 6, ..., 12
    while (n_padded < n) n_padded += 2;
    Eigen::VectorXd c_padded(n_padded);
    c_padded.setZero();
    for (int i = 0; i < n; ++i) c_padded[i] = c[i];
    Segment seg(n_padded, 1);
    seg[0] = Polynomial(c_padded);
    const std::vector<double> values = seg.evaluateBatch(ts, 1);
    size_t i_lo = 0, i_hi = 0;
    for (size_t i = 0; i < ts.size(); ++i) {
      if (values[i] < values[i_lo]) i_lo = i;
      if (values[i] > values[i_hi]) i_hi = i;
    }
    // the analytic extremum is at least as extreme as any sample, and close to the sampled one
    const double scale = std::max(std::fabs(values[i_lo]), std::fabs(values[i_hi])) + 1.0;
    EXPECT_TRUE(lo.second <= values[i_lo] + 1e-9 * scale);
    EXPECT_TRUE(hi.second >= values[i_hi] - 1e-9 * scale);
    // ref approxEqual on the extremum TIMES (kEqualityResolution = 1e-2); two nearly equal extrema far apart
    // in time are told apart by value instead
    EXPECT_TRUE(std::fabs(lo.first - ts[i_lo]) < 1.0e-2 || std::fabs(lo.second - values[i_lo]) < 1e-3 * scale);
    EXPECT_TRUE(std::fabs(hi.first - ts[i_hi]) < 1.0e-2 || std::fabs(hi.second - values[i_hi]) < 1e-3 * scale);
  }
}

int main() {
  struct Case {
    const char* name;
    void (*fn)();
  };
  const Case cases[] = {
      {"PathPlanning_TestVertexGeneration", testVertexGeneration},
      {"PathPlanning_A_matrix_inversion", testMappingMatrixInversion},
      {"PathPlanningUnconstrained_1D_10_segments", [] { testUnconstrained(1, 10, 12, 2.0); }},
      {"PathPlanningUnconstrained_1D_50_segments", [] { testUnconstrained(1, 50, 123, 2.0); }},
      {"PathPlanningUnconstrained_1D_100_segments", [] { testUnconstrained(1, 100, 1234, 5.0); }},
      {"PathPlanningUnconstrained_1D_100_segments_high_segment_times",
       [] { testUnconstrained(1, 100, 12345, 5.0, 50.0); }},
      {"PathPlanningUnconstrained_3D_100_segments", [] { testUnconstrained(3, 100, 12345, 5.0); }},
      {"2_vertices_setup", testTwoVerticesSetup},
      {"2_vertices_rand", testTwoVerticesRand},
      {"ConstraintPacking", testConstraintPacking},
      {"EvaluateRange_and_Batch", testRangeAndBatch},
      {"Polynomial_Convolution_and_FindMinMax", testPolynomialConvolutionAndMinMax},
      {"PathOptimization_1D_segment_extrema_of_magnitude", [] { testExtremaOfMagnitude(1, 100, 1234); }},
      {"PathOptimization3D_segment_extrema_of_magnitude", [] { testExtremaOfMagnitude(3, 100, 978); }},
  };
  for (const Case& c : cases) {
    const int before = g_failures;
    c.fn();
    std::printf("[%s] %s\n", g_failures == before ? "  OK  " : "FAILED", c.name);
  }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures == 0 ? 0 : 1;
}
