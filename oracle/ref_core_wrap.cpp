// TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE's own value classes -- Vertex,
// Polynomial, Segment, Trajectory, createRandomVertices, estimateSegmentTimes -- compiled from where
// they lie under /root/reference (src/vertex.cpp, polynomial.cpp, segment.cpp, trajectory.cpp,
// motion_defines.cpp, rpoly.cpp) against the Eigen / glog stand-ins in oracle/ref_shim/ into
// oracle/_ref/libmav_ref_core.so (oracle/Makefile, target `ref`).  tests/test_reference_pinning.py
// checks the oracle (oracle/minsnap_oracle.c, extrema_oracle.c) against these functions; nothing of
// the product links or loads this file.
//
// Layouts: positions [K+1][D]; coeffs [K][D][N] increasing powers; times [K].
#include <cstdint>
#include <utility>
#include <vector>

#include "mav_trajectory_generation/polynomial.h"
#include "mav_trajectory_generation/segment.h"
#include "mav_trajectory_generation/trajectory.h"
#include "mav_trajectory_generation/vertex.h"

namespace mtg = mav_trajectory_generation;

#define REFC_API extern "C" __attribute__((visibility("default")))

namespace {
Eigen::VectorXd vec(const double* p, int n) {
  Eigen::VectorXd v(n);
  for (int i = 0; i < n; ++i) v[i] = p[i];
  return v;
}
mtg::Segment make_segment(int N, int D, const double* c, double T) {
  mtg::Segment s(N, D);
  for (int d = 0; d < D; ++d) s[d] = mtg::Polynomial(N, vec(c + d * N, N));
  s.setTime(T);
  return s;
}
mtg::Trajectory make_trajectory(int N, int K, int D, const double* c, const double* times) {
  mtg::Segment::Vector segs;
  for (int k = 0; k < K; ++k) segs.push_back(make_segment(N, D, c + (size_t)k * D * N, times[k]));
  mtg::Trajectory t;
  t.setSegments(segs);
  return t;
}
std::vector<int> dim_list(const int* dims, int n) { return std::vector<int>(dims, dims + n); }
}  // namespace

// ref createRandomVertices (src/vertex.cpp:27-79): positions of the K+1 vertices and the number of
// constraints every vertex carries.
REFC_API int refc_create_random_positions(int max_derivative, int K, int D, const double* pos_min, const double* pos_max,
                                          uint64_t seed, double* positions, int* n_constraints) {
  mtg::Vertex::Vector v = mtg::createRandomVertices(max_derivative, (size_t)K, vec(pos_min, D), vec(pos_max, D), (size_t)seed);
  if ((int)v.size() != K + 1) return -1;
  for (int i = 0; i <= K; ++i) {
    Eigen::VectorXd p;
    if (!v[i].getConstraint(mtg::derivative_order::POSITION, &p)) return -2;
    for (int d = 0; d < D; ++d) positions[i * D + d] = p[d];
    if (n_constraints) n_constraints[i] = (int)v[i].getNumberOfConstraints();
  }
  return 0;
}

// ref estimateSegmentTimes (src/vertex.cpp:162-178)
REFC_API int refc_estimate_segment_times(int K, int D, const double* positions, double v_max, double a_max, double magic,
                                         double* times) {
  mtg::Vertex::Vector v;
  for (int i = 0; i <= K; ++i) {
    mtg::Vertex x((size_t)D);
    x.addConstraint(mtg::derivative_order::POSITION, vec(positions + i * D, D));
    v.push_back(x);
  }
  std::vector<double> t = mtg::estimateSegmentTimes(v, v_max, a_max, magic);
  for (int i = 0; i < K; ++i) times[i] = t[i];
  return (int)t.size();
}

// ref computeBaseCoefficients (src/polynomial.cpp:140-155)
REFC_API void refc_base_coefficients(int n, double* out) {
  Eigen::MatrixXd m = mtg::computeBaseCoefficients(n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[i * n + j] = m(i, j);
}

// ref Polynomial::baseCoeffsWithTime (polynomial.h:215-233)
REFC_API void refc_base_coeffs_with_time(int N, int derivative, double t, double* out) {
  Eigen::VectorXd c = mtg::Polynomial::baseCoeffsWithTime(N, derivative, t);
  for (int i = 0; i < N; ++i) out[i] = c[i];
}

// ref Polynomial::evaluate(t, derivative) (polynomial.h:138-151)
REFC_API double refc_polynomial_evaluate(int N, const double* c, double t, int derivative) {
  return mtg::Polynomial(N, vec(c, N)).evaluate(t, derivative);
}

// ref Polynomial::evaluate(t, VectorXd*) (polynomial.h:120-134): derivatives 0..n_deriv-1
REFC_API void refc_polynomial_evaluate_all(int N, const double* c, double t, int n_deriv, double* out) {
  Eigen::VectorXd r(n_deriv);
  mtg::Polynomial(N, vec(c, N)).evaluate(t, &r);
  for (int i = 0; i < n_deriv; ++i) out[i] = r[i];
}

// ref Polynomial::getCoefficients(derivative) (polynomial.h:100-115)
REFC_API void refc_polynomial_get_coefficients(int N, const double* c, int derivative, double* out) {
  Eigen::VectorXd r = mtg::Polynomial(N, vec(c, N)).getCoefficients(derivative);
  for (int i = 0; i < N; ++i) out[i] = r[i];
}

// ref Polynomial::convolve (src/polynomial.cpp:157-175)
REFC_API int refc_convolve(const double* a, int na, const double* b, int nb, double* out) {
  Eigen::VectorXd r = mtg::Polynomial::convolve(vec(a, na), vec(b, nb));
  for (long i = 0; i < r.size(); ++i) out[i] = r[i];
  return (int)r.size();
}

// ref Polynomial::computeMinMax (src/polynomial.cpp:95-108): out = {t_min, v_min, t_max, v_max}
REFC_API int refc_polynomial_min_max(int N, const double* c, double t_start, double t_end, int derivative, double* out) {
  std::pair<double, double> mn, mx;
  const bool ok = mtg::Polynomial(N, vec(c, N)).computeMinMax(t_start, t_end, derivative, &mn, &mx);
  out[0] = mn.first; out[1] = mn.second; out[2] = mx.first; out[3] = mx.second;
  return ok ? 1 : 0;
}

// ref Segment::evaluate (src/segment.cpp:51-58)
REFC_API void refc_segment_evaluate(int N, int D, const double* c, double T, double t, int derivative, double* out) {
  Eigen::VectorXd r = make_segment(N, D, c, T).evaluate(t, derivative);
  for (int d = 0; d < D; ++d) out[d] = r[d];
}

// ref Trajectory::evaluate (src/trajectory.cpp:41-66)
REFC_API void refc_trajectory_evaluate(int N, int K, int D, const double* c, const double* times, double t, int derivative,
                                       double* out) {
  Eigen::VectorXd r = make_trajectory(N, K, D, c, times).evaluate(t, derivative);
  for (int d = 0; d < D; ++d) out[d] = r[d];
}

REFC_API double refc_trajectory_max_time(int N, int K, int D, const double* c, const double* times) {
  return make_trajectory(N, K, D, c, times).getMaxTime();
}

// ref Trajectory::evaluateRange (src/trajectory.cpp:68-128): returns the number of samples the reference emits
REFC_API int refc_trajectory_evaluate_range(int N, int K, int D, const double* c, const double* times, double t_start,
                                            double t_end, double dt, int derivative, int max_out, double* out,
                                            double* sample_times) {
  std::vector<Eigen::VectorXd> res;
  std::vector<double> ts;
  make_trajectory(N, K, D, c, times).evaluateRange(t_start, t_end, dt, derivative, &res, &ts);
  const int n = (int)res.size();
  for (int i = 0; i < n && i < max_out; ++i) {
    for (int d = 0; d < D; ++d) out[(size_t)i * D + d] = res[i][d];
    sample_times[i] = ts[i];
  }
  return n;
}

// ref Segment::computeMinMaxMagnitudeCandidates (src/segment.cpp:132-156): times and values of the candidates
REFC_API int refc_segment_minmax_candidates(int N, int D, const double* c, double T, int derivative, double t_start,
                                            double t_end, const int* dims, int n_dims, int max_out, double* cand_times,
                                            double* cand_values) {
  std::vector<mtg::Extremum> cand;
  const bool ok = make_segment(N, D, c, T).computeMinMaxMagnitudeCandidates(derivative, t_start, t_end,
                                                                             dim_list(dims, n_dims), &cand);
  if (!ok) return -1;
  for (int i = 0; i < (int)cand.size() && i < max_out; ++i) {
    cand_times[i] = cand[i].time;
    cand_values[i] = cand[i].value;
  }
  return (int)cand.size();
}

// ref Trajectory::computeMinMaxMagnitude (src/trajectory.cpp:181-217): out = {t_min, v_min, seg_min, t_max, v_max, seg_max}
REFC_API int refc_trajectory_minmax_magnitude(int N, int K, int D, const double* c, const double* times, int derivative,
                                              const int* dims, int n_dims, double* out) {
  mtg::Extremum mn, mx;
  const bool ok = make_trajectory(N, K, D, c, times).computeMinMaxMagnitude(derivative, dim_list(dims, n_dims), &mn, &mx);
  out[0] = mn.time; out[1] = mn.value; out[2] = mn.segment_idx;
  out[3] = mx.time; out[4] = mx.value; out[5] = mx.segment_idx;
  return ok ? 1 : 0;
}
