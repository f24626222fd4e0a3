// PolynomialOptimization<N>: unconstrained QP of Richter, Bry, Roy (ISRR 2013) for piecewise
// polynomial trajectories -- the C++ mirror of ref LIN.h / LIN.i for the hot path.  Same names,
// argument meaning and error behaviour (programmer errors CHECK-abort, oddities LOG(WARNING),
// the bool returns are always true).  All numerics happen on the GPU through the C ABI:
//
//   setupFromVertices   -> minsnap_reorder_host                    (ref LIN.i:46-99, 171-250)
//   solveLinear         -> minsnap_solve_host                      (ref LIN.i:328-369)
//                          (n_free == 0: minsnap_coeffs_from_constraints_host, LIN.i:333-339)
//   setFreeConstraints  -> minsnap_coeffs_from_constraints_host    (ref LIN.i:505-514, 252-273)
//   computeCost         -> minsnap_cost_host                       (ref LIN.i:113-130)
//   setupMappingMatrix / invertMappingMatrix / computeQuadraticCostJacobian / getA / getAInverse /
//   getR                -> minsnap_segment_matrices_host           (ref LIN.i:101-169, 573-589)
//
//   computeMaximumOfMagnitude / computeSegmentMaximumMagnitudeCandidates
//                       -> minsnap_extrema_host                    (ref LIN.i:378-503; the real roots are
//                          isolated on the GPU instead of by the reference's Jenkins-Traub routine)
// Additive: PolynomialOptimizationBatch<N> below solves many independent problems per call.
#ifndef MAV_TRAJECTORY_GENERATION_POLYNOMIAL_OPTIMIZATION_LINEAR_H_
#define MAV_TRAJECTORY_GENERATION_POLYNOMIAL_OPTIMIZATION_LINEAR_H_

#include <cstdint>
#include <ostream>
#include <vector>

#include "mav_trajectory_generation/extremum.h"
#include "mav_trajectory_generation/minsnap_gpu.h"
#include "mav_trajectory_generation/motion_defines.h"
#include "mav_trajectory_generation/polynomial.h"
#include "mav_trajectory_generation/segment.h"
#include "mav_trajectory_generation/trajectory.h"
#include "mav_trajectory_generation/vertex.h"

namespace mav_trajectory_generation {

template <int _N = 10>
class PolynomialOptimization {
  static_assert(_N % 2 == 0, "The number of coefficients has to be even.");

 public:
  enum { N = _N };
  static constexpr int kHighestDerivativeToOptimize = N / 2 - 1;
  typedef Eigen::Matrix<double, N, N> SquareMatrix;
  typedef std::vector<SquareMatrix, Eigen::aligned_allocator<SquareMatrix> > SquareMatrixVector;

  explicit PolynomialOptimization(size_t dimension)
      : dimension_(dimension),
        derivative_to_optimize_(derivative_order::kINVALID),
        n_vertices_(0),
        n_segments_(0),
        n_all_constraints_(0),
        n_fixed_constraints_(0),
        n_free_constraints_(0) {
    fixed_constraints_compact_.resize(dimension_);
    free_constraints_compact_.resize(dimension_);
  }

  bool setupFromVertices(const Vertex::Vector& vertices, const std::vector<double>& segment_times,
                         int derivative_to_optimize = kHighestDerivativeToOptimize) {
    CHECK(derivative_to_optimize >= 0 && derivative_to_optimize <= kHighestDerivativeToOptimize)
        << "You tried to optimize the " << derivative_to_optimize << "th derivative of position on a " << N
        << "th order polynomial. This is not possible, you either need a higher order polynomial or a smaller "
           "derivative to optimize.";
    derivative_to_optimize_ = derivative_to_optimize;
    vertices_ = vertices;
    segment_times_ = segment_times;
    n_vertices_ = vertices.size();
    n_segments_ = n_vertices_ - 1;
    segments_.resize(n_segments_, Segment(N, static_cast<int>(dimension_)));
    CHECK(n_vertices_ == segment_times.size() + 1) << "Size of times must be one less than positions.";

    // Constraints on derivatives the polynomial order cannot carry are dropped with a warning.
    for (size_t v = 0; v < n_vertices_; ++v) {
      Vertex kept(dimension_);
      bool all_valid = true;
      for (Vertex::Constraints::const_iterator it = vertices_[v].cBegin(); it != vertices_[v].cEnd(); ++it) {
        if (it->first > kHighestDerivativeToOptimize) {
          all_valid = false;
          LOG(WARNING) << "Invalid constraint on vertex " << v << ": maximum possible derivative is "
                       << kHighestDerivativeToOptimize << ", but was set to " << it->first << ". Ignoring constraint";
        } else {
          kept.addConstraint(it->first, it->second);
        }
      }
      if (!all_valid) vertices_[v] = kept;
    }
    updateSegmentTimes(segment_times);
    setupConstraintReordering();
    return true;
  }

  bool setupFromPositons(const std::vector<double>& positions, const std::vector<double>& times) {
    CHECK_EQ(dimension_, static_cast<size_t>(1));
    Vertex::Vector vertices;
    for (size_t i = 0; i < positions.size(); ++i) {
      Vertex v(1);
      if (i == 0 || i + 1 == positions.size()) v.makeStartOrEnd(positions[i], kHighestDerivativeToOptimize);
      else v.addConstraint(derivative_order::POSITION, positions[i]);
      vertices.push_back(v);
    }
    return setupFromVertices(vertices, times, kHighestDerivativeToOptimize);
  }

  // ---- per-segment matrices (closed forms evaluated on the GPU) ---------------------------------
  static void setupMappingMatrix(double segment_time, SquareMatrix* A) {
    CHECK_NOTNULL(A);
    double buf[N * N];
    gpu::check(minsnap_segment_matrices_host(1, N, 0, &segment_time, buf, nullptr, nullptr, nullptr),
               "minsnap_segment_matrices_host");
    fillSquare(buf, A);
  }
  // A = [A(0); A(T)] determines T through its position row at T: A(N/2, 1) = T.
  static void invertMappingMatrix(const SquareMatrix& mapping_matrix, SquareMatrix* inverse_mapping_matrix) {
    CHECK_NOTNULL(inverse_mapping_matrix);
    const double segment_time = mapping_matrix(N / 2, 1);
    CHECK(mapping_matrix(N / 2, 0) == 1.0 && segment_time > 0.0) << "not a mapping matrix [A(0); A(T)]";
    double buf[N * N];
    gpu::check(minsnap_segment_matrices_host(1, N, 0, &segment_time, nullptr, buf, nullptr, nullptr),
               "minsnap_segment_matrices_host");
    fillSquare(buf, inverse_mapping_matrix);
  }
  static void computeQuadraticCostJacobian(int derivative, double t, SquareMatrix* cost_jacobian) {
    CHECK_LT(derivative, static_cast<int>(N));
    CHECK_NOTNULL(cost_jacobian);
    CHECK(derivative <= kHighestDerivativeToOptimize) << "cost matrices are built for derivatives 0.." << N / 2 - 1;
    double buf[N * N];
    gpu::check(minsnap_segment_matrices_host(1, N, derivative, &t, nullptr, nullptr, buf, nullptr),
               "minsnap_segment_matrices_host");
    fillSquare(buf, cost_jacobian);
  }

  // 0.5 * sum over segments and dimensions of c^T Q c for the current segments.
  double computeCost() const {
    CHECK(n_segments_ == segments_.size());
    std::vector<double> coeffs = packCoefficients();
    double cost = 0.0;
    gpu::check(minsnap_cost_host(1, static_cast<int>(n_segments_), static_cast<int>(dimension_), N,
                                 derivative_to_optimize_, coeffs.data(), segment_times_.data(), &cost),
               "minsnap_cost_host");
    return cost;
  }

  // ---- extrema of |p^(Derivative)| (ref LIN.i:378-503) ---------------------------------------------
  // Appends to `candidates` the times inside [t_start, t_stop] (t_start >= 0) at which the magnitude of
  // the derivative may be extremal: the real roots of sum_dim p^(k) p^(k+1), or of p^(k+1) for one
  // dimension, in ascending order.
  template <int Derivative>
  static bool computeSegmentMaximumMagnitudeCandidates(const Segment& segment, double t_start, double t_stop,
                                                       std::vector<double>* candidates) {
    CHECK(candidates);
    static_assert(N - Derivative - 1 > 0, "N-Derivative-1 has to be greater 0");
    CHECK_EQ(segment.N(), static_cast<int>(N));
    CHECK_GE(t_start, 0.0) << "the candidate search runs over [0, t_stop]";
    if (t_start > t_stop) return true;
    const int D = segment.D();
    std::vector<double> coeffs = segment.packCoefficients();
    const int max_roots = minsnap_extrema_max_roots(N, Derivative, D);
    std::vector<double> times(max_roots + 2);
    int32_t n_roots = 0;
    gpu::check(minsnap_extrema_host(1, 1, D, N, coeffs.data(), &t_stop, Derivative,
                                    gpu::extremaMode(MINSNAP_EXTREMA_OPTIMIZATION), 0, nullptr, nullptr, nullptr, nullptr,
                                    nullptr, nullptr, times.data(), nullptr, &n_roots),
               "minsnap_extrema_host");
    for (int i = 0; i < n_roots; ++i)
      if (times[2 + i] >= t_start) candidates->push_back(times[2 + i]);
    return true;
  }

  // The same candidates found by sampling every dt and watching the squared magnitude turn around
  // (ref LIN.i:439-468); the samples are evaluated on the GPU in one launch.
  template <int Derivative>
  static void computeSegmentMaximumMagnitudeCandidatesBySampling(const Segment& segment, double t_start, double t_stop,
                                                                 double dt, std::vector<double>* candidates) {
    CHECK(candidates);
    CHECK_GT(dt, 0.0);
    std::vector<double> ts;
    ts.push_back(t_start - dt);
    ts.push_back(t_start);
    for (double t = t_start + dt; t < t_stop + 2 * dt; t += dt) ts.push_back(t);
    const std::vector<double> samples = segment.evaluateBatch(ts, Derivative + 1);
    const int D = segment.D();
    auto squared_norm = [&](size_t m) {
      double s = 0.0;
      for (int d = 0; d < D; ++d) {
        const double v = samples[(m * (Derivative + 1) + Derivative) * D + d];
        s += v * v;
      }
      return s;
    };
    auto sgn = [](double x) { return (0.0 < x) - (x < 0.0); };
    double value_old = squared_norm(1);
    double direction = value_old - squared_norm(0);
    for (size_t m = 2; m < ts.size(); ++m) {
      const double value_new = squared_norm(m);
      const double direction_new = value_new - value_old;
      if (sgn(direction) != sgn(direction_new)) candidates->push_back(ts[m] - dt);
      value_old = value_new;
      direction = direction_new;
    }
  }

  // Largest magnitude of the derivative over the solved trajectory: per segment its start and the
  // candidate times above, plus the end of the last segment; a candidate replaces the incumbent
  // only when strictly larger; `candidates`, when given, receives every candidate in that order.
  template <int Derivative>
  Extremum computeMaximumOfMagnitude(std::vector<Extremum>* candidates) const {
    static_assert(N - Derivative - 1 > 0, "N-Derivative-1 has to be greater 0");
    if (candidates != nullptr) candidates->clear();
    Extremum extremum;
    if (n_segments_ == 0) return extremum;
    CHECK(n_segments_ == segments_.size());
    const int K = static_cast<int>(n_segments_), D = static_cast<int>(dimension_);
    std::vector<double> coeffs = packCoefficients();
    std::vector<double> durations(n_segments_);
    for (size_t i = 0; i < n_segments_; ++i) durations[i] = segments_[i].getTime();
    const int stride = minsnap_extrema_max_roots(N, Derivative, D) + 2;
    std::vector<double> times, values;
    std::vector<int32_t> counts;
    if (candidates != nullptr) {
      times.resize(static_cast<size_t>(K) * stride);
      values.resize(static_cast<size_t>(K) * stride);
      counts.resize(n_segments_);
    }
    int32_t segment_idx = 0;
    gpu::check(minsnap_extrema_host(1, K, D, N, coeffs.data(), durations.data(), Derivative,
                                    gpu::extremaMode(MINSNAP_EXTREMA_OPTIMIZATION), 0, &extremum.time, &extremum.value,
                                    &segment_idx, nullptr, nullptr, nullptr,
                                    candidates ? times.data() : nullptr, candidates ? values.data() : nullptr,
                                    candidates ? counts.data() : nullptr),
               "minsnap_extrema_host");
    extremum.segment_idx = segment_idx;
    if (candidates != nullptr) {
      for (int s = 0; s < K; ++s) {
        const double* t = &times[static_cast<size_t>(s) * stride];
        const double* v = &values[static_cast<size_t>(s) * stride];
        candidates->emplace_back(t[0], v[0], s);
        for (int i = 0; i < counts[s]; ++i) candidates->emplace_back(t[2 + i], v[2 + i], s);
        if (s == K - 1) candidates->emplace_back(t[1], v[1], s);
      }
    }
    return extremum;
  }

  void updateSegmentTimes(const std::vector<double>& segment_times) {
    const size_t n_segment_times = segment_times.size();
    CHECK(n_segment_times == n_segments_) << "Number of segment times (" << n_segment_times
                                          << ") does not match number of segments (" << n_segments_ << ")";
    segment_times_ = segment_times;
    for (size_t i = 0; i < n_segments_; ++i)
      CHECK_GT(segment_times[i], 0) << "Segment times need to be greater than zero";
  }

  bool solveLinear() {
    CHECK(derivative_to_optimize_ >= 0 && derivative_to_optimize_ <= kHighestDerivativeToOptimize);
    const int K = static_cast<int>(n_segments_), D = static_cast<int>(dimension_);
    if (n_free_constraints_ == 0) {
      LOG(WARNING) << "No free constraints set in the vertices. Polynomial can not be optimized. Outputting fully "
                      "constrained polynomial.";
      updateSegmentsFromCompactConstraints();
      return true;
    }
    std::vector<double> fixed = interleave(fixed_constraints_compact_, n_fixed_constraints_);
    std::vector<double> coeffs(static_cast<size_t>(K) * D * N), free_values(n_free_constraints_ * D);
    int32_t status = 0;
    if (hasStandardStructure()) {
      // createRandomVertices structure (ends fully fixed, interior position only), N = 10, snap:
      // the single-launch kernel.  Column order of d_f: vertex 0 (derivatives 0..4), interior
      // positions, vertex K (derivatives 0..4).
      const int h = N / 2;
      std::vector<double> positions(static_cast<size_t>(K + 1) * D), ends(static_cast<size_t>(2) * (h - 1) * D);
      for (int d = 0; d < D; ++d) {
        positions[d] = fixed[d];
        for (int v = 1; v < K; ++v) positions[static_cast<size_t>(v) * D + d] = fixed[static_cast<size_t>(h - 1 + v) * D + d];
        positions[static_cast<size_t>(K) * D + d] = fixed[static_cast<size_t>(h + K - 1) * D + d];
        for (int c = 1; c < h; ++c) {
          ends[static_cast<size_t>(c - 1) * D + d] = fixed[static_cast<size_t>(c) * D + d];
          ends[static_cast<size_t>(h - 1 + c - 1) * D + d] = fixed[static_cast<size_t>(h + K - 1 + c) * D + d];
        }
      }
      gpu::check(minsnap_solve_standard_host(1, K, D, N, derivative_to_optimize_, positions.data(), ends.data(),
                                             segment_times_.data(), 0.0, 0.0, 0.0, nullptr, coeffs.data(),
                                             free_values.data(), nullptr, &status),
                 "minsnap_solve_standard_host");
    } else {
      gpu::check(minsnap_solve_host(1, K, D, N, derivative_to_optimize_, fixed_mask_.data(), fixed.data(),
                                    segment_times_.data(), coeffs.data(), free_values.data(), nullptr, &status,
                                    nullptr),
                 "minsnap_solve_host");
    }
    if (status != MINSNAP_STATUS_OK)
      LOG(WARNING) << "solveLinear: GPU status word " << status
                   << " (1 = R_pp not positive definite, 2 = bad segment time, 4 = non-finite coefficient)";
    for (int d = 0; d < D; ++d) {
      free_constraints_compact_[d].resize(static_cast<long>(n_free_constraints_));
      for (size_t c = 0; c < n_free_constraints_; ++c) free_constraints_compact_[d][c] = free_values[c * D + d];
    }
    storeCoefficients(coeffs);
    return true;
  }

  void getTrajectory(Trajectory* trajectory) const {
    CHECK_NOTNULL(trajectory);
    trajectory->setSegments(segments_);
  }
  void getSegments(Segment::Vector* segments) const {
    CHECK_NOTNULL(segments);
    *segments = segments_;
  }
  void getSegmentTimes(std::vector<double>* segment_times) const {
    CHECK(segment_times != nullptr);
    *segment_times = segment_times_;
  }
  void getFreeConstraints(std::vector<Eigen::VectorXd>* free_constraints) const {
    CHECK(free_constraints != nullptr);
    *free_constraints = free_constraints_compact_;
  }
  void setFreeConstraints(const std::vector<Eigen::VectorXd>& free_constraints) {
    CHECK(free_constraints.size() == dimension_);
    for (const Eigen::VectorXd& v : free_constraints) CHECK(static_cast<size_t>(v.size()) == n_free_constraints_);
    free_constraints_compact_ = free_constraints;
    updateSegmentsFromCompactConstraints();
  }
  void getFixedConstraints(std::vector<Eigen::VectorXd>* fixed_constraints) const {
    CHECK(fixed_constraints != nullptr);
    *fixed_constraints = fixed_constraints_compact_;
  }

  size_t getDimension() const { return dimension_; }
  size_t getNumberSegments() const { return n_segments_; }
  size_t getNumberAllConstraints() const { return n_all_constraints_; }
  size_t getNumberFixedConstraints() const { return n_fixed_constraints_; }
  size_t getNumberFreeConstraints() const { return n_free_constraints_; }

  // ---- accessors for the internal matrices (dense), built from the device outputs ----------------
  void getAInverse(Eigen::MatrixXd* A_inv) const { blockDiagonal(A_inv, 1); }
  void getA(Eigen::MatrixXd* A) const {
    for (size_t i = 0; i < n_segments_; ++i)
      CHECK_GT(segment_times_[i], 0) << "Segment times need to be greater than zero";
    blockDiagonal(A, 0);
  }
  // Reordering matrix C of [1]: one 1 per row, row r -> column col_of_row[r].
  void getM(Eigen::MatrixXd* M) const {
    CHECK_NOTNULL(M);
    M->resize(static_cast<long>(n_all_constraints_), static_cast<long>(n_fixed_constraints_ + n_free_constraints_));
    M->setZero();
    for (size_t r = 0; r < n_all_constraints_; ++r) (*M)(static_cast<long>(r), col_of_row_[r]) = 1.0;
  }
  // Row-normalised transpose of M (ref LIN.i:560-571).
  void getMpinv(Eigen::MatrixXd* M_pinv) const {
    CHECK_NOTNULL(M_pinv);
    const long n_cols = static_cast<long>(n_fixed_constraints_ + n_free_constraints_);
    M_pinv->resize(n_cols, static_cast<long>(n_all_constraints_));
    M_pinv->setZero();
    std::vector<int> hits(static_cast<size_t>(n_cols), 0);
    for (size_t r = 0; r < n_all_constraints_; ++r) ++hits[static_cast<size_t>(col_of_row_[r])];
    for (size_t r = 0; r < n_all_constraints_; ++r)
      (*M_pinv)(col_of_row_[r], static_cast<long>(r)) = 1.0 / hits[static_cast<size_t>(col_of_row_[r])];
  }
  // R = C^T blockdiag(H_i) C (ref LIN.i:297-326): the H_i come from the GPU, the scatter through
  // the index map is pure bookkeeping.
  void getR(Eigen::MatrixXd* R) const {
    CHECK_NOTNULL(R);
    const long n_cols = static_cast<long>(n_fixed_constraints_ + n_free_constraints_);
    std::vector<double> H(n_segments_ * N * N);
    gpu::check(minsnap_segment_matrices_host(static_cast<long>(n_segments_), N, derivative_to_optimize_,
                                             segment_times_.data(), nullptr, nullptr, nullptr, H.data()),
               "minsnap_segment_matrices_host");
    R->resize(n_cols, n_cols);
    R->setZero();
    for (size_t i = 0; i < n_segments_; ++i)
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c)
          (*R)(col_of_row_[i * N + r], col_of_row_[i * N + c]) += H[(i * N + r) * N + c];
  }

  void printReorderingMatrix(std::ostream& stream) const {
    Eigen::MatrixXd M;
    getM(&M);
    stream << "Mapping matrix:\n" << M << std::endl;
  }

 private:
  // True for the constraint structure createRandomVertices produces (ref src/vertex.cpp:59,71-76)
  // when the fast standard-mask entry point covers the problem (N = 10, snap, D <= 3).
  bool hasStandardStructure() const {
    if (N != 10 || derivative_to_optimize_ != N / 2 - 1 || dimension_ > 3 || n_segments_ < 2) return false;
    const size_t h = N / 2;
    for (size_t v = 0; v < n_vertices_; ++v)
      for (size_t c = 0; c < h; ++c) {
        const bool want = c == 0 || v == 0 || v + 1 == n_vertices_;
        if ((fixed_mask_[v * h + c] != 0) != want) return false;
      }
    return true;
  }

  // row-major [N][N] buffer of the C ABI -> matrix (independent of the matrix storage order)
  static void fillSquare(const double* buf, SquareMatrix* m) {
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) (*m)(r, c) = buf[r * N + c];
  }

  // Fixed/free partition and the reordering index map (ref LIN.i:171-250), computed on the GPU.
  void setupConstraintReordering() {
    const int h = N / 2;
    const int K = static_cast<int>(n_segments_);
    fixed_mask_.assign(n_vertices_ * h, 0);
    for (size_t v = 0; v < n_vertices_; ++v)
      for (int c = 0; c < h; ++c) fixed_mask_[v * h + c] = vertices_[v].hasConstraint(c) ? 1 : 0;
    col_of_row_.assign(static_cast<size_t>(N) * K, 0);
    int32_t counts[2] = {0, 0};
    gpu::check(minsnap_reorder_host(N, K, 1, fixed_mask_.data(), col_of_row_.data(), counts), "minsnap_reorder_host");
    n_fixed_constraints_ = static_cast<size_t>(counts[0]);
    n_free_constraints_ = static_cast<size_t>(counts[1]);
    n_all_constraints_ = static_cast<size_t>(N) * K;
    // d_f per dimension, in column order = (vertex, derivative)-sorted fixed constraints
    for (Eigen::VectorXd& df : fixed_constraints_compact_) df.resize(static_cast<long>(n_fixed_constraints_));
    long col = 0;
    for (size_t v = 0; v < n_vertices_; ++v)
      for (int c = 0; c < h; ++c) {
        Eigen::VectorXd value;
        if (!vertices_[v].getConstraint(c, &value)) continue;
        for (size_t d = 0; d < dimension_; ++d) fixed_constraints_compact_[d][col] = value[static_cast<long>(d)];
        ++col;
      }
    for (Eigen::VectorXd& dp : free_constraints_compact_) {
      dp.resize(static_cast<long>(n_free_constraints_));
      dp.setZero();
    }
  }

  // [D] vectors of length n -> [n][D] interleaved (C-ABI layout)
  std::vector<double> interleave(const std::vector<Eigen::VectorXd>& per_dim, size_t n) const {
    std::vector<double> out(n * dimension_);
    for (size_t d = 0; d < dimension_; ++d)
      for (size_t c = 0; c < n; ++c) out[c * dimension_ + d] = per_dim[d][static_cast<long>(c)];
    return out;
  }

  void updateSegmentsFromCompactConstraints() {
    const int K = static_cast<int>(n_segments_), D = static_cast<int>(dimension_);
    std::vector<double> fixed = interleave(fixed_constraints_compact_, n_fixed_constraints_);
    std::vector<double> free_values = interleave(free_constraints_compact_, n_free_constraints_);
    std::vector<double> coeffs(static_cast<size_t>(K) * D * N);
    gpu::check(minsnap_coeffs_from_constraints_host(1, K, D, N, fixed_mask_.data(), fixed.data(), free_values.data(),
                                                    segment_times_.data(), coeffs.data()),
               "minsnap_coeffs_from_constraints_host");
    storeCoefficients(coeffs);
  }

  void storeCoefficients(const std::vector<double>& coeffs) {
    const int D = static_cast<int>(dimension_);
    for (size_t i = 0; i < n_segments_; ++i) {
      Segment& segment = segments_[i];
      segment.setTime(segment_times_[i]);
      for (int d = 0; d < D; ++d) {
        Eigen::VectorXd c(N);
        for (int j = 0; j < N; ++j) c[j] = coeffs[(i * D + d) * N + j];
        segment[d] = Polynomial(N, c);
      }
    }
  }

  std::vector<double> packCoefficients() const {
    const int D = static_cast<int>(dimension_);
    std::vector<double> coeffs(n_segments_ * D * N);
    for (size_t i = 0; i < n_segments_; ++i)
      for (int d = 0; d < D; ++d) {
        const Eigen::VectorXd c = segments_[i][d].getCoefficients(0);
        for (int j = 0; j < N; ++j) coeffs[(i * D + d) * N + j] = c[j];
      }
    return coeffs;
  }

  // which = 0: A, 1: A^-1, per segment on the block diagonal
  void blockDiagonal(Eigen::MatrixXd* out, int which) const {
    CHECK_NOTNULL(out);
    std::vector<double> blocks(n_segments_ * N * N);
    gpu::check(minsnap_segment_matrices_host(static_cast<long>(n_segments_), N, 0, segment_times_.data(),
                                             which == 0 ? blocks.data() : nullptr, which == 1 ? blocks.data() : nullptr,
                                             nullptr, nullptr),
               "minsnap_segment_matrices_host");
    out->resize(static_cast<long>(N * n_segments_), static_cast<long>(N * n_segments_));
    out->setZero();
    for (size_t i = 0; i < n_segments_; ++i)
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) (*out)(static_cast<long>(i * N + r), static_cast<long>(i * N + c)) = blocks[(i * N + r) * N + c];
  }

  Vertex::Vector vertices_;
  Segment::Vector segments_;
  std::vector<uint8_t> fixed_mask_;     // [(K+1)][N/2], 1 = fixed
  std::vector<int32_t> col_of_row_;     // [N*K], the reordering matrix as an index map
  std::vector<Eigen::VectorXd> fixed_constraints_compact_;
  std::vector<Eigen::VectorXd> free_constraints_compact_;
  std::vector<double> segment_times_;
  size_t dimension_;
  int derivative_to_optimize_;
  size_t n_vertices_;
  size_t n_segments_;
  size_t n_all_constraints_;
  size_t n_fixed_constraints_;
  size_t n_free_constraints_;
};

// -------------------------------------------------------------------------------------------------
// Additive batched host API: B independent problems with the createRandomVertices constraint
// structure (ends fix derivatives 0..N/2-1, interior vertices fix position) in one call.
// positions [B][K+1][D], times [B][K] -> trajectories.  Thin wrapper over
// minsnap_solve_standard_host; use the C ABI directly to keep results on the device.
// -------------------------------------------------------------------------------------------------
template <int _N = 10>
class PolynomialOptimizationBatch {
 public:
  enum { N = _N };
  PolynomialOptimizationBatch(size_t dimension, size_t n_segments)
      : D_(static_cast<int>(dimension)), K_(static_cast<int>(n_segments)) {}

  // times empty: estimated on the device from (v_max, a_max, magic) like estimateSegmentTimes.
  bool solve(const std::vector<double>& positions, std::vector<double> times, double v_max = 0.0, double a_max = 0.0,
             double magic = 6.5, int derivative_to_optimize = N / 2 - 1) {
    CHECK_EQ(positions.size() % (static_cast<size_t>(K_ + 1) * D_), static_cast<size_t>(0));
    B_ = static_cast<long>(positions.size() / (static_cast<size_t>(K_ + 1) * D_));
    const bool have_times = !times.empty();
    if (have_times) CHECK_EQ(times.size(), static_cast<size_t>(B_) * K_);
    times_.assign(static_cast<size_t>(B_) * K_, 0.0);
    coeffs_.assign(static_cast<size_t>(B_) * K_ * D_ * N, 0.0);
    cost_.assign(static_cast<size_t>(B_), 0.0);
    status_.assign(static_cast<size_t>(B_), 0);
    gpu::check(minsnap_solve_standard_host(B_, K_, D_, N, derivative_to_optimize, positions.data(), nullptr,
                                           have_times ? times.data() : nullptr, v_max, a_max, magic, times_.data(),
                                           coeffs_.data(), nullptr, cost_.data(), status_.data()),
               "minsnap_solve_standard_host");
    return true;
  }

  long size() const { return B_; }
  double cost(long b) const { return cost_[static_cast<size_t>(b)]; }
  int32_t status(long b) const { return status_[static_cast<size_t>(b)]; }
  const std::vector<double>& coefficients() const { return coeffs_; }   // [B][K][D][N]
  const std::vector<double>& segmentTimes() const { return times_; }   // [B][K]

  // Largest magnitude of a derivative of every solved trajectory in one launch: what a loop over
  // PolynomialOptimization<N>::computeMaximumOfMagnitude<Derivative>(nullptr) would return (ref LIN.i:470-503).
  template <int Derivative>
  std::vector<Extremum> computeMaximumOfMagnitude() const {
    static_assert(N - Derivative - 1 > 0, "N-Derivative-1 has to be greater 0");
    std::vector<Extremum> result(static_cast<size_t>(B_));
    if (B_ == 0) return result;
    std::vector<double> time(static_cast<size_t>(B_)), value(static_cast<size_t>(B_));
    std::vector<int32_t> segment(static_cast<size_t>(B_));
    gpu::check(minsnap_extrema_host(B_, K_, D_, N, coeffs_.data(), times_.data(), Derivative,
                                    gpu::extremaMode(MINSNAP_EXTREMA_OPTIMIZATION), 0, time.data(), value.data(),
                                    segment.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
               "minsnap_extrema_host");
    for (size_t b = 0; b < result.size(); ++b) result[b] = Extremum(time[b], value[b], segment[b]);
    return result;
  }

  void getTrajectory(long b, Trajectory* trajectory) const {
    CHECK_NOTNULL(trajectory);
    Segment::Vector segments(static_cast<size_t>(K_), Segment(N, D_));
    for (int i = 0; i < K_; ++i) {
      segments[static_cast<size_t>(i)].setTime(times_[static_cast<size_t>(b) * K_ + i]);
      for (int d = 0; d < D_; ++d) {
        Eigen::VectorXd c(N);
        for (int j = 0; j < N; ++j) c[j] = coeffs_[((static_cast<size_t>(b) * K_ + i) * D_ + d) * N + j];
        segments[static_cast<size_t>(i)][static_cast<size_t>(d)] = Polynomial(N, c);
      }
    }
    trajectory->setSegments(segments);
  }

 private:
  int D_, K_;
  long B_ = 0;
  std::vector<double> times_, coeffs_, cost_;
  std::vector<int32_t> status_;
};

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_POLYNOMIAL_OPTIMIZATION_LINEAR_H_
