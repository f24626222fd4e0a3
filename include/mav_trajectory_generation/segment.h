// Segment: D polynomials sharing one duration (mirror of ref include/mav_trajectory_generation/
// segment.h:43-119, src/segment.cpp:27-199).
#ifndef MAV_TRAJECTORY_GENERATION_SEGMENT_H_
#define MAV_TRAJECTORY_GENERATION_SEGMENT_H_

#include <cmath>
#include <cstdint>
#include <limits>
#include <ostream>
#include <vector>

#include "mav_trajectory_generation/extremum.h"
#include "mav_trajectory_generation/motion_defines.h"
#include "mav_trajectory_generation/polynomial.h"

namespace mav_trajectory_generation {

constexpr double kNumNSecPerSec = 1.0e9;
constexpr double kNumSecPerNsec = 1.0e-9;

class Segment {
 public:
  typedef std::vector<Segment> Vector;

  Segment(int N, int D) : time_(0.0), N_(N), D_(D) { polynomials_.resize(D_, Polynomial(N_)); }
  Segment(const Segment& segment) = default;
  Segment& operator=(const Segment& segment) = default;

  bool operator==(const Segment& rhs) const {
    if (D_ != rhs.D_ || time_ != rhs.time_) return false;
    for (int i = 0; i < D_; ++i)
      if (polynomials_[i] != rhs.polynomials_[i]) return false;
    return true;
  }
  bool operator!=(const Segment& rhs) const { return !operator==(rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  double getTime() const { return time_; }
  uint64_t getTimeNSec() const { return static_cast<uint64_t>(kNumNSecPerSec * time_); }
  void setTime(double time_sec) { time_ = time_sec; }
  void setTimeNSec(uint64_t time_ns) { time_ = time_ns * kNumSecPerNsec; }

  Polynomial& operator[](size_t idx) {
    CHECK_LT(idx, static_cast<size_t>(D_));
    return polynomials_[idx];
  }
  const Polynomial& operator[](size_t idx) const {
    CHECK_LT(idx, static_cast<size_t>(D_));
    return polynomials_[idx];
  }
  const Polynomial::Vector& getPolynomialsRef() const { return polynomials_; }

  // D values of the requested derivative at time t inside the segment (ref src/segment.cpp:51-58),
  // evaluated on the GPU.  Like the reference, t is NOT clamped to [0, getTime()].
  Eigen::VectorXd evaluate(double t, int derivative = derivative_order::POSITION) const {
    Eigen::VectorXd result(D_);
    result.setZero();
    if (derivative >= N_) return result;
    std::vector<double> out = evaluateBatch(std::vector<double>(1, t), derivative + 1);
    for (int d = 0; d < D_; ++d) result[d] = out[static_cast<size_t>(derivative) * D_ + d];
    return result;
  }

  // Additive batched form: derivatives 0..n_deriv-1 at many instants in one launch;
  // returns [times.size()][n_deriv][D].
  std::vector<double> evaluateBatch(const std::vector<double>& times, int n_deriv) const {
    std::vector<double> coeffs(static_cast<size_t>(D_) * N_);
    for (int d = 0; d < D_; ++d)
      for (int j = 0; j < N_; ++j) coeffs[static_cast<size_t>(d) * N_ + j] = polynomials_[d].getCoefficients(0)[j];
    double span = 1.0;  // a single-segment trajectory long enough to contain every instant
    for (double t : times) span = std::max(span, 2.0 * std::fabs(t) + 1.0);
    std::vector<double> out(times.size() * static_cast<size_t>(n_deriv) * D_);
    gpu::check(minsnap_sample_at_host(1, 1, D_, N_, coeffs.data(), &span, static_cast<int>(times.size()),
                                      times.data(), 0, n_deriv, out.data(), nullptr),
               "minsnap_sample_at_host");
    return out;
  }

  // ---- extrema of the magnitude of a derivative over [t_start, t_end] (ref src/segment.cpp:82-199) ----
  // Candidate times: t_start, t_end, then the real roots inside the range of d/dt |p^(derivative)|^2
  // (one dimension: of p^(derivative+1)), ascending.  The roots are isolated on the GPU
  // (minsnap_extrema_host) instead of by the reference's Jenkins-Traub routine.  t_start >= 0.
  bool computeMinMaxMagnitudeCandidateTimes(int derivative, double t_start, double t_end,
                                            const std::vector<int>& dimensions,
                                            std::vector<double>* candidate_times) const {
    CHECK_NOTNULL(candidate_times);
    std::vector<Extremum> candidates;
    candidate_times->clear();
    if (!computeMinMaxMagnitudeCandidates(derivative, t_start, t_end, dimensions, &candidates)) return false;
    for (const Extremum& c : candidates) candidate_times->push_back(c.time);
    return true;
  }

  bool computeMinMaxMagnitudeCandidates(int derivative, double t_start, double t_end,
                                        const std::vector<int>& dimensions,
                                        std::vector<Extremum>* candidates) const {
    CHECK_NOTNULL(candidates);
    candidates->clear();
    if (dimensions.empty()) {
      LOG(WARNING) << "No dimensions specified." << std::endl;
      return false;
    }
    uint32_t mask = 0;
    if (!gpu::dimensionMask(dimensions, D_, &mask)) {
      LOG(WARNING) << "Specified dimensions are out of bounds [0.." << D_ - 1 << "]." << std::endl;
      return false;
    }
    if (N_ - derivative - 1 < 0) {
      LOG(WARNING) << "N - derivative - 1 has to be at least 0.";
      return false;
    }
    if (t_start > t_end) {
      LOG(WARNING) << "t_start is greater than t_end.";
      return false;
    }
    CHECK_GE(t_start, 0.0) << "the candidate search runs over [0, t_end]";
    std::vector<double> coeffs = packCoefficients();
    const int n_dims = __builtin_popcount(mask);
    const int max_roots = derivative <= N_ - 2 ? minsnap_extrema_max_roots(N_, derivative, n_dims) : 0;
    std::vector<double> times(max_roots + 2), values(max_roots + 2);
    int32_t n_roots = 0;
    if (derivative <= N_ - 2) {
      gpu::check(minsnap_extrema_host(1, 1, D_, N_, coeffs.data(), &t_end, derivative,
                                      gpu::extremaMode(MINSNAP_EXTREMA_TRAJECTORY), mask, nullptr, nullptr, nullptr,
                                      nullptr, nullptr, nullptr, times.data(), values.data(), &n_roots),
                 "minsnap_extrema_host");
    } else {
      times[1] = t_end;   // the highest derivative is a constant: no roots, the range ends only
      values[0] = values[1] = magnitude(t_end, derivative, dimensions);
    }
    candidates->reserve(static_cast<size_t>(n_roots) + 2);
    candidates->emplace_back(t_start, t_start == 0.0 ? values[0] : magnitude(t_start, derivative, dimensions), 0);
    candidates->emplace_back(t_end, values[1], 0);
    for (int i = 0; i < n_roots; ++i)
      if (times[2 + i] >= t_start) candidates->emplace_back(times[2 + i], values[2 + i], 0);
    return true;
  }

  bool selectMinMaxMagnitudeFromCandidates(double t_start, double t_end, int derivative,
                                           const std::vector<int>& dimensions,
                                           const std::vector<Extremum>& candidates, Extremum* minimum,
                                           Extremum* maximum) const {
    CHECK_NOTNULL(minimum);
    CHECK_NOTNULL(maximum);
    if (t_start > t_end) {
      LOG(WARNING) << "t_start is greater than t_end.";
      return false;
    }
    minimum->value = std::numeric_limits<double>::max();
    maximum->value = std::numeric_limits<double>::lowest();
    for (const Extremum& candidate : candidates) {
      if (candidate.time < t_start || candidate.time > t_end) continue;
      if (*maximum < candidate) *maximum = candidate;
      if (candidate < *minimum) *minimum = candidate;
    }
    const Extremum ends[2] = {Extremum(t_start, magnitude(t_start, derivative, dimensions), 0),
                              Extremum(t_end, magnitude(t_end, derivative, dimensions), 0)};
    for (const Extremum& e : ends) {
      if (*maximum < e) *maximum = e;
      if (e < *minimum) *minimum = e;
    }
    return true;
  }

  // coefficients [D][N] in the C-ABI layout
  std::vector<double> packCoefficients() const {
    std::vector<double> coeffs(static_cast<size_t>(D_) * N_);
    for (int d = 0; d < D_; ++d) {
      const Eigen::VectorXd c = polynomials_[d].getCoefficients(0);
      for (int j = 0; j < N_; ++j) coeffs[static_cast<size_t>(d) * N_ + j] = c[j];
    }
    return coeffs;
  }

 protected:
  // sqrt of the sum over the listed dimensions of p^(derivative)(t)^2 (evaluated on the GPU)
  double magnitude(double t, int derivative, const std::vector<int>& dimensions) const {
    const Eigen::VectorXd v = evaluate(t, derivative);
    double s = 0.0;
    for (int dim : dimensions) s += v[dim] * v[dim];
    return std::sqrt(s);
  }

  Polynomial::Vector polynomials_;
  double time_;

 private:
  int N_;
  int D_;
};

inline void printSegment(std::ostream& stream, const Segment& s, int derivative) {
  CHECK(derivative >= 0 && derivative < s.N());
  stream << "t: " << s.getTime() << std::endl;
  stream << " coefficients for " << positionDerivativeToString(derivative) << ": " << std::endl;
  for (int i = 0; i < s.D(); ++i) stream << s[i].getCoefficients(derivative) << std::endl;
}

inline std::ostream& operator<<(std::ostream& stream, const Segment& s) {
  printSegment(stream, s, derivative_order::POSITION);
  return stream;
}

inline std::ostream& operator<<(std::ostream& stream, const std::vector<Segment>& segments) {
  for (const Segment& s : segments) stream << s << std::endl;
  return stream;
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_SEGMENT_H_
