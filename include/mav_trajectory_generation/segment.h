// Segment: D polynomials sharing one duration (mirror of ref include/mav_trajectory_generation/
// segment.h:43-119, src/segment.cpp:27-80; the extremum members are out of scope).
#ifndef MAV_TRAJECTORY_GENERATION_SEGMENT_H_
#define MAV_TRAJECTORY_GENERATION_SEGMENT_H_

#include <cstdint>
#include <ostream>
#include <vector>

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

 protected:
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
