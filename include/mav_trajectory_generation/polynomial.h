// Polynomial value type of the hot path (mirror of ref include/mav_trajectory_generation/
// polynomial.h:34-242, src/polynomial.cpp:27-175; computeRoots / the *FromRoots members, which hand
// complex roots around, are not provided).  Coefficients are
// stored with increasing powers: c_0 + c_1 t + ... + c_{N-1} t^{N-1}.
//
// evaluate() runs on the GPU through the C ABI (minsnap_sample_at_host, rows a17 of SURVEY.md
// section 8); there is no host Horner.  For many instants use Trajectory::evaluateBatch or the
// batched C-ABI entry points -- one call per instant pays a kernel launch each time.
#ifndef MAV_TRAJECTORY_GENERATION_POLYNOMIAL_H_
#define MAV_TRAJECTORY_GENERATION_POLYNOMIAL_H_

#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>
#include <vector>

#include "mav_trajectory_generation/minsnap_gpu.h"

namespace mav_trajectory_generation {

// b(d, j) = j!/(j-d)!  (ref computeBaseCoefficients, src/polynomial.cpp:140-155): how often
// t^j survives d differentiations.  Table lookups of the reference become this closed form.
inline double baseCoefficient(int derivative, int power) {
  if (power < derivative) return 0.0;
  double b = 1.0;
  for (int q = 0; q < derivative; ++q) b *= static_cast<double>(power - q);
  return b;
}

class Polynomial {
 public:
  typedef std::vector<Polynomial> Vector;

  static constexpr int kMaxN = 12;
  static constexpr int kMaxConvolutionSize = 2 * kMaxN - 2;

  explicit Polynomial(int N) : N_(N), coefficients_(N) { coefficients_.setZero(); }
  Polynomial(int N, const Eigen::VectorXd& coeffs) : N_(N), coefficients_(coeffs) {
    CHECK_EQ(N_, static_cast<int>(coeffs.size())) << "Number of coefficients has to match.";
  }
  explicit Polynomial(const Eigen::VectorXd& coeffs) : N_(static_cast<int>(coeffs.size())), coefficients_(coeffs) {}

  int N() const { return N_; }

  bool operator==(const Polynomial& rhs) const { return coefficients_ == rhs.coefficients_; }
  bool operator!=(const Polynomial& rhs) const { return !operator==(rhs); }
  Polynomial operator+(const Polynomial& rhs) const { return Polynomial(coefficients_ + rhs.coefficients_); }
  Polynomial& operator+=(const Polynomial& rhs) {
    coefficients_ += rhs.coefficients_;
    return *this;
  }
  Polynomial operator*(const double& rhs) const { return Polynomial(coefficients_ * rhs); }

  void setCoefficients(const Eigen::VectorXd& coeffs) {
    CHECK_EQ(N_, static_cast<int>(coeffs.size())) << "Number of coefficients has to match.";
    coefficients_ = coeffs;
  }

  // Coefficients of the derivative-th derivative, padded with zeros to length N
  // (ref polynomial.h:100-115).
  Eigen::VectorXd getCoefficients(int derivative = 0) const {
    CHECK_LE(derivative, N_);
    if (derivative == 0) return coefficients_;
    Eigen::VectorXd result(N_);
    result.setZero();
    for (int j = derivative; j < N_; ++j) result[j - derivative] = baseCoefficient(derivative, j) * coefficients_[j];
    return result;
  }

  // Derivatives 0 .. result->size()-1 at time t (ref polynomial.h:120-136).
  void evaluate(double t, Eigen::VectorXd* result) const {
    CHECK_NOTNULL(result);
    CHECK_LE(static_cast<int>(result->size()), N_);
    sample(t, static_cast<int>(result->size()), result->data());
  }

  // One derivative at time t; zero when derivative >= N (ref polynomial.h:138-151).
  double evaluate(double t, int derivative) const {
    if (derivative >= N_) return 0.0;
    std::vector<double> out(static_cast<size_t>(derivative) + 1);
    sample(t, derivative + 1, out.data());
    return out[static_cast<size_t>(derivative)];
  }

  // Row of the mapping matrix: c[j] = b(d,j) t^(j-d)  (ref polynomial.h:215-242).  Plain data
  // marshalling for callers that build constraint rows; the solver itself never forms A.
  static void baseCoeffsWithTime(int N, int derivative, double t, Eigen::VectorXd* coeffs) {
    CHECK_LT(derivative, N);
    CHECK_GE(derivative, 0);
    coeffs->resize(N, 1);
    coeffs->setZero();
    (*coeffs)[derivative] = baseCoefficient(derivative, derivative);
    if (std::abs(t) < std::numeric_limits<double>::epsilon()) return;
    double t_power = t;
    for (int j = derivative + 1; j < N; ++j) {
      (*coeffs)[j] = baseCoefficient(derivative, j) * t_power;
      t_power = t_power * t;
    }
  }
  static Eigen::VectorXd baseCoeffsWithTime(int N, int derivative, double t) {
    Eigen::VectorXd c(N);
    baseCoeffsWithTime(N, derivative, t, &c);
    return c;
  }

  static inline int getConvolutionLength(int data_size, int kernel_size) { return data_size + kernel_size - 1; }

  // Discrete convolution = coefficients of the product polynomial (ref src/polynomial.cpp:157-175, same
  // summation order).  A handful of operations on caller data: computed where it is called.
  static Eigen::VectorXd convolve(const Eigen::VectorXd& data, const Eigen::VectorXd& kernel) {
    const int n_data = static_cast<int>(data.size()), n_kernel = static_cast<int>(kernel.size());
    Eigen::VectorXd out(getConvolutionLength(n_data, n_kernel));
    out.setZero();
    for (int i = 0; i < static_cast<int>(out.size()); ++i)
      for (int j = std::min(n_kernel - 1, i); j >= std::max(0, i - (n_data - 1)); --j) out[i] += kernel[j] * data[i - j];
    return out;
  }
  Polynomial operator*(const Polynomial& rhs) const { return Polynomial(convolve(coefficients_, rhs.coefficients_)); }

  // ---- extrema of one derivative over [t_start, t_end] (ref src/polynomial.cpp:57-138) --------------
  // Candidates: t_start, t_end, then the real roots of the next derivative inside the range (ascending),
  // isolated on the GPU (minsnap_extrema_host) instead of by Jenkins-Traub.  0 <= t_start <= t_end.
  bool computeMinMaxCandidates(double t_start, double t_end, int derivative, std::vector<double>* candidates) const {
    CHECK_NOTNULL(candidates);
    candidates->clear();
    if (N_ - derivative - 1 < 0) {
      LOG(WARNING) << "N - derivative - 1 has to be at least 0.";
      return false;
    }
    if (t_start > t_end) {
      LOG(WARNING) << "t_start is greater than t_end.";
      return false;
    }
    CHECK_GE(t_start, 0.0) << "the candidate search runs over [0, t_end]";
    candidates->push_back(t_start);
    candidates->push_back(t_end);
    const int Np = minsnapSupportedN(N_) ? N_ : paddedN(N_);
    CHECK(Np > 0) << "Polynomial::computeMinMaxCandidates: unsupported number of coefficients " << N_;
    if (derivative > Np - 2) return true;   // the next derivative vanishes identically: no roots
    std::vector<double> c(static_cast<size_t>(Np), 0.0);
    for (int j = 0; j < N_; ++j) c[static_cast<size_t>(j)] = coefficients_[j];
    const int max_roots = minsnap_extrema_max_roots(Np, derivative, 1);
    std::vector<double> times(static_cast<size_t>(max_roots) + 2);
    int32_t n_roots = 0;
    gpu::check(minsnap_extrema_host(1, 1, 1, Np, c.data(), &t_end, derivative,
                                    gpu::extremaMode(MINSNAP_EXTREMA_TRAJECTORY), 1u, nullptr, nullptr, nullptr,
                                    nullptr, nullptr, nullptr, times.data(), nullptr, &n_roots),
               "minsnap_extrema_host");
    for (int i = 0; i < n_roots; ++i)
      if (times[static_cast<size_t>(2 + i)] >= t_start) candidates->push_back(times[static_cast<size_t>(2 + i)]);
    return true;
  }

  // Smallest and largest (signed) value of the derivative among the candidates: (time, value) pairs.
  bool selectMinMaxFromCandidates(const std::vector<double>& candidates, int derivative,
                                  std::pair<double, double>* minimum, std::pair<double, double>* maximum) const {
    CHECK_NOTNULL(minimum);
    CHECK_NOTNULL(maximum);
    if (candidates.empty()) {
      LOG(WARNING) << "Cannot find extrema from an empty candidates vector.";
      return false;
    }
    minimum->first = maximum->first = candidates[0];
    minimum->second = std::numeric_limits<double>::max();
    maximum->second = std::numeric_limits<double>::lowest();
    for (const double t : candidates) {
      const double value = evaluate(t, derivative);
      if (value < minimum->second) *minimum = std::make_pair(t, value);
      if (value > maximum->second) *maximum = std::make_pair(t, value);
    }
    return true;
  }

  bool computeMinMax(double t_start, double t_end, int derivative, std::pair<double, double>* minimum,
                     std::pair<double, double>* maximum) const {
    std::vector<double> candidates;
    if (!computeMinMaxCandidates(t_start, t_end, derivative, &candidates)) return false;
    return selectMinMaxFromCandidates(candidates, derivative, minimum, maximum);
  }

 private:
  // GPU evaluation of derivatives 0..n_deriv-1: a single-segment, single-dimension trajectory
  // whose duration safely contains t.
  void sample(double t, int n_deriv, double* out) const {
    if (!minsnapSupportedN(N_)) {
      // orders the kernels are not built for: pad to the next built order (exact: zero tail)
      const int Np = paddedN(N_);
      CHECK(Np > 0) << "Polynomial::evaluate: unsupported number of coefficients " << N_;
      Eigen::VectorXd padded(Np);
      padded.setZero();
      for (int j = 0; j < N_; ++j) padded[j] = coefficients_[j];
      Polynomial(Np, padded).sample(t, n_deriv, out);
      return;
    }
    const double duration = 2.0 * std::fabs(t) + 1.0;
    gpu::check(minsnap_sample_at_host(1, 1, 1, N_, coefficients_.data(), &duration, 1, &t, 0, n_deriv, out, nullptr),
               "minsnap_sample_at_host");
  }
  static bool minsnapSupportedN(int N) { return N == 4 || N == 6 || N == 8 || N == 10 || N == 12; }
  static int paddedN(int N) {
    for (int c : {4, 6, 8, 10, 12})
      if (N <= c) return c;
    return 0;
  }

  int N_;
  Eigen::VectorXd coefficients_;
};

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_POLYNOMIAL_H_
