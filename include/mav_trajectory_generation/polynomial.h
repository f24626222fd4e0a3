// Polynomial value type of the hot path (mirror of ref include/mav_trajectory_generation/
// polynomial.h:34-151, 215-242; the root-finding members are out of scope).  Coefficients are
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
