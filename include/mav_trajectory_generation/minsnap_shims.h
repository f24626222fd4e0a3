// Minimal stand-ins for the two third-party libraries whose types and macros appear in the
// reference's public headers (Eigen3: vertex.h:45,64,85, segment.h:71, trajectory.h:86-94,
// LIN.h:180-214; glog: CHECK*/LOG everywhere).  Neither library is installed in this image.
// When the real headers are available they are used instead and nothing below is defined.
#ifndef MAV_TRAJECTORY_GENERATION_MINSNAP_SHIMS_H_
#define MAV_TRAJECTORY_GENERATION_MINSNAP_SHIMS_H_

#include <cmath>
#include <cstdlib>
#include <initializer_list>
#include <iostream>
#include <sstream>
#include <vector>

// ---------------------------------------------------------------------------------------
// glog
// ---------------------------------------------------------------------------------------
#if defined(MINSNAP_USE_GLOG) || (defined(__has_include) && __has_include(<glog/logging.h>))
#include <glog/logging.h>
#else
namespace minsnap_shim {
class LogMessage {
 public:
  LogMessage(const char* severity, const char* file, int line, bool fatal) : fatal_(fatal) {
    stream_ << severity << " " << file << ":" << line << "] ";
  }
  ~LogMessage() {
    std::cerr << stream_.str() << std::endl;
    if (fatal_) std::abort();
  }
  std::ostream& stream() { return stream_; }

 private:
  std::ostringstream stream_;
  bool fatal_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
template <typename T>
T* CheckNotNull(const char* file, int line, const char* expr, T* ptr) {
  if (ptr == nullptr) LogMessage("F", file, line, true).stream() << "Check failed: '" << expr << "' must be non NULL";
  return ptr;
}
}  // namespace minsnap_shim
#define MINSNAP_LOG_INFO ::minsnap_shim::LogMessage("I", __FILE__, __LINE__, false).stream()
#define MINSNAP_LOG_WARNING ::minsnap_shim::LogMessage("W", __FILE__, __LINE__, false).stream()
#define MINSNAP_LOG_ERROR ::minsnap_shim::LogMessage("E", __FILE__, __LINE__, false).stream()
#define MINSNAP_LOG_FATAL ::minsnap_shim::LogMessage("F", __FILE__, __LINE__, true).stream()
#define LOG(severity) MINSNAP_LOG_##severity
#define CHECK(cond) \
  (cond) ? (void)0 : ::minsnap_shim::Voidify() & ::minsnap_shim::LogMessage("F", __FILE__, __LINE__, true).stream() << "Check failed: " #cond " "
#define MINSNAP_CHECK_OP(a, b, op) CHECK((a)op(b)) << "(" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) MINSNAP_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) MINSNAP_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) MINSNAP_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) MINSNAP_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) MINSNAP_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) MINSNAP_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(ptr) ::minsnap_shim::CheckNotNull(__FILE__, __LINE__, #ptr, (ptr))
#endif

// ---------------------------------------------------------------------------------------
// Eigen
// ---------------------------------------------------------------------------------------
#if defined(MINSNAP_USE_EIGEN) || (defined(__has_include) && (__has_include(<Eigen/Core>) || __has_include(<eigen3/Eigen/Core>)))
#if defined(__has_include) && __has_include(<eigen3/Eigen/Core>)
#include <eigen3/Eigen/Core>
#else
#include <Eigen/Core>
#endif
#else
#define MINSNAP_EIGEN_SHIM 1
namespace Eigen {

const int Dynamic = -1;

// Dense column-major-agnostic matrix of doubles with just the operations the hot path's API
// and its tests use.  Storage is row-major; element access is (row, col).
class MatrixXd {
 public:
  MatrixXd() : rows_(0), cols_(0) {}
  MatrixXd(long rows, long cols) : rows_(rows), cols_(cols), v_(static_cast<size_t>(rows * cols), 0.0) {}
  long rows() const { return rows_; }
  long cols() const { return cols_; }
  long size() const { return rows_ * cols_; }
  void resize(long rows, long cols) {
    rows_ = rows;
    cols_ = cols;
    v_.assign(static_cast<size_t>(rows * cols), 0.0);
  }
  void setZero() { v_.assign(v_.size(), 0.0); }
  double& operator()(long r, long c) { return v_[static_cast<size_t>(r * cols_ + c)]; }
  double operator()(long r, long c) const { return v_[static_cast<size_t>(r * cols_ + c)]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  MatrixXd transpose() const {
    MatrixXd t(cols_, rows_);
    for (long r = 0; r < rows_; ++r)
      for (long c = 0; c < cols_; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
  MatrixXd operator*(const MatrixXd& o) const {
    MatrixXd out(rows_, o.cols_);
    for (long r = 0; r < rows_; ++r)
      for (long k = 0; k < cols_; ++k) {
        const double a = (*this)(r, k);
        if (a == 0.0) continue;
        for (long c = 0; c < o.cols_; ++c) out(r, c) += a * o(k, c);
      }
    return out;
  }
  MatrixXd operator-(const MatrixXd& o) const {
    MatrixXd out(rows_, cols_);
    for (size_t i = 0; i < v_.size(); ++i) out.v_[i] = v_[i] - o.v_[i];
    return out;
  }
  double maxAbs() const {
    double m = 0.0;
    for (double x : v_) m = std::fabs(x) > m ? std::fabs(x) : m;
    return m;
  }

 protected:
  long rows_, cols_;
  std::vector<double> v_;
};

class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(long n) : v_(static_cast<size_t>(n), 0.0) {}
  VectorXd(std::initializer_list<double> init) : v_(init) {}
  static VectorXd Constant(long n, double value) {
    VectorXd r(n);
    r.v_.assign(static_cast<size_t>(n), value);
    return r;
  }
  static VectorXd Zero(long n, long = 1) { return VectorXd(n); }
  long size() const { return static_cast<long>(v_.size()); }
  long rows() const { return size(); }
  long cols() const { return 1; }
  void resize(long n, long = 1) { v_.assign(static_cast<size_t>(n), 0.0); }
  void setZero() { v_.assign(v_.size(), 0.0); }
  double& operator[](long i) { return v_[static_cast<size_t>(i)]; }
  double operator[](long i) const { return v_[static_cast<size_t>(i)]; }
  double& operator()(long i) { return v_[static_cast<size_t>(i)]; }
  double operator()(long i) const { return v_[static_cast<size_t>(i)]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  bool operator==(const VectorXd& o) const { return v_ == o.v_; }
  bool operator!=(const VectorXd& o) const { return !(v_ == o.v_); }
  VectorXd operator-(const VectorXd& o) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[i] - o.v_[i];
    return r;
  }
  VectorXd operator+(const VectorXd& o) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[i] + o.v_[i];
    return r;
  }
  VectorXd operator*(double s) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[i] * s;
    return r;
  }
  VectorXd& operator+=(const VectorXd& o) {
    for (long i = 0; i < size(); ++i) v_[i] += o.v_[i];
    return *this;
  }
  double squaredNorm() const {
    double s = 0.0;
    for (double x : v_) s += x * x;
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  double maxAbs() const {
    double m = 0.0;
    for (double x : v_) m = std::fabs(x) > m ? std::fabs(x) : m;
    return m;
  }
  bool isZero(double tol) const { return maxAbs() <= tol; }
  VectorXd head(long n) const {
    VectorXd r(n);
    for (long i = 0; i < n; ++i) r[i] = v_[i];
    return r;
  }
  VectorXd tail(long n) const {
    VectorXd r(n);
    for (long i = 0; i < n; ++i) r[i] = v_[size() - n + i];
    return r;
  }
  // comma initialiser:  v << 1, 2, 3;
  class CommaInit {
   public:
    CommaInit(VectorXd* v, double first) : v_(v), i_(0) { (*v_)[i_++] = first; }
    CommaInit& operator,(double x) {
      (*v_)[i_++] = x;
      return *this;
    }

   private:
    VectorXd* v_;
    long i_;
  };
  CommaInit operator<<(double first) { return CommaInit(this, first); }

 private:
  std::vector<double> v_;
};

inline VectorXd operator*(const MatrixXd& m, const VectorXd& x) {
  VectorXd out(m.rows());
  for (long r = 0; r < m.rows(); ++r) {
    double acc = 0.0;
    for (long c = 0; c < m.cols(); ++c) acc += m(r, c) * x[c];
    out[r] = acc;
  }
  return out;
}

inline std::ostream& operator<<(std::ostream& s, const VectorXd& v) {
  for (long i = 0; i < v.size(); ++i) s << (i ? " " : "") << v[i];
  return s;
}
inline std::ostream& operator<<(std::ostream& s, const MatrixXd& m) {
  for (long r = 0; r < m.rows(); ++r) {
    for (long c = 0; c < m.cols(); ++c) s << (c ? " " : "") << m(r, c);
    s << "\n";
  }
  return s;
}

// Fixed-size matrix (only the square double case PolynomialOptimization<N>::SquareMatrix needs).
template <typename Scalar, int Rows, int Cols>
class Matrix : public MatrixXd {
 public:
  Matrix() : MatrixXd(Rows, Cols) {}
  Matrix(const MatrixXd& m) : MatrixXd(m) {}
};

template <typename T>
class aligned_allocator : public std::allocator<T> {
 public:
  template <typename U>
  struct rebind {
    typedef aligned_allocator<U> other;
  };
};

}  // namespace Eigen
#endif  // Eigen

#endif  // MAV_TRAJECTORY_GENERATION_MINSNAP_SHIMS_H_
