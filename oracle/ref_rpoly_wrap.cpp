// TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's own Jenkins-Traub root finder, compiled together with
// /root/reference/mav_trajectory_generation/src/rpoly.cpp (never copied into this repository)
// into oracle/_ref/librpoly_ref.so.  Used by tests/ to pin the oracle's real-root finder and the
// GPU extrema kernel against the roots the reference itself computes.
// The reference keeps its working storage in namespace-scope globals (src/rpoly.cpp:120-125):
// these entry points are NOT re-entrant; call them from one thread.
#include <complex>

#include "mav_trajectory_generation/rpoly.h"

extern "C" {

// ref: int findRootsJenkinsTraub(const double*, int, double*, double*, int[]) (rpoly.h:41-42).
__attribute__((visibility("default"))) int ref_rpoly_decreasing(const double* coefficients_decreasing, int degree,
                                                                double* roots_real, double* roots_imag) {
  return mav_trajectory_generation::findRootsJenkinsTraub(coefficients_decreasing, degree, roots_real, roots_imag,
                                                          nullptr);
}

// ref: bool findRootsJenkinsTraub(const Eigen::VectorXd& increasing, Eigen::VectorXcd*) (src/rpoly.cpp:57-99),
// including its removal of trailing coefficients below machine epsilon.  Returns the number of
// roots, or -1 when the reference reports failure.
__attribute__((visibility("default"))) int ref_rpoly_increasing(const double* coefficients_increasing, int n,
                                                                double* roots_real, double* roots_imag) {
  Eigen::VectorXd c(n);
  for (int i = 0; i < n; ++i) c[i] = coefficients_increasing[i];
  Eigen::VectorXcd roots;
  if (!mav_trajectory_generation::findRootsJenkinsTraub(c, &roots)) return -1;
  for (std::size_t i = 0; i < roots.size(); ++i) {
    roots_real[i] = roots[i].real();
    roots_imag[i] = roots[i].imag();
  }
  return (int)roots.size();
}

}  // extern "C"
