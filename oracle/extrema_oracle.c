/*
 * extrema_oracle.c -- TEST INFRASTRUCTURE ONLY (same rules as minsnap_oracle.c: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load it).
 *
 * CPU restatement of the reference's "extrema of the magnitude of a derivative" path
 * (SURVEY.md section 8 (f) 1).  Citations relative to /root/reference/mav_trajectory_generation/:
 *   LIN.i = include/mav_trajectory_generation/impl/polynomial_optimization_linear_impl.h
 *   - derivative coefficients           include/mav_trajectory_generation/polynomial.h:100-117
 *   - Polynomial::convolve              src/polynomial.cpp:157-175
 *   - candidate polynomial, D > 1       LIN.i:389-408, src/segment.cpp:91-117
 *   - candidate polynomial, D == 1      LIN.i:412-416, src/segment.cpp:124-129, src/polynomial.cpp:57-78
 *   - trailing-coefficient removal      src/rpoly.cpp:44-75 (absolute threshold: machine epsilon)
 *   - real roots inside the range       LIN.i:423-434, src/polynomial.cpp:27-55
 *   - maximum over a trajectory         LIN.i:470-503 (mode 0)
 *   - minimum and maximum               src/trajectory.cpp:181-217, src/segment.cpp:133-199 (mode 1)
 *
 * ROOT FINDER.  The reference finds the roots with a C translation of TOMS 493 (Jenkins-Traub,
 * src/rpoly.cpp:127-820) and keeps those whose imaginary part is at most machine epsilon.  This
 * oracle does NOT restate Jenkins-Traub: the real roots of a polynomial inside an interval are a
 * mathematical object, and the oracle computes them by isolating them between the critical points
 * of the polynomial (the real roots of its derivative, found recursively down to the linear
 * derivative) and refining each bracket with a safeguarded Newton iteration, in long double
 * arithmetic whatever the build.  It is PINNED against the reference's own root finder: the
 * reference translation unit src/rpoly.cpp is compiled where it lies (oracle/Makefile, target
 * _ref/librpoly_ref.so, with a stand-in for the few Eigen vector operations its wrappers use)
 * and tests/test_extrema_oracle.py compares root sets and extremum values on the reference's
 * test seeds; it is also checked the way the reference's own tests check extrema (against dense
 * sampling, test/test_polynomial_optimization.cpp:418-507 and test/test_polynomial.cpp:79-128).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef ORC_LONG_DOUBLE
typedef long double real;
#define R_SQRT sqrtl
#define R_FABS fabsl
#else
typedef double real;
#define R_SQRT sqrt
#define R_FABS fabs
#endif
typedef long double xreal; /* the root finder always works in extended precision */

#define API __attribute__((visibility("default")))
#define EXT_MAX_N 24                 /* polynomial coefficients per dimension */
#define EXT_MAX_G (2 * EXT_MAX_N)    /* coefficients of the candidate polynomial */

/* b(k, j) = j (j-1) ... (j-k+1): row k of the reference's base table, src/polynomial.cpp:140-155 */
static real falling(int k, int j) {
  real r = 1;
  for (int q = 0; q < k; ++q) r *= (real)(j - q);
  return j >= k ? r : 0;
}

/* polynomial.h:100-117: coefficients of the k-th derivative, zero padded to N entries */
static void derivative_coefficients(int N, const real* c, int k, real* out) {
  for (int j = 0; j < N; ++j) out[j] = 0;
  for (int j = 0; j + k < N; ++j) out[j] = c[j + k] * falling(k, j + k);
}

/* src/polynomial.cpp:157-175: out[i] = sum_j kernel[j] data[i-j], j from high to low */
static void convolve(const real* data, int n_data, const real* kernel, int n_kernel, real* out) {
  const int n_out = n_data + n_kernel - 1;
  for (int i = 0; i < n_out; ++i) {
    const int data_idx = i - n_kernel + 1;
    const int lower = data_idx < 0 ? -data_idx : 0;
    const int upper = n_kernel < n_data - data_idx ? n_kernel : n_data - data_idx;
    real acc = 0;
    for (int kernel_idx = lower; kernel_idx < upper; ++kernel_idx)
      acc += kernel[n_kernel - 1 - kernel_idx] * data[data_idx + kernel_idx];
    out[i] = acc;
  }
}

/* Horner evaluation of a derivative, polynomial.h:138-151 */
static real poly_eval(int N, const real* c, real t, int k) {
  if (k >= N) return 0;
  real r = falling(k, N - 1) * c[N - 1];
  for (int j = N - 2; j >= k; --j) {
    r *= t;
    r += falling(k, j) * c[j];
  }
  return r;
}

/* The polynomial whose real roots are the candidate times of one segment.
 * seg_coeffs [D][N]; dims: the dimensions that take part.  Returns the number of coefficients
 * written to g (increasing powers), before the removal of trailing coefficients. */
API int orc_candidate_polynomial(int N, int D, const real* seg_coeffs, int derivative, const int* dims,
                                 int n_dims, real* g) {
  (void)D;
  const int n_d = N - derivative, n_dd = n_d - 1;
  if (n_dims > 1) {
    const int len = n_d + n_dd - 1;
    real d[EXT_MAX_N], dd[EXT_MAX_N], part[EXT_MAX_G];
    for (int i = 0; i < len; ++i) g[i] = 0;
    for (int q = 0; q < n_dims; ++q) {
      const real* c = seg_coeffs + (size_t)dims[q] * N;
      derivative_coefficients(N, c, derivative, d);
      derivative_coefficients(N, c, derivative + 1, dd);
      convolve(d, n_d, dd, n_dd, part);
      for (int i = 0; i < len; ++i) g[i] += part[i];
    }
    return len;
  }
  real dd[EXT_MAX_N];
  derivative_coefficients(N, seg_coeffs + (size_t)dims[0] * N, derivative + 1, dd);
  for (int i = 0; i < n_dd; ++i) g[i] = dd[i];
  return n_dd;
}

/* src/rpoly.cpp:44-55: index of the last coefficient whose magnitude reaches machine epsilon.
 * NOTE (reference behaviour, reproduced on purpose): the threshold is ABSOLUTE.  A min-snap segment
 * of duration T has g_i ~ L^2 / T^(3+i), so for T above roughly 12 s the highest coefficients of g
 * fall below 2.2e-16 although g_i T^i is of the size of g itself; the reference then solves a
 * truncated polynomial and misses or misplaces extrema near the end of long segments
 * (tests/test_extrema_oracle.py::test_reference_truncation_misses_extrema shows a case).
 * keep_small != 0 removes exact zeros only (what a caller who wants the true extrema asks for). */
static int last_nonzero(const real* c, int n, int keep_small) {
  for (int i = n - 1; i >= 0; --i)
    if (keep_small ? (c[i] != 0) : (R_FABS(c[i]) >= (real)DBL_EPSILON)) return i;
  return -1;
}
API int orc_last_nonzero_coefficient(const real* c, int n) { return last_nonzero(c, n, 0); }

/* ---- real roots of a polynomial inside [t0, t1] ----------------------------------------- */
/* value, slope and a running bound of the rounding error of the value */
static void eval_pair(const xreal* d, int deg, xreal t, xreal* f, xreal* fp, xreal* bound) {
  const xreal at = fabsl(t);
  xreal a = d[deg], b = 0, e = fabsl(a);
  for (int j = deg - 1; j >= 0; --j) {
    b = b * t + a;
    a = a * t + d[j];
    e = e * at + fabsl(a);
  }
  *f = a;
  *fp = b;
  if (bound) *bound = 2 * LDBL_EPSILON * e;
}

static xreal refine_bracket(const xreal* d, int deg, xreal lo, xreal hi, int lo_negative) {
  xreal x = 0.5L * (lo + hi);
  for (int it = 0; it < 300; ++it) {
    xreal f, fp, bound;
    eval_pair(d, deg, x, &f, &fp, &bound);
    if (fabsl(f) <= bound) return x; /* inside the rounding error of the evaluation */
    if ((f < 0) == lo_negative) lo = x; else hi = x;
    xreal next = fp != 0 ? x - f / fp : lo - 1;
    if (!(next > lo && next < hi)) next = 0.5L * (lo + hi);
    if (next == lo || next == hi) return next;
    if (fabsl(next - x) <= 2 * LDBL_EPSILON * fabsl(next)) return next;
    x = next;
  }
  return x;
}

/* g: n coefficients (increasing powers), leading one non-zero.  Writes the real roots inside
 * [t0, t1] in ascending order and returns how many (at most n-1). */
API int orc_real_roots_in_range(const real* g, int n, real t0_in, real t1_in, real* roots_out) {
  const int deg = n - 1;
  if (deg < 1) return 0;
  const xreal t0 = t0_in, t1 = t1_in;
  if (t0 > t1) return 0;
  xreal prev[EXT_MAX_G], cur[EXT_MAX_G], d[EXT_MAX_G];
  int n_prev = 0;
  for (int m = deg - 1; m >= 0; --m) {
    /* d = coefficients of g^(m)/m!: C(j+m, m) g[j+m] */
    const int dm = deg - m;
    for (int j = 0; j <= dm; ++j) {
      xreal binom = 1;
      for (int q = 1; q <= m; ++q) binom = binom * (xreal)(j + q) / (xreal)q;
      d[j] = binom * (xreal)g[j + m];
    }
    int n_cur = 0;
    /* separators: t0, the roots of g^(m+1) strictly inside, t1 */
    xreal left = t0, f_left, tmp;
    eval_pair(d, dm, left, &f_left, &tmp, 0);
    if (f_left == 0) cur[n_cur++] = left;
    for (int i = 0; i <= n_prev; ++i) {
      const xreal right = i < n_prev ? prev[i] : t1;
      if (!(right > left)) continue;
      xreal f_right;
      eval_pair(d, dm, right, &f_right, &tmp, 0);
      if (f_right == 0) {
        cur[n_cur++] = right;
      } else if (f_left != 0 && (f_left < 0) != (f_right < 0)) {
        cur[n_cur++] = refine_bracket(d, dm, left, right, f_left < 0);
      }
      left = right;
      f_left = f_right;
    }
    memcpy(prev, cur, sizeof(xreal) * (size_t)n_cur);
    n_prev = n_cur;
  }
  for (int i = 0; i < n_prev; ++i) roots_out[i] = (real)prev[i];
  return n_prev;
}

static real magnitude_at(int N, const real* seg_coeffs, const int* dims, int n_dims, real t, int derivative) {
  real s = 0;
  for (int q = 0; q < n_dims; ++q) {
    const real v = poly_eval(N, seg_coeffs + (size_t)dims[q] * N, t, derivative);
    s += v * v;
  }
  return R_SQRT(s);
}

/* Candidate times of one segment (roots only, ascending).  Returns the count. */
API int orc_segment_candidate_roots(int N, int D, const real* seg_coeffs, int derivative, const int* dims,
                                    int n_dims, real t_start, real t_end, int keep_small, real* roots) {
  real g[EXT_MAX_G];
  const int len = orc_candidate_polynomial(N, D, seg_coeffs, derivative, dims, n_dims, g);
  const int last = last_nonzero(g, len, keep_small);
  if (last < 1) return 0; /* all zero, or a constant: no roots (src/rpoly.cpp:61-75) */
  return orc_real_roots_in_range(g, last + 1, t_start, t_end, roots);
}

/* mode 0: PolynomialOptimization::computeMaximumOfMagnitude (LIN.i:470-503): per segment the
 *         candidates are t = 0 and the roots; the end of the last segment closes the list; a
 *         candidate replaces the incumbent only when strictly larger; the incumbent starts as
 *         Extremum() = (0, 0, 0).  Only the maximum is produced.
 * mode 1: Trajectory::computeMinMaxMagnitude (src/trajectory.cpp:181-217): per segment the
 *         candidates are start, end and the roots; strict comparisons, first one wins.
 * out[6] = max time, max value, max segment, min time, min value, min segment.
 * cand_times [K][max_cand] / cand_count [K]: the roots per segment (optional). */
API void orc_minmax_magnitude(int mode, int K, int D, int N, const real* coeffs, const real* times,
                              int derivative, const int* dims, int n_dims, int keep_small, real* out,
                              real* cand_times, int32_t* cand_count, int max_cand) {
  real best_max_t = 0, best_max_v = mode == 0 ? 0 : -DBL_MAX;
  real best_min_t = 0, best_min_v = DBL_MAX;
  int best_max_s = 0, best_min_s = 0;
  for (int s = 0; s < K; ++s) {
    const real* sc = coeffs + (size_t)s * D * N;
    const real T = times[s];
    real roots[EXT_MAX_G];
    const int n_roots = orc_segment_candidate_roots(N, D, sc, derivative, dims, n_dims, 0, T, keep_small, roots);
    if (cand_count) cand_count[s] = n_roots;
    if (cand_times)
      for (int i = 0; i < n_roots && i < max_cand; ++i) cand_times[(size_t)s * max_cand + i] = roots[i];
    if (mode == 0) {
      for (int i = -1; i < n_roots; ++i) {
        const real t = i < 0 ? 0 : roots[i];
        const real v = magnitude_at(N, sc, dims, n_dims, t, derivative);
        if (best_max_v < v) { best_max_v = v; best_max_t = t; best_max_s = s; }
      }
      if (s == K - 1) {
        const real v = magnitude_at(N, sc, dims, n_dims, T, derivative);
        if (best_max_v < v) { best_max_v = v; best_max_t = T; best_max_s = s; }
      }
    } else {
      real seg_max_t = 0, seg_max_v = -DBL_MAX, seg_min_t = 0, seg_min_v = DBL_MAX;
      for (int i = -2; i < n_roots; ++i) {
        const real t = i == -2 ? 0 : i == -1 ? T : roots[i];
        const real v = magnitude_at(N, sc, dims, n_dims, t, derivative);
        if (seg_max_v < v) { seg_max_v = v; seg_max_t = t; }
        if (v < seg_min_v) { seg_min_v = v; seg_min_t = t; }
      }
      if (seg_min_v < best_min_v) { best_min_v = seg_min_v; best_min_t = seg_min_t; best_min_s = s; }
      if (seg_max_v > best_max_v) { best_max_v = seg_max_v; best_max_t = seg_max_t; best_max_s = s; }
    }
  }
  out[0] = best_max_t; out[1] = best_max_v; out[2] = (real)best_max_s;
  out[3] = best_min_t; out[4] = best_min_v; out[5] = (real)best_min_s;
}

/* Magnitude of a derivative at a segment-local time (for re-evaluating reported extrema). */
API real orc_segment_magnitude(int N, int D, const real* seg_coeffs, int derivative, const int* dims, int n_dims,
                               real t) {
  (void)D;
  return magnitude_at(N, seg_coeffs, dims, n_dims, t, derivative);
}
