/*
 * minsnap_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the reference's minimum-snap hot path, written to be
 * the checker for the CUDA kernels.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; nothing under
 * mav_trajectory_generation_cmake_b200/ does.
 *
 * It follows the reference's ORDER OF OPERATIONS (form A, invert it block-wise with a
 * partial-pivot LU, form Q with pow(), H = Ainv^T Q Ainv, R = C^T H C, partition, solve
 * R_pp with an orthogonal factorisation, coefficients = Ainv * d), not the closed forms
 * the GPU kernels use.  Citations are relative to /root/reference/mav_trajectory_generation/:
 *   LIN.i = include/mav_trajectory_generation/impl/polynomial_optimization_linear_impl.h
 *
 * PARITY PINNING.  The reference cannot be compiled here (Eigen3, glog, NLopt are absent
 * and there is no network), and the arithmetic of solveLinear() lives in Eigen's SparseQR
 * (unpinned version, LIN.i:355-364).  This oracle is pinned by:
 *   - the reference's one known-answer vector (test/test_polynomial_optimization.cpp:733-737),
 *     which exercises base table, A^-1, reordering and coefficient recovery;
 *   - the reference's property tests re-expressed in tests/ (A^-1 vs full inverse :194-204,
 *     checkPath :73-131, checkCost :133-152, ConstraintPacking :777-836);
 *   - bit-exact agreement of the input generator with libstdc++'s std::mt19937 +
 *     std::uniform_real_distribution (what the reference's createRandomVertices runs on).
 * The element-level output of the QR solve itself is "PARITY UNPINNED" by any reference
 * fixture: it is checked against the extended-precision build of this same file
 * (-DORC_LONG_DOUBLE) and against the QP's optimality condition R_pp d_p + R_pf d_f = 0,
 * whose solution is unique when R_pp is positive definite.
 *
 * Build: see oracle/Makefile (two shared objects: double and long double arithmetic).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_LONG_DOUBLE
typedef long double real;
#define R_POW powl
#define R_EXP expl
#define R_SQRT sqrtl
#define R_FABS fabsl
#else
typedef double real;
#define R_POW pow
#define R_EXP exp
#define R_SQRT sqrt
#define R_FABS fabs
#endif

#define ORC_MAX_N 12
#define ORC_BASE_N 22 /* Polynomial::kMaxConvolutionSize, polynomial.h:46-50 */

#define API __attribute__((visibility("default")))

API int orc_real_bytes(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------------------------
 * Falling-factorial table, src/polynomial.cpp:140-155 (computeBaseCoefficients):
 * row 0 is ones; row n is row n-1 times (i - n + 1).
 * ---------------------------------------------------------------------------------- */
static real g_base[ORC_BASE_N][ORC_BASE_N];
static int g_base_ready = 0;

static void base_init(void) {
  if (g_base_ready) return;
  for (int i = 0; i < ORC_BASE_N; ++i) g_base[0][i] = 1;
  for (int n = 1; n < ORC_BASE_N; ++n)
    for (int i = 0; i < ORC_BASE_N; ++i)
      g_base[n][i] = (i >= n - 1) ? (real)(i - n + 1) * g_base[n - 1][i] : 0;
  g_base_ready = 1;
}

API void orc_base_coefficients(int n, real* out) {
  base_init();
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) out[r * n + c] = g_base[r][c];
}

/* polynomial.h:215-233 (baseCoeffsWithTime): c[d] = b(d,d); c[j] = b(d,j) t^(j-d) with a
 * running power; if |t| < eps only c[d] is set. */
API void orc_base_coeffs_with_time(int N, int derivative, real t, real* c) {
  base_init();
  for (int j = 0; j < N; ++j) c[j] = 0;
  c[derivative] = g_base[derivative][derivative];
  if (R_FABS(t) < (real)2.220446049250313e-16) return;
  real t_power = t;
  for (int j = derivative + 1; j < N; ++j) {
    c[j] = g_base[derivative][j] * t_power;
    t_power = t_power * t;
  }
}

/* LIN.i:101-111 (setupMappingMatrix): A = [A(0); A(T)], row-major N x N. */
API void orc_setup_mapping_matrix(int N, real T, real* A) {
  const int h = N / 2;
  for (int i = 0; i < h; ++i) {
    orc_base_coeffs_with_time(N, i, 0, A + i * N);
    orc_base_coeffs_with_time(N, i, T, A + (i + h) * N);
  }
}

/* Inverse of a small dense matrix through a partial-pivot LU, the algorithm Eigen's
 * Matrix<double,5,5>::inverse() dispatches to for sizes above 4 (LIN.i:160-161). */
static int lu_inverse(int n, const real* M, real* Minv) {
  real a[ORC_MAX_N * ORC_MAX_N];
  int perm[ORC_MAX_N];
  memcpy(a, M, sizeof(real) * n * n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    real best = R_FABS(a[k * n + k]);
    for (int r = k + 1; r < n; ++r)
      if (R_FABS(a[r * n + k]) > best) { best = R_FABS(a[r * n + k]); p = r; }
    if (best == 0) return 1;
    if (p != k) {
      for (int c = 0; c < n; ++c) { real t = a[k * n + c]; a[k * n + c] = a[p * n + c]; a[p * n + c] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    for (int r = k + 1; r < n; ++r) {
      a[r * n + k] /= a[k * n + k];
      for (int c = k + 1; c < n; ++c) a[r * n + c] -= a[r * n + k] * a[k * n + c];
    }
  }
  for (int col = 0; col < n; ++col) {
    real x[ORC_MAX_N];
    for (int r = 0; r < n; ++r) x[r] = (perm[r] == col) ? 1 : 0;
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < r; ++c) x[r] -= a[r * n + c] * x[c];
    for (int r = n - 1; r >= 0; --r) {
      for (int c = r + 1; c < n; ++c) x[r] -= a[r * n + c] * x[c];
      x[r] /= a[r * n + r];
    }
    for (int r = 0; r < n; ++r) Minv[r * n + col] = x[r];
  }
  return 0;
}

API int orc_dense_inverse(int n, const real* M, real* Minv) { return lu_inverse(n, M, Minv); }

/* LIN.i:132-169 (invertMappingMatrix): with A = [diag 0; C D],
 * A^-1 = [diag^-1 0; -D^-1 C diag^-1  D^-1]. */
API int orc_invert_mapping_matrix(int N, const real* A, real* Ainv) {
  const int h = N / 2;
  real dinv[ORC_MAX_N], C[ORC_MAX_N * ORC_MAX_N], Dm[ORC_MAX_N * ORC_MAX_N] = {0}, Di[ORC_MAX_N * ORC_MAX_N];
  for (int i = 0; i < h; ++i) dinv[i] = (real)1 / A[i * N + i];
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < h; ++c) {
      C[r * h + c] = A[(h + r) * N + c];
      Dm[r * h + c] = A[(h + r) * N + h + c];
    }
  if (lu_inverse(h, Dm, Di)) return 1;
  for (int i = 0; i < N * N; ++i) Ainv[i] = 0;
  for (int i = 0; i < h; ++i) Ainv[i * N + i] = dinv[i];
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < h; ++c) {
      /* (-D^-1 * C) first, then times the diagonal matrix (left-to-right product). */
      real acc = 0;
      for (int k = 0; k < h; ++k) acc += (-Di[r * h + k]) * C[k * h + c];
      Ainv[(h + r) * N + c] = acc * dinv[c];
      Ainv[(h + r) * N + h + c] = Di[r * h + c];
    }
  return 0;
}

/* LIN.i:573-589 (computeQuadraticCostJacobian). */
API void orc_quadratic_cost_jacobian(int N, int derivative, real t, real* Q) {
  base_init();
  for (int i = 0; i < N * N; ++i) Q[i] = 0;
  for (int col = 0; col < N - derivative; ++col)
    for (int row = 0; row < N - derivative; ++row) {
      real exponent = (real)((N - 1 - derivative) * 2 + 1 - row - col);
      Q[(N - 1 - row) * N + (N - 1 - col)] = g_base[derivative][N - 1 - row] *
                                             g_base[derivative][N - 1 - col] *
                                             R_POW(t, exponent) * (real)2.0 / exponent;
    }
}

/* H = Ainv^T Q Ainv as the reference forms it in constructR (LIN.i:305-308). */
API int orc_segment_hessian(int N, int derivative, real T, real* H) {
  real A[ORC_MAX_N * ORC_MAX_N], Ai[ORC_MAX_N * ORC_MAX_N], Q[ORC_MAX_N * ORC_MAX_N], tmp[ORC_MAX_N * ORC_MAX_N];
  orc_setup_mapping_matrix(N, T, A);
  if (orc_invert_mapping_matrix(N, A, Ai)) return 1;
  orc_quadratic_cost_jacobian(N, derivative, T, Q);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      real acc = 0;
      for (int k = 0; k < N; ++k) acc += Ai[k * N + r] * Q[k * N + c];
      tmp[r * N + c] = acc;
    }
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      real acc = 0;
      for (int k = 0; k < N; ++k) acc += tmp[r * N + k] * Ai[k * N + c];
      H[r * N + c] = acc;
    }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Constraint reordering, LIN.h:272-289 + LIN.i:171-250.
 * fixed_mask[(K+1)][N/2]: non-zero when the vertex holds a constraint on that derivative.
 * Row order ("all_constraints"): vertex 0 once, interior vertices twice, vertex K once,
 * derivative 0..N/2-1 inside each occurrence.  Column order: the std::set ordering by
 * (vertex, derivative) -- every fixed constraint first, then every free one.
 * ---------------------------------------------------------------------------------- */
API int orc_constraint_reordering(int N, int K, const uint8_t* fixed_mask, int32_t* col_of_row,
                                  int32_t* n_fixed_out, int32_t* n_free_out) {
  const int h = N / 2;
  const int nv = K + 1;
  int32_t* col_of_constraint = (int32_t*)malloc(sizeof(int32_t) * nv * h);
  int n_fixed = 0, n_free = 0;
  for (int i = 0; i < nv * h; ++i) (fixed_mask[i] ? ++n_fixed : ++n_free);
  int cf = 0, cp = n_fixed;
  for (int v = 0; v < nv; ++v)
    for (int c = 0; c < h; ++c) col_of_constraint[v * h + c] = fixed_mask[v * h + c] ? cf++ : cp++;
  int row = 0;
  for (int v = 0; v < nv; ++v) {
    const int occ = (v == 0 || v == K) ? 1 : 2;
    for (int o = 0; o < occ; ++o)
      for (int c = 0; c < h; ++c) col_of_row[row++] = col_of_constraint[v * h + c];
  }
  free(col_of_constraint);
  *n_fixed_out = n_fixed;
  *n_free_out = n_free;
  return row == N * K ? 0 : 1;
}

/* Dense Householder QR solve of a square system M x = b for nrhs right-hand sides
 * (stands in for Eigen::SparseQR, LIN.i:355-364: same class of backward-stable
 * orthogonal factorisation, no pivoting needed for the well-conditioned R_pp). */
static int qr_solve(int n, real* M /* n*n row-major, destroyed */, real* Bm /* n*nrhs */, int nrhs) {
  real* v = (real*)malloc(sizeof(real) * n);
  int status = 0;
  for (int k = 0; k < n; ++k) {
    real norm2 = 0;
    for (int r = k; r < n; ++r) norm2 += M[r * n + k] * M[r * n + k];
    real norm = R_SQRT(norm2);
    if (norm == 0) { status = 1; continue; }
    real alpha = (M[k * n + k] > 0) ? -norm : norm;
    for (int r = k; r < n; ++r) v[r] = M[r * n + k];
    v[k] -= alpha;
    real vnorm2 = 0;
    for (int r = k; r < n; ++r) vnorm2 += v[r] * v[r];
    if (vnorm2 == 0) continue;
    for (int c = k; c < n; ++c) {
      real dot = 0;
      for (int r = k; r < n; ++r) dot += v[r] * M[r * n + c];
      real f = (real)2 * dot / vnorm2;
      for (int r = k; r < n; ++r) M[r * n + c] -= f * v[r];
    }
    for (int c = 0; c < nrhs; ++c) {
      real dot = 0;
      for (int r = k; r < n; ++r) dot += v[r] * Bm[r * nrhs + c];
      real f = (real)2 * dot / vnorm2;
      for (int r = k; r < n; ++r) Bm[r * nrhs + c] -= f * v[r];
    }
  }
  for (int c = 0; c < nrhs; ++c)
    for (int r = n - 1; r >= 0; --r) {
      real acc = Bm[r * nrhs + c];
      for (int j = r + 1; j < n; ++j) acc -= M[r * n + j] * Bm[j * nrhs + c];
      if (M[r * n + r] == 0) { status = 1; Bm[r * nrhs + c] = 0; }
      else Bm[r * nrhs + c] = acc / M[r * n + r];
    }
  free(v);
  return status;
}

/* ------------------------------------------------------------------------------------
 * setupFromVertices + solveLinear for one trajectory (LIN.i:46-99, 275-295, 297-369,
 * 252-273, 113-130).
 *   fixed_mask   [(K+1)][N/2]
 *   vertex_values[(K+1)][N/2][D]   value of every constraint (ignored where mask == 0)
 *   times        [K]
 * outputs (any may be NULL):
 *   coeffs [K][D][N], d_fixed [D][n_fixed], d_free [D][n_free], cost (0.5 sum c^T Q c),
 *   R_dense [(n_fixed+n_free)^2] row-major.
 * returns 0, 1 = singular factorisation, 2 = bad argument.
 * ---------------------------------------------------------------------------------- */
API int orc_solve_linear(int N, int K, int D, int derivative, const uint8_t* fixed_mask,
                         const real* vertex_values, const real* times, real* coeffs, real* d_fixed,
                         real* d_free, real* cost, real* R_dense) {
  if (N < 2 || N > ORC_MAX_N || (N & 1) || K < 1 || D < 1 || derivative < 0 || derivative > N / 2 - 1) return 2;
  for (int i = 0; i < K; ++i)
    if (!(times[i] > 0)) return 2; /* CHECK_GT(segment_time, 0), LIN.i:287 */
  const int h = N / 2;
  const int rows = N * K;
  int status = 0;
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * rows);
  int32_t n_fixed, n_free;
  orc_constraint_reordering(N, K, fixed_mask, col, &n_fixed, &n_free);
  const int n_all = n_fixed + n_free;

  /* updateSegmentTimes, LIN.i:275-295 */
  real* Ainv = (real*)malloc(sizeof(real) * K * N * N);
  real* Q = (real*)malloc(sizeof(real) * K * N * N);
  real A[ORC_MAX_N * ORC_MAX_N];
  for (int i = 0; i < K; ++i) {
    orc_quadratic_cost_jacobian(N, derivative, times[i], Q + i * N * N);
    orc_setup_mapping_matrix(N, times[i], A);
    if (orc_invert_mapping_matrix(N, A, Ainv + i * N * N)) status = 1;
  }

  /* d_f in column order, LIN.i:228-246 */
  real* dall = (real*)calloc((size_t)D * n_all, sizeof(real)); /* [D][n_all] = [d_f; d_p] */
  {
    int cf = 0;
    for (int v = 0; v <= K; ++v)
      for (int c = 0; c < h; ++c)
        if (fixed_mask[v * h + c]) {
          for (int d = 0; d < D; ++d) dall[d * n_all + cf] = vertex_values[(v * h + c) * D + d];
          ++cf;
        }
  }

  if (n_free > 0) {
    /* constructR, LIN.i:297-326: R = C^T blockdiag(H_i) C */
    real* R = (real*)calloc((size_t)n_all * n_all, sizeof(real));
    real H[ORC_MAX_N * ORC_MAX_N], tmp[ORC_MAX_N * ORC_MAX_N];
    for (int i = 0; i < K; ++i) {
      const real* Ai = Ainv + i * N * N;
      const real* Qi = Q + i * N * N;
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          real acc = 0;
          for (int k = 0; k < N; ++k) acc += Ai[k * N + r] * Qi[k * N + c];
          tmp[r * N + c] = acc;
        }
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          real acc = 0;
          for (int k = 0; k < N; ++k) acc += tmp[r * N + k] * Ai[k * N + c];
          H[r * N + c] = acc;
        }
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) R[col[i * N + r] * n_all + col[i * N + c]] += H[r * N + c];
    }
    if (R_dense) memcpy(R_dense, R, sizeof(real) * n_all * n_all);

    /* solveLinear, LIN.i:350-365 */
    real* Rpp = (real*)malloc(sizeof(real) * n_free * n_free);
    real* rhs = (real*)malloc(sizeof(real) * n_free * D);
    for (int r = 0; r < n_free; ++r) {
      for (int c = 0; c < n_free; ++c) Rpp[r * n_free + c] = R[(n_fixed + r) * n_all + n_fixed + c];
      for (int d = 0; d < D; ++d) {
        real acc = 0;
        for (int c = 0; c < n_fixed; ++c) acc += (-R[(n_fixed + r) * n_all + c]) * dall[d * n_all + c];
        rhs[r * D + d] = acc;
      }
    }
    if (qr_solve(n_free, Rpp, rhs, D)) status = 1;
    for (int r = 0; r < n_free; ++r)
      for (int d = 0; d < D; ++d) dall[d * n_all + n_fixed + r] = rhs[r * D + d];
    free(Rpp); free(rhs); free(R);
  } else if (R_dense) {
    /* fully constrained (LIN.i:333-339): R is still defined; build it for the accessor. */
    real H[ORC_MAX_N * ORC_MAX_N];
    for (int i = 0; i < n_all * n_all; ++i) R_dense[i] = 0;
    for (int i = 0; i < K; ++i) {
      orc_segment_hessian(N, derivative, times[i], H);
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) R_dense[col[i * N + r] * n_all + col[i * N + c]] += H[r * N + c];
    }
  }

  /* updateSegmentsFromCompactConstraints, LIN.i:252-273; computeCost, LIN.i:113-130 */
  real total = 0;
  for (int d = 0; d < D; ++d)
    for (int i = 0; i < K; ++i) {
      real dseg[ORC_MAX_N], c[ORC_MAX_N];
      for (int r = 0; r < N; ++r) dseg[r] = dall[d * n_all + col[i * N + r]];
      for (int r = 0; r < N; ++r) {
        real acc = 0;
        for (int k = 0; k < N; ++k) acc += Ainv[i * N * N + r * N + k] * dseg[k];
        c[r] = acc;
      }
      if (coeffs)
        for (int r = 0; r < N; ++r) coeffs[(i * D + d) * N + r] = c[r];
      real partial = 0;
      for (int r = 0; r < N; ++r) {
        real acc = 0;
        for (int k = 0; k < N; ++k) acc += Q[i * N * N + r * N + k] * c[k];
        partial += c[r] * acc;
      }
      total += partial;
    }
  if (cost) *cost = (real)0.5 * total;
  if (d_fixed)
    for (int d = 0; d < D; ++d)
      for (int c = 0; c < n_fixed; ++c) d_fixed[d * n_fixed + c] = dall[d * n_all + c];
  if (d_free)
    for (int d = 0; d < D; ++d)
      for (int c = 0; c < n_free; ++c) d_free[d * n_free + c] = dall[d * n_all + n_fixed + c];
  free(col); free(Ainv); free(Q); free(dall);
  return status;
}

/* updateSegmentsFromCompactConstraints alone (LIN.i:252-273), used by setFreeConstraints
 * (LIN.i:505-514): d_all[D][n_all] -> coeffs[K][D][N]. */
API int orc_coeffs_from_constraints(int N, int K, int D, const int32_t* col_of_row, int n_all,
                                    const real* d_all, const real* times, real* coeffs) {
  real A[ORC_MAX_N * ORC_MAX_N], Ai[ORC_MAX_N * ORC_MAX_N];
  for (int i = 0; i < K; ++i) {
    orc_setup_mapping_matrix(N, times[i], A);
    if (orc_invert_mapping_matrix(N, A, Ai)) return 1;
    for (int d = 0; d < D; ++d)
      for (int r = 0; r < N; ++r) {
        real acc = 0;
        for (int k = 0; k < N; ++k) acc += Ai[r * N + k] * d_all[d * n_all + col_of_row[i * N + k]];
        coeffs[(i * D + d) * N + r] = acc;
      }
  }
  return 0;
}

/* computeCost from coefficients (LIN.i:113-130). coeffs[K][D][N]. */
API real orc_compute_cost(int N, int K, int D, int derivative, const real* coeffs, const real* times) {
  real Q[ORC_MAX_N * ORC_MAX_N];
  real total = 0;
  for (int i = 0; i < K; ++i) {
    orc_quadratic_cost_jacobian(N, derivative, times[i], Q);
    for (int d = 0; d < D; ++d) {
      const real* c = coeffs + (i * D + d) * N;
      real partial = 0;
      for (int r = 0; r < N; ++r) {
        real acc = 0;
        for (int k = 0; k < N; ++k) acc += Q[r * N + k] * c[k];
        partial += c[r] * acc;
      }
      total += partial;
    }
  }
  return (real)0.5 * total;
}

/* ------------------------------------------------------------------------------------
 * Evaluation: polynomial.h:138-151 (Horner over a base-table row), src/segment.cpp:51-58,
 * src/trajectory.cpp:41-66.
 * ---------------------------------------------------------------------------------- */
API real orc_polynomial_evaluate(int N, const real* c, real t, int derivative) {
  base_init();
  if (derivative >= N) return 0;
  real result = g_base[derivative][N - 1] * c[N - 1];
  for (int j = N - 2; j >= derivative; --j) {
    result *= t;
    result += g_base[derivative][j] * c[j];
  }
  return result;
}

/* Trajectory::evaluate: coeffs[K][D][N], times[K]; out[D]. Returns the segment index used,
 * or -1 when t is past the end (the reference logs an error and returns zeros). */
API int orc_trajectory_evaluate(int N, int K, int D, const real* coeffs, const real* times, real t,
                                int derivative, real* out) {
  real accumulated = 0;
  int i;
  for (i = 0; i < K; ++i) {
    accumulated += times[i];
    if (accumulated > t) break;
  }
  if (t > accumulated || i >= K) { /* i >= K: t == max time, out of bounds in the reference */
    for (int d = 0; d < D; ++d) out[d] = 0;
    return -1;
  }
  accumulated -= times[i];
  for (int d = 0; d < D; ++d)
    out[d] = orc_polynomial_evaluate(N, coeffs + (i * D + d) * N, t - accumulated, derivative);
  return i;
}

/* Batch of sample instants for one trajectory: out[M][n_deriv][D], derivative 0..n_deriv-1. */
API void orc_trajectory_sample(int N, int K, int D, const real* coeffs, const real* times, int M,
                               const real* t, int n_deriv, real* out) {
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < n_deriv; ++k)
      orc_trajectory_evaluate(N, K, D, coeffs, times, t[m], k, out + ((size_t)m * n_deriv + k) * D);
}

/* Trajectory::evaluateRange, src/trajectory.cpp:68-128. Writes at most max_out samples of
 * out[.][D] (and sample_times[.] when non-NULL); returns the number the reference would
 * produce. */
API int orc_trajectory_evaluate_range(int N, int K, int D, const real* coeffs, const real* times,
                                      real t_start, real t_end, real dt, int derivative, int max_out,
                                      real* out, real* sample_times) {
  real accumulated = 0;
  int i;
  for (i = 0; i < K; ++i) {
    accumulated += times[i];
    if (accumulated > t_start) break;
  }
  if (t_start > accumulated || i >= K) return 0;
  accumulated -= times[i];
  real time_in_segment = t_start - accumulated;
  int n = 0;
  while (accumulated < t_end) {
    if (time_in_segment > times[i]) {
      time_in_segment = time_in_segment - times[i];
      ++i;
      if (i >= K) break;
      continue;
    }
    if (n < max_out) {
      for (int d = 0; d < D; ++d)
        out[(size_t)n * D + d] = orc_polynomial_evaluate(N, coeffs + (i * D + d) * N, time_in_segment, derivative);
      if (sample_times) sample_times[n] = accumulated;
    }
    ++n;
    time_in_segment += dt;
    accumulated += dt;
  }
  return n;
}

/* ------------------------------------------------------------------------------------
 * Inputs: src/vertex.cpp:162-178 (estimateSegmentTimes) and :27-79 (createRandomVertices,
 * std::mt19937 + one std::uniform_real_distribution<double> per dimension).
 * ---------------------------------------------------------------------------------- */
API void orc_estimate_segment_times(int K, int D, const real* positions /* [K+1][D] */, real v_max,
                                    real a_max, real magic, real* times) {
  for (int i = 0; i < K; ++i) {
    real s = 0;
    for (int d = 0; d < D; ++d) {
      real diff = positions[(i + 1) * D + d] - positions[i * D + d];
      s += diff * diff;
    }
    real distance = R_SQRT(s);
    times[i] = distance / v_max * 2 * ((real)1.0 + magic * v_max / a_max * R_EXP(-distance / v_max * 2));
  }
}

typedef struct { uint32_t mt[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt19937_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* libstdc++ std::generate_canonical<double, 53>(mt19937): two 32-bit draws, low word first,
 * accumulated in double, divided by 2^64; then a + (b - a) * u. Always double arithmetic
 * (the inputs must be bit-identical in both precision builds). */
static double mt_uniform(mt19937_t* g, double a, double b) {
  double sum = (double)mt_next(g);
  sum += (double)mt_next(g) * 4294967296.0;
  double u = sum / 18446744073709551616.0;
  if (u >= 1.0) u = nextafter(1.0, 0.0);
  return u * (b - a) + a;
}

/* positions[K+1][D] in double; the caller applies the mask (ends fixed up to
 * maximum_derivative with zero derivatives, interior position only). */
API void orc_create_random_positions(int K, int D, const double* pos_min, const double* pos_max,
                                     uint64_t seed, double* positions) {
  mt19937_t g;
  mt_seed(&g, (uint32_t)seed); /* std::mt19937(size_t) truncates to result_type */
  const double min_distance = 0.2;
  double last[16], pos[16];
  for (int d = 0; d < D; ++d) last[d] = mt_uniform(&g, pos_min[d], pos_max[d]);
  for (int d = 0; d < D; ++d) positions[d] = last[d];
  for (int i = 1; i <= K; ++i) {
    for (;;) {
      double s = 0;
      for (int d = 0; d < D; ++d) {
        pos[d] = mt_uniform(&g, pos_min[d], pos_max[d]);
        s += (pos[d] - last[d]) * (pos[d] - last[d]);
      }
      if (sqrt(s) > min_distance) break;
    }
    for (int d = 0; d < D; ++d) positions[i * D + d] = last[d] = pos[d];
  }
}

/* Raw generator access so tests can compare with std::mt19937 directly. */
API void orc_mt19937_uniform(uint64_t seed, double a, double b, int n, double* out) {
  mt19937_t g;
  mt_seed(&g, (uint32_t)seed);
  for (int i = 0; i < n; ++i) out[i] = mt_uniform(&g, a, b);
}

/* ------------------------------------------------------------------------------------
 * Batched driver used as the CPU baseline (bench.py): solves problems [b0, b1) of a batch
 * with the standard createRandomVertices mask, one problem at a time on the calling thread,
 * OpenMP across problems (single-threaded per problem, as the reference is).
 * positions[B][K+1][D], times[B][K] -> coeffs[B][K][D][N].  Double build only.
 * ---------------------------------------------------------------------------------- */
#ifndef ORC_LONG_DOUBLE
API int orc_solve_batch_standard(int N, int K, int D, int derivative, int max_fixed_derivative,
                                 long B, const double* positions, const double* times,
                                 double* coeffs, double* cost, int n_threads) {
  const int h = N / 2;
  int bad = 0;
  (void)n_threads;
  base_init(); /* before the parallel region: the table is shared and read-only afterwards */
  uint8_t* mask = (uint8_t*)calloc((size_t)(K + 1) * h, 1);
  for (int v = 0; v <= K; ++v) {
    mask[v * h] = 1;
    if (v == 0 || v == K)
      for (int c = 1; c <= max_fixed_derivative && c < h; ++c) mask[v * h + c] = 1;
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads) reduction(| : bad)
#endif
  for (long b = 0; b < B; ++b) {
    double* vals = (double*)calloc((size_t)(K + 1) * h * D, sizeof(double));
    for (int v = 0; v <= K; ++v)
      for (int d = 0; d < D; ++d) vals[(v * h) * D + d] = positions[(b * (K + 1) + v) * D + d];
    double c;
    int st = orc_solve_linear(N, K, D, derivative, mask, vals, times + b * K,
                              coeffs ? coeffs + (size_t)b * K * D * N : NULL, NULL, NULL, &c, NULL);
    if (cost) cost[b] = c;
    bad |= st;
    free(vals);
  }
  free(mask);
  return bad;
}
#endif
