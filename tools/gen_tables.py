#!/usr/bin/env python3
"""Generate mav_trajectory_generation_cmake_b200/csrc/minsnap_tables.h.

Exact-rational (fractions.Fraction) evaluation of the two constant matrices the
CUDA kernels use instead of forming A, A^-1, Q and H = A^-T Q A^-1 per segment
(note A):

    A1inv[N]        = A(T=1)^-1              (mapping matrix inverse at unit time)
    H1[N][delta]    = A1inv^T Q(T=1) A1inv   (end-point-derivative Hessian at unit time)

with, for a segment of duration T and row derivative orders k_r = r mod N/2,

    Ainv_T[i][r] = A1inv[i][r] * T^(k_r - i)
    H_T[r][s]    = H1[r][s]    * T^(k_r + k_s + 1 - 2*delta)

Definitions follow the reference (citations relative to /root/reference):
  * base coefficients b(d,j) = j!/(j-d)!           mav_trajectory_generation/src/polynomial.cpp:140-155
  * rows of A                                      include/mav_trajectory_generation/polynomial.h:215-233
  * A = [A(0); A(T)]                               impl/polynomial_optimization_linear_impl.h:101-111
  * Q(i,j) = 2 b(d,i) b(d,j) T^e / e               impl/polynomial_optimization_linear_impl.h:573-589

Every table entry is the correctly rounded double of an exact rational number.
Run:  python tools/gen_tables.py
"""
from fractions import Fraction
import os

NS = (2, 4, 6, 8, 10, 12)
OUT = os.path.join(os.path.dirname(__file__), "..", "mav_trajectory_generation_cmake_b200",
                   "csrc", "minsnap_tables.h")


def falling(d, j):
    """b(d, j) = j * (j-1) * ... * (j-d+1); zero for j < d."""
    if j < d:
        return 0
    r = 1
    for q in range(d):
        r *= (j - q)
    return r


def mapping_matrix_unit(N):
    h = N // 2
    A = [[Fraction(0)] * N for _ in range(N)]
    for d in range(h):
        A[d][d] = Fraction(falling(d, d))          # row at t = 0
        for j in range(d, N):
            A[h + d][j] = Fraction(falling(d, j))  # row at t = 1
    return A


def cost_matrix_unit(N, delta):
    Q = [[Fraction(0)] * N for _ in range(N)]
    for i in range(delta, N):
        for j in range(delta, N):
            e = i + j - 2 * delta + 1
            Q[i][j] = Fraction(2 * falling(delta, i) * falling(delta, j), e)
    return Q


def inverse(M):
    n = len(M)
    a = [row[:] + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(M)]
    for c in range(n):
        p = next(r for r in range(c, n) if a[r][c] != 0)
        a[c], a[p] = a[p], a[c]
        piv = a[c][c]
        a[c] = [x / piv for x in a[c]]
        for r in range(n):
            if r != c and a[r][c] != 0:
                f = a[r][c]
                a[r] = [x - f * y for x, y in zip(a[r], a[c])]
    return [row[n:] for row in a]


def matmul(X, Y):
    return [[sum(X[i][k] * Y[k][j] for k in range(len(Y))) for j in range(len(Y[0]))]
            for i in range(len(X))]


def transpose(X):
    return [list(r) for r in zip(*X)]


def lit(x):
    f = float(x)  # Fraction -> float is correctly rounded
    return repr(f)


def emit_matrix(name, M, fh):
    n = len(M)
    # no `const` on purpose: a non-const __constant__ array cannot be folded into immediate
    # moves by the compiler, so kernels read it as c[bank][offset] operands of DFMA/DMUL
    fh.write("MINSNAP_TABLE_QUAL MINSNAP_TABLE_CONST double %s[%d] = {\n" % (name, n * n))
    for row in M:
        fh.write("    " + ", ".join(lit(x) for x in row) + ",\n")
    fh.write("};\n")


def recovery_roles_n10(Ai):
    """Per-lane-role constants of the coefficient recovery in minsnap_standard_tm.cuh (N = 10): a top-down
    lane's new vertex vector STARTS its segment and the previous one ends it; a bottom-up lane has it the other
    way round and odd derivatives change sign.  Layout per role (54 entries, 53 used):
      [(i-5) 9 + 0]     = A1inv[i][5]                                          (position difference)
      [(i-5) 9 + 1 + a] = role ? (-1)^(a+1) A1inv[i][6+a] : A1inv[i][1+a]      (applied to T^k new_k)
      [(i-5) 9 + 5 + a] = role ? (-1)^(a+1) A1inv[i][1+a] : A1inv[i][6+a]      (applied to T^k old_k)
      [45 + a]          = role ? 0 : A1inv[1+a][1+a]                           (c_k from new_k)
      [49 + a]          = role ? (-1)^(a+1) A1inv[1+a][1+a] : 0                (c_k from old_k)"""
    out = []
    for role in (0, 1):
        t = [Fraction(0)] * 54
        for i in range(5, 10):
            t[(i - 5) * 9] = Ai[i][5]
            for a in range(4):
                sgn = -1 if (role and a % 2 == 0) else 1   # derivative k = a + 1 odd
                t[(i - 5) * 9 + 1 + a] = sgn * (Ai[i][6 + a] if role else Ai[i][1 + a])
                t[(i - 5) * 9 + 5 + a] = sgn * (Ai[i][1 + a] if role else Ai[i][6 + a])
        for a in range(4):
            sgn = -1 if (role and a % 2 == 0) else 1
            if role:
                t[49 + a] = sgn * Ai[1 + a][1 + a]
            else:
                t[45 + a] = Ai[1 + a][1 + a]
        out.append(t)
    return out


def cost_form_n10(H):
    """Packed lower triangle of the per-segment cost form in the kernels' variable order
    u = [dp, T^k start_k (k = 1..4), T^k end_k (k = 1..4)] (rows of H1: dp -> 5, start k -> k, end k -> 5 + k):
    entry tri(r, s) = H1[row r][row s] on the diagonal and 2 H1[row r][row s] below it, so that
    u^T H u = sum_r u_r (sum_{s <= r} table[tri(r, s)] u_s) with 54 multiply-adds instead of 90."""
    rows = [5, 1, 2, 3, 4, 6, 7, 8, 9]
    out = []
    for r in range(9):
        for s in range(r + 1):
            v = H[rows[r]][rows[s]]
            out.append(v if r == s else 2 * v)
    return out


def main():
    with open(OUT, "w") as fh:
        fh.write("// GENERATED by tools/gen_tables.py -- do not edit.\n")
        fh.write("// Unit-time constants: A1inv[N] = A(1)^-1, H1[N][delta] = A1inv^T Q(1) A1inv.\n")
        fh.write("// Each literal is the correctly rounded double of an exact rational.\n")
        fh.write("#pragma once\n#ifndef MINSNAP_TABLE_QUAL\n#define MINSNAP_TABLE_QUAL static\n#endif\n#ifndef MINSNAP_TABLE_CONST\n#define MINSNAP_TABLE_CONST\n#endif\n#ifndef MINSNAP_TABLE_GLOBAL_QUAL\n#define MINSNAP_TABLE_GLOBAL_QUAL static\n#endif\n\nnamespace minsnap_tables {\n\n")
        for N in NS:
            A = mapping_matrix_unit(N)
            Ai = inverse(A)
            # sanity: A * Ai == I exactly
            I = matmul(A, Ai)
            assert all(I[i][j] == (1 if i == j else 0) for i in range(N) for j in range(N))
            emit_matrix("kA1inv_N%d" % N, Ai, fh)
            if N == 10:
                # global-memory table (coalesced loads; a lane-indexed read of a __constant__ array serialises)
                fh.write("MINSNAP_TABLE_GLOBAL_QUAL MINSNAP_TABLE_CONST double kRecoveryRoles_N10[108] = {\n")
                for t in recovery_roles_n10(Ai):
                    for r0 in range(0, 54, 9):
                        fh.write("    " + ", ".join(lit(x) for x in t[r0:r0 + 9]) + ",\n")
                fh.write("};\n")
            AiT = transpose(Ai)
            for delta in range(N // 2):
                H = matmul(matmul(AiT, cost_matrix_unit(N, delta)), Ai)
                assert all(H[i][j] == H[j][i] for i in range(N) for j in range(N))
                emit_matrix("kH1_N%d_d%d" % (N, delta), H, fh)
                if N == 10 and delta == 4:
                    t = cost_form_n10(H)
                    fh.write("MINSNAP_TABLE_QUAL MINSNAP_TABLE_CONST double kCostForm_N10_d4[45] = {\n")
                    for r in range(9):
                        fh.write("    " + ", ".join(lit(x) for x in t[r * (r + 1) // 2:(r + 1) * (r + 2) // 2]) + ",\n")
                    fh.write("};\n")
                    # the same table in global memory, for kernels that stage it in shared memory by cp.async
                    fh.write("MINSNAP_TABLE_GLOBAL_QUAL MINSNAP_TABLE_CONST double kCostFormGlobal_N10_d4[46] = {\n")
                    for r in range(9):
                        fh.write("    " + ", ".join(lit(x) for x in t[r * (r + 1) // 2:(r + 1) * (r + 2) // 2]) + ",\n")
                    fh.write("    0.0,\n};\n")
            fh.write("\n")
        fh.write("}  // namespace minsnap_tables\n")
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
