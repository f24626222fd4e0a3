// Warp-specialised form of the standard-mask solver (N = 10, snap): a CTA of D+1 warps owns a
// batch of 16 trajectories (thread pair per trajectory in every warp, "burn at both ends" as in
// minsnap_standard_fast.cuh, whose algebra and local-coordinate conventions it shares).
//
//   warp 0      MATRIX warp.  The block-tridiagonal factorisation does not depend on the
//               right-hand sides: per eliminated block it forms S = L D L^T, S^-1 and
//               Z = S^-1 SE, publishes S^-1 and Z in shared memory and signals a named barrier
//               (bar.arrive); then the Schur update S' = D' - SE^T Z.  Last: the middle block.
//   warp 1+d    DIMENSION warp d.  Carries dimension d only: forward g' = b' - Z^T g (waits on
//               the block's barrier with bar.sync), the middle block x_m = S_m^-1 g_m, back
//               substitution x = S^-1 g - Z x_next, then coefficient recovery (ref
//               updateSegmentsFromCompactConstraints, LIN.i:252-273) and the cost terms of its
//               dimension, storing each polynomial as 80 contiguous bytes.
//
// Versus the two-lane kernel this keeps the shared memory per batch about equal but runs D+1
// warps on it, each with a dependency chain ~4x shorter and ~128 registers: 16 resident warps
// per SM instead of 6.  Total instruction count per trajectory is unchanged.
#pragma once
#include "minsnap_standard_fast.cuh"

namespace minsnap {
namespace ws {

using fast::FastParams;
using fast::TimePowers;
using fast::fast_rcp;
using fast::kF;
using fast::kN;
using fast::kPairsPerWarp;
using fast::tri;

constexpr int kStride = 32;   // doubles between consecutive slots ([slot][lane])

#define H1T(r, s) (minsnap_tables::kH1_N10_d4[(r) * 10 + (s)])
#define A1T(i, r) (minsnap_tables::kA1inv_N10[(i) * 10 + (r)])

__device__ __forceinline__ void named_barrier_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_barrier_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Shared-memory carving (doubles).  Per lane and eliminated block: Z (16) and S^-1 (10, lower
// triangle); one more S^-1 for the middle block; per dimension and block: g, later x (4).
template <int D>
struct Layout {
  int mA, z_off, s_off, g_off, slots, pos_off, time_off, red_off, total;
  __host__ __device__ explicit Layout(int K) {
    mA = (K - 1) / 2;
    z_off = 0;
    s_off = z_off + 16 * mA;
    g_off = s_off + 10 * (mA + 1);
    slots = g_off + 4 * D * mA;
    pos_off = slots * kStride;
    time_off = pos_off + ((kPairsPerWarp * (K + 1) * D + 1) & ~1);
    red_off = time_off + ((kPairsPerWarp * K + 1) & ~1);
    total = red_off + (D + D + 2) * 32 / 1;   // cost partials [D][32] + status words [(D+1)][32] as doubles
  }
};

// S^-1 (lower triangle) from S = L D L^T.
__device__ __forceinline__ void ldlt4_inverse(const double (&l)[10], const double (&inv)[4], double (&si)[10]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double col[4] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0, c == 3 ? 1.0 : 0.0};
    fast::ldlt4_solve(l, inv, col);
#pragma unroll
    for (int r = c; r < 4; ++r) si[tri(r, c)] = col[r];
  }
}

__device__ __forceinline__ double sym(const double (&m)[10], int a, int b) { return a >= b ? m[tri(a, b)] : m[tri(b, a)]; }

template <int D, bool kCoeffs>
__global__ void __launch_bounds__((D + 1) * 32, 3) solve_standard_ws_kernel(FastParams p) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_threads = (D + 1) * 32;
  const int K = p.K;
  const Layout<D> lay(K);
  const int mA = lay.mA;
  const int nb = K - 1;
  const int nB = nb - mA - 1;
  const int side = lane & 1;
  const int q = lane >> 1;
  const int my_n = side ? nB : mA;
  double* const slot = smem + lane;                // slot s of this lane: slot[s * kStride]
  double* const pos_s = smem + lay.pos_off;
  double* const time_s = smem + lay.time_off;
  double* const red = smem + lay.red_off;
  const double flip[kF] = {side ? -1.0 : 1.0, 1.0, side ? -1.0 : 1.0, 1.0};
  const int per_pos = (K + 1) * D;

  const long n_problems = p.sweep_S > 0 ? p.B * p.sweep_S : p.B;
  const long base = (long)blockIdx.x * kPairsPerWarp;
  if (base >= n_problems) return;
  const int n_here = (int)min((long)kPairsPerWarp, n_problems - base);
  const long prob = base + q;
  const bool active = q < n_here;

  // ---- stage the inputs of the batch (all warps), cp.async -----------------------------------
  {
    const int tid = threadIdx.x;
    if (p.sweep_S > 0) {
      for (int e = tid; e < n_here * per_pos; e += n_threads) {
        const int r = e / per_pos, o = e - r * per_pos;
        __pipeline_memcpy_async(pos_s + e, p.positions + ((base + r) / p.sweep_S) * per_pos + o, 8);
      }
    } else if (p.aligned16) {
      const double* src = p.positions + base * per_pos;
      const int n = n_here * per_pos, n2 = n >> 1;
      for (int e = tid; e < n2; e += n_threads) __pipeline_memcpy_async(pos_s + 2 * e, src + 2 * e, 16);
      if ((n & 1) && tid == 0) __pipeline_memcpy_async(pos_s + n - 1, src + n - 1, 8);
    } else {
      const double* src = p.positions + base * per_pos;
      for (int e = tid; e < n_here * per_pos; e += n_threads) __pipeline_memcpy_async(pos_s + e, src + e, 8);
    }
    if (p.times) {
      const double* src = p.times + base * K;
      const int n = n_here * K;
      if (p.aligned16) {
        for (int e = tid; e < (n >> 1); e += n_threads) __pipeline_memcpy_async(time_s + 2 * e, src + 2 * e, 16);
        if ((n & 1) && tid == 0) __pipeline_memcpy_async(time_s + n - 1, src + n - 1, 8);
      } else {
        for (int e = tid; e < n; e += n_threads) __pipeline_memcpy_async(time_s + e, src + e, 8);
      }
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    if (!p.times) {
      for (int e = tid; e < n_here * K; e += n_threads) {
        const int r = e / K, o = e - r * K;
        const double* p0 = pos_s + r * per_pos + o * D;
        double s2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double diff = p0[D + d] - p0[d];
          s2 += diff * diff;
        }
        const double distance = sqrt(s2);
        const double T = distance / p.v_max * 2 * (1.0 + p.magic * p.v_max / p.a_max * exp(-distance / p.v_max * 2));
        time_s[e] = T;
        if (p.times_out) p.times_out[(base + r) * K + o] = T;
      }
      __syncthreads();
    }
  }

  const double* my_pos = pos_s + q * per_pos;
  const double* my_time = time_s + q * K;
  auto local_T = [&](int j) { return my_time[side ? K - 1 - j : j]; };
  int status = 0;

  if (warp == 0) {
    // =========================== MATRIX warp ==================================================
    if (nb > 0) {
      TimePowers tp_prev, tp_next;
      tp_prev.set(local_T(0));
      tp_next.set(local_T(1));
      double S[10];
      fast::diag_block(tp_prev, tp_next, S);
      double C[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) C[i] = 0.0;
      for (int j = 1; j <= mA; ++j) {
        if (j <= my_n) {
          double L[10], inv[4], Si[10];
          if (!fast::ldlt4(S, L, inv)) status |= 1;
          ldlt4_inverse(L, inv, Si);
          double* sb = slot + (lay.s_off + (j - 1) * 10) * kStride;
#pragma unroll
          for (int i = 0; i < 10; ++i) sb[i * kStride] = Si[i];
          double E[kF][kF], Z[kF][kF];
          fast::coupling_block(tp_next, E);
          double* zb = slot + (lay.z_off + (j - 1) * 16) * kStride;
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int b = 0; b < kF; ++b) {
              double acc = sym(Si, a, 0) * E[0][b];
#pragma unroll
              for (int r = 1; r < kF; ++r) acc = fma(sym(Si, a, r), E[r][b], acc);
              Z[a][b] = acc;
              zb[(a * kF + b) * kStride] = acc;
            }
          if (j < my_n) {
            tp_prev = tp_next;
            tp_next.set(local_T(j + 1));
            fast::diag_block(tp_prev, tp_next, S);
#pragma unroll
            for (int a = 0; a < kF; ++a)
#pragma unroll
              for (int b = 0; b <= a; ++b) {
                double acc = S[tri(a, b)];
#pragma unroll
                for (int r = 0; r < kF; ++r) acc = fma(-E[r][a], Z[r][b], acc);
                S[tri(a, b)] = acc;
              }
          } else {
#pragma unroll
            for (int a = 0; a < kF; ++a)
#pragma unroll
              for (int b = 0; b <= a; ++b) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < kF; ++r) acc = fma(E[r][a], Z[r][b], acc);
                C[tri(a, b)] = acc;
              }
          }
        }
        __threadfence_block();
        named_barrier_arrive(j, n_threads);      // S^-1_j and Z_j of both sides are published
      }
      // middle block in the coordinates of the top-down lane (both lanes, identical arithmetic)
#pragma unroll
      for (int a = 0; a < kF; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b)
          if ((a + b) & 1) C[tri(a, b)] = side ? -C[tri(a, b)] : C[tri(a, b)];
      const int m = mA + 1;
      TimePowers ta, tb;
      ta.set(my_time[m - 1]);
      tb.set(my_time[m]);
      double Sm[10];
      fast::diag_block(ta, tb, Sm);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const double other = __shfl_xor_sync(0xffffffffu, C[i], 1);
        const double cA = side ? other : C[i];
        const double cB = side ? C[i] : other;
        Sm[i] = (Sm[i] - cA) - cB;
      }
      double L[10], inv[4], Si[10];
      if (!fast::ldlt4(Sm, L, inv)) status |= 1;
      ldlt4_inverse(L, inv, Si);
      double* sb = slot + (lay.s_off + mA * 10) * kStride;
#pragma unroll
      for (int i = 0; i < 10; ++i) sb[i * kStride] = Si[i];
      __threadfence_block();
      named_barrier_arrive(mA + 1, n_threads);
    }
    reinterpret_cast<int*>(red + 2 * D * 32)[lane] = status;
  } else {
    // =========================== DIMENSION warp ===============================================
    const int d = warp - 1;
    auto local_p = [&](int j) { return my_pos[(side ? K - j : j) * D + d]; };
    const double* bd_src = nullptr;   // boundary derivatives of the lane's end of the chain, this dimension
    if (p.end_derivatives && active)
      bd_src = p.end_derivatives + ((p.sweep_S > 0 ? prob / p.sweep_S : prob) * 2 + side) * (kF * D) + d;
    auto bd = [&](int a) { return bd_src ? flip[a] * bd_src[a * D] : 0.0; };
    // P[a+1] = T^(a-6), a = 0..3: the time powers the right-hand side needs
    auto rhs_powers = [&](double T, double (&pw)[kF]) {
      const double i1 = fast_rcp(T);
      const double i2 = i1 * i1, i3 = i2 * i1;
      pw[3] = i3; pw[2] = i2 * i2; pw[1] = pw[2] * i1; pw[0] = i3 * i3;
    };
    auto rhs_block = [&](const double (&pprev)[kF], const double (&pnext)[kF], double dprev, double dnext,
                         double (&out)[kF]) {
#pragma unroll
      for (int a = 0; a < kF; ++a)
        out[a] = -fma(H1T(6 + a, 5) * pprev[a], dprev, H1T(1 + a, 5) * pnext[a] * dnext);
    };
    double xm[kF] = {0.0, 0.0, 0.0, 0.0};   // middle block solution, local coordinates
    double* gx = slot + (lay.g_off + d * 4 * mA) * kStride;   // g_j, later x_j, of this dimension: gx[((j-1)*4 + a) * kStride]
    if (nb > 0) {
      double pw_prev[kF], pw_next[kF], g[kF];
      rhs_powers(local_T(0), pw_prev);
      rhs_powers(local_T(1), pw_next);
      double dp_prev = local_p(1) - local_p(0);
      double dp_next = local_p(2) - local_p(1);
      rhs_block(pw_prev, pw_next, dp_prev, dp_next, g);
      if (bd_src) {
        TimePowers t0;
        t0.set(local_T(0));
        double E0[kF][kF];
        fast::coupling_block(t0, E0);
#pragma unroll
        for (int b = 0; b < kF; ++b)
#pragma unroll
          for (int a = 0; a < kF; ++a) g[b] = fma(-E0[a][b], bd(a), g[b]);
      }
      double c[kF] = {0.0, 0.0, 0.0, 0.0};
      for (int j = 1; j <= mA; ++j) {
        if (j <= my_n) {
#pragma unroll
          for (int a = 0; a < kF; ++a) gx[((j - 1) * 4 + a) * kStride] = g[a];
        }
        named_barrier_sync(j, n_threads);        // Z_j is visible
        if (j <= my_n) {
          const double* zb = slot + (lay.z_off + (j - 1) * 16) * kStride;
          double zt[kF] = {0.0, 0.0, 0.0, 0.0};   // Z_j^T g_j
#pragma unroll
          for (int r = 0; r < kF; ++r)
#pragma unroll
            for (int a = 0; a < kF; ++a) zt[a] = fma(zb[(r * kF + a) * kStride], g[r], zt[a]);
          if (j < my_n) {
#pragma unroll
            for (int a = 0; a < kF; ++a) pw_prev[a] = pw_next[a];
            rhs_powers(local_T(j + 1), pw_next);
            dp_prev = dp_next;
            dp_next = local_p(j + 2) - local_p(j + 1);
            rhs_block(pw_prev, pw_next, dp_prev, dp_next, g);
#pragma unroll
            for (int a = 0; a < kF; ++a) g[a] -= zt[a];
          } else {
#pragma unroll
            for (int a = 0; a < kF; ++a) c[a] = zt[a];
          }
        }
      }
      // middle block
#pragma unroll
      for (int a = 0; a < kF; ++a) c[a] *= flip[a];
      const int m = mA + 1;
      double gm[kF];
      {
        double pa[kF], pb[kF];
        rhs_powers(my_time[m - 1], pa);
        rhs_powers(my_time[m], pb);
        rhs_block(pa, pb, my_pos[m * D + d] - my_pos[(m - 1) * D + d], my_pos[(m + 1) * D + d] - my_pos[m * D + d], gm);
        if (p.end_derivatives && active) {
          const long rec = p.sweep_S > 0 ? prob / p.sweep_S : prob;
          if (m - 1 == 0) {
            const double* src = p.end_derivatives + (rec * 2 + 0) * (kF * D) + d;
            TimePowers ta;
            ta.set(my_time[m - 1]);
            double E0[kF][kF];
            fast::coupling_block(ta, E0);
#pragma unroll
            for (int b = 0; b < kF; ++b)
#pragma unroll
              for (int a = 0; a < kF; ++a) gm[b] = fma(-E0[a][b], src[a * D], gm[b]);
          }
          if (m + 1 == K) {
            const double* src = p.end_derivatives + (rec * 2 + 1) * (kF * D) + d;
            TimePowers tb;
            tb.set(my_time[m]);
            double E1[kF][kF];
            fast::coupling_block(tb, E1);
#pragma unroll
            for (int a = 0; a < kF; ++a)
#pragma unroll
              for (int b = 0; b < kF; ++b) gm[a] = fma(-E1[a][b], src[b * D], gm[a]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < kF; ++a) {
        const double other = __shfl_xor_sync(0xffffffffu, c[a], 1);
        const double cA = side ? other : c[a];
        const double cB = side ? c[a] : other;
        gm[a] = (gm[a] - cA) - cB;
      }
      named_barrier_sync(mA + 1, n_threads);     // S_m^-1 is visible
      {
        const double* sb = slot + (lay.s_off + mA * 10) * kStride;
        double Si[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) Si[i] = sb[i * kStride];
#pragma unroll
        for (int a = 0; a < kF; ++a) {
          double acc = sym(Si, a, 0) * gm[0];
#pragma unroll
          for (int r = 1; r < kF; ++r) acc = fma(sym(Si, a, r), gm[r], acc);
          xm[a] = flip[a] * acc;
        }
      }
      if (p.free_out && active && side == 0 && p.sweep_S == 0) {
        double* dst = p.free_out + (prob * (long)nb + mA) * (kF * D) + d;
#pragma unroll
        for (int a = 0; a < kF; ++a) dst[a * D] = xm[a];
      }
      // back substitution outwards: x_j = S_j^-1 g_j - Z_j x_{j+1}, stored over g_j
      double x_far[kF] = {xm[0], xm[1], xm[2], xm[3]};
      for (int j = mA; j >= 1; --j) {
        const int jj = side ? j - (mA - nB) : j;
        if (jj >= 1) {
          const double* sb = slot + (lay.s_off + (jj - 1) * 10) * kStride;
          const double* zb = slot + (lay.z_off + (jj - 1) * 16) * kStride;
          double Si[10], gj[kF], x[kF];
#pragma unroll
          for (int i = 0; i < 10; ++i) Si[i] = sb[i * kStride];
#pragma unroll
          for (int a = 0; a < kF; ++a) gj[a] = gx[((jj - 1) * 4 + a) * kStride];
#pragma unroll
          for (int a = 0; a < kF; ++a) {
            double acc = sym(Si, a, 0) * gj[0];
#pragma unroll
            for (int r = 1; r < kF; ++r) acc = fma(sym(Si, a, r), gj[r], acc);
#pragma unroll
            for (int b = 0; b < kF; ++b) acc = fma(-zb[(a * kF + b) * kStride], x_far[b], acc);
            x[a] = acc;
          }
#pragma unroll
          for (int a = 0; a < kF; ++a) {
            gx[((jj - 1) * 4 + a) * kStride] = x[a];
            x_far[a] = x[a];
          }
          if (p.free_out && active && p.sweep_S == 0) {
            const int v = side ? K - jj : jj;
            double* dst = p.free_out + (prob * (long)nb + (v - 1)) * (kF * D) + d;
#pragma unroll
            for (int a = 0; a < kF; ++a) dst[a * D] = flip[a] * x[a];
          }
        }
      }
    } else if (p.end_derivatives && active) {
      // K == 1: the far end of the only segment is the other boundary
      const long rec = p.sweep_S > 0 ? prob / p.sweep_S : prob;
      const double* src = p.end_derivatives + (rec * 2 + 1) * (kF * D) + d;
#pragma unroll
      for (int a = 0; a < kF; ++a) xm[a] = src[a * D];
    }

    // ---- coefficient recovery / cost of this dimension, one local segment per step ----------
    double cost_acc = 0.0;
    int nonfinite = 0;
    const int top = nb > 0 ? mA : 0;
    for (int step = 0; step <= top; ++step) {
      const int j = top - step;
      const int jj = side ? j - (mA - nB) : j;
      const bool mine = active && jj >= 0 && (nb > 0 || side == 0);
      double x_near[kF], x_far[kF];
#pragma unroll
      for (int a = 0; a < kF; ++a) {
        x_near[a] = (jj >= 1) ? gx[((jj - 1) * 4 + a) * kStride] : bd(a);
        x_far[a] = (jj >= 0 && jj + 1 <= my_n) ? gx[(jj * 4 + a) * kStride] : xm[a];
      }
      const int jc = jj >= 0 ? jj : 0;
      const int seg = side ? K - 1 - jc : jc;
      const double T = my_time[seg];
      if (!(T > 0.0)) status |= 2;
      double u[2 * kF + 1], ds[kF];
      {
        const double T2 = T * T, T3 = T2 * T, T4 = T2 * T2;
        const double tk[kF] = {T, T2, T3, T4};
#pragma unroll
        for (int a = 0; a < kF; ++a) {
          const double s_val = side ? flip[a] * x_far[a] : x_near[a];
          const double e_val = side ? flip[a] * x_near[a] : x_far[a];
          ds[a] = s_val;
          u[1 + a] = tk[a] * s_val;
          u[1 + kF + a] = tk[a] * e_val;
        }
        u[0] = my_pos[(seg + 1) * D + d] - my_pos[seg * D + d];
      }
      const double i1 = fast_rcp(T);
      const double i2 = i1 * i1, i4 = i2 * i2, i5 = i4 * i1;
      if (kCoeffs) {
        const double ipow[5] = {i5, i5 * i1, i5 * i2, i4 * i4, i4 * i5};
        double cf[kN];
        cf[0] = my_pos[seg * D + d];
#pragma unroll
        for (int a = 0; a < kF; ++a) cf[1 + a] = A1T(1 + a, 1 + a) * ds[a];
#pragma unroll
        for (int i = 5; i < kN; ++i) {
          double acc = A1T(i, 5) * u[0];
#pragma unroll
          for (int a = 0; a < kF; ++a) {
            acc = fma(A1T(i, 1 + a), u[1 + a], acc);
            acc = fma(A1T(i, 6 + a), u[1 + kF + a], acc);
          }
          cf[i] = acc * ipow[i - 5];
        }
        double chk = 0.0;
#pragma unroll
        for (int i = 0; i < kN; ++i) chk = fma(cf[i], 0.0, chk);
        if (chk != 0.0) nonfinite = 1;
        if (mine) {
          double* dst = p.coeffs + ((prob * K + seg) * D + d) * kN;
          if (p.aligned16) {
#pragma unroll
            for (int i = 0; i < kN; i += 2) __stcs(reinterpret_cast<double2*>(dst + i), make_double2(cf[i], cf[i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < kN; ++i) __stcs(dst + i, cf[i]);
          }
        }
      }
      if (p.cost && mine) {
        const double i7 = i5 * i2;
        double qd = 0.0;
#pragma unroll
        for (int r = 0; r < 2 * kF + 1; ++r) {
          const int hr = r == 0 ? 5 : (r <= kF ? r : r + 1);
          double row = 0.0;
#pragma unroll
          for (int s = 0; s < 2 * kF + 1; ++s) {
            const int hs = s == 0 ? 5 : (s <= kF ? s : s + 1);
            row = fma(H1T(hr, hs), u[s], row);
          }
          qd = fma(row, u[r], qd);
        }
        cost_acc = fma(qd, i7, cost_acc);
      }
    }
    if (nonfinite) status |= 4;
    red[d * 32 + lane] = cost_acc;
    reinterpret_cast<int*>(red + 2 * D * 32)[warp * 32 + lane] = status;
  }

  // ---- combine the per-dimension cost terms and status words (fixed order: deterministic) -----
  __syncthreads();
  if (warp == 0) {
    const int* st = reinterpret_cast<const int*>(red + 2 * D * 32);
    int s_all = 0;
#pragma unroll
    for (int w = 0; w <= D; ++w) s_all |= st[w * 32 + lane];
    s_all |= __shfl_xor_sync(0xffffffffu, s_all, 1);
    double c_all = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) c_all += red[d * 32 + lane];
    c_all += __shfl_xor_sync(0xffffffffu, c_all, 1);
    if (active && side == 0) {
      if (p.cost) p.cost[prob] = 0.5 * c_all;
      if (p.status) p.status[prob] = s_all;
    }
  }
}

#undef H1T
#undef A1T

template <int D, bool kCoeffs>
inline cudaError_t launch_d(const FastParams& p, cudaStream_t stream) {
  const Layout<D> lay(p.K);
  const size_t smem = (size_t)lay.total * sizeof(double);
  if (smem > kMaxDynamicSmem) return cudaErrorInvalidConfiguration;
  auto kernel = solve_standard_ws_kernel<D, kCoeffs>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long n_problems = p.sweep_S > 0 ? p.B * p.sweep_S : p.B;
  const long grid = (n_problems + kPairsPerWarp - 1) / kPairsPerWarp;
  if (grid > 2147483647L) return cudaErrorInvalidConfiguration;
  kernel<<<(unsigned)grid, (D + 1) * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

// named barriers 1..mA+1 must exist (16 per CTA, 0 is __syncthreads)
inline bool supported(int K, int D, int N, int derivative) {
  return fast::supported(K, D, N, derivative) && (K - 1) / 2 + 1 <= 15;
}

}  // namespace ws
}  // namespace minsnap
