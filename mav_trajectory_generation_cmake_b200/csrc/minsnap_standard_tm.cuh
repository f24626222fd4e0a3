// Standard-mask solve, second generation (N = 10, snap, even K <= 10): the thread-pair
// elimination of minsnap_standard_fast.cuh with the two costs ncu attributed to it removed --
// the per-lane coefficient stores (every lane of a store instruction in its own 128-byte line)
// and the lane-serial coefficient recovery behind them.
//
//  * Solve phase (a thread PAIR per trajectory, 16 trajectories per warp): unchanged algebra --
//    lane 2q eliminates top-down, lane 2q+1 bottom-up, they meet at the middle block.  The Z (4x4)
//    of every eliminated block is lane-private scratch, written once in the forward sweep and read
//    once in the back substitution: it now lives in TENSOR MEMORY (tcgen05.st / tcgen05.ld, shape
//    32x32b: thread i of a warp owns TMEM lane 32 (warp % 4) + i, a block is 32 consecutive 32-bit
//    columns).  No MMA is involved: TMEM is used as what it physically is, a 128-lane x 512-column
//    register-file extension next to the SM, and it frees the shared memory the next phase needs.
//    w -> x of every interior vertex goes to shared memory as X[trajectory][vertex][derivative][dim]
//    in actual orientation.
//  * Recovery phase (a thread per SEGMENT, in output order): task t = 32 i + lane of a warp is
//    segment t mod K of trajectory t / K, so the 32 tasks of an iteration produce 32 x 80 D
//    CONTIGUOUS bytes of the coefficient array.  Each lane evaluates c = A^-1 d for its segment
//    one dimension at a time (10 live coefficients), parks them in a shared-memory tile, and one
//    elected lane hands the whole tile to the TMA (cp.async.bulk.global.shared::cta, SASS UBLKCP):
//    5 bulk copies of 7,680 bytes per 16 trajectories instead of 640 store instructions.  The
//    9-term dot products of 32 independent segments run at FP64 throughput, not at the latency of
//    one lane's dependency chain.
//
// Control flow is uniform across the warp (even K: both lanes of a pair eliminate the same
// number of blocks), which the warp-collective tcgen05.ld/st require.  Other shapes (odd K,
// K = 1, K > 10, the cost sweep, unaligned outputs) keep the first-generation kernel.
//
// Arithmetic is that of minsnap_standard_fast.cuh (same closed-form blocks, same 2x2-Schur inverse,
// same summation orders).
#pragma once
#include "minsnap_standard_fast.cuh"

namespace minsnap {
namespace tm {

using fast::kF;
using fast::kN;
using fast::kPairsPerWarp;
using fast::FastParams;
using fast::TimePowers;
using fast::tri;

constexpr int kWarpsPerCta = 4;   // one warp per TMEM lane quarter
constexpr int kMaxK = 10;

#define H1T(r, s) (minsnap_tables::kH1_N10_d4[(r) * 10 + (s)])
#define A1T(i, r) (minsnap_tables::kA1inv_N10[(i) * 10 + (r)])

// ---- tensor memory -----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_slot));
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one 4x4 block = 16 doubles = 32 columns of the calling thread's TMEM lane
__device__ __forceinline__ void tmem_store(uint32_t taddr, const double (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :: "r"(taddr), "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])), "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])), "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])), "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7])), "r"(__double2loint(v[8])), "r"(__double2hiint(v[8])), "r"(__double2loint(v[9])), "r"(__double2hiint(v[9])), "r"(__double2loint(v[10])), "r"(__double2hiint(v[10])), "r"(__double2loint(v[11])), "r"(__double2hiint(v[11])), "r"(__double2loint(v[12])), "r"(__double2hiint(v[12])), "r"(__double2loint(v[13])), "r"(__double2hiint(v[13])), "r"(__double2loint(v[14])), "r"(__double2hiint(v[14])), "r"(__double2loint(v[15])), "r"(__double2hiint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_load(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ double pair_to_double(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// ---- bulk copy shared -> global (TMA, non-tensor form) --------------------------------------
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, uint32_t bytes) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline int stored_blocks(int K) {
  const int mA = (K - 1) / 2;
  return mA > 0 ? mA - 1 : 0;
}
inline int tmem_columns(int K) {
  const int need = stored_blocks(K) * 32;
  int cols = 32;
  while (cols < need) cols <<= 1;
  return cols;
}
// per-warp shared memory, in doubles: positions [16][(K+1) D], times [16][K], X [4 D][16 K + 1], the
// copy-out tile [32][10 D], 16 status words.
// X is element-major: entry e = (derivative, dimension) of the vertex that STARTS segment t = traj K + seg of
// the batch sits at X[e][t]; the vertex that ends it at X[e][t + 1].  Column traj K (vertex 0 of a trajectory,
// which is also the end vertex K of the trajectory before it) holds the end derivatives when they are zero,
// i.e. zeros that are written once.  Consecutive lanes of the recovery phase read consecutive words.
template <int D>
struct WarpSmem {
  size_t pos, tim, x, x_pitch, tile, flags, total;
  __host__ __device__ explicit WarpSmem(int K) {
    size_t o = 0;
    pos = o; o += ((size_t)kPairsPerWarp * (K + 1) * D + 1) & ~(size_t)1;
    tim = o; o += ((size_t)kPairsPerWarp * K + 1) & ~(size_t)1;
    x_pitch = (size_t)kPairsPerWarp * K + 1;
    x = o; o += ((size_t)kF * D * x_pitch + 1) & ~(size_t)1;
    tile = o; o += (size_t)32 * D * kN;
    flags = o; o += 8;
    total = o;
  }
};

template <int D, bool kCost, bool kBoundary>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) solve_standard_tm_kernel(FastParams p, int tmem_cols) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint32_t tmem_base_slot;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int K = p.K;
  const int per_pos = (K + 1) * D;
  constexpr int kVec = kF * D;    // doubles per vertex vector
  constexpr int kTile = D * kN;   // doubles per segment

  // tensor memory: one allocation per CTA, every warp works in its own lane quarter
  if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)tmem_cols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp & 3) << 21);   // lane field (bits 31..16) = 32 (warp % 4)

  const WarpSmem<D> lay(K);
  double* wbase = smem + (size_t)warp * lay.total;
  double* pos_s = wbase + lay.pos;      // [16][per_pos]
  double* time_s = wbase + lay.tim;     // [16][K]; a segment's entry is replaced by its cost term once recovered
  double* x_s = wbase + lay.x;          // [kVec][16 K + 1]: w of stored blocks, then x of every interior vertex
  double* tile_s = wbase + lay.tile;    // [32][kTile]
  int* flags_s = reinterpret_cast<int*>(wbase + lay.flags);   // [16] status bits found by the recovery phase
  const int x_pitch = (int)lay.x_pitch;

  const int side = lane & 1;   // 0: top-down lane, 1: bottom-up lane
  const int q = lane >> 1;     // pair index within the warp
  const int nb = K - 1;        // unknown blocks (odd)
  const int mA = nb / 2;       // blocks eliminated by either lane; the middle block is vertex mA + 1
  const double flip[kF] = {side ? -1.0 : 1.0, 1.0, side ? -1.0 : 1.0, 1.0};   // (-1)^k, k = 1..4, bottom-up lane
  const unsigned inv_k = 65536u / (unsigned)K + 1u;   // t / K == (t * inv_k) >> 16 for t < 16 K (K <= 16)
  // the boundary columns of X (vertex 0 of every trajectory, and the one past the last) stay zero
  for (int e = lane; e < kVec * (kPairsPerWarp + 1); e += kWarp) x_s[(e / (kPairsPerWarp + 1)) * x_pitch + (e % (kPairsPerWarp + 1)) * K] = 0.0;

  const long pairs_per_cta = (long)kWarpsPerCta * kPairsPerWarp;
  const long stride = (long)gridDim.x * pairs_per_cta;
  for (long base = (long)blockIdx.x * pairs_per_cta + (long)warp * kPairsPerWarp; base < p.B; base += stride) {
    const int n_here = (int)min((long)kPairsPerWarp, p.B - base);
    const long prob = base + q;
    const bool active = q < n_here;
    // ---- inputs: cp.async, every chunk in flight before the single wait ------------------------
    __syncwarp();   // the previous batch's readers are done with the inputs
    fast::async_copy_doubles(pos_s, p.positions + base * per_pos, n_here * per_pos, lane, p.aligned16);
    if (p.times) fast::async_copy_doubles(time_s, p.times + base * K, n_here * K, lane, p.aligned16);
    __pipeline_commit();
    if (lane < kPairsPerWarp) flags_s[lane] = 0;
    __pipeline_wait_prior(0);
    __syncwarp();
    if (!p.times) {
      // ref estimateSegmentTimes (src/vertex.cpp:162-178), same expression as minsnap_estimate_segment_times
      for (int e = lane; e < n_here * K; e += kWarp) {
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
        time_s[r * K + o] = T;
        if (p.times_out) p.times_out[(base + r) * K + o] = T;
      }
      __syncwarp();
    }
    // idle pairs of a ragged last batch run on the first pair's inputs (nothing of theirs is stored)
    const int qi = active ? q : 0;
    const double* my_pos = pos_s + qi * per_pos;
    const double* my_time = time_s + qi * K;
    double* my_x = x_s + q * K;   // column of vertex v of this trajectory: my_x[v], entry e at + e x_pitch
    // local chain: vertex j <-> actual vertex (side ? K - j : j); segment j <-> actual (side ? K-1-j : j)
    auto local_T = [&](int j) { return my_time[side ? K - 1 - j : j]; };
    auto local_p = [&](int j, int d) { return my_pos[(side ? K - j : j) * D + d]; };
    auto x_slot = [&](int j) { return my_x + (side ? K - j : j); };   // local vertex j (1 <= j <= K-1)

    int status = 0;
    const double* bd_src = nullptr;   // boundary derivatives of the lane's end of the chain (actual coordinates)
    if (kBoundary && p.end_derivatives && active) bd_src = p.end_derivatives + (prob * 2 + side) * kVec;
    auto bd = [&](int a, int d) { return bd_src ? flip[a] * bd_src[a * D + d] : 0.0; };

    {
      double xm[kF][D];             // middle block solution, local coordinates
      double Z[kF][kF], w[kF][D];   // after the forward sweep: the lane's LAST block (never leaves the registers)
#pragma unroll
      for (int a = 0; a < kF; ++a) {
#pragma unroll
        for (int b = 0; b < kF; ++b) Z[a][b] = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) w[a][d] = 0.0;
      }
      auto rhs_block = [&](const TimePowers& tprev, const TimePowers& tnext, const double (&dprev)[D],
                           const double (&dnext)[D], double (&out)[kF][D]) {
#pragma unroll
        for (int a = 0; a < kF; ++a) {
          const double ce = H1T(6 + a, 5) * tprev.P[a + 1];   // end-free row of the previous segment
          const double cs = H1T(1 + a, 5) * tnext.P[a + 1];   // start-free row of the next segment
#pragma unroll
          for (int d = 0; d < D; ++d) out[a][d] = -fma(ce, dprev[d], cs * dnext[d]);
        }
      };
      TimePowers tp_prev, tp_next;
      // ---- forward sweep ------------------------------------------------------------------
      double S[10], g[kF][D];
      double dp_prev[D], dp_next[D];
      tp_prev.set(local_T(0));
      tp_next.set(local_T(1));
#pragma unroll
      for (int d = 0; d < D; ++d) {
        dp_prev[d] = local_p(1, d) - local_p(0, d);
        dp_next[d] = local_p(2, d) - local_p(1, d);
      }
      fast::diag_block(tp_prev, tp_next, S);
      rhs_block(tp_prev, tp_next, dp_prev, dp_next, g);
      if (bd_src) {
        double E0[kF][kF];
        fast::coupling_block(tp_prev, E0);
#pragma unroll
        for (int b = 0; b < kF; ++b)
#pragma unroll
          for (int d = 0; d < D; ++d) {
            double acc = g[b][d];
#pragma unroll
            for (int a = 0; a < kF; ++a) acc = fma(-E0[a][b], bd(a, d), acc);
            g[b][d] = acc;
          }
      }
      for (int j = 1; j <= mA; ++j) {
        // here: tp_prev = segment j-1, tp_next = segment j, S/g = reduced block j
        const bool more = j < mA;
        TimePowers tp_new;
        double dp_new[D];
        tp_new.set(local_T(more ? j + 1 : j));
#pragma unroll
        for (int d = 0; d < D; ++d) dp_new[d] = more ? local_p(j + 2, d) - local_p(j + 1, d) : 0.0;
        double Si[10];
        if (!fast::spd4_inverse(S, Si)) status |= 1;
        double E[kF][kF];
        fast::coupling_block(tp_next, E);
#pragma unroll
        for (int b = 0; b < kF; ++b) {
          const double in[4] = {E[0][b], E[1][b], E[2][b], E[3][b]};
          double col[4];
          fast::sym4_apply(Si, in, col);
#pragma unroll
          for (int a = 0; a < kF; ++a) Z[a][b] = col[a];
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double in[4] = {g[0][d], g[1][d], g[2][d], g[3][d]};
          double col[4];
          fast::sym4_apply(Si, in, col);
#pragma unroll
          for (int a = 0; a < kF; ++a) w[a][d] = col[a];
        }
        if (more) {
          // Z of block j goes to tensor memory, w to the vertex's slot of X (x replaces it later)
          {
            const double zv[16] = {Z[0][0], Z[0][1], Z[0][2], Z[0][3], Z[1][0], Z[1][1], Z[1][2], Z[1][3],
                                   Z[2][0], Z[2][1], Z[2][2], Z[2][3], Z[3][0], Z[3][1], Z[3][2], Z[3][3]};
            tmem_store(taddr + (uint32_t)((j - 1) * 32), zv);
          }
          double* ws = x_slot(j);
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) ws[(a * D + d) * x_pitch] = w[a][d];
          // advance to block j+1: D_{j+1} - E^T Z,  b_{j+1} - E^T w
          tp_prev = tp_next;
          tp_next = tp_new;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            dp_prev[d] = dp_next[d];
            dp_next[d] = dp_new[d];
          }
          fast::diag_block(tp_prev, tp_next, S);
          rhs_block(tp_prev, tp_next, dp_prev, dp_next, g);
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
              double acc = S[tri(a, b)];
#pragma unroll
              for (int r = 0; r < kF; ++r) acc = fma(-E[r][a], Z[r][b], acc);
              S[tri(a, b)] = acc;
            }
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) {
              double acc = g[a][d];
#pragma unroll
              for (int r = 0; r < kF; ++r) acc = fma(-E[r][a], w[r][d], acc);
              g[a][d] = acc;
            }
        }
      }
      if (mA > 1) tmem_wait_st();   // the stores are complete before the back substitution loads them

      // Schur contribution of this lane to the middle block, from its last eliminated block
      // (tp_next holds the powers of the lane's last local segment): C = E^T Z, c = E^T w
      double C[10], c[kF][D];
#pragma unroll
      for (int i = 0; i < 10; ++i) C[i] = 0.0;
#pragma unroll
      for (int a = 0; a < kF; ++a)
#pragma unroll
        for (int d = 0; d < D; ++d) c[a][d] = 0.0;
      if (mA >= 1) {
        double E[kF][kF];
        fast::coupling_block(tp_next, E);
#pragma unroll
        for (int r = 0; r < kF; ++r)
#pragma unroll
          for (int a = 0; a < kF; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) C[tri(a, b)] = fma(E[r][a], Z[r][b], C[tri(a, b)]);
#pragma unroll
            for (int d = 0; d < D; ++d) c[a][d] = fma(E[r][a], w[r][d], c[a][d]);
          }
      }
      // ---- middle block, solved by both lanes in the coordinates of the top-down lane -------
#pragma unroll
      for (int a = 0; a < kF; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b)
          if ((a + b) & 1) C[tri(a, b)] = side ? -C[tri(a, b)] : C[tri(a, b)];
#pragma unroll
        for (int d = 0; d < D; ++d) c[a][d] *= flip[a];
      }
      double Sm[10], gm[kF][D];
      {
        const int m = mA + 1;   // actual middle vertex
        TimePowers ta, tb;
        ta.set(my_time[m - 1]);
        tb.set(my_time[m]);
        double da[D], db[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          da[d] = my_pos[m * D + d] - my_pos[(m - 1) * D + d];
          db[d] = my_pos[(m + 1) * D + d] - my_pos[m * D + d];
        }
        fast::diag_block(ta, tb, Sm);
        rhs_block(ta, tb, da, db, gm);
        if (kBoundary && p.end_derivatives && active && K == 2) {
          // both boundary couplings reach the middle block directly
          const double* src = p.end_derivatives + (prob * 2 + 0) * kVec;
          double E0[kF][kF];
          fast::coupling_block(ta, E0);
#pragma unroll
          for (int b = 0; b < kF; ++b)
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
              for (int a = 0; a < kF; ++a) gm[b][d] = fma(-E0[a][b], src[a * D + d], gm[b][d]);
          src += kVec;
          double E1[kF][kF];
          fast::coupling_block(tb, E1);
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
              for (int b = 0; b < kF; ++b) gm[a][d] = fma(-E1[a][b], src[b * D + d], gm[a][d]);
        }
      }
      {
        // subtract the top-down contribution first, then the bottom-up one, on both lanes
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const double other = __shfl_xor_sync(0xffffffffu, C[i], 1);
          const double cA = side ? other : C[i];
          const double cB = side ? C[i] : other;
          Sm[i] = (Sm[i] - cA) - cB;
        }
#pragma unroll
        for (int a = 0; a < kF; ++a)
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const double other = __shfl_xor_sync(0xffffffffu, c[a][d], 1);
            const double cA = side ? other : c[a][d];
            const double cB = side ? c[a][d] : other;
            gm[a][d] = (gm[a][d] - cA) - cB;
          }
        double Si[10];
        if (!fast::spd4_inverse(Sm, Si)) status |= 1;
        double* ms_ = my_x + (mA + 1);   // the middle vertex mA + 1, actual coordinates = the top-down lane's
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double in[4] = {gm[0][d], gm[1][d], gm[2][d], gm[3][d]};
          double col[4];
          fast::sym4_apply(Si, in, col);
#pragma unroll
          for (int a = 0; a < kF; ++a) {
            xm[a][d] = flip[a] * col[a];   // -> local coordinates
            if (side == 0) ms_[(a * D + d) * x_pitch] = col[a];
          }
        }
      }
      if (p.free_out && active && side == 0) {
        double* dst = p.free_out + (prob * (long)nb + mA) * kVec;
#pragma unroll
        for (int a = 0; a < kF; ++a)
#pragma unroll
          for (int d = 0; d < D; ++d) dst[a * D + d] = xm[a][d];
      }

      // ---- back substitution outwards: x_j = w_j - Z_j x_{j+1}, stored over w_j in actual orientation ----
      for (int j = mA; j >= 1; --j) {
        double* xs = x_slot(j);
        if (j < mA) {
          uint32_t z[32];
          tmem_load(taddr + (uint32_t)((j - 1) * 32), z);
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) w[a][d] = xs[(a * D + d) * x_pitch];
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) Z[e / 4][e % 4] = pair_to_double(z[2 * e], z[2 * e + 1]);
        }
        double x_near[kF][D];
#pragma unroll
        for (int a = 0; a < kF; ++a)
#pragma unroll
          for (int d = 0; d < D; ++d) {
            double acc = w[a][d];
#pragma unroll
            for (int b = 0; b < kF; ++b) acc = fma(-Z[a][b], xm[b][d], acc);
            x_near[a][d] = acc;
          }
#pragma unroll
        for (int a = 0; a < kF; ++a)
#pragma unroll
          for (int d = 0; d < D; ++d) {
            xm[a][d] = x_near[a][d];   // becomes the far vector of the next step
            xs[(a * D + d) * x_pitch] = flip[a] * x_near[a][d];
          }
        if (p.free_out && active) {
          const int v = side ? K - j : j;   // actual vertex
          double* dst = p.free_out + (prob * (long)nb + (v - 1)) * kVec;
#pragma unroll
          for (int a = 0; a < kF; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) dst[a * D + d] = flip[a] * x_near[a][d];
        }
      }
    }
    status |= __shfl_xor_sync(0xffffffffu, status, 1);
    __syncwarp();   // X is complete

    // ---- recovery (ref updateSegmentsFromCompactConstraints, LIN.i:252-273): one thread per segment, in
    //      output order; task t = 32 it + lane is segment t mod K of trajectory t / K of this batch ----
    const int n_tasks = n_here * K;
    for (int it = 0; it * 32 < kPairsPerWarp * K; ++it) {
      const int t = it * 32 + lane;
      const int traj = (int)(((unsigned)t * inv_k) >> 16);
      const int seg = t - traj * K;
      const bool valid = t < n_tasks;
      const double T = time_s[t];
      const double* tp = pos_s + (t + traj) * D;   // positions of vertices seg, seg + 1 of trajectory traj
      const double* xp = x_s + t;                   // start vertex at xp[e x_pitch], end vertex one word further
      int flags = (T > 0.0) ? 0 : 2;   // MINSNAP_STATUS_BAD_TIME
      const double T2 = T * T, T3 = T2 * T, T4 = T2 * T2;
      const double tk[kF] = {T, T2, T3, T4};
      const double i1 = fast::fast_rcp(T);
      const double i2 = i1 * i1, i4 = i2 * i2, i5 = i4 * i1;
      const double ipow[5] = {i5, i5 * i1, i5 * i2, i4 * i4, i4 * i5};   // T^-5 .. T^-9
      double u[2 * kF + 1][D];   // [dp, T^k start_k (k = 1..4), T^k end_k (k = 1..4)]
      double cf[D][kN];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double p0 = tp[d];
        u[0][d] = tp[D + d] - p0;
        cf[d][0] = p0;
      }
#pragma unroll
      for (int a = 0; a < kF; ++a)
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double s_val = xp[(a * D + d) * x_pitch];
          double e_val = xp[(a * D + d) * x_pitch + 1];
          if (kBoundary && p.end_derivatives && valid) {
            // end derivatives given by the caller (global memory): the boundary columns of X hold zeros
            if (seg == 0) s_val = p.end_derivatives[((base + traj) * 2 + 0) * kVec + a * D + d];
            if (seg == K - 1) e_val = p.end_derivatives[((base + traj) * 2 + 1) * kVec + a * D + d];
          }
          u[1 + a][d] = tk[a] * s_val;
          u[1 + kF + a][d] = tk[a] * e_val;
          cf[d][1 + a] = A1T(1 + a, 1 + a) * s_val;
        }
      // row-outer, dimension-inner: every table constant is fetched once per segment, not once per dimension
#pragma unroll
      for (int i = 5; i < kN; ++i) {
        double acc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = A1T(i, 5) * u[0][d];
#pragma unroll
        for (int a = 0; a < kF; ++a) {
#pragma unroll
          for (int d = 0; d < D; ++d) acc[d] = fma(A1T(i, 1 + a), u[1 + a][d], acc[d]);
#pragma unroll
          for (int d = 0; d < D; ++d) acc[d] = fma(A1T(i, 6 + a), u[1 + kF + a][d], acc[d]);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) cf[d][i] = acc[d] * ipow[i - 5];
      }
      // non-finite detection on the exponent fields of c_9 (scaled by T^-9: overflows first) and c_4
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int e9 = __double2hiint(cf[d][kN - 1]) & 0x7ff00000, e4 = __double2hiint(cf[d][kF]) & 0x7ff00000;
        if (e9 == 0x7ff00000 || e4 == 0x7ff00000) flags |= 4;
      }
      // the tile is free once the previous bulk copy has read it
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
      double2* t2 = reinterpret_cast<double2*>(tile_s + lane * kTile);
#pragma unroll
      for (int e = 0; e < kTile; e += 2) t2[e / 2] = make_double2(cf[e / kN][e % kN], cf[(e + 1) / kN][(e + 1) % kN]);
      if (kCost) {
        // u^T Hred1 u * T^-7: the per-segment quadratic form in scaled variables
        double qsum = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double qd = 0.0;
#pragma unroll
          for (int r = 0; r < 2 * kF + 1; ++r) {
            const int hr = r == 0 ? 5 : (r <= kF ? r : r + 1);   // row of H1: dp -> 5, start k -> k, end k -> 5 + k
            double row = 0.0;
#pragma unroll
            for (int s = 0; s < 2 * kF + 1; ++s) {
              const int hs = s == 0 ? 5 : (s <= kF ? s : s + 1);
              row = fma(H1T(hr, hs), u[s][d], row);
            }
            qd = fma(row, u[r][d], qd);
          }
          qsum += qd;
        }
        time_s[t] = qsum * (i5 * i2);   // only this task read the entry
      }
      if (flags && valid) atomicOr(flags_s + traj, flags);
      // the 32 segments of this iteration are 32 x 80 D contiguous bytes of the coefficient array
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int n_valid = min(32, n_tasks - it * 32);
        if (n_valid > 0)
          bulk_store(p.coeffs + (base * K + (long)it * 32) * kTile, tile_s, (uint32_t)(n_valid * kTile * sizeof(double)));
        bulk_commit();
      }
    }
    __syncwarp();
    if (kCost && p.cost && lane < n_here) {
      double acc = 0.0;
      for (int s = 0; s < K; ++s) acc += time_s[lane * K + s];
      p.cost[base + lane] = 0.5 * acc;
    }
    if (p.status && active && side == 0) p.status[prob] = status | flags_s[q];
  }
  if (lane == 0) bulk_wait_read_all();   // shared memory stays valid until the last copy has read it
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

#undef H1T
#undef A1T

inline bool supported(int K, int D, int N, int derivative) {
  return N == 10 && derivative == 4 && D >= 1 && D <= 3 && K >= 2 && K <= kMaxK && (K % 2) == 0;
}

template <int D>
inline cudaError_t launch_d(FastParams p, cudaStream_t stream) {
  const size_t smem = WarpSmem<D>(p.K).total * sizeof(double) * kWarpsPerCta;
  if (smem > kMaxDynamicSmem) return cudaErrorInvalidConfiguration;
  const int cols = tmem_columns(p.K);
  void (*kernel)(FastParams, int) =
      p.end_derivatives ? (p.cost ? solve_standard_tm_kernel<D, true, true> : solve_standard_tm_kernel<D, false, true>)
                        : (p.cost ? solve_standard_tm_kernel<D, true, false> : solve_standard_tm_kernel<D, false, false>);
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long per_cta = (long)kWarpsPerCta * kPairsPerWarp;
  long grid = (p.B + per_cta - 1) / per_cta;
  // one batch of 16 trajectories per warp while the grid stays modest (the hardware CTA scheduler
  // balances the SMs); very large batches loop
  const long max_grid = 148L * 64;
  if (grid > max_grid) grid = max_grid;
  kernel<<<(int)grid, kWarpsPerCta * 32, smem, stream>>>(p, cols);
  return cudaGetLastError();
}

inline cudaError_t launch(const FastParams& p, int D, cudaStream_t stream) {
  switch (D) {
    case 1: return launch_d<1>(p, stream);
    case 2: return launch_d<2>(p, stream);
    case 3: return launch_d<3>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tm
}  // namespace minsnap
