// Sampling kernels: position through snap (or any derivative range) of solved trajectories.
//
// Reference rows: a17 Polynomial::evaluate (polynomial.h:138-151),
// a18 Segment::evaluate (src/segment.cpp:51-58), a19 Trajectory::evaluate
// (src/trajectory.cpp:41-66), a20 Trajectory::evaluateRange (src/trajectory.cpp:68-128).
//
// HBM-write-bound: per sample the kernel writes n_deriv*D doubles (120 B at 5 x 3) and does
// about 2*sum_k(N-1-k)*D flops.  A CTA owns one trajectory (or a run of its samples): the
// coefficients are read from HBM once, multiplied by the falling factorials b(k,j) once
// (the product the reference forms inside every evaluate call) and kept in shared memory;
// each warp then evaluates 32 consecutive instants, transposes the 32 x (n_deriv*D) results
// through shared memory and streams them out with fully coalesced stores.
#include "minsnap_device.cuh"
#include "minsnap_launch.h"

namespace minsnap {

struct SampleKernelParams {
  long B;
  int K, D, M, n_deriv, chunks, chunk_len;
  const double* coeffs;
  const double* times;
  const double* t_in;
  long t_stride;
  double* out;
  double* t_out;
  int32_t* segment;
  bool out_aligned16 = false;
};

constexpr int kSampleThreads = 128;
constexpr int kSampleWarps = kSampleThreads / kWarp;

// First segment whose running end time exceeds t (strict), as the reference's linear scan
// (src/trajectory.cpp:45-57); K when t is at or past the end.  acc_end is non-decreasing.
__device__ inline int find_segment(const double* acc_end, int K, double t) {
  if (K <= 16) {
    int i = 0;
    while (i < K && !(acc_end[i] > t)) ++i;
    return i;
  }
  int lo = 0, hi = K;  // invariant: acc_end[lo-1] <= t, acc_end[hi] > t (hi == K: none)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (acc_end[mid] > t) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Horner evaluation of derivative k of one polynomial: ref polynomial.h:138-151 with
// bc[j] = b(k,j) * c[j] (one rounding, as the reference's row[j] * coefficients_[j]).
template <int N>
__device__ inline double horner_scaled(const double* bc, int k, double t) {
  double r = bc[N - 1];
#pragma unroll
  for (int j = N - 2; j >= 0; --j)
    if (j >= k) r = fma(r, t, bc[j]);
  return r;
}

template <int N>
__device__ inline double horner_raw(const double* c, int k, double t) {
  double r = falling_factorial(k, N - 1) * c[N - 1];
#pragma unroll
  for (int j = N - 2; j >= 0; --j)
    if (j >= k) r = fma(r, t, falling_factorial(k, j) * c[j]);
  return r;
}

// Shared memory: acc_end[K] | start[K] | stage[warps][32][n_deriv*D] | (kStaged) bc[n_deriv][K][D][N]
template <int N, bool kStaged>
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(SampleKernelParams p) {
  extern __shared__ double smem[];
  const int K = p.K, D = p.D, nd = p.n_deriv;
  const int per = nd * D;
  double* acc_end = smem;
  double* seg_start = acc_end + K;
  double* stage = seg_start + K;
  double* bc = stage + kSampleWarps * kWarp * per;
  const int lane = threadIdx.x & 31;
  const int warp = uniform_warp_index();
  double* my_stage = stage + warp * kWarp * per;

  const long n_work = p.B * p.chunks;
  for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
    const long b = w / p.chunks;
    const int chunk = (int)(w - b * p.chunks);
    const double* cb = p.coeffs + b * (long)(K * D * N);
    const double* Tb = p.times + b * K;
    __syncthreads();  // previous work item fully consumed the shared tables
    if (threadIdx.x == 0) {
      // running end time of every segment, accumulated left to right like the reference,
      // and the start time it derives from it: (acc + T_i) - T_i  (src/trajectory.cpp:46-63)
      double acc = 0.0;
      for (int i = 0; i < K; ++i) {
        acc += Tb[i];
        acc_end[i] = acc;
        seg_start[i] = acc - Tb[i];
      }
    }
    if (kStaged) {
      const int n_c = K * D * N;
      for (int e = threadIdx.x; e < nd * n_c; e += blockDim.x) {
        const int k = e / n_c;
        const int rest = e - k * n_c;
        bc[e] = falling_factorial(k, rest % N) * cb[rest];
      }
    }
    __syncthreads();
    const double total = acc_end[K - 1];
    const double dt = total / (double)p.M;
    const int m_begin = chunk * p.chunk_len;
    const int m_end = min(p.M, m_begin + p.chunk_len);
    for (int m0 = m_begin + warp * kWarp; m0 < m_end; m0 += kSampleWarps * kWarp) {
      const int m = m0 + lane;
      const bool valid = m < m_end;
      double t = 0.0;
      int seg = -1;
      if (valid) {
        t = p.t_in ? p.t_in[b * p.t_stride + m] : (double)m * dt;
        seg = find_segment(acc_end, K, t);
        if (seg >= K || !(t == t)) seg = -1;  // past the end (or NaN): zeros, as the reference's error path
      }
      if (valid) {
        const double tl = seg >= 0 ? t - seg_start[seg] : 0.0;
        for (int k = 0; k < nd; ++k)
          for (int dim = 0; dim < D; ++dim) {
            double v = 0.0;
            if (seg >= 0 && k < N) {
              if (kStaged) v = horner_scaled<N>(bc + ((size_t)(k * K + seg) * D + dim) * N, k, tl);
              else v = horner_raw<N>(cb + ((size_t)seg * D + dim) * N, k, tl);
            }
            my_stage[lane * per + k * D + dim] = v;
          }
        if (p.t_out) p.t_out[b * (long)p.M + m] = t;
        if (p.segment) p.segment[b * (long)p.M + m] = seg;
      }
      __syncwarp();
      const int n_valid = min(kWarp, m_end - m0);
      double* dst = p.out + (b * (long)p.M + m0) * per;
      for (int e = lane; e < n_valid * per; e += kWarp) __stcs(dst + e, my_stage[e]);
      __syncwarp();
    }
  }
}

// =========================================================================================
// Derivatives 0..ND-1 in ONE pass over the raw coefficients (complete Horner scheme /
// repeated synthetic division): after pass k, a[k] = p^(k)(t) / k!.  Same 35 FMAs per
// dimension as five separate Horner runs over b(k,j) c_j (ref polynomial.h:138-151), but only
// N coefficient loads per polynomial instead of sum_k (N-k): the kernel stops being bound by
// shared-memory loads and becomes HBM-write bound.  Stores are 16 bytes per lane.
// Shared memory: acc_end[K] | seg_start[K] | coef[K][D][N] | stage[warps][32][ND*D] (16-B aligned)
// =========================================================================================
template <int N, int ND>
__global__ void __launch_bounds__(kSampleThreads) sample_all_derivatives_kernel(SampleKernelParams p) {
  extern __shared__ __align__(16) double smem[];
  const int K = p.K, D = p.D;
  const int per = ND * D;
  const int n_c = K * D * N;
  double* acc_end = smem;
  double* seg_start = acc_end + K;
  double* coef = seg_start + K;
  double* stage = coef + n_c + ((2 * K + n_c) & 1);   // keep the staging area 16-byte aligned
  const int lane = threadIdx.x & 31;
  const int warp = uniform_warp_index();
  double* my_stage = stage + (size_t)warp * kWarp * per + ((warp * kWarp * per) & 1);
  constexpr double kFactorial[6] = {1.0, 1.0, 2.0, 6.0, 24.0, 120.0};

  const long n_work = p.B * p.chunks;
  for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
    const long b = w / p.chunks;
    const int chunk = (int)(w - b * p.chunks);
    const double* cb = p.coeffs + b * (long)n_c;
    const double* Tb = p.times + b * K;
    __syncthreads();
    for (int e = threadIdx.x; e < n_c; e += blockDim.x) coef[e] = cb[e];
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int i = 0; i < K; ++i) {
        acc += Tb[i];
        acc_end[i] = acc;
        seg_start[i] = acc - Tb[i];   // (acc + T_i) - T_i, as the reference (src/trajectory.cpp:46-63)
      }
    }
    __syncthreads();
    const double total = acc_end[K - 1];
    const double dt = total / (double)p.M;
    const int m_begin = chunk * p.chunk_len;
    const int m_end = min(p.M, m_begin + p.chunk_len);
    for (int m0 = m_begin + warp * kWarp; m0 < m_end; m0 += kSampleWarps * kWarp) {
      const int m = m0 + lane;
      const bool valid = m < m_end;
      if (valid) {
        const double t = p.t_in ? p.t_in[b * p.t_stride + m] : (double)m * dt;
        int seg = find_segment(acc_end, K, t);
        if (seg >= K || !(t == t)) seg = -1;
        const double tl = seg >= 0 ? t - seg_start[seg] : 0.0;
        const double* cs = coef + (size_t)(seg >= 0 ? seg : 0) * D * N;
        for (int dim = 0; dim < D; ++dim) {
          double a[N];
#pragma unroll
          for (int j = 0; j < N; ++j) a[j] = cs[dim * N + j];
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            if (k < N) {
#pragma unroll
              for (int j = N - 2; j >= k; --j) a[j] = fma(a[j + 1], tl, a[j]);
              double f = kFactorial[k < 6 ? k : 5];
              if (k >= 6)
                for (int q = 6; q <= k; ++q) f *= (double)q;
              my_stage[lane * per + k * D + dim] = seg >= 0 ? a[k] * f : 0.0;
            } else {
              my_stage[lane * per + k * D + dim] = 0.0;
            }
          }
        }
        if (p.t_out) p.t_out[b * (long)p.M + m] = t;
        if (p.segment) p.segment[b * (long)p.M + m] = seg;
      }
      __syncwarp();
      const int n_valid = min(kWarp, m_end - m0);
      const long first = (b * (long)p.M + m0) * per;
      double* dst = p.out + first;
      const int count = n_valid * per;
      if (((first | count) & 1) == 0 && p.out_aligned16) {
        const double2* src2 = reinterpret_cast<const double2*>(my_stage);
        double2* dst2 = reinterpret_cast<double2*>(dst);
        for (int e = lane; e < count / 2; e += kWarp) __stcs(dst2 + e, src2[e]);
      } else {
        for (int e = lane; e < count; e += kWarp) __stcs(dst + e, my_stage[e]);
      }
      __syncwarp();
    }
  }
}

template <int N>
static cudaError_t launch_sample_n(const SampleArgs& a, cudaStream_t stream) {
  if (a.B == 0 || a.M == 0) return cudaSuccess;
  const int per = a.n_deriv * a.D;
  // fast kernel: n_deriv <= 5 and the raw coefficients of one trajectory fit shared memory
  {
    const size_t n_c = (size_t)a.K * a.D * N;
    const size_t doubles = 2 * (size_t)a.K + n_c + 1 + (size_t)kSampleWarps * (kWarp * per + 1);
    const size_t smem_fast = doubles * sizeof(double);
    if (a.n_deriv <= 5 && smem_fast <= 96 * 1024) {
      SampleKernelParams p;
      p.B = a.B; p.K = a.K; p.D = a.D; p.M = a.M; p.n_deriv = a.n_deriv;
      int chunk_len = a.M;
      if (a.B < sm_count() * 8) {
        const long want = (sm_count() * 8 + a.B - 1) / a.B;
        chunk_len = (int)((a.M + want - 1) / want);
        if (chunk_len < kSampleThreads) chunk_len = kSampleThreads;
        chunk_len = (chunk_len + kWarp - 1) / kWarp * kWarp;
      }
      p.chunk_len = chunk_len;
      p.chunks = (a.M + chunk_len - 1) / chunk_len;
      p.coeffs = a.d_coeffs; p.times = a.d_times; p.t_in = a.d_t; p.t_stride = a.t_stride;
      p.out = a.d_out; p.t_out = a.d_t_out; p.segment = a.d_segment;
      p.out_aligned16 = reinterpret_cast<uintptr_t>(a.d_out) % 16 == 0;
      long grid = a.B * p.chunks;
      if (grid > (1L << 30)) grid = 1L << 30;
      void (*kernel)(SampleKernelParams) = nullptr;
      switch (a.n_deriv) {
        case 1: kernel = sample_all_derivatives_kernel<N, 1>; break;
        case 2: kernel = sample_all_derivatives_kernel<N, 2>; break;
        case 3: kernel = sample_all_derivatives_kernel<N, 3>; break;
        case 4: kernel = sample_all_derivatives_kernel<N, 4>; break;
        default: kernel = sample_all_derivatives_kernel<N, 5>; break;
      }
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);
      if (e != cudaSuccess) return e;
      kernel<<<(unsigned)grid, kSampleThreads, smem_fast, stream>>>(p);
      return cudaGetLastError();
    }
  }
  const size_t base = sizeof(double) * ((size_t)2 * a.K + (size_t)kSampleWarps * kWarp * per);
  const size_t staged = base + sizeof(double) * (size_t)a.n_deriv * a.K * a.D * N;
  const bool use_staged = staged <= 64 * 1024;
  const size_t smem = use_staged ? staged : base;
  if (smem > kMaxDynamicSmem) return cudaErrorInvalidConfiguration;
  SampleKernelParams p;
  p.B = a.B; p.K = a.K; p.D = a.D; p.M = a.M; p.n_deriv = a.n_deriv;
  // one CTA per trajectory unless that leaves the machine idle: then cut M into chunks
  int chunk_len = a.M;
  if (a.B < sm_count() * 8) {
    const long want = (sm_count() * 8 + a.B - 1) / a.B;
    chunk_len = (int)((a.M + want - 1) / want);
    const int min_len = kSampleThreads * (use_staged ? 4 : 1);
    if (chunk_len < min_len) chunk_len = min_len;
    chunk_len = (chunk_len + kWarp - 1) / kWarp * kWarp;
  }
  p.chunk_len = chunk_len;
  p.chunks = (a.M + chunk_len - 1) / chunk_len;
  p.coeffs = a.d_coeffs; p.times = a.d_times; p.t_in = a.d_t; p.t_stride = a.t_stride;
  p.out = a.d_out; p.t_out = a.d_t_out; p.segment = a.d_segment;
  long grid = a.B * p.chunks;
  const long max_grid = 1L << 30;
  if (grid > max_grid) grid = max_grid;
  cudaError_t e;
  if (use_staged) {
    auto kernel = sample_kernel<N, true>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)grid, kSampleThreads, smem, stream>>>(p);
  } else {
    auto kernel = sample_kernel<N, false>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)grid, kSampleThreads, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t stream) {
  switch (a.N) {
    case 4: return launch_sample_n<4>(a, stream);
    case 6: return launch_sample_n<6>(a, stream);
    case 8: return launch_sample_n<8>(a, stream);
    case 10: return launch_sample_n<10>(a, stream);
    case 12: return launch_sample_n<12>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

// =========================================================================================
// a20. evaluateRange (ref: src/trajectory.cpp:68-128).  The sample instants are defined by a
// sequential accumulation (time_in_segment += dt; accumulated_time += dt), so lane 0 of the
// warp that owns the trajectory walks 32 steps ahead, then the 32 lanes evaluate and store.
// =========================================================================================
template <int N>
__global__ void __launch_bounds__(128) evaluate_range_kernel(long B, int K, int D, const double* __restrict__ coeffs,
                                                             const double* __restrict__ times, double t_start,
                                                             double t_end, double dt, int derivative, int max_samples,
                                                             double* __restrict__ out, double* __restrict__ t_out,
                                                             int32_t* __restrict__ count) {
  __shared__ int s_seg[4][kWarp];
  __shared__ double s_tl[4][kWarp];
  __shared__ double s_acc[4][kWarp];
  const int lane = threadIdx.x & 31;
  const int wib = uniform_warp_index();
  const long warp = blockIdx.x * 4L + wib;
  const long n_warps = gridDim.x * 4L;
  for (long b = warp; b < B; b += n_warps) {
    const double* cb = coeffs + b * (long)(K * D * N);
    const double* Tb = times + b * K;
    // walker state (meaningful in lane 0 only)
    double acc = 0.0, tis = 0.0;
    int i = 0;
    bool alive = false;
    if (lane == 0) {
      for (i = 0; i < K; ++i) {
        acc += Tb[i];
        if (acc > t_start) break;
      }
      if (!(t_start > acc) && i < K) {
        acc -= Tb[i];
        tis = t_start - acc;
        alive = true;
      }
    }
    int n_emitted = 0;
    for (;;) {
      int n_batch = 0;
      if (lane == 0 && alive) {
        while (n_batch < kWarp) {
          if (!(acc < t_end)) { alive = false; break; }
          if (tis > Tb[i]) {
            tis = tis - Tb[i];
            ++i;
            if (i >= K) { alive = false; break; }
            continue;
          }
          s_seg[wib][n_batch] = i;
          s_tl[wib][n_batch] = tis;
          s_acc[wib][n_batch] = acc;
          ++n_batch;
          tis += dt;
          acc += dt;
        }
      }
      n_batch = __shfl_sync(0xffffffffu, n_batch, 0);
      __syncwarp();
      if (n_batch == 0) break;
      const int m = n_emitted + lane;
      if (lane < n_batch && m < max_samples) {
        const int seg = s_seg[wib][lane];
        const double tl = s_tl[wib][lane];
        for (int dim = 0; dim < D; ++dim)
          out[(b * (long)max_samples + m) * D + dim] =
              derivative < N ? horner_raw<N>(cb + ((size_t)seg * D + dim) * N, derivative, tl) : 0.0;
        if (t_out) t_out[b * (long)max_samples + m] = s_acc[wib][lane];
      }
      n_emitted += n_batch;
      __syncwarp();
      if (!__shfl_sync(0xffffffffu, (int)alive, 0)) break;
    }
    if (count && lane == 0) count[b] = n_emitted;
  }
}

cudaError_t launch_evaluate_range(long B, int K, int D, int N, const double* d_coeffs, const double* d_times,
                                  double t_start, double t_end, double dt, int derivative, int max_samples,
                                  double* d_out, double* d_t_out, int32_t* d_count, cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  long grid = (B + 3) / 4;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
#define MINSNAP_ER(N_)                                                                                         \
  case N_:                                                                                                     \
    evaluate_range_kernel<N_><<<(int)grid, 128, 0, stream>>>(B, K, D, d_coeffs, d_times, t_start, t_end, dt,   \
                                                             derivative, max_samples, d_out, d_t_out, d_count); \
    break;
  switch (N) {
    MINSNAP_ER(4) MINSNAP_ER(6) MINSNAP_ER(8) MINSNAP_ER(10) MINSNAP_ER(12)
    default: return cudaErrorInvalidValue;
  }
#undef MINSNAP_ER
  return cudaGetLastError();
}

}  // namespace minsnap
