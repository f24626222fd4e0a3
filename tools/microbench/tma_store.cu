// Microbenchmark: throughput of small cp.async.bulk shared->global copies (one per lane), the
// pattern a per-lane copy-out of 240-byte coefficient blocks would use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store tma_store.cu && ./tma_store
#include <cstdio>
#include <cuda_runtime.h>

template <int BYTES, bool FENCE>
__global__ void bulk_store_kernel(double* out, int iters, long stride_doubles) {
  extern __shared__ __align__(128) double tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* my = tile + (warp * 32 + lane) * (BYTES / 8);
  for (int e = 0; e < BYTES / 8; ++e) my[e] = lane + e;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long gwarp = (long)blockIdx.x * (blockDim.x >> 5) + warp;
  double* dst = out + (gwarp * 32 + lane) * stride_doubles;
  const unsigned src = (unsigned)__cvta_generic_to_shared(my);
  for (int it = 0; it < iters; ++it) {
    if (FENCE) {   // the solver's pattern: rewrite the piece, proxy fence, then copy
#pragma unroll
      for (int e = 0; e < BYTES / 8; e += 2) *reinterpret_cast<double2*>(my + e) = make_double2(it + e, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (it & 3) * (BYTES / 8)),
                 "r"(src), "n"(BYTES)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int BYTES, bool FENCE = false>
void run(int warps_per_cta, int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm, iters = 200;
  const long stride = 150;  // 1200-byte chunks like the solver's lanes
  double* out;
  cudaMalloc(&out, (size_t)grid * warps_per_cta * 32 * stride * 8 + 4096);
  const size_t smem = (size_t)warps_per_cta * 32 * BYTES;
  cudaFuncSetAttribute(bulk_store_kernel<BYTES, FENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  bulk_store_kernel<BYTES, FENCE><<<grid, warps_per_cta * 32, smem>>>(out, 10, stride);
  cudaEventRecord(a);
  bulk_store_kernel<BYTES, FENCE><<<grid, warps_per_cta * 32, smem>>>(out, iters, stride);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)grid * warps_per_cta * 32 * iters;
  printf("fence %d bytes/op %4d warps/SM %2d: %8.3f ms  %7.2f ns/op/SM  (%5.1f cycles/op/SM at 1.9 GHz)  %7.1f GB/s  err=%s\n", (int)FENCE, BYTES,
         warps_per_cta * ctas_per_sm, ms, ms * 1e6 / (ops / 148), ms * 1e6 / (ops / 148) * 1.9, ops * BYTES / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<240>(1, 8);
  run<240, true>(1, 8);
  run<240>(4, 2);
  run<480>(1, 8);
  run<80>(1, 8);
  run<1200>(1, 4);
  return 0;
}
