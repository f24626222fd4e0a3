// Dependent-chain latencies on the target GPU (cycles): DFMA, DMUL, DADD, rcp seed, LDS, SHFL.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0, t1;
  const int N = 512;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * a;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  int idx = threadIdx.x & 63;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { idx = (int)sm[idx & 63] & 63; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  // two independent DFMA chains (ILP 2) and four (ILP 4)
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { y0 = fma(y0, a, b); y1 = fma(y1, a, b); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { y0 = fma(y0, a, b); y1 = fma(y1, a, b); y2 = fma(y2, a, b); y3 = fma(y3, a, b); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  out[threadIdx.x] = x + idx + y0 + y1 + y2 + y3;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
  for (int warps = 1; warps <= 8; warps *= 2) {
    lat<<<1, 32 * warps>>>(out, cyc, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("warps/SM=%d (per SMSP %.2f): DFMA %.1f DMUL %.1f DADD %.1f RCP64H %.1f LDS(dep,+cvt) %.1f SHFL64 %.1f | 2xDFMA %.1f 4xDFMA %.1f cyc/iter\n",
           warps, warps / 4.0, h[0] / 512.0, h[1] / 512.0, h[2] / 512.0, h[3] / 512.0, h[4] / 512.0, h[5] / 512.0, h[6] / 512.0, h[7] / 512.0);
  }
  return 0;
}
