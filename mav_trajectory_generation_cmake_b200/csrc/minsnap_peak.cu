// Measurement helper: dependency-free DFMA loop.  Gives the FP64 roofline denominator that
// MEASURED_PEAKS.json does not carry (it only has HBM copy bandwidth and bf16 GEMM rate).
#include "minsnap_device.cuh"
#include "minsnap_launch.h"

namespace minsnap {

constexpr int kPeakChains = 8;
constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, double a, double b) {
  double acc[kPeakChains];
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) acc[c] = (double)(threadIdx.x + c);
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) s += acc[c];
  if (s == 123.456) sink[0] = s;  // never true; keeps the loop alive
}

cudaError_t run_fp64_peak(int repeats, double* tflops) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  double* sink = nullptr;
  e = cudaMalloc(&sink, sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  const int grid = sms * 8, block = 256;
  dfma_peak_kernel<<<grid, block>>>(sink, 0.999999, 1e-9);  // warm-up
  double best = 0.0;
  for (int r = 0; r < repeats; ++r) {
    cudaEventRecord(t0);
    dfma_peak_kernel<<<grid, block>>>(sink, 0.999999, 1e-9);
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flops = 2.0 * kPeakChains * (double)kPeakIters * (double)grid * block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(sink);
  *tflops = best;
  return e;
}

}  // namespace minsnap
