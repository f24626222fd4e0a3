// Fast route for the standard mask (N = 10, snap): placeholder until the thread-pair kernel
// lands; every shape currently takes the generic route.
#pragma once
#include "minsnap_device.cuh"
#include "minsnap_launch.h"

namespace minsnap {
namespace fast {

inline bool supported(int, int, int, int) { return false; }
inline bool sweep_supported(int, int, int, int) { return false; }
inline cudaError_t launch(const StandardSolveArgs&, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t launch_sweep(const SweepArgs&, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace fast
}  // namespace minsnap
