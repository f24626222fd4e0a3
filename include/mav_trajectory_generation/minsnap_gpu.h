// Glue between the C++ mirror of the reference API and the C ABI (minsnap_b200.h).
// Error convention of the reference: programmer errors abort through glog CHECK; a failing GPU
// call is treated the same way -- there is no CPU path to fall back to.
#ifndef MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_
#define MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_

#include <cstdint>
#include <vector>

#include "mav_trajectory_generation/minsnap_shims.h"
#include "minsnap_b200.h"

namespace mav_trajectory_generation {
namespace gpu {

inline void check(int rc, const char* what) {
  CHECK(rc == MINSNAP_OK) << what << " failed: " << minsnap_error_string(rc) << " " << minsnap_last_cuda_error()
                          << " (this build has no CPU fallback)";
}

// Extrema (computeMaximumOfMagnitude, computeMinMaxMagnitude...): the reference drops trailing
// coefficients of the candidate polynomial below machine epsilon (absolute, src/rpoly.cpp:44-55),
// which loses extrema near the end of segments longer than about 12 s.  The default reproduces
// the reference; keepSmallCoefficients(true) switches every later call of this process to the
// exact candidate polynomial (MINSNAP_EXTREMA_KEEP_SMALL_COEFFICIENTS).
inline bool& keepSmallCoefficientsFlag() {
  static bool keep = false;
  return keep;
}
inline void keepSmallCoefficients(bool keep) { keepSmallCoefficientsFlag() = keep; }
inline int extremaMode(int mode) {
  return keepSmallCoefficientsFlag() ? (mode | MINSNAP_EXTREMA_KEEP_SMALL_COEFFICIENTS) : mode;
}

// dimensions {0, 2} -> bit mask; false when one is out of [0, D)
inline bool dimensionMask(const std::vector<int>& dimensions, int D, uint32_t* mask) {
  *mask = 0;
  for (int dim : dimensions) {
    if (dim < 0 || dim >= D || dim >= 32) return false;
    *mask |= 1u << dim;
  }
  return true;
}

}  // namespace gpu
}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_
