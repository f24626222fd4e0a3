// Glue between the C++ mirror of the reference API and the C ABI (minsnap_b200.h).
// Error convention of the reference: programmer errors abort through glog CHECK; a failing GPU
// call is treated the same way -- there is no CPU path to fall back to.
#ifndef MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_
#define MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_

#include "mav_trajectory_generation/minsnap_shims.h"
#include "minsnap_b200.h"

namespace mav_trajectory_generation {
namespace gpu {

inline void check(int rc, const char* what) {
  CHECK(rc == MINSNAP_OK) << what << " failed: " << minsnap_error_string(rc) << " " << minsnap_last_cuda_error()
                          << " (this build has no CPU fallback)";
}

}  // namespace gpu
}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_MINSNAP_GPU_H_
