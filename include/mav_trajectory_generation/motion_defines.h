// Derivative-order constants of the hot path's API (ref: include/mav_trajectory_generation/
// motion_defines.h:28-40, src/motion_defines.cpp).  Header-only.
#ifndef MAV_TRAJECTORY_GENERATION_MOTION_DEFINES_H_
#define MAV_TRAJECTORY_GENERATION_MOTION_DEFINES_H_

#include <string>

namespace mav_trajectory_generation {

namespace derivative_order {
static constexpr int POSITION = 0;
static constexpr int VELOCITY = 1;
static constexpr int ACCELERATION = 2;
static constexpr int JERK = 3;
static constexpr int SNAP = 4;

static constexpr int ORIENTATION = 0;
static constexpr int ANGULAR_VELOCITY = 1;
static constexpr int ANGULAR_ACCELERATION = 2;

static constexpr int kINVALID = -1;
}  // namespace derivative_order

inline std::string positionDerivativeToString(int derivative) {
  static const char* const names[] = {"position", "velocity", "acceleration", "jerk", "snap"};
  return (derivative >= 0 && derivative <= 4) ? names[derivative] : "invalid";
}

inline int positionDerivativeToInt(const std::string& name) {
  for (int k = 0; k <= 4; ++k)
    if (positionDerivativeToString(k) == name) return k;
  return derivative_order::kINVALID;
}

inline std::string orintationDerivativeToString(int derivative) {
  static const char* const names[] = {"orientation", "angular_velocity", "angular_acceleration"};
  return (derivative >= 0 && derivative <= 2) ? names[derivative] : "invalid";
}

inline int orientationDerivativeToInt(const std::string& name) {
  for (int k = 0; k <= 2; ++k)
    if (orintationDerivativeToString(k) == name) return k;
  return derivative_order::kINVALID;
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_MOTION_DEFINES_H_
