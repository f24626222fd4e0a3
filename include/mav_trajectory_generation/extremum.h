// Extremum: where and how large the magnitude of a derivative gets.  API mirror of the reference's
// include/mav_trajectory_generation/extremum.h:27-49 (same member names, ordering by value only, same
// printed form); produced by minsnap_extrema[_host] through Segment / Trajectory /
// PolynomialOptimization.
#ifndef MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
#define MAV_TRAJECTORY_GENERATION_EXTREMUM_H_

#include <ostream>

namespace mav_trajectory_generation {

struct Extremum {
  // seconds since the start of segment `segment_idx`
  double time = 0.0;
  // |p^(k)| there
  double value = 0.0;
  int segment_idx = 0;

  Extremum() = default;
  Extremum(double at_time, double magnitude, int segment) : time(at_time), value(magnitude), segment_idx(segment) {}
};

// Extrema compare by magnitude alone: the fold "a later candidate wins only when strictly larger" of
// computeMaximumOfMagnitude / computeMinMaxMagnitude is written with these.
inline bool operator<(const Extremum& a, const Extremum& b) { return a.value < b.value; }
inline bool operator>(const Extremum& a, const Extremum& b) { return b < a; }

inline std::ostream& operator<<(std::ostream& os, const Extremum& e) {
  os << "time: " << e.time << ", value: " << e.value << ", segment idx: " << e.segment_idx << '\n';
  return os.flush();
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
