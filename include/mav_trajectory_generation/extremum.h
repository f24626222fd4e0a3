// Extremum: where and how large a derivative's magnitude gets (mirror of ref
// include/mav_trajectory_generation/extremum.h:27-49).  `time` is relative to the start of the
// segment `segment_idx`; ordering compares the values only.
#ifndef MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
#define MAV_TRAJECTORY_GENERATION_EXTREMUM_H_

#include <ostream>

namespace mav_trajectory_generation {

struct Extremum {
  double time = 0.0;
  double value = 0.0;
  int segment_idx = 0;

  Extremum() = default;
  Extremum(double _time, double _value, int _segment_idx) : time(_time), value(_value), segment_idx(_segment_idx) {}

  bool operator<(const Extremum& rhs) const { return value < rhs.value; }
  bool operator>(const Extremum& rhs) const { return value > rhs.value; }
};

inline std::ostream& operator<<(std::ostream& stream, const Extremum& e) {
  return stream << "time: " << e.time << ", value: " << e.value << ", segment idx: " << e.segment_idx << std::endl;
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
