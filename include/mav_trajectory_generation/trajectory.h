// Trajectory: a sequence of segments (mirror of ref include/mav_trajectory_generation/
// trajectory.h:31-110, src/trajectory.cpp:27-179; computeMinMaxMagnitude is out of scope).
// evaluate / evaluateRange run on the GPU through the C ABI (rows a19, a20).
#ifndef MAV_TRAJECTORY_GENERATION_TRAJECTORY_H_
#define MAV_TRAJECTORY_GENERATION_TRAJECTORY_H_

#include <limits>
#include <vector>

#include "mav_trajectory_generation/extremum.h"
#include "mav_trajectory_generation/segment.h"

namespace mav_trajectory_generation {

class Trajectory {
 public:
  Trajectory() : D_(0), N_(0), max_time_(0.0) {}

  bool operator==(const Trajectory& rhs) const { return segments_ == rhs.segments_; }
  bool operator!=(const Trajectory& rhs) const { return !operator==(rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  int K() const { return static_cast<int>(segments_.size()); }

  bool empty() const { return segments_.empty(); }
  void clear() {
    segments_.clear();
    D_ = 0;
    N_ = 0;
  }

  void setSegments(const Segment::Vector& segments) {
    CHECK(!segments.empty());
    segments_ = segments;
    D_ = segments_.front().D();
    N_ = segments_.front().N();
    max_time_ = 0.0;  // cached end time: left-to-right sum, as the reference accumulates it
    for (const Segment& segment : segments) {
      CHECK_EQ(segment.D(), D_);
      max_time_ += segment.getTime();
    }
  }
  void getSegments(Segment::Vector* segments) const {
    CHECK_NOTNULL(segments);
    *segments = segments_;
  }
  const Segment::Vector& segments() const { return segments_; }

  double getMinTime() const { return 0.0; }
  double getMaxTime() const { return max_time_; }

  Trajectory getTrajectoryWithSingleDimension(int dimension) const {
    CHECK_LT(dimension, D_);
    Segment::Vector segments;
    segments.reserve(segments_.size());
    for (const Segment& s : segments_) {
      Segment one(N_, 1);
      one[0] = s[dimension];
      one.setTime(s.getTime());
      segments.push_back(one);
    }
    Trajectory traj;
    traj.setSegments(segments);
    return traj;
  }

  Trajectory getTrajectoryWithAppendedDimension(const Trajectory& trajectory_to_append) const {
    if (N_ == 0 || D_ == 0) return trajectory_to_append;
    if (trajectory_to_append.N() == 0 || trajectory_to_append.D() == 0) return *this;
    CHECK_EQ(N_, trajectory_to_append.N());
    CHECK_EQ(static_cast<int>(segments_.size()), trajectory_to_append.K());
    Segment::Vector segments;
    segments.reserve(segments_.size());
    for (size_t k = 0; k < segments_.size(); ++k) {
      Segment both(N_, D_ + trajectory_to_append.D());
      both.setTime(segments_[k].getTime());
      for (int d = 0; d < D_; ++d) both[d] = segments_[k][d];
      for (int d = 0; d < trajectory_to_append.D(); ++d) both[D_ + d] = trajectory_to_append.segments()[k][d];
      segments.push_back(both);
    }
    Trajectory traj;
    traj.setSegments(segments);
    return traj;
  }

  // The segment is the first whose running end time exceeds t (a vertex instant belongs to the
  // segment on its right); an instant past the end logs an error and yields zeros
  // (ref src/trajectory.cpp:41-66).
  Eigen::VectorXd evaluate(double t, int derivative = derivative_order::POSITION) const {
    Eigen::VectorXd result = Eigen::VectorXd::Zero(D_);
    if (segments_.empty() || derivative >= N_) return result;
    std::vector<int32_t> segment_index;
    std::vector<double> out = evaluateBatch(std::vector<double>(1, t), derivative + 1, &segment_index);
    if (segment_index[0] < 0) {
      LOG(ERROR) << "Time out of range of the trajectory!";
      return result;
    }
    for (int d = 0; d < D_; ++d) result[d] = out[static_cast<size_t>(derivative) * D_ + d];
    return result;
  }

  // Additive batched form: derivatives 0..n_deriv-1 at many instants in one launch; returns
  // [times.size()][n_deriv][D] and optionally the segment each instant fell into (-1: outside).
  std::vector<double> evaluateBatch(const std::vector<double>& times, int n_deriv,
                                    std::vector<int32_t>* segment_index = nullptr) const {
    std::vector<double> coeffs, durations;
    pack(&coeffs, &durations);
    std::vector<double> out(times.size() * static_cast<size_t>(n_deriv) * D_);
    std::vector<int32_t> seg(times.size());
    gpu::check(minsnap_sample_at_host(1, K(), D_, N_, coeffs.data(), durations.data(), static_cast<int>(times.size()),
                                      times.data(), 0, n_deriv, out.data(), seg.data()),
               "minsnap_sample_at_host");
    if (segment_index) *segment_index = seg;
    return out;
  }

  // Minimum and maximum of the magnitude of a derivative over the listed dimensions and the whole
  // trajectory (ref src/trajectory.cpp:181-217): per segment the candidates are its start, its end
  // and the real roots of d/dt |p^(derivative)|^2 inside it; the first strictly smaller / larger
  // candidate wins.  One GPU call for all segments (minsnap_extrema_host).
  bool computeMinMaxMagnitude(int derivative, const std::vector<int>& dimensions, Extremum* minimum,
                              Extremum* maximum) const {
    CHECK_NOTNULL(minimum);
    CHECK_NOTNULL(maximum);
    minimum->value = std::numeric_limits<double>::max();
    maximum->value = std::numeric_limits<double>::lowest();
    if (segments_.empty()) return true;
    if (dimensions.empty()) {
      LOG(WARNING) << "No dimensions specified." << std::endl;
      return false;
    }
    uint32_t mask = 0;
    if (!gpu::dimensionMask(dimensions, D_, &mask)) {
      LOG(WARNING) << "Specified dimensions are out of bounds [0.." << D_ - 1 << "]." << std::endl;
      return false;
    }
    if (derivative < 0 || N_ - derivative - 2 < 0) {
      LOG(WARNING) << "N - derivative - 1 has to be at least 0.";
      return false;
    }
    std::vector<double> coeffs, durations;
    pack(&coeffs, &durations);
    int32_t min_segment = 0, max_segment = 0;
    gpu::check(minsnap_extrema_host(1, K(), D_, N_, coeffs.data(), durations.data(), derivative,
                                    gpu::extremaMode(MINSNAP_EXTREMA_TRAJECTORY), mask, &maximum->time,
                                    &maximum->value, &max_segment, &minimum->time, &minimum->value, &min_segment,
                                    nullptr, nullptr, nullptr),
               "minsnap_extrema_host");
    minimum->segment_idx = min_segment;
    maximum->segment_idx = max_segment;
    return true;
  }

  // Samples one derivative from t_start to t_end every dt, accumulating the sample time the way
  // the reference does (ref src/trajectory.cpp:68-128).
  void evaluateRange(double t_start, double t_end, double dt, int derivative,
                     std::vector<Eigen::VectorXd>* result, std::vector<double>* sampling_times = nullptr) const {
    CHECK_NOTNULL(result);
    result->clear();
    if (sampling_times != nullptr) sampling_times->clear();
    if (segments_.empty()) return;
    // The reference's loop runs its accumulated time from the START OF THE SEGMENT that contains t_start
    // (src/trajectory.cpp:104-127), so it can emit more than (t_end - t_start) / dt samples; the kernel reports
    // the exact count and the call is repeated with that capacity when the first guess was too small.
    int capacity = static_cast<int>((t_end - t_start) / dt) + 4;
    std::vector<double> coeffs, durations;
    pack(&coeffs, &durations);
    std::vector<double> out, ts;
    int32_t count = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      out.assign(static_cast<size_t>(capacity) * D_, 0.0);
      ts.assign(static_cast<size_t>(capacity), 0.0);
      gpu::check(minsnap_evaluate_range_host(K(), D_, N_, coeffs.data(), durations.data(), t_start, t_end, dt,
                                             derivative, capacity, out.data(), ts.data(), &count),
                 "minsnap_evaluate_range_host");
      if (count <= capacity) break;
      capacity = count;
    }
    if (count == 0 && t_start > max_time_) LOG(ERROR) << "Start time out of range of the trajectory!";
    const int n = count < capacity ? count : capacity;
    result->reserve(static_cast<size_t>(n));
    for (int m = 0; m < n; ++m) {
      Eigen::VectorXd v(D_);
      for (int d = 0; d < D_; ++d) v[d] = out[static_cast<size_t>(m) * D_ + d];
      result->push_back(v);
      if (sampling_times != nullptr) sampling_times->push_back(ts[static_cast<size_t>(m)]);
    }
  }

 private:
  // coefficients [K][D][N] and durations [K] in the C-ABI layout
  void pack(std::vector<double>* coeffs, std::vector<double>* durations) const {
    coeffs->resize(segments_.size() * static_cast<size_t>(D_) * N_);
    durations->resize(segments_.size());
    size_t o = 0;
    for (size_t k = 0; k < segments_.size(); ++k) {
      (*durations)[k] = segments_[k].getTime();
      for (int d = 0; d < D_; ++d) {
        const Eigen::VectorXd c = segments_[k][d].getCoefficients(0);
        for (int j = 0; j < N_; ++j) (*coeffs)[o++] = c[j];
      }
    }
  }

  int D_;
  int N_;
  double max_time_;
  Segment::Vector segments_;
};

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_TRAJECTORY_H_
