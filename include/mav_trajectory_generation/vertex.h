// Vertex: the support point of a path and its derivative constraints (mirror of ref
// include/mav_trajectory_generation/vertex.h:42-145 and src/vertex.cpp).  A derivative that is
// present in the constraint map is FIXED at this vertex, an absent one is FREE (optimised).
#ifndef MAV_TRAJECTORY_GENERATION_VERTEX_H_
#define MAV_TRAJECTORY_GENERATION_VERTEX_H_

#include <map>
#include <ostream>
#include <utility>
#include <vector>

#include "mav_trajectory_generation/minsnap_gpu.h"
#include "mav_trajectory_generation/motion_defines.h"
#include "mav_trajectory_generation/polynomial.h"

namespace mav_trajectory_generation {

class Vertex {
 public:
  typedef std::vector<Vertex> Vector;
  typedef Eigen::VectorXd ConstraintValue;
  typedef std::pair<int, ConstraintValue> Constraint;
  typedef std::map<int, ConstraintValue> Constraints;

  explicit Vertex(size_t dimension) : D_(static_cast<int>(dimension)) {}

  int D() const { return D_; }

  // Same value in every dimension.
  void addConstraint(int derivative_order, double value) {
    constraints_[derivative_order] = ConstraintValue::Constant(D_, value);
  }
  // One value per dimension; the size has to match D().
  void addConstraint(int type, const Eigen::VectorXd& constraint) {
    CHECK_EQ(static_cast<long>(constraint.rows()), static_cast<long>(D_));
    constraints_[type] = constraint;
  }
  bool removeConstraint(int type) { return constraints_.erase(type) > 0; }

  // Position constraint plus zero derivatives 1..up_to_derivative (start / goal vertices).
  void makeStartOrEnd(const Eigen::VectorXd& constraint, int up_to_derivative) {
    addConstraint(derivative_order::POSITION, constraint);
    for (int k = 1; k <= up_to_derivative; ++k) constraints_[k] = ConstraintValue::Zero(D_);
  }
  void makeStartOrEnd(double value, int up_to_derivative) {
    makeStartOrEnd(Eigen::VectorXd::Constant(D_, value), up_to_derivative);
  }

  bool hasConstraint(int derivative_order) const { return constraints_.count(derivative_order) > 0; }
  bool getConstraint(int derivative_order, Eigen::VectorXd* constraint) const {
    CHECK_NOTNULL(constraint);
    Constraints::const_iterator it = constraints_.find(derivative_order);
    if (it == constraints_.end()) return false;
    *constraint = it->second;
    return true;
  }

  Constraints::const_iterator cBegin() const { return constraints_.begin(); }
  Constraints::const_iterator cEnd() const { return constraints_.end(); }
  size_t getNumberOfConstraints() const { return constraints_.size(); }

  bool isEqualTol(const Vertex& rhs, double tol) const {
    if (constraints_.size() != rhs.constraints_.size()) return false;
    for (Constraints::const_iterator it = cBegin(); it != cEnd(); ++it) {
      Constraints::const_iterator other = rhs.constraints_.find(it->first);
      if (other == rhs.constraints_.end()) return false;
      if (!((it->second - other->second).isZero(tol))) return false;
    }
    return true;
  }

 private:
  int D_;
  Constraints constraints_;
};

inline std::ostream& operator<<(std::ostream& stream, const Vertex& v) {
  stream << "constraints: " << std::endl;
  for (Vertex::Constraints::const_iterator it = v.cBegin(); it != v.cEnd(); ++it)
    stream << "  type: " << positionDerivativeToString(it->first) << "  value: [" << it->second << "]" << std::endl;
  return stream;
}

inline std::ostream& operator<<(std::ostream& stream, const std::vector<Vertex>& vertices) {
  for (const Vertex& v : vertices) stream << v << std::endl;
  return stream;
}

// t = 2 d / v_max (1 + magic v_max / a_max exp(-2 d / v_max)), d = distance between consecutive
// vertex positions (ref src/vertex.cpp:162-178).  Evaluated by minsnap_estimate_segment_times.
inline std::vector<double> estimateSegmentTimes(const Vertex::Vector& vertices, double v_max, double a_max,
                                                double magic_fabian_constant = 6.5) {
  CHECK_GE(vertices.size(), static_cast<size_t>(2));
  const int D = vertices.front().D();
  const int K = static_cast<int>(vertices.size()) - 1;
  std::vector<double> positions(static_cast<size_t>(K + 1) * D);
  for (int v = 0; v <= K; ++v) {
    Eigen::VectorXd p;
    CHECK(vertices[v].getConstraint(derivative_order::POSITION, &p)) << "vertex " << v << " has no position";
    for (int d = 0; d < D; ++d) positions[static_cast<size_t>(v) * D + d] = p[d];
  }
  std::vector<double> segment_times(static_cast<size_t>(K));
  gpu::check(minsnap_estimate_segment_times_host(1, K, D, positions.data(), v_max, a_max, magic_fabian_constant,
                                                 segment_times.data()),
             "minsnap_estimate_segment_times_host");
  return segment_times;
}

// Random vertices inside [minimum_position, maximum_position]: start and goal fix derivatives
// 0..maximum_derivative (zero derivatives), interior vertices fix position only; consecutive
// vertices are at least 0.2 apart (ref src/vertex.cpp:27-79).  Host-only workload generator
// (std::mt19937 + std::uniform_real_distribution, via minsnap_random_positions_host).
inline Vertex::Vector createRandomVertices(int maximum_derivative, size_t n_segments,
                                           const Eigen::VectorXd& minimum_position,
                                           const Eigen::VectorXd& maximum_position, size_t seed = 0) {
  CHECK_GE(static_cast<int>(n_segments), 1);
  CHECK_EQ(minimum_position.size(), maximum_position.size());
  CHECK_GT(maximum_derivative, 0);
  const int D = static_cast<int>(minimum_position.size());
  const int K = static_cast<int>(n_segments);
  std::vector<double> positions(static_cast<size_t>(K + 1) * D);
  gpu::check(minsnap_random_positions_host(1, K, D, minimum_position.data(), maximum_position.data(),
                                           static_cast<uint64_t>(seed), positions.data()),
             "minsnap_random_positions_host");
  Vertex::Vector vertices;
  vertices.reserve(static_cast<size_t>(K) + 1);
  for (int v = 0; v <= K; ++v) {
    Eigen::VectorXd p(D);
    for (int d = 0; d < D; ++d) p[d] = positions[static_cast<size_t>(v) * D + d];
    Vertex vertex(D);
    if (v == 0 || v == K) vertex.makeStartOrEnd(p, maximum_derivative);
    else vertex.addConstraint(derivative_order::POSITION, p);
    vertices.push_back(vertex);
  }
  return vertices;
}

inline Vertex::Vector createRandomVertices1D(int maximum_derivative, size_t n_segments, double minimum_position,
                                             double maximum_position, size_t seed = 0) {
  return createRandomVertices(maximum_derivative, n_segments, Eigen::VectorXd::Constant(1, minimum_position),
                              Eigen::VectorXd::Constant(1, maximum_position), seed);
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_VERTEX_H_
