// The reference's own timing procedure, through the C++ mirror of its API (one trajectory per call, B = 1):
// ref src/polynomial_timing_evaluation.cpp:93-127 -- for K in {2, 10, 50, 100} segments, 1000 times:
// random 3-D vertices (createRandomVerticesPath, :34-91, seed 1, average distance 5), estimateSegmentTimes
// (v_max = a_max = 2, 6.5), PolynomialOptimization<10>::setupFromVertices + solveLinear inside the timer.
// Added: the per-call cost of the scalar evaluate() calls of the mirror (one host<->device round trip each)
// next to the batched forms (evaluateBatch, evaluateRange).  Prints one JSON object.
//
//   tests/cpp/bin/timing_evaluation [repetitions]        (built by tests/cpp/build_cpp_tests.py)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "mav_trajectory_generation/polynomial_optimization_linear.h"
#include "mav_trajectory_generation/trajectory.h"
#include "mav_trajectory_generation/vertex.h"

namespace mtg = mav_trajectory_generation;

// ref createRandomVerticesPath (src/polynomial_timing_evaluation.cpp:34-91): unit-box directions scaled by a
// uniform random length in [0, 2 average_distance]; the reference keeps the LAST STEP (not the last position)
// as the base of the next vertex, which is reproduced here.
static mtg::Vertex::Vector randomVerticesPath(int dimension, size_t n_segments, double average_distance,
                                             int maximum_derivative, size_t seed) {
  std::mt19937 generator(seed);
  std::vector<std::uniform_real_distribution<double> > axis(dimension, std::uniform_real_distribution<double>(-1, 1));
  std::uniform_real_distribution<double> random_distance(0, 2 * average_distance);
  Eigen::VectorXd last(dimension);
  for (int i = 0; i < dimension; ++i) last[i] = axis[i](generator);
  mtg::Vertex::Vector vertices;
  vertices.push_back(mtg::Vertex(dimension));
  vertices.front().makeStartOrEnd(last, maximum_derivative);
  for (size_t v = 1; v <= n_segments; ++v) {
    Eigen::VectorXd step(dimension);
    do {
      for (int d = 0; d < dimension; ++d) step[d] = axis[d](generator);
    } while (!(step.norm() > 0.2));
    step = step * (1.0 / step.norm()) * random_distance(generator);
    mtg::Vertex vertex(dimension);
    vertex.addConstraint(mtg::derivative_order::POSITION, step + last);
    vertices.push_back(vertex);
    last = step;
  }
  vertices.back().makeStartOrEnd(last, maximum_derivative);
  return vertices;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? std::atoi(argv[1]) : 1000;
  const int segment_counts[4] = {2, 10, 50, 100};
  std::printf("{\"procedure\": \"ref src/polynomial_timing_evaluation.cpp:93-127 through the C++ mirror (B = 1)\", \"repetitions\": %d, \"setup_and_solve_ms\": {", reps);
  mtg::Trajectory trajectory;
  for (int j = 0; j < 4; ++j) {
    const int K = segment_counts[j];
    double total = 0.0, worst = 0.0;
    for (int i = 0; i < reps + 3; ++i) {
      mtg::Vertex::Vector vertices = randomVerticesPath(3, K, 5.0, mtg::derivative_order::SNAP, 1);
      std::vector<double> times = mtg::estimateSegmentTimes(vertices, 2.0, 2.0, 6.5);
      const double t0 = now_ms();
      mtg::PolynomialOptimization<10> opt(3);
      opt.setupFromVertices(vertices, times, mtg::derivative_order::SNAP);
      opt.solveLinear();
      const double dt = now_ms() - t0;
      if (i >= 3) {   // three untimed warm-up calls (context creation, first launch)
        total += dt;
        worst = dt > worst ? dt : worst;
      }
      if (K == 10 && i == 0) opt.getTrajectory(&trajectory);
    }
    std::printf("%s\"%d\": {\"mean\": %.6f, \"max\": %.6f}", j ? ", " : "", K, total / reps, worst);
  }
  std::printf("}, ");
  // scalar evaluate(): one round trip per call
  const int n_eval = 2000;
  const double T = trajectory.getMaxTime();
  double sink = 0.0;
  for (int i = 0; i < 20; ++i) sink += trajectory.evaluate(0.5 * T, 0)[0];
  double t0 = now_ms();
  for (int i = 0; i < n_eval; ++i) sink += trajectory.evaluate(T * (i + 0.5) / n_eval, 1)[0];
  const double scalar_us = (now_ms() - t0) * 1e3 / n_eval;
  // the same instants in one call
  std::vector<double> instants(100000);
  for (size_t i = 0; i < instants.size(); ++i) instants[i] = T * (i + 0.5) / instants.size();
  sink += trajectory.evaluateBatch(instants, 5)[0];
  t0 = now_ms();
  sink += trajectory.evaluateBatch(instants, 5)[0];
  const double batch_us = (now_ms() - t0) * 1e3 / instants.size();
  // the reference's test helper walks a trajectory at dt = 1e-3 with scalar evaluate calls (test/...:61-71):
  // evaluateRange does the same walk in one launch
  std::vector<Eigen::VectorXd> range;
  trajectory.evaluateRange(0.0, T, 1e-3, 0, &range);
  t0 = now_ms();
  trajectory.evaluateRange(0.0, T, 1e-3, 0, &range);
  const double range_ms = now_ms() - t0;
  std::printf("\"scalar_evaluate_us_per_call\": %.3f, \"evaluate_batch_us_per_instant\": %.5f, "
              "\"evaluate_range\": {\"samples\": %zu, \"ms\": %.4f, \"us_per_sample\": %.5f}, \"checksum\": %.6g}\n",
              scalar_us, batch_us, range.size(), range_ms, range_ms * 1e3 / (range.size() ? range.size() : 1), sink);
  return 0;
}
