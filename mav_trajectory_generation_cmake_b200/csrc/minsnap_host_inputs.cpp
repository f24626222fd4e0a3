// Host-only input generator of the C ABI: the batched form of the reference's
// createRandomVertices (ref: src/vertex.cpp:27-79).  Row a3 of SURVEY.md section 8 keeps
// this on the host: it is the synthetic-workload generator, run once, and it must draw
// from std::mt19937 + std::uniform_real_distribution<double> exactly like the reference so
// that the oracle and the GPU solve identical problems.
#include "../../include/minsnap_b200.h"

#include <cmath>
#include <random>
#include <vector>

extern "C" int minsnap_random_positions_host(long B, int K, int D, const double* h_pos_min, const double* h_pos_max,
                                             uint64_t base_seed, double* h_positions) {
  if (B < 0 || K < 1 || D < 1 || !h_pos_min || !h_pos_max || !h_positions) return MINSNAP_ERR_ARG;
  const double min_distance = 0.2;
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {
    std::mt19937 generator(static_cast<std::mt19937::result_type>(base_seed + static_cast<uint64_t>(b)));
    std::vector<std::uniform_real_distribution<double>> axis;
    for (int d = 0; d < D; ++d) axis.emplace_back(h_pos_min[d], h_pos_max[d]);
    double* out = h_positions + static_cast<size_t>(b) * (K + 1) * D;
    for (int d = 0; d < D; ++d) out[d] = axis[d](generator);
    for (int v = 1; v <= K; ++v) {
      double* cur = out + static_cast<size_t>(v) * D;
      const double* prev = cur - D;
      double dist2;
      do {  // reject points closer than min_distance to the previous vertex
        dist2 = 0.0;
        for (int d = 0; d < D; ++d) {
          cur[d] = axis[d](generator);
          dist2 += (cur[d] - prev[d]) * (cur[d] - prev[d]);
        }
      } while (!(std::sqrt(dist2) > min_distance));
    }
  }
  return MINSNAP_OK;
}
