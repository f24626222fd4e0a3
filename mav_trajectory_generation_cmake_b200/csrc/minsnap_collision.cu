// Collision cost of solved trajectories against a signed-distance grid.
//
// ref: getCostAndGradientCollision (NL.i:1523-1709, cost and collision flag), getCostAndGradientPotentialESDF
// (NL.i:1713-1806), getDistanceSDF / getNeighborsSDF (NL.i:1808-1905), getCostPotential (NL.i:2319-2345),
// lerp / triLerp (NL.i:2435-2464); NL.i = include/mav_trajectory_generation/impl/polynomial_optimization_nonlinear_impl.h.
//
// The reference walks a trajectory at a fixed time increment, integrates path length and time between the
// samples it charges (a sample is charged once the path since the last charged one reaches the map resolution),
// and adds  potential(position) |velocity| time_sum  per charged sample.  The walk is sequential in its two
// running sums; everything else is not.  One WARP per trajectory: the 32 lanes evaluate 32 consecutive samples
// (position and velocity by Horner's rule on the segment's coefficients, which every lane holds), the running
// sums are advanced over those 32 step lengths by every lane redundantly (shuffles, no shared memory, no
// divergence), and only the lanes whose sample is charged touch the grid: 8 values for the reference's
// two-cell trilinear stencil, or one for the discrete lookup.  Nothing but the coefficients, the segment times
// and the touched grid cells is read; one cost and one flag per trajectory are written -- the 120 bytes per
// sample that a sampled trajectory would cost never exist.
//
// Sample times advance by repeated addition (t += dt) exactly as the reference's loop does, the first sample of
// a walk and the first sample after a segment whose correction left time_sum negative are skipped as in the
// reference, and triLerp keeps the reference's blending order.  The map is the dense grid of minsnap_b200.h.
#include "minsnap_device.cuh"
#include "minsnap_launch.h"

namespace minsnap {

namespace {

struct GridView {
  const double* data;
  int nx, ny, nz;
  double ox, oy, oz, res, oob;
};

__device__ __forceinline__ int cell_of(double x, double origin, double res) { return (int)floor((x - origin) / res); }

__device__ __forceinline__ bool grid_get_safe(const GridView& g, int i, int j, int k, double* v) {
  if (i < 0 || j < 0 || k < 0 || i >= g.nx || j >= g.ny || k >= g.nz) return false;
  *v = __ldg(g.data + ((size_t)i * g.ny + j) * g.nz + k);
  return true;
}

__device__ __forceinline__ double grid_get(const GridView& g, double x, double y, double z) {
  double v;
  if (grid_get_safe(g, cell_of(x, g.ox, g.res), cell_of(y, g.oy, g.res), cell_of(z, g.oz, g.res), &v)) return v;
  return g.oob;
}

// ref lerp, NL.i:2435-2439
__device__ __forceinline__ double lerp(double x, double x1, double x2, double q00, double q01) {
  return ((x2 - x) / (x2 - x1)) * q00 + ((x - x1) / (x2 - x1)) * q01;
}

// ref getDistanceSDF, NL.i:1843-1905, with triLerp's blending order (NL.i:2451-2464)
__device__ double distance_continuous(const GridView& g, double x, double y, double z) {
  const int ix = cell_of(x, g.ox, g.res), iy = cell_of(y, g.oy, g.res), iz = cell_of(z, g.oz, g.res);
  double q[8];
  bool valid = true;
  int n = 0;
#pragma unroll
  for (int a = -1; a <= 1; a += 2)
#pragma unroll
    for (int b = -1; b <= 1; b += 2)
#pragma unroll
      for (int c = -1; c <= 1; c += 2) valid &= grid_get_safe(g, ix + a, iy + b, iz + c, &q[n++]);
  if (!valid) return grid_get(g, x, y, z);
  const double x0 = g.ox + (ix - 1 + 0.5) * g.res, x1 = g.ox + (ix + 1 + 0.5) * g.res;
  const double y0 = g.oy + (iy - 1 + 0.5) * g.res, y1 = g.oy + (iy + 1 + 0.5) * g.res;
  const double z0 = g.oz + (iz - 1 + 0.5) * g.res, z1 = g.oz + (iz + 1 + 0.5) * g.res;
  const double x00 = lerp(x, x0, x1, q[0], q[4]);
  const double x10 = lerp(x, x0, x1, q[2], q[6]);
  const double x01 = lerp(x, x0, x1, q[1], q[5]);
  const double x11 = lerp(x, x0, x1, q[3], q[7]);
  const double r0 = lerp(y, y0, y1, x00, x01);
  const double r1 = lerp(y, y0, y1, x10, x11);
  return lerp(z, z0, z1, r0, r1);
}

// ref getCostPotential, NL.i:2319-2345
__device__ __forceinline__ double cost_potential(const CollisionArgs& a, double d, bool* hit) {
  *hit = false;
  double cost = 0.0;
  d -= a.robot_radius;
  if (d <= 0.0) {
    cost = a.coll_pot_multiplier * (-d) + 0.5 * a.epsilon;
    *hit = true;
  } else if (d <= a.epsilon) {
    const double e = d - a.epsilon;
    cost = 0.5 * 1.0 / a.epsilon * e * e;
  }
  return cost;
}

// ref getCostAndGradientPotentialESDF, NL.i:1713-1753 (value) and, with grad != null, NL.i:1756-1785: the central
// difference of the potential over +-map_resolution along each axis (continuous or discrete distance as the
// CENTRE position decides)
__device__ double potential(const CollisionArgs& a, const GridView& g, double x, double y, double z, bool* hit,
                            double* grad = nullptr) {
  const double inc = a.map_resolution;
  const bool valid_state = !(x < a.min_bound[0] + inc || x > a.max_bound[0] - inc || y < a.min_bound[1] + inc ||
                             y > a.max_bound[1] - inc || z < a.min_bound[2] + inc || z > a.max_bound[2] - inc);
  const bool cont = valid_state && a.use_continuous_distance;
  const double d = cont ? distance_continuous(g, x, y, z) : grid_get(g, x, y, z);
  if (grad) {
    for (int k = 0; k < 3; ++k) {
      const double ex = k == 0 ? inc : 0.0, ey = k == 1 ? inc : 0.0, ez = k == 2 ? inc : 0.0;
      const double dl = cont ? distance_continuous(g, x - ex, y - ey, z - ez) : grid_get(g, x - ex, y - ey, z - ez);
      const double dr = cont ? distance_continuous(g, x + ex, y + ey, z + ez) : grid_get(g, x + ex, y + ey, z + ez);
      bool h;
      const double cl = cost_potential(a, dl, &h), cr = cost_potential(a, dr, &h);
      grad[k] = (cr - cl) / (2.0 * inc);
    }
  }
  return cost_potential(a, d, hit);
}

// kGrad adds the gradient w.r.t. the free derivatives (NL.i:1666-1686).  The reference adds, per charged sample,
//   |v| time_sum dc/dx_k (T^T L_pp) + time_sum c v_k / |v| (T^T V L_pp)      to axis k's gradient,
// two row vectors through L = A^-1 M per sample.  Both are linear in the sample's monomial vector, so the lanes
// accumulate in COEFFICIENT space instead -- gc[k][n] += alpha_k t^n + beta_k n t^(n-1), 20 FMAs per axis -- and
// once per segment the warp folds the 3 x N sums and applies A^-T in closed form
// (A^-1_T[n][r] = A1inv[n][r] T^(k_r - n)): row r of the segment's end-point constraints receives
// T^(k_r) sum_n A1inv[n][r] T^-n gc[k][n] and adds it to its free column.  Same sum, reassociated.
template <int N, bool kGrad>
__global__ void __launch_bounds__(128) collision_cost_kernel(CollisionArgs a) {
  const int lane = threadIdx.x & 31;
  const long warp = blockIdx.x * (long)(blockDim.x >> 5) + uniform_warp_index();
  const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  GridView g;
  g.data = a.d_sdf; g.nx = a.nx; g.ny = a.ny; g.nz = a.nz;
  g.ox = a.origin[0]; g.oy = a.origin[1]; g.oz = a.origin[2]; g.res = a.resolution; g.oob = a.oob_value;
  const int K = a.K;
  const double dt = a.dt, limit = a.map_resolution;
  for (long b = warp; b < a.B; b += n_warps) {
    const double* cb = a.d_coeffs + b * (long)K * 3 * N;
    const double* tb = a.d_times + b * K;
    double J = 0.0;           // this lane's charged samples
    bool collided = false;
    int charged = 0;
    // the walk's running state, identical in every lane
    double time_sum = -1.0, dist_sum = 0.0;
    double px = 0.0, py = 0.0, pz = 0.0;   // previous sample's position
    for (int i = 0; i < K; ++i) {
      const double T = tb[i];
      double c[3][N], dc[3][N - 1];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int n = 0; n < N; ++n) c[k][n] = __ldg(cb + ((long)i * 3 + k) * N + n);
#pragma unroll
        for (int n = 0; n + 1 < N; ++n) dc[k][n] = (n + 1) * c[k][n + 1];
      }
      double gc[kGrad ? 3 : 1][kGrad ? N : 1];   // this lane's share of dJ/dcoefficients of segment i
      if (kGrad) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int n = 0; n < N; ++n) gc[k][n] = 0.0;
      }
      double t0 = 0.0, t_end = 0.0;
      while (true) {
        // this lane's sample time: t0 advanced `lane` times by dt, the reference's repeated addition
        double t = t0;
        for (int j = 0; j < lane; ++j) t += dt;
        const bool valid = t < T;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const int n_valid = __popc(vmask);   // the valid samples are lanes 0 .. n_valid-1 (t increases with the lane)
        double x = 0.0, y = 0.0, z = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
        if (valid) {
          x = c[0][N - 1]; y = c[1][N - 1]; z = c[2][N - 1];
          vx = dc[0][N - 2]; vy = dc[1][N - 2]; vz = dc[2][N - 2];
#pragma unroll
          for (int n = N - 2; n >= 0; --n) {
            x = fma(x, t, c[0][n]); y = fma(y, t, c[1][n]); z = fma(z, t, c[2][n]);
          }
#pragma unroll
          for (int n = N - 3; n >= 0; --n) {
            vx = fma(vx, t, dc[0][n]); vy = fma(vy, t, dc[1][n]); vz = fma(vz, t, dc[2][n]);
          }
        }
        // path length from the previous sample
        double qx = __shfl_up_sync(0xffffffffu, x, 1), qy = __shfl_up_sync(0xffffffffu, y, 1),
               qz = __shfl_up_sync(0xffffffffu, z, 1);
        if (lane == 0) { qx = px; qy = py; qz = pz; }
        const double ddx = x - qx, ddy = y - qy, ddz = z - qz;
        const double step = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        // the reference's running sums over these samples, advanced by every lane alike
        bool mine = false;
        double my_ts = 0.0;
        for (int s = 0; s < n_valid; ++s) {
          const double step_s = __shfl_sync(0xffffffffu, step, s);
          if (time_sum < 0) {   // first sample of the walk, or of a segment whose predecessor left time_sum negative
            time_sum = 0.0;
            continue;
          }
          time_sum += dt;
          dist_sum += step_s;
          if (dist_sum < limit) continue;
          if (lane == s) { mine = true; my_ts = time_sum; }
          dist_sum = 0.0;
          time_sum = 0.0;
        }
        if (n_valid > 0) {
          px = __shfl_sync(0xffffffffu, x, n_valid - 1);
          py = __shfl_sync(0xffffffffu, y, n_valid - 1);
          pz = __shfl_sync(0xffffffffu, z, n_valid - 1);
        }
        if (mine) {
          bool hit;
          double gp[3];
          const double cp = potential(a, g, x, y, z, &hit, kGrad ? gp : nullptr);
          collided |= hit;
          const double vnorm = sqrt(vx * vx + vy * vy + vz * vz);
          J += cp * vnorm * my_ts;
          ++charged;
          if (kGrad && vnorm > 1e-6) {   // ref NL.i:1668: slower samples keep their cost and drop their gradient
            const double v3[3] = {vx, vy, vz};
            double alpha[3], beta[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              alpha[k] = vnorm * my_ts * gp[k];
              beta[k] = my_ts * cp * v3[k] / vnorm;
            }
            double tp = 1.0, tm = 0.0;   // t^n and n t^(n-1)
#pragma unroll
            for (int n = 0; n < N; ++n) {
#pragma unroll
              for (int k = 0; k < 3; ++k) gc[k][n] = fma(alpha[k], tp, fma(beta[k], tm, gc[k][n]));
              tm = (n + 1) * tp;
              tp *= t;
            }
          }
        }
        if (n_valid < 32) {
          t_end = __shfl_sync(0xffffffffu, t, n_valid);   // the loop variable when the reference's loop ends
          break;
        }
        t0 = __shfl_sync(0xffffffffu, t, 31) + dt;
      }
      time_sum += -dt + (T - t_end);   // ref NL.i:1698: make sure the dt is correct for the next segment
      if (kGrad) {
        // fold the lanes' sums (fixed-order butterfly), then lanes 3 r + k map row r, axis k through A^-T
        const double iT = 1.0 / T;
        double s[3][N];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double ip = 1.0;
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double v = gc[k][n];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            s[k][n] = v * ip;   // T^-n gc[k][n]
            ip *= iT;
          }
        }
        __syncwarp();   // the previous segment's additions to shared columns are visible
        if (lane < 3 * N) {
          const int r = lane / 3, k = lane - 3 * r;
          const int order = r % (N / 2);
          int col;
          if (a.d_col_of_row) {
            col = a.d_col_of_row[i * N + r] - a.n_fixed;
          } else {   // standard mask: interior vertices free in derivatives 1 .. N/2-1
            const int vertex = i + (r >= N / 2 ? 1 : 0);
            col = (order >= 1 && vertex >= 1 && vertex <= K - 1) ? (vertex - 1) * (N / 2 - 1) + (order - 1) : -1;
          }
          if (col >= 0 && col < a.n_free) {
            double acc = 0.0;
#pragma unroll
            for (int n = 0; n < N; ++n) {
              const double sk = k == 0 ? s[0][n] : (k == 1 ? s[1][n] : s[2][n]);
              acc = fma(minsnap_tables::kA1inv_N10[n * N + r], sk, acc);
            }
            double tk = 1.0;
            for (int e = 0; e < order; ++e) tk *= T;
            double* dst = a.d_grad_free + (b * (long)a.n_free + col) * 3 + k;
            *dst += acc * tk;
          }
        }
        __syncwarp();
      }
    }
    // fixed-order butterfly: the sum does not depend on timing
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      J += __shfl_xor_sync(0xffffffffu, J, off);
      charged += __shfl_xor_sync(0xffffffffu, charged, off);
    }
    const bool any = __any_sync(0xffffffffu, collided);
    if (lane == 0) {
      a.d_cost[b] = J;
      if (a.d_is_collision) a.d_is_collision[b] = any ? 1 : 0;
      if (a.d_charged) a.d_charged[b] = charged;
    }
  }
}

}  // namespace

cudaError_t launch_collision_cost(const CollisionArgs& a, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  if (a.N != 10) return cudaErrorInvalidConfiguration;
  const int threads = 128;
  long grid = (a.B * 32 + threads - 1) / threads;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (grid > (long)sms * 32) grid = (long)sms * 32;
  if (a.d_grad_free)
    collision_cost_kernel<10, true><<<(int)grid, threads, 0, stream>>>(a);
  else
    collision_cost_kernel<10, false><<<(int)grid, threads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace minsnap
