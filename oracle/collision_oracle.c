/*
 * collision_oracle.c -- TEST INFRASTRUCTURE ONLY (see minsnap_oracle.c).
 *
 * CPU restatement of the reference's collision cost of a polynomial trajectory:
 *   getCostAndGradientCollision        NL.i:1523-1709   (NL.i = include/mav_trajectory_generation/impl/
 *   getCostAndGradientPotentialESDF    NL.i:1713-1806            polynomial_optimization_nonlinear_impl.h)
 *   getNeighborsSDF / getDistanceSDF   NL.i:1808-1905
 *   getCostPotential                   NL.i:2319-2345
 *   lerp / triLerp                     NL.i:2435-2464
 * in the reference's order of operations: positions as sum_n pow(t, n) c_n, velocities through the derivative
 * of the coefficients, t advanced by repeated addition of dt, the running path-length / time integrals with
 * their resets, and the end-of-segment correction  time_sum += -dt + (T_i - t).
 *
 * The map classes of the reference (voxblox::EsdfMap, sdf_tools::SignedDistanceField) are un-vendored
 * dependencies.  The map here is the plain dense grid the C ABI defines (include/minsnap_b200.h):
 *   value(i, j, k) = data[(i ny + j) nz + k], the distance at the CENTRE of cell (i, j, k);
 *   cell of a point  i = floor((x - origin_x) / resolution)   (likewise j, k);
 *   centre of a cell  origin + (i + 0.5) resolution;
 *   a point outside the grid reads `oob_value` (sdf_tools: the default value of the field).
 * Two reference behaviours are reproduced on purpose: the continuous distance interpolates between the cells
 * idx-1 and idx+1 (a two-cell stencil, NL.i:1814-1841), and triLerp blends its x-interpolants of (y0,z0) and
 * (y0,z1) with the Y weight before blending with the Z weight (NL.i:2456-2463).  (The reference's map classes
 * return float distances; this grid is double.)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define API __attribute__((visibility("default")))

typedef struct {
  const double* data;
  int nx, ny, nz;
  double origin[3];
  double resolution;
  double oob_value;
} Grid;

typedef struct {
  double min_bound[3], max_bound[3];
  int use_continuous_distance;
  double map_resolution;   /* increment of the numeric potential gradient and the path-length threshold */
  double epsilon, robot_radius, coll_pot_multiplier;
} PotentialParams;

static int cell_of(const Grid* g, double x, int axis) { return (int)floor((x - g->origin[axis]) / g->resolution); }

static int grid_get_safe(const Grid* g, int i, int j, int k, double* value) {
  if (i < 0 || j < 0 || k < 0 || i >= g->nx || j >= g->ny || k >= g->nz) return 0;
  *value = g->data[((size_t)i * g->ny + j) * g->nz + k];
  return 1;
}

/* discrete lookup: sdf_->Get(x, y, z) */
static double grid_get(const Grid* g, const double* p) {
  double v;
  if (grid_get_safe(g, cell_of(g, p[0], 0), cell_of(g, p[1], 1), cell_of(g, p[2], 2), &v)) return v;
  return g->oob_value;
}

/* NL.i:2435-2439 */
static double lerp(double x, double x1, double x2, double q00, double q01) {
  return ((x2 - x) / (x2 - x1)) * q00 + ((x - x1) / (x2 - x1)) * q01;
}

/* NL.i:2451-2464, argument order and blending order as in the reference */
static double tri_lerp(double x, double y, double z, double q000, double q001, double q010, double q011, double q100,
                       double q101, double q110, double q111, double x1, double x2, double y1, double y2, double z1,
                       double z2) {
  const double x00 = lerp(x, x1, x2, q000, q100);
  const double x10 = lerp(x, x1, x2, q010, q110);
  const double x01 = lerp(x, x1, x2, q001, q101);
  const double x11 = lerp(x, x1, x2, q011, q111);
  const double r0 = lerp(y, y1, y2, x00, x01);
  const double r1 = lerp(y, y1, y2, x10, x11);
  return lerp(z, z1, z2, r0, r1);
}

/* NL.i:1843-1905 getDistanceSDF */
static double distance_continuous(const Grid* g, const double* p) {
  const int ix = cell_of(g, p[0], 0), iy = cell_of(g, p[1], 1), iz = cell_of(g, p[2], 2);
  double q[8];
  int valid = 1, n = 0;
  for (int a = -1; a <= 1; a += 2)
    for (int b = -1; b <= 1; b += 2)
      for (int c = -1; c <= 1; c += 2) valid &= grid_get_safe(g, ix + a, iy + b, iz + c, &q[n++]);   /* q000, q001, ... q111 */
  if (!valid) return grid_get(g, p);
  const double r = g->resolution;
  const double x0 = g->origin[0] + (ix - 1 + 0.5) * r, x1 = g->origin[0] + (ix + 1 + 0.5) * r;
  const double y0 = g->origin[1] + (iy - 1 + 0.5) * r, y1 = g->origin[1] + (iy + 1 + 0.5) * r;
  const double z0 = g->origin[2] + (iz - 1 + 0.5) * r, z1 = g->origin[2] + (iz + 1 + 0.5) * r;
  return tri_lerp(p[0], p[1], p[2], q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], x0, x1, y0, y1, z0, z1);
}

/* NL.i:2319-2345 getCostPotential */
static double cost_potential(const PotentialParams* pp, double d, int* is_collision) {
  *is_collision = 0;
  double cost = 0.0;
  d -= pp->robot_radius;
  if (d <= 0.0) {
    cost = pp->coll_pot_multiplier * (-d) + 0.5 * pp->epsilon;
    *is_collision = 1;
  } else if (d <= pp->epsilon) {
    const double e = d - pp->epsilon;
    cost = 0.5 * 1.0 / pp->epsilon * e * e;
  }
  return cost;
}

static double distance_at(const Grid* g, const PotentialParams* pp, const double* p, int valid_state) {
  return (valid_state && pp->use_continuous_distance) ? distance_continuous(g, p) : grid_get(g, p);
}

/* NL.i:1713-1806 getCostAndGradientPotentialESDF; gradient may be NULL */
static double potential(const Grid* g, const PotentialParams* pp, const double* p, double* gradient, int* is_collision) {
  const double inc = pp->map_resolution;
  int valid_state = 1;
  for (int k = 0; k < 3; ++k)
    if (p[k] < pp->min_bound[k] + inc || p[k] > pp->max_bound[k] - inc) valid_state = 0;
  const double cost = cost_potential(pp, distance_at(g, pp, p, valid_state), is_collision);
  if (gradient) {
    for (int k = 0; k < 3; ++k) {
      double lo[3] = {p[0], p[1], p[2]}, hi[3] = {p[0], p[1], p[2]};
      lo[k] -= inc;
      hi[k] += inc;
      int cl, cr;
      const double left = cost_potential(pp, distance_at(g, pp, lo, valid_state), &cl);
      const double right = cost_potential(pp, distance_at(g, pp, hi, valid_state), &cr);
      gradient[k] = (right - left) / (2.0 * inc);
    }
  }
  return cost;
}

API double orc_collision_potential(const double* data, int nx, int ny, int nz, const double* origin, double resolution,
                                   double oob_value, const double* min_bound, const double* max_bound,
                                   int use_continuous_distance, double map_resolution, double epsilon, double robot_radius,
                                   double coll_pot_multiplier, const double* position, double* gradient, int* is_collision) {
  Grid g = {data, nx, ny, nz, {origin[0], origin[1], origin[2]}, resolution, oob_value};
  PotentialParams pp = {{min_bound[0], min_bound[1], min_bound[2]}, {max_bound[0], max_bound[1], max_bound[2]},
                        use_continuous_distance, map_resolution, epsilon, robot_radius, coll_pot_multiplier};
  return potential(&g, &pp, position, gradient, is_collision);
}

/* NL.i:1523-1709 getCostAndGradientCollision, cost and collision flag (gradients == NULL branch).
 * coeffs [K][3][N] increasing powers, times [K].  Also reports how many samples were charged. */
API double orc_collision_cost(int N, int K, const double* coeffs, const double* times, const double* data, int nx, int ny,
                              int nz, const double* origin, double resolution, double oob_value, const double* min_bound,
                              const double* max_bound, int use_continuous_distance, double dt, double map_resolution,
                              double epsilon, double robot_radius, double coll_pot_multiplier, int* is_collision,
                              int* n_charged) {
  Grid g = {data, nx, ny, nz, {origin[0], origin[1], origin[2]}, resolution, oob_value};
  PotentialParams pp = {{min_bound[0], min_bound[1], min_bound[2]}, {max_bound[0], max_bound[1], max_bound[2]},
                        use_continuous_distance, map_resolution, epsilon, robot_radius, coll_pot_multiplier};
  double J_c = 0.0;
  int any_collision = 0, charged = 0;
  double prev[3] = {0.0, 0.0, 0.0};
  double time_sum = -1.0, dist_sum = 0.0, t = 0.0;
  for (int i = 0; i < K; ++i) {
    for (t = 0.0; t < times[i]; t += dt) {
      double pos[3], vel[3];
      for (int k = 0; k < 3; ++k) {
        const double* c = coeffs + ((size_t)i * 3 + k) * N;
        double p = 0.0, v = 0.0;
        for (int n = 0; n < N; ++n) p += pow(t, n) * c[n];                     /* T^T p_k */
        for (int n = 0; n + 1 < N; ++n) v += pow(t, n) * ((n + 1) * c[n + 1]);   /* T^T V p_k */
        pos[k] = p;
        vel[k] = v;
      }
      if (time_sum < 0) {   /* numerical integration: skip the first sample */
        time_sum = 0.0;
        prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
        continue;
      }
      time_sum += dt;
      {
        const double dx = pos[0] - prev[0], dy = pos[1] - prev[1], dz = pos[2] - prev[2];
        dist_sum += sqrt(dx * dx + dy * dy + dz * dz);
      }
      prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
      if (dist_sum < map_resolution) continue;
      int hit;
      const double c = potential(&g, &pp, pos, 0, &hit);
      if (hit) any_collision = 1;
      J_c += c * sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]) * time_sum;
      ++charged;
      dist_sum = 0.0;
      time_sum = 0.0;
    }
    time_sum += -dt + (times[i] - t);   /* make sure the dt is correct for the next segment */
  }
  if (is_collision) *is_collision = any_collision;
  if (n_charged) *n_charged = charged;
  return J_c;
}

/* NL.i:1523-1709 getCostAndGradientCollision with gradients != NULL: the cost as above plus, per charged sample
 * whose speed exceeds 1e-6 (NL.i:1668; slower samples are charged but their gradient is dropped), for every axis k
 *   grad_c[k] += ( |v| time_sum dc/dx_k  T_all^T L_pp  +  time_sum c v_k / |v|  T_all^T V_all L_pp )^T      (NL.i:1672-1679)
 * dc/dx the central difference of the potential over +-map_resolution (getCostAndGradientPotentialESDF,
 * NL.i:1756-1785), L = A^-1 M the map from [d_f; d_p] to the coefficients (NL.i:200-222), L_pp its free columns,
 * V the derivative matrix (V[n][n+1] = n + 1).  The caller passes the per-segment A^-1 [K][N][N] (the oracle's own
 * invertMappingMatrix) and the reference's constraint index map; L_pp and V_all L_pp are formed densely once, as
 * the reference holds them, and every sample takes its row-vector products with the rows of its segment (the
 * other entries of T_all are zero).  grad [n_free][3]. */
API double orc_collision_cost_gradient(int N, int K, const double* coeffs, const double* times, const double* ainv,
                                       const int32_t* col_of_row, int n_fixed, int n_free, const double* data, int nx,
                                       int ny, int nz, const double* origin, double resolution, double oob_value,
                                       const double* min_bound, const double* max_bound, int use_continuous_distance,
                                       double dt, double map_resolution, double epsilon, double robot_radius,
                                       double coll_pot_multiplier, double* Lpp /* scratch [K N][n_free] */,
                                       double* VLpp /* scratch [K N][n_free] */, double* grad, int* is_collision,
                                       int* n_charged) {
  Grid g = {data, nx, ny, nz, {origin[0], origin[1], origin[2]}, resolution, oob_value};
  PotentialParams pp = {{min_bound[0], min_bound[1], min_bound[2]}, {max_bound[0], max_bound[1], max_bound[2]},
                        use_continuous_distance, map_resolution, epsilon, robot_radius, coll_pot_multiplier};
  /* L_pp = (A^-1 M)[:, free]: M maps compact column col_of_row[i N + r] onto row r of segment i */
  for (size_t e = 0; e < (size_t)K * N * n_free; ++e) Lpp[e] = VLpp[e] = 0.0;
  for (int i = 0; i < K; ++i)
    for (int r = 0; r < N; ++r) {
      const int col = col_of_row[i * N + r] - n_fixed;
      if (col < 0) continue;
      for (int n = 0; n < N; ++n) Lpp[((size_t)i * N + n) * n_free + col] += ainv[((size_t)i * N + n) * N + r];
    }
  for (int i = 0; i < K; ++i)
    for (int n = 0; n + 1 < N; ++n)
      for (int col = 0; col < n_free; ++col)
        VLpp[((size_t)i * N + n) * n_free + col] = (n + 1) * Lpp[((size_t)i * N + n + 1) * n_free + col];
  for (int e = 0; e < n_free * 3; ++e) grad[e] = 0.0;

  double J_c = 0.0;
  int any_collision = 0, charged = 0;
  double prev[3] = {0.0, 0.0, 0.0};
  double time_sum = -1.0, dist_sum = 0.0, t = 0.0;
  for (int i = 0; i < K; ++i) {
    for (t = 0.0; t < times[i]; t += dt) {
      double Tv[32];
      for (int n = 0; n < N; ++n) Tv[n] = pow(t, n);
      double pos[3], vel[3];
      for (int k = 0; k < 3; ++k) {
        const double* c = coeffs + ((size_t)i * 3 + k) * N;
        double p = 0.0, v = 0.0;
        for (int n = 0; n < N; ++n) p += Tv[n] * c[n];
        for (int n = 0; n + 1 < N; ++n) v += Tv[n] * ((n + 1) * c[n + 1]);
        pos[k] = p;
        vel[k] = v;
      }
      if (time_sum < 0) {
        time_sum = 0.0;
        prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
        continue;
      }
      time_sum += dt;
      {
        const double dx = pos[0] - prev[0], dy = pos[1] - prev[1], dz = pos[2] - prev[2];
        dist_sum += sqrt(dx * dx + dy * dy + dz * dz);
      }
      prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
      if (dist_sum < map_resolution) continue;
      int hit;
      double gpot[3];
      const double c = potential(&g, &pp, pos, gpot, &hit);
      if (hit) any_collision = 1;
      const double vnorm = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
      J_c += c * vnorm * time_sum;
      ++charged;
      if (vnorm > 1e-6) {
        for (int col = 0; col < n_free; ++col) {
          double tl = 0.0, tvl = 0.0;   /* T_all^T L_pp and T_all^T V_all L_pp, column col */
          for (int n = 0; n < N; ++n) {
            tl += Tv[n] * Lpp[((size_t)i * N + n) * n_free + col];
            tvl += Tv[n] * VLpp[((size_t)i * N + n) * n_free + col];
          }
          for (int k = 0; k < 3; ++k)
            grad[col * 3 + k] += vnorm * time_sum * gpot[k] * tl + time_sum * c * vel[k] / vnorm * tvl;
        }
      }
      dist_sum = 0.0;
      time_sum = 0.0;
    }
    time_sum += -dt + (times[i] - t);
  }
  if (is_collision) *is_collision = any_collision;
  if (n_charged) *n_charged = charged;
  return J_c;
}
