"""ctypes front-end of the CPU oracle (oracle/minsnap_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm;
the product package never imports this module.

Two builds of the same C file are exposed: ``f64`` (double arithmetic: the parity oracle and
the CPU baseline) and ``ld`` (x87 long double arithmetic: extended-precision ground truth).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "lib")


def build(force=False):
    """Compile both oracle shared objects with oracle/Makefile (gcc, seconds)."""
    need = force or not all(
        os.path.exists(os.path.join(_LIBDIR, n)) for n in ("liboracle_f64.so", "liboracle_ld.so"))
    src_m = max(os.path.getmtime(os.path.join(_HERE, n)) for n in ("minsnap_oracle.c", "extrema_oracle.c", "collision_oracle.c"))
    if not need:
        need = any(os.path.getmtime(os.path.join(_LIBDIR, n)) < src_m
                   for n in ("liboracle_f64.so", "liboracle_ld.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])


class Oracle:
    def __init__(self, precision="f64"):
        build()
        assert precision in ("f64", "ld")
        self.lib = C.CDLL(os.path.join(_LIBDIR, "liboracle_%s.so" % precision))
        self.real = np.float64 if precision == "f64" else np.longdouble
        self.creal = C.c_double if precision == "f64" else C.c_longdouble
        assert self.lib.orc_real_bytes() == np.dtype(self.real).itemsize
        self.lib.orc_polynomial_evaluate.restype = self.creal
        self.lib.orc_compute_cost.restype = self.creal
        self.lib.orc_segment_magnitude.restype = self.creal
        self.precision = precision

    # -- helpers -------------------------------------------------------------------
    def _r(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=self.real))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    # -- small matrices ------------------------------------------------------------
    def base_coefficients(self, n):
        out = np.zeros((n, n), self.real)
        self.lib.orc_base_coefficients(n, self._p(out))
        return out

    def mapping_matrix(self, N, T):
        A = np.zeros((N, N), self.real)
        self.lib.orc_setup_mapping_matrix(C.c_int(N), self.creal(T), self._p(A))
        return A

    def invert_mapping_matrix(self, A):
        A = self._r(A)
        N = A.shape[0]
        Ai = np.zeros((N, N), self.real)
        st = self.lib.orc_invert_mapping_matrix(N, self._p(A), self._p(Ai))
        assert st == 0
        return Ai

    def dense_inverse(self, M):
        M = self._r(M)
        n = M.shape[0]
        Mi = np.zeros((n, n), self.real)
        assert self.lib.orc_dense_inverse(n, self._p(M), self._p(Mi)) == 0
        return Mi

    def cost_matrix(self, N, derivative, T):
        Q = np.zeros((N, N), self.real)
        self.lib.orc_quadratic_cost_jacobian(C.c_int(N), C.c_int(derivative), self.creal(T), self._p(Q))
        return Q

    def segment_hessian(self, N, derivative, T):
        H = np.zeros((N, N), self.real)
        assert self.lib.orc_segment_hessian(C.c_int(N), C.c_int(derivative), self.creal(T), self._p(H)) == 0
        return H

    # -- reordering ----------------------------------------------------------------
    def reorder(self, N, K, mask):
        mask = np.ascontiguousarray(np.asarray(mask, np.uint8).reshape(K + 1, N // 2))
        col = np.zeros(N * K, np.int32)
        nf = C.c_int32()
        npf = C.c_int32()
        st = self.lib.orc_constraint_reordering(N, K, self._p(mask), self._p(col), C.byref(nf), C.byref(npf))
        assert st == 0
        return col, nf.value, npf.value

    # -- solve ---------------------------------------------------------------------
    def solve(self, N, K, D, derivative, mask, vertex_values, times, want_R=False):
        """mask[(K+1)][N/2], vertex_values[(K+1)][N/2][D], times[K].
        Returns dict(coeffs[K][D][N], d_fixed[D][nf], d_free[D][np], cost, status[, R])."""
        h = N // 2
        mask = np.ascontiguousarray(np.asarray(mask, np.uint8).reshape(K + 1, h))
        vals = self._r(vertex_values).reshape(K + 1, h, D)
        times = self._r(times).reshape(K)
        n_fixed = int(mask.sum())
        n_free = (K + 1) * h - n_fixed
        coeffs = np.zeros((K, D, N), self.real)
        d_fixed = np.zeros((D, n_fixed), self.real)
        d_free = np.zeros((D, n_free), self.real)
        cost = np.zeros(1, self.real)
        R = np.zeros((n_fixed + n_free, n_fixed + n_free), self.real) if want_R else None
        st = self.lib.orc_solve_linear(N, K, D, derivative, self._p(mask), self._p(vals), self._p(times),
                                       self._p(coeffs), self._p(d_fixed), self._p(d_free), self._p(cost),
                                       self._p(R) if want_R else None)
        out = dict(coeffs=coeffs, d_fixed=d_fixed, d_free=d_free, cost=cost[0], status=st)
        if want_R:
            out["R"] = R
        return out

    def coeffs_from_constraints(self, N, K, D, col_of_row, d_all, times):
        col = np.ascontiguousarray(np.asarray(col_of_row, np.int32))
        d_all = self._r(d_all)
        n_all = d_all.shape[1]
        times = self._r(times)
        coeffs = np.zeros((K, D, N), self.real)
        st = self.lib.orc_coeffs_from_constraints(N, K, D, self._p(col), n_all, self._p(d_all),
                                                  self._p(times), self._p(coeffs))
        assert st == 0
        return coeffs

    def compute_cost(self, N, K, D, derivative, coeffs, times):
        coeffs = self._r(coeffs)
        times = self._r(times)
        return self.lib.orc_compute_cost(N, K, D, derivative, self._p(coeffs), self._p(times))

    # -- evaluation ----------------------------------------------------------------
    def polynomial_evaluate(self, c, t, derivative):
        c = self._r(c)
        return self.lib.orc_polynomial_evaluate(C.c_int(c.shape[0]), self._p(c), self.creal(t), C.c_int(derivative))

    def trajectory_evaluate(self, coeffs, times, t, derivative):
        coeffs = self._r(coeffs)
        K, D, N = coeffs.shape
        times = self._r(times)
        out = np.zeros(D, self.real)
        seg = self.lib.orc_trajectory_evaluate(N, K, D, self._p(coeffs), self._p(times), self.creal(t),
                                               C.c_int(derivative), self._p(out))
        return out, seg

    def trajectory_sample(self, coeffs, times, t, n_deriv):
        coeffs = self._r(coeffs)
        K, D, N = coeffs.shape
        times = self._r(times)
        t = self._r(t)
        M = t.shape[0]
        out = np.zeros((M, n_deriv, D), self.real)
        self.lib.orc_trajectory_sample(N, K, D, self._p(coeffs), self._p(times), M, self._p(t), n_deriv, self._p(out))
        return out

    def trajectory_evaluate_range(self, coeffs, times, t_start, t_end, dt, derivative, max_out=1 << 20):
        coeffs = self._r(coeffs)
        K, D, N = coeffs.shape
        times = self._r(times)
        out = np.zeros((max_out, D), self.real)
        st = np.zeros(max_out, self.real)
        n = self.lib.orc_trajectory_evaluate_range(N, K, D, self._p(coeffs), self._p(times), self.creal(t_start),
                                                   self.creal(t_end), self.creal(dt), C.c_int(derivative),
                                                   C.c_int(max_out), self._p(out), self._p(st))
        n = min(n, max_out)
        return out[:n].copy(), st[:n].copy()

    # -- inputs --------------------------------------------------------------------
    def estimate_segment_times(self, positions, v_max, a_max, magic=6.5):
        positions = self._r(positions)
        K = positions.shape[0] - 1
        D = positions.shape[1]
        times = np.zeros(K, self.real)
        self.lib.orc_estimate_segment_times(K, D, self._p(positions), self.creal(v_max), self.creal(a_max),
                                            self.creal(magic), self._p(times))
        return times

    def create_random_positions(self, K, pos_min, pos_max, seed):
        pos_min = np.ascontiguousarray(np.asarray(pos_min, np.float64))
        pos_max = np.ascontiguousarray(np.asarray(pos_max, np.float64))
        D = pos_min.shape[0]
        out = np.zeros((K + 1, D), np.float64)
        self.lib.orc_create_random_positions(K, D, self._p(pos_min), self._p(pos_max), C.c_uint64(seed), self._p(out))
        return out

    def mt19937_uniform(self, seed, a, b, n):
        out = np.zeros(n, np.float64)
        self.lib.orc_mt19937_uniform(C.c_uint64(seed), C.c_double(a), C.c_double(b), n, self._p(out))
        return out

    # -- CPU baseline --------------------------------------------------------------
    def solve_batch_standard(self, positions, times, N=10, derivative=4, max_fixed_derivative=4,
                             n_threads=1, want_coeffs=True):
        """positions[B][K+1][D], times[B][K] (float64) -> coeffs[B][K][D][N], cost[B], status."""
        assert self.precision == "f64"
        positions = np.ascontiguousarray(positions, np.float64)
        times = np.ascontiguousarray(times, np.float64)
        B, K1, D = positions.shape
        K = K1 - 1
        coeffs = np.zeros((B, K, D, N), np.float64) if want_coeffs else None
        cost = np.zeros(B, np.float64)
        st = self.lib.orc_solve_batch_standard(N, K, D, derivative, max_fixed_derivative, C.c_long(B),
                                               self._p(positions), self._p(times),
                                               self._p(coeffs) if want_coeffs else None, self._p(cost),
                                               C.c_int(n_threads))
        return coeffs, cost, st


    # -- extrema of the magnitude of a derivative (oracle/extrema_oracle.c) ------------------
    @staticmethod
    def _dims(D, dims):
        d = np.ascontiguousarray(np.arange(D) if dims is None else np.asarray(dims), dtype=np.int32)
        return d

    def candidate_polynomial(self, seg_coeffs, derivative, dims=None):
        """seg_coeffs [D][N] -> coefficients (increasing) of the polynomial whose real roots are
        the candidate times, before trailing-coefficient removal."""
        c = self._r(seg_coeffs)
        D, N = c.shape
        d = self._dims(D, dims)
        g = np.zeros(2 * N, self.real)
        n = self.lib.orc_candidate_polynomial(N, D, self._p(c), C.c_int(derivative), self._p(d), len(d), self._p(g))
        return g[:n].copy()

    def last_nonzero_coefficient(self, g):
        g = self._r(g)
        return self.lib.orc_last_nonzero_coefficient(self._p(g), len(g))

    def real_roots_in_range(self, g, t0, t1):
        g = self._r(g)
        out = np.zeros(max(len(g), 1), self.real)
        n = self.lib.orc_real_roots_in_range(self._p(g), len(g), self.creal(t0), self.creal(t1), self._p(out))
        return out[:n].copy()

    def segment_candidate_roots(self, seg_coeffs, derivative, t_start, t_end, dims=None, keep_small=False):
        c = self._r(seg_coeffs)
        D, N = c.shape
        d = self._dims(D, dims)
        out = np.zeros(2 * N, self.real)
        n = self.lib.orc_segment_candidate_roots(N, D, self._p(c), C.c_int(derivative), self._p(d), len(d),
                                                 self.creal(t_start), self.creal(t_end), C.c_int(int(keep_small)),
                                                 self._p(out))
        return out[:n].copy()

    def minmax_magnitude(self, coeffs, times, derivative, mode, dims=None, want_candidates=False,
                         keep_small=False):
        """coeffs [K][D][N], times [K].  mode 0: computeMaximumOfMagnitude; mode 1:
        Trajectory::computeMinMaxMagnitude.  Returns dict(max=(t, v, seg), min=(t, v, seg)[,
        candidates=list of per-segment root arrays])."""
        c = self._r(coeffs)
        K, D, N = c.shape
        tm = self._r(times).reshape(K)
        d = self._dims(D, dims)
        out = np.zeros(6, self.real)
        max_cand = 2 * N
        ct = np.zeros((K, max_cand), self.real)
        cc = np.zeros(K, np.int32)
        self.lib.orc_minmax_magnitude(C.c_int(mode), K, D, N, self._p(c), self._p(tm), C.c_int(derivative),
                                      self._p(d), len(d), C.c_int(int(keep_small)), self._p(out), self._p(ct),
                                      self._p(cc), max_cand)
        r = {"max": (out[0], out[1], int(out[2])), "min": (out[3], out[4], int(out[5]))}
        if want_candidates:
            r["candidates"] = [ct[s, :cc[s]].copy() for s in range(K)]
        return r

    def segment_magnitude(self, seg_coeffs, derivative, t, dims=None):
        c = self._r(seg_coeffs)
        D, N = c.shape
        d = self._dims(D, dims)
        return self.lib.orc_segment_magnitude(N, D, self._p(c), C.c_int(derivative), self._p(d), len(d),
                                              self.creal(t))


def collision_cost(coeffs, times, sdf, origin, resolution, min_bound, max_bound, dt=0.1, map_resolution=None,
                   epsilon=0.5, robot_radius=0.5, coll_pot_multiplier=1.0, use_continuous_distance=True, oob_value=0.0):
    """oracle/collision_oracle.c: ref getCostAndGradientCollision (NL.i:1523-1709), one trajectory.
    coeffs [K][3][N], times [K], sdf [nx][ny][nz] -> (cost, is_collision, n_charged)."""
    build()
    lib = C.CDLL(os.path.join(_LIBDIR, "liboracle_f64.so"))
    lib.orc_collision_cost.restype = C.c_double
    d = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    c, t, g = d(coeffs), d(times), d(sdf)
    K, D, N = c.shape
    assert D == 3
    org, lo, hi = d(origin), d(min_bound), d(max_bound)
    if map_resolution is None:
        map_resolution = resolution
    hit, charged = C.c_int(), C.c_int()
    cost = lib.orc_collision_cost(N, K, p(c), p(t), p(g), g.shape[0], g.shape[1], g.shape[2], p(org), C.c_double(resolution),
                                  C.c_double(oob_value), p(lo), p(hi), int(bool(use_continuous_distance)), C.c_double(dt),
                                  C.c_double(map_resolution), C.c_double(epsilon), C.c_double(robot_radius),
                                  C.c_double(coll_pot_multiplier), C.byref(hit), C.byref(charged))
    return cost, hit.value, charged.value


def collision_cost_gradient(coeffs, times, col_of_row, n_fixed, n_free, sdf, origin, resolution, min_bound, max_bound,
                            dt=0.1, map_resolution=None, epsilon=0.5, robot_radius=0.5, coll_pot_multiplier=1.0,
                            use_continuous_distance=True, oob_value=0.0):
    """oracle/collision_oracle.c: ref getCostAndGradientCollision with gradients (NL.i:1523-1709), one trajectory.
    coeffs [K][3][N], times [K], col_of_row [K N] (the reference's constraint index map) ->
    (cost, grad [n_free][3] w.r.t. the free derivatives, is_collision, n_charged)."""
    build()
    lib = C.CDLL(os.path.join(_LIBDIR, "liboracle_f64.so"))
    lib.orc_collision_cost_gradient.restype = C.c_double
    d = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    c, t, g = d(coeffs), d(times), d(sdf)
    K, D, N = c.shape
    assert D == 3
    orc = Oracle("f64")
    ainv = np.stack([orc.invert_mapping_matrix(orc.mapping_matrix(N, float(T))) for T in t])
    ainv = d(ainv)
    cr = np.ascontiguousarray(np.asarray(col_of_row, np.int32))
    org, lo, hi = d(origin), d(min_bound), d(max_bound)
    if map_resolution is None:
        map_resolution = resolution
    grad = np.zeros((n_free, 3))
    s1, s2 = np.zeros((K * N, max(n_free, 1))), np.zeros((K * N, max(n_free, 1)))
    hit, charged = C.c_int(), C.c_int()
    cost = lib.orc_collision_cost_gradient(N, K, p(c), p(t), p(ainv), p(cr), int(n_fixed), int(n_free), p(g), g.shape[0],
                                           g.shape[1], g.shape[2], p(org), C.c_double(resolution), C.c_double(oob_value),
                                           p(lo), p(hi), int(bool(use_continuous_distance)), C.c_double(dt),
                                           C.c_double(map_resolution), C.c_double(epsilon), C.c_double(robot_radius),
                                           C.c_double(coll_pot_multiplier), p(s1), p(s2), p(grad), C.byref(hit),
                                           C.byref(charged))
    return cost, grad, hit.value, charged.value


def collision_potential(position, sdf, origin, resolution, min_bound, max_bound, map_resolution=None, epsilon=0.5,
                        robot_radius=0.5, coll_pot_multiplier=1.0, use_continuous_distance=True, oob_value=0.0):
    """ref getCostAndGradientPotentialESDF (NL.i:1713-1806): (cost, numeric gradient [3], is_collision)."""
    build()
    lib = C.CDLL(os.path.join(_LIBDIR, "liboracle_f64.so"))
    lib.orc_collision_potential.restype = C.c_double
    d = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    g, org, lo, hi, pos = d(sdf), d(origin), d(min_bound), d(max_bound), d(position)
    if map_resolution is None:
        map_resolution = resolution
    grad = np.zeros(3)
    hit = C.c_int()
    cost = lib.orc_collision_potential(p(g), g.shape[0], g.shape[1], g.shape[2], p(org), C.c_double(resolution),
                                       C.c_double(oob_value), p(lo), p(hi), int(bool(use_continuous_distance)),
                                       C.c_double(map_resolution), C.c_double(epsilon), C.c_double(robot_radius),
                                       C.c_double(coll_pot_multiplier), p(pos), p(grad), C.byref(hit))
    return cost, grad, hit.value


def standard_mask(K, N=10, max_fixed_derivative=4):
    """Mask produced by createRandomVertices (src/vertex.cpp:59,71-76): end vertices fix
    derivatives 0..max_fixed_derivative, interior vertices fix position only."""
    h = N // 2
    m = np.zeros((K + 1, h), np.uint8)
    m[:, 0] = 1
    m[0, : max_fixed_derivative + 1] = 1
    m[K, : max_fixed_derivative + 1] = 1
    return m


def vertex_values_from_positions(positions, N=10):
    """[K+1][D] positions -> [K+1][N/2][D] constraint-value table (derivatives zero)."""
    positions = np.asarray(positions)
    K1, D = positions.shape
    v = np.zeros((K1, N // 2, D), positions.dtype)
    v[:, 0, :] = positions
    return v


class ReferenceRpoly:
    """The reference's own Jenkins-Traub root finder (oracle/_ref/librpoly_ref.so, built by
    oracle/Makefile from /root/reference/.../src/rpoly.cpp).  Not re-entrant.  TEST USE ONLY."""

    PATH = os.path.join(_HERE, "_ref", "librpoly_ref.so")

    @classmethod
    def available(cls):
        if not os.path.exists(cls.PATH) and os.path.exists("/root/reference/mav_trajectory_generation/src/rpoly.cpp"):
            subprocess.call(["make", "-C", _HERE, "-s", "ref"])
        return os.path.exists(cls.PATH)

    def __init__(self):
        assert self.available()
        self.lib = C.CDLL(self.PATH)

    def roots_increasing(self, coeffs_increasing):
        """ref findRootsJenkinsTraub(VectorXd increasing, VectorXcd*) -> complex roots or None."""
        c = np.ascontiguousarray(np.asarray(coeffs_increasing, np.float64))
        re = np.zeros(len(c) + 1)
        im = np.zeros(len(c) + 1)
        n = self.lib.ref_rpoly_increasing(c.ctypes.data_as(C.c_void_p), len(c), re.ctypes.data_as(C.c_void_p),
                                          im.ctypes.data_as(C.c_void_p))
        if n < 0:
            return None
        return re[:n] + 1j * im[:n]

    def real_roots_in_range(self, coeffs_increasing, t0, t1):
        """Selection rule of the reference (LIN.i:423-434 / src/polynomial.cpp:41-52)."""
        r = self.roots_increasing(coeffs_increasing)
        if r is None:
            return None
        keep = [z.real for z in r if abs(z.imag) <= np.finfo(np.float64).eps and t0 <= z.real <= t1]
        return np.array(sorted(keep))


class ReferenceCore:
    """The reference's own value classes (oracle/_ref/libmav_ref_core.so, built by oracle/Makefile from
    /root/reference/.../src/{vertex,polynomial,segment,trajectory,motion_defines,rpoly}.cpp against the
    Eigen / glog stand-ins of oracle/ref_shim/).  TEST USE ONLY: pins the oracle to the reference itself."""

    PATH = os.path.join(_HERE, "_ref", "libmav_ref_core.so")

    @classmethod
    def available(cls):
        if not os.path.exists(cls.PATH) and os.path.exists("/root/reference/mav_trajectory_generation/src/vertex.cpp"):
            subprocess.call(["make", "-C", _HERE, "-s", "ref"])
        return os.path.exists(cls.PATH)

    def __init__(self):
        assert self.available()
        self.lib = C.CDLL(self.PATH)
        self.lib.refc_polynomial_evaluate.restype = C.c_double
        self.lib.refc_trajectory_max_time.restype = C.c_double

    @staticmethod
    def _d(a):
        return np.ascontiguousarray(np.asarray(a, np.float64))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def create_random_positions(self, max_derivative, K, pos_min, pos_max, seed):
        lo, hi = self._d(pos_min), self._d(pos_max)
        D = lo.shape[0]
        pos = np.zeros((K + 1, D))
        nc = np.zeros(K + 1, np.int32)
        rc = self.lib.refc_create_random_positions(max_derivative, K, D, self._p(lo), self._p(hi), C.c_uint64(seed),
                                                   self._p(pos), self._p(nc))
        assert rc == 0
        return pos, nc

    def estimate_segment_times(self, positions, v_max, a_max, magic=6.5):
        pos = self._d(positions)
        K, D = pos.shape[0] - 1, pos.shape[1]
        t = np.zeros(K)
        n = self.lib.refc_estimate_segment_times(K, D, self._p(pos), C.c_double(v_max), C.c_double(a_max),
                                                 C.c_double(magic), self._p(t))
        assert n == K
        return t

    def base_coefficients(self, n):
        out = np.zeros((n, n))
        self.lib.refc_base_coefficients(n, self._p(out))
        return out

    def base_coeffs_with_time(self, N, derivative, t):
        out = np.zeros(N)
        self.lib.refc_base_coeffs_with_time(N, derivative, C.c_double(t), self._p(out))
        return out

    def polynomial_evaluate(self, c, t, derivative):
        c = self._d(c)
        return self.lib.refc_polynomial_evaluate(len(c), self._p(c), C.c_double(t), derivative)

    def polynomial_evaluate_all(self, c, t, n_deriv):
        c = self._d(c)
        out = np.zeros(n_deriv)
        self.lib.refc_polynomial_evaluate_all(len(c), self._p(c), C.c_double(t), n_deriv, self._p(out))
        return out

    def polynomial_get_coefficients(self, c, derivative):
        c = self._d(c)
        out = np.zeros(len(c))
        self.lib.refc_polynomial_get_coefficients(len(c), self._p(c), derivative, self._p(out))
        return out

    def convolve(self, a, b):
        a, b = self._d(a), self._d(b)
        out = np.zeros(len(a) + len(b))
        n = self.lib.refc_convolve(self._p(a), len(a), self._p(b), len(b), self._p(out))
        return out[:n].copy()

    def polynomial_min_max(self, c, t_start, t_end, derivative):
        c = self._d(c)
        out = np.zeros(4)
        ok = self.lib.refc_polynomial_min_max(len(c), self._p(c), C.c_double(t_start), C.c_double(t_end), derivative,
                                              self._p(out))
        return bool(ok), (out[0], out[1]), (out[2], out[3])

    def segment_evaluate(self, seg_coeffs, T, t, derivative):
        c = self._d(seg_coeffs)
        D, N = c.shape
        out = np.zeros(D)
        self.lib.refc_segment_evaluate(N, D, self._p(c), C.c_double(T), C.c_double(t), derivative, self._p(out))
        return out

    def trajectory_evaluate(self, coeffs, times, t, derivative):
        c, tm = self._d(coeffs), self._d(times)
        K, D, N = c.shape
        out = np.zeros(D)
        self.lib.refc_trajectory_evaluate(N, K, D, self._p(c), self._p(tm), C.c_double(t), derivative, self._p(out))
        return out

    def trajectory_max_time(self, coeffs, times):
        c, tm = self._d(coeffs), self._d(times)
        K, D, N = c.shape
        return self.lib.refc_trajectory_max_time(N, K, D, self._p(c), self._p(tm))

    def trajectory_evaluate_range(self, coeffs, times, t_start, t_end, dt, derivative, max_out=1 << 20):
        c, tm = self._d(coeffs), self._d(times)
        K, D, N = c.shape
        out = np.zeros((max_out, D))
        ts = np.zeros(max_out)
        n = self.lib.refc_trajectory_evaluate_range(N, K, D, self._p(c), self._p(tm), C.c_double(t_start), C.c_double(t_end),
                                                    C.c_double(dt), derivative, max_out, self._p(out), self._p(ts))
        n = min(n, max_out)
        return out[:n].copy(), ts[:n].copy()

    def segment_minmax_candidates(self, seg_coeffs, T, derivative, t_start, t_end, dims=None):
        c = self._d(seg_coeffs)
        D, N = c.shape
        d = np.ascontiguousarray(np.arange(D) if dims is None else np.asarray(dims), dtype=np.int32)
        ct, cv = np.zeros(4 * N), np.zeros(4 * N)
        n = self.lib.refc_segment_minmax_candidates(N, D, self._p(c), C.c_double(T), derivative, C.c_double(t_start),
                                                    C.c_double(t_end), self._p(d), len(d), 4 * N, self._p(ct), self._p(cv))
        if n < 0:
            return None
        return ct[:n].copy(), cv[:n].copy()

    def trajectory_minmax_magnitude(self, coeffs, times, derivative, dims=None):
        c, tm = self._d(coeffs), self._d(times)
        K, D, N = c.shape
        d = np.ascontiguousarray(np.arange(D) if dims is None else np.asarray(dims), dtype=np.int32)
        out = np.zeros(6)
        ok = self.lib.refc_trajectory_minmax_magnitude(N, K, D, self._p(c), self._p(tm), derivative, self._p(d), len(d),
                                                       self._p(out))
        return bool(ok), (out[0], out[1], int(out[2])), (out[3], out[4], int(out[5]))
