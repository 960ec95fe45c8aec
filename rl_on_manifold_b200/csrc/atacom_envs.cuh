// Constraint callbacks of the reference's environment families, evaluated in registers.
//
// Each functor returns what the reference obtains from the `fun / J / b` callbacks of its
// ViabilityConstraints (atacom/constraints.py:11-31) at one (q, dq): c(q), J(q), b(q, dq),
// equality rows first.  The reference evaluates J three times per step per constraint and
// crosses into pinocchio ~30 times; here one forward-kinematics pass
// serves all rows.
#pragma once

#include "atacom_core.cuh"

namespace atacom {

// Layout-identical to AtacomParams (include/atacom_b200.h) for T = float.
template <typename T>
struct ParamsT {
  T K_f[4];
  T K_g[16];
  T K_c[20];
  T K_q[8];
  T vel_max[8];
  T acc_max[8];
  T dt;
  T rref_tol;
  int32_t variant;
  int32_t bias_mode;
  int32_t clip_acc;
  int32_t reserved;
  T env[24];
};

enum { VARIANT_ATACOM = 0, VARIANT_EC = 1 };
enum { BIAS_JDOT_QDOT = 0, BIAS_OMEGA_X_V = 1 };

template <typename T, class D>
struct RawConstraints {
  static constexpr int C1 = at_least_1<D::C>::value;
  T c[C1];
  T J[C1][D::n];
  T b[C1];
};

template <typename T> ATACOM_HD void sincos_t(T x, T* s, T* c);
template <> ATACOM_HD void sincos_t<float>(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = ::sinf(x);
  *c = ::cosf(x);
#endif
}
template <> ATACOM_HD void sincos_t<double>(double x, double* s, double* c) {
  *s = ::sin(x);
  *c = ::cos(x);
}

// ----------------------------------------------------------------------------- circle (env A / E)
// circle_atacom.py:47-69: f = q0^2 + q1^2 - 1, g = -q1 - 0.5
struct CircleEnv {
  using D = Dims<2, 1, 1>;
  template <typename T>
  static ATACOM_HD void eval(const ParamsT<T>&, const T* q, const T* dq, RawConstraints<T, D>& R) {
    R.c[0] = q[0] * q[0] + q[1] * q[1] - T(1);
    R.J[0][0] = T(2) * q[0];
    R.J[0][1] = T(2) * q[1];
    R.b[0] = T(2) * dq[0] * dq[0] + T(2) * dq[1] * dq[1];
    R.c[1] = -q[1] - T(0.5);
    R.J[1][0] = T(0);
    R.J[1][1] = T(-1);
    R.b[1] = T(0);
  }
};

// ----------------------------------------------------------------------------- planar 3R (env H)
// atacom_air_hockey.py:78-107.  env[] = l1 l2 l3 base_x base_y qmax[3] half_len half_wid
struct PlanarEnv {
  using D = Dims<3, 0, 6>;
  template <typename T>
  static ATACOM_HD void eval(const ParamsT<T>& P, const T* q, const T* dq, RawConstraints<T, D>& R) {
    T th = T(0), om = T(0);
    T lc[3], ls[3], w[3];
    ATACOM_UNROLL
    for (int i = 0; i < 3; ++i) {
      th += q[i];
      om += dq[i];
      T sn, cs;
      sincos_t<T>(th, &sn, &cs);
      lc[i] = P.env[i] * cs;
      ls[i] = P.env[i] * sn;
      w[i] = om;
    }
    const T x = P.env[3] + lc[0] + lc[1] + lc[2];
    const T y = P.env[4] + ls[0] + ls[1] + ls[2];
    T Jx[3], Jy[3];
    Jx[2] = -ls[2];
    Jy[2] = lc[2];
    Jx[1] = Jx[2] - ls[1];
    Jy[1] = Jy[2] + lc[1];
    Jx[0] = Jx[1] - ls[0];
    Jy[0] = Jy[1] + lc[0];
    T bx, by;
    if (P.bias_mode == BIAS_JDOT_QDOT) {
      bx = -(lc[0] * w[0] * w[0] + lc[1] * w[1] * w[1] + lc[2] * w[2] * w[2]);
      by = -(ls[0] * w[0] * w[0] + ls[1] * w[1] * w[1] + ls[2] * w[2] * w[2]);
    } else {  // pinocchio classical acceleration with zero spatial acceleration: omega x v
      const T vx = Jx[0] * dq[0] + Jx[1] * dq[1] + Jx[2] * dq[2];
      const T vy = Jy[0] * dq[0] + Jy[1] * dq[1] + Jy[2] * dq[2];
      bx = -om * vy;
      by = om * vx;
    }
    const T hl = P.env[8], hw = P.env[9];
    R.c[0] = -x - hl;
    R.c[1] = -y - hw;
    R.c[2] = y - hw;
    R.b[0] = -bx;
    R.b[1] = -by;
    R.b[2] = by;
    ATACOM_UNROLL
    for (int j = 0; j < 3; ++j) {
      R.J[0][j] = -Jx[j];
      R.J[1][j] = -Jy[j];
      R.J[2][j] = Jy[j];
      const T qm = P.env[5 + j];
      R.c[3 + j] = q[j] * q[j] - qm * qm;
      R.b[3 + j] = T(2) * dq[j] * dq[j];
      ATACOM_UNROLL
      for (int i = 0; i < 3; ++i) R.J[3 + j][i] = (i == j) ? T(2) * q[j] : T(0);
    }
  }
};

// ----------------------------------------------------------------------------- iiwa (env 7H)
// iiwa_hit_atacom.py:70-139 on the chain of urdf/iiwa_1.urdf:69-301 (all joints revolute about
// local z).  The fixed joint rotations are signed permutations:
//   rpy (pi/2, 0, pi) and (-pi/2, pi, 0) -> [[-1,0,0],[0,0,1],[0,1,0]]   (joints 2,3,5,7)
//   rpy (pi/2, 0, 0)                     -> [[ 1,0,0],[0,0,-1],[0,1,0]]  (joints 4,6)
// env[] = base_x half_len half_wid height z4_min z7_min qmax[7]
template <typename T> struct Vec3 { T x, y, z; };
template <typename T> ATACOM_HD Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> ATACOM_HD Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> ATACOM_HD Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> ATACOM_HD Vec3<T> operator*(T s, const Vec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }

template <int NJ>
struct IiwaEnv {
  static_assert(NJ == 6 || NJ == 7, "iiwa: 6 (isolated joint 7) or 7 controlled joints");
  using D = Dims<NJ, 1, 5 + NJ>;

  template <typename T>
  static ATACOM_HD void eval(const ParamsT<T>& P, const T* q, const T* dq, RawConstraints<T, D>& R) {
    // joint placement in the parent frame: translation along z (kind 0) or along y (kind 1)
    constexpr double len[7] = {0.1575, 0.2025, 0.2045, 0.2155, 0.1845, 0.2155, 0.081};
    constexpr int along_y[7] = {0, 0, 1, 0, 1, 0, 1};
    constexpr int rot_kind[7] = {0, 1, 1, 2, 1, 2, 1};  // 0: I, 1: A, 2: Bm (see header comment)
    constexpr double tip_len = 0.585;                   // env_base.py:148-149

    // frame axes as columns ex, ey, ez and origin o, all in the robot base frame
    Vec3<T> ex{T(1), T(0), T(0)}, ey{T(0), T(1), T(0)}, ez{T(0), T(0), T(1)}, o{T(0), T(0), T(0)};
    Vec3<T> zax[7], org[7];
    ATACOM_UNROLL
    for (int i = 0; i < 7; ++i) {
      o = o + T(len[i]) * (along_y[i] ? ey : ez);
      // R <- R * Rc: columns of the new frame expressed with the old ones
      Vec3<T> nx, ny, nz;
      if (rot_kind[i] == 0) { nx = ex; ny = ey; nz = ez; }
      else if (rot_kind[i] == 1) { nx = T(-1) * ex; ny = ez; nz = ey; }
      else { nx = ex; ny = ez; nz = T(-1) * ey; }
      // R <- R * Rz(q_i)
      if (i < NJ) {
        T sn, cs;
        sincos_t<T>(q[i], &sn, &cs);
        ex = cs * nx + sn * ny;
        ey = cs * ny - sn * nx;
      } else {
        ex = nx;
        ey = ny;
      }
      ez = nz;
      zax[i] = ez;
      org[i] = o;
    }
    const Vec3<T> tip = o + T(tip_len) * ez;

    // velocity-product recursion (zero joint acceleration): angular velocity w, angular
    // acceleration al, acceleration of each joint origin; sampled at link 4, link 7 and the tip
    Vec3<T> w{T(0), T(0), T(0)}, al = w, ao = w, oprev = w;
    Vec3<T> acc4 = w, acc7 = w, w4 = w, w7 = w;
    ATACOM_UNROLL
    for (int i = 0; i < 7; ++i) {
      const Vec3<T> rr = org[i] - oprev;
      ao = ao + cross(al, rr) + cross(w, cross(w, rr));
      if (i == 3) { acc4 = ao; w4 = w; }
      if (i == 6) { acc7 = ao; w7 = w; }
      if (i < NJ) {
        const Vec3<T> zq = dq[i] * zax[i];
        al = al + cross(w, zq);
        w = w + zq;
      }
      oprev = org[i];
    }
    const Vec3<T> rt = tip - oprev;
    Vec3<T> acct = ao + cross(al, rt) + cross(w, cross(w, rt));

    // world-aligned linear Jacobians: column j = z_j x (p - o_j)
    T Jt[3][NJ], J4z[NJ], J7z[NJ];
    Vec3<T> vt{T(0), T(0), T(0)}, v4 = vt, v7 = vt;
    ATACOM_UNROLL
    for (int j = 0; j < NJ; ++j) {
      const Vec3<T> ct = cross(zax[j], tip - org[j]);
      Jt[0][j] = ct.x; Jt[1][j] = ct.y; Jt[2][j] = ct.z;
      vt = vt + dq[j] * ct;
      Vec3<T> c4{T(0), T(0), T(0)}, c7 = c4;
      if (j < 3) c4 = cross(zax[j], org[3] - org[j]);
      if (j < 6) c7 = cross(zax[j], org[6] - org[j]);
      J4z[j] = c4.z;
      J7z[j] = c7.z;
      v4 = v4 + dq[j] * c4;
      v7 = v7 + dq[j] * c7;
    }
    if (P.bias_mode == BIAS_OMEGA_X_V) {
      // link-4 / link-7 frames are the child frames of joints 4 / 7: their angular velocity
      // includes that joint's own rate
      const Vec3<T> w4f = w4 + dq[3] * zax[3];
      const Vec3<T> w7f = (NJ == 7) ? w7 + dq[NJ == 7 ? 6 : 0] * zax[6] : w7;
      acct = cross(w, vt);
      acc4 = cross(w4f, v4);
      acc7 = cross(w7f, v7);
    }

    const T xw = tip.x + P.env[0];
    const T hl = P.env[1], hw = P.env[2];
    R.c[0] = tip.z - P.env[3];
    R.b[0] = acct.z;
    R.c[1] = -xw - hl;      R.b[1] = -acct.x;
    R.c[2] = -tip.y - hw;   R.b[2] = -acct.y;
    R.c[3] = tip.y - hw;    R.b[3] = acct.y;
    R.c[4] = P.env[4] - org[3].z;  R.b[4] = -acc4.z;
    R.c[5] = P.env[5] - org[6].z;  R.b[5] = -acc7.z;
    ATACOM_UNROLL
    for (int j = 0; j < NJ; ++j) {
      R.J[0][j] = Jt[2][j];
      R.J[1][j] = -Jt[0][j];
      R.J[2][j] = -Jt[1][j];
      R.J[3][j] = Jt[1][j];
      R.J[4][j] = -J4z[j];
      R.J[5][j] = -J7z[j];
      const T qm = P.env[6 + j];
      R.c[6 + j] = q[j] * q[j] - qm * qm;
      R.b[6 + j] = T(2) * dq[j] * dq[j];
      ATACOM_UNROLL
      for (int i = 0; i < NJ; ++i) R.J[6 + j][i] = (i == j) ? T(2) * q[j] : T(0);
    }
  }
};

// ----------------------------------------------------------------------------- shared tail
// constraints.py:33-43 (fun / K_J / b), atacom.py:151-196 (Jc, psi, c), :130-137 (assembly,
// slack integration, acceleration truncation); error_correction_wrapper.py:117-134 for VARIANT_EC.
template <typename T, class D>
ATACOM_HD uint8_t step_from_raw(const ParamsT<T>& P, const RawConstraints<T, D>& R, const T* dq, const T* s,
                                const T* alpha, T* ddq, T* s_out, T* w_dbg) {
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  constexpr int C1 = at_least_1<C>::value;
  T A[C1][n], r[C1];
  const bool ec = P.variant == VARIANT_EC;
  ATACOM_UNROLL
  for (int i = 0; i < C; ++i) {
    const T K = i < F ? P.K_f[i < F ? i : 0] : P.K_g[i >= F ? i - F : 0];
    T Jdq = T(0);
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) {
      Jdq += R.J[i][j] * dq[j];
      A[i][j] = K * R.J[i][j];
    }
    T ct = R.c[i] + K * Jdq;                                    // constraints.py:37
    if (i >= F) ct += T(0.5) * s[i >= F ? i - F : 0] * s[i >= F ? i - F : 0];   // atacom.py:195
    const T psi = Jdq + K * R.b[i];                             // constraints.py:43
    r[i] = (ec ? T(0) : psi) + P.K_c[i] * ct;
  }
  T w_mn[N], w_null[N];
  uint8_t st = project_dense<T, D>(&A[0][0], &A[F < C ? F : 0][0], s, r, alpha, P.rref_tol, !ec, w_mn, w_null);
  if (ec) {
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) w_null[j] = alpha[j];           // error_correction_wrapper.py:127
  }
  bool finite = true;
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) {
    const T wz = w_mn[n + i] + w_null[n + i];
    s_out[i] = s[i] + wz * P.dt;                                // atacom.py:135
    finite = finite && (wz - wz == T(0));
  }
  ATACOM_UNROLL
  for (int j = 0; j < n; ++j) {
    T a = w_mn[j] + w_null[j];
    finite = finite && (a - a == T(0));
    if (P.clip_acc) {                                           // atacom.py:117-121
      const T am = P.acc_max[j];
      T up = -P.K_q[j] * (dq[j] - P.vel_max[j]);
      up = up < am ? up : am;
      up = up > -am ? up : -am;
      T lo = -P.K_q[j] * (dq[j] + P.vel_max[j]);
      lo = lo > -am ? lo : -am;
      lo = lo < am ? lo : am;
      a = a > lo ? a : lo;
      a = a < up ? a : up;
    }
    ddq[j] = a;
  }
  if (w_dbg) {
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      w_dbg[i] = w_mn[i];
      w_dbg[N + i] = w_null[i];
    }
  }
  if (!finite) st |= ST_NONFINITE;
  (void)k;
  return st;
}

// atacom.py:145-149
template <typename T, class D>
ATACOM_HD void slack_from_raw(const ParamsT<T>& P, const RawConstraints<T, D>& R, const T* dq, T* s) {
  constexpr int n = D::n, F = D::F, G = D::G;
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) {
    T Jdq = T(0);
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) Jdq += R.J[F + i][j] * dq[j];
    const T ct = R.c[F + i] + P.K_g[i] * Jdq;
    const T v = T(-2) * ct;
    s[i] = v > T(0) ? num<T>::sqrt(v) : T(0);
  }
}

// ----------------------------------------------------------------------------- point reach (env C)
// collision_avoidance_atacom.py:29-47,72-127.  env[] = radius^2, K, K_c.  No K_g factor in Jc,
// no alpha_max scaling, no acceleration truncation; bp + bq as written (positions, not velocities).
template <int G_>
struct PointReachEnv {
  using D = Dims<2, 0, G_>;
  template <typename T>
  static ATACOM_HD uint8_t step(const ParamsT<T>& P, const T* q, const T* dq, const T* p, const T* dp,
                                const T* s, const T* action, T* w_out, T* s_out, T* w_dbg) {
    constexpr int G = G_, N = 2 + G_;
    const T rad2 = P.env[0], K = P.env[1], Kc = P.env[2];
    T A[G][2], r[G];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      const T px = p[2 * i], py = p[2 * i + 1];
      const T dx = q[0] - px, dy = q[1] - py;
      const T c_o = rad2 - (dx * dx + dy * dy);
      A[i][0] = T(-2) * dx;
      A[i][1] = T(-2) * dy;
      const T dc = T(2) * (dx * dp[2 * i] + dy * dp[2 * i + 1]) + A[i][0] * dq[0] + A[i][1] * dq[1];
      const T bp = (T(-2) * px + T(2) * q[0]) * px + (T(-2) * py + T(2) * q[1]) * py;
      const T bq = (T(-2) * q[0] + T(2) * px) * q[0] + (T(-2) * q[1] + T(2) * py) * q[1];
      const T c = c_o + T(0.5) * s[i] * s[i] + K * dc;
      r[i] = dc + K * (bp + bq) + Kc * c;
    }
    T w_mn[N], w_null[N];
    uint8_t st = project_dense<T, D>((const T*)nullptr, &A[0][0], s, r, action, P.rref_tol, true, w_mn, w_null);
    bool finite = true;
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      const T w = w_mn[i] + w_null[i];
      finite = finite && (w - w == T(0));
      if (i < 2) w_out[i] = w;
      else s_out[i - 2] = s[i - 2] + w * P.dt;
      if (w_dbg) { w_dbg[i] = w_mn[i]; w_dbg[N + i] = w_null[i]; }
    }
    if (!finite) st |= ST_NONFINITE;
    return st;
  }
  template <typename T>
  static ATACOM_HD void slack_init(const ParamsT<T>& P, const T* q, const T* p, T* s) {
    ATACOM_UNROLL
    for (int i = 0; i < G_; ++i) {
      const T dx = q[0] - p[2 * i], dy = q[1] - p[2 * i + 1];
      const T v = T(-2) * (P.env[0] - (dx * dx + dy * dy));
      s[i] = v > T(0) ? num<T>::sqrt(v) : T(0);
    }
  }
};

}  // namespace atacom
