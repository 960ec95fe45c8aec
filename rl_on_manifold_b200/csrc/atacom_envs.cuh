// Constraint callbacks of the reference's environment families, evaluated in registers.
//
// Each functor returns what the reference obtains from the `fun / J / b` callbacks of its
// ViabilityConstraints (atacom/constraints.py:11-31) at one (q, dq): c(q), J(q), b(q, dq),
// equality rows first.  The reference evaluates J three times per step per constraint and
// crosses into pinocchio ~30 times; here one forward-kinematics pass
// serves all rows.
//
// Precision: two scalar types.  HP (double on the device) carries the geometry and everything
// that enters the residual  c(q) + K J dq + s^2/2 : the reference multiplies that residual, a
// cancellation of O(1) terms that is ~0 on the constraint manifold, by K_c = 100..240
// (atacom.py:181), so fp32 rounding of any term shows up at 1e-4 in the output.  T (float on the
// device) carries the Jacobian entries of Jc, the velocity-product terms b(q,dq) and the whole
// projection algebra, where fp32 is backward stable to ~2e-6.
#pragma once

#include "atacom_core.cuh"
#include "atacom_structured.cuh"
#include "atacom_dual.cuh"

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
  int32_t basis_mode;
  double env[24];
};

enum { VARIANT_ATACOM = 0, VARIANT_EC = 1 };
enum { BIAS_JDOT_QDOT = 0, BIAS_OMEGA_X_V = 1 };
enum { BASIS_LAPACK = 0, BASIS_CANONICAL = 1 };

// The environment functors hand their results to a SINK, one value at a time and as early as it is
// known: put_c(i, c_i(q)), put_Jdq(i, (J dq)_i), put_J(i, j, J_ij), put_b(i, b_i(q, dq)), i over
// the C rows (equalities first).  RawConstraints is the sink that just keeps them (JT_ is the type the
// Jacobian entries are kept in); DualSink (below) folds them into the operands of the dual projection on
// the fly, so nothing of size C x n is ever live in registers.
template <typename T, typename HP, class D, typename JT_ = T>
struct RawConstraints {
  using JT = JT_;
  static constexpr int C1 = at_least_1<D::C>::value;
  HP c[C1];        // c(q)
  HP Jdq[C1];      // J(q) dq
  JT J[C1][D::n];  // J(q)
  T b[C1];         // b(q, dq)
  ATACOM_HD void put_c(int i, HP v) { c[i] = v; }
  ATACOM_HD void put_Jdq(int i, HP v) { Jdq[i] = v; }
  ATACOM_HD void put_J(int i, int j, JT v) { J[i][j] = v; }
  ATACOM_HD void put_b(int i, T v) { b[i] = v; }
};

// Sink that keeps only what the constraint statistics need (AtacomEnvWrapper._update_constraint_stats,
// atacom.py:201-205): max over the rows of |c_f(q)| and c_g(q).  Everything else the functor computes
// for it is dead code and disappears at compile time.
template <typename T, typename HP, class D>
struct StatsSink {
  using JT = HP;
  HP cmax;
  ATACOM_HD StatsSink() : cmax(HP(-1e300)) {}
  ATACOM_HD void put_c(int i, HP v) {
    const HP a = (i < D::F) ? (v < HP(0) ? -v : v) : v;
    cmax = a > cmax ? a : cmax;
  }
  ATACOM_HD void put_Jdq(int, HP) {}
  ATACOM_HD void put_J(int, int, HP) {}
  ATACOM_HD void put_b(int, T) {}
};

// sin and cos of a joint angle in double precision, ~1e-16 absolute: Cody-Waite reduction by
// multiples of pi/2 and Taylor polynomials on [-pi/4, pi/4] (|x| < ~1e5; joint angles are < 3.1).
// About 30 FP64 operations, a fraction of the library sincos(double).
ATACOM_HD void sincos_hp(double x, double* sn, double* cs) {
  const double kd = ::rint(x * 0.63661977236758134308);  // 2/pi
  double r = ::fma(-kd, 1.57079632679489655800e+00, x);  // pi/2 high
  r = ::fma(-kd, 6.12323399573676603587e-17, r);         // pi/2 low
  const double r2 = r * r;
  double ps = -7.64716373181981647590e-13;  // -1/15!
  ps = ::fma(ps, r2, 1.60590438368216145994e-10);
  ps = ::fma(ps, r2, -2.50521083854417187751e-08);
  ps = ::fma(ps, r2, 2.75573192239858906526e-06);
  ps = ::fma(ps, r2, -1.98412698412698412698e-04);
  ps = ::fma(ps, r2, 8.33333333333333333333e-03);
  ps = ::fma(ps, r2, -1.66666666666666666667e-01);
  const double sr = ::fma(ps * r2, r, r);
  double pc = 4.77947733238738529744e-14;  // 1/16!
  pc = ::fma(pc, r2, -1.14707455977297247139e-11);
  pc = ::fma(pc, r2, 2.08767569878680989792e-09);
  pc = ::fma(pc, r2, -2.75573192239858906526e-07);
  pc = ::fma(pc, r2, 2.48015873015873015873e-05);
  pc = ::fma(pc, r2, -1.38888888888888888889e-03);
  pc = ::fma(pc, r2, 4.16666666666666666667e-02);
  pc = ::fma(pc, r2, -0.5);
  const double cr = ::fma(pc, r2, 1.0);
  const int k = static_cast<int>(kd) & 3;
  const double s0 = (k & 1) ? cr : sr;
  const double c0 = (k & 1) ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}

// ----------------------------------------------------------------------------- circle (env A / E)
// circle_atacom.py:47-69: f = q0^2 + q1^2 - 1, g = -q1 - 0.5
struct CircleEnv {
  using D = Dims<2, 1, 1>;
  static constexpr int NDIAG = 0;   // trailing inequality rows with a diagonal Jacobian
  template <typename T, typename HP, class Out>
  static ATACOM_HD void eval(const ParamsT<T>&, const T* q, const T* dq, Out& R) {
    using JT = typename Out::JT;
    const HP q0 = q[0], q1 = q[1], d0 = dq[0], d1 = dq[1];
    R.put_c(0, q0 * q0 + q1 * q1 - HP(1));
    R.put_Jdq(0, HP(2) * (q0 * d0 + q1 * d1));
    R.put_J(0, 0, JT(2) * JT(q[0]));
    R.put_J(0, 1, JT(2) * JT(q[1]));
    R.put_b(0, T(2) * dq[0] * dq[0] + T(2) * dq[1] * dq[1]);
    R.put_c(1, -q1 - HP(0.5));
    R.put_Jdq(1, -d1);
    R.put_J(1, 0, JT(0));
    R.put_J(1, 1, JT(-1));
    R.put_b(1, T(0));
  }
};

// ----------------------------------------------------------------------------- planar 3R (env H)
// atacom_air_hockey.py:78-107.  env[] = l1 l2 l3 base_x base_y qmax[3] half_len half_wid
struct PlanarEnv {
  using D = Dims<3, 0, 6>;
  static constexpr int NDIAG = 3;   // joint-limit rows q_j^2 - qmax_j^2
  template <typename T, typename HP, class Out>
  static ATACOM_HD void eval(const ParamsT<T>& P, const T* q, const T* dq, Out& R) {
    using JT = typename Out::JT;
    HP th = HP(0), om = HP(0);
    HP lc[3], ls[3], w[3];
    ATACOM_UNROLL
    for (int i = 0; i < 3; ++i) {
      th += HP(q[i]);
      om += HP(dq[i]);
      double sn, cs;
      sincos_hp(static_cast<double>(th), &sn, &cs);
      lc[i] = HP(P.env[i] * cs);
      ls[i] = HP(P.env[i] * sn);
      w[i] = om;
    }
    const HP x = HP(P.env[3]) + lc[0] + lc[1] + lc[2];
    const HP y = HP(P.env[4]) + ls[0] + ls[1] + ls[2];
    HP Jx[3], Jy[3];
    Jx[2] = -ls[2];
    Jy[2] = lc[2];
    Jx[1] = Jx[2] - ls[1];
    Jy[1] = Jy[2] + lc[1];
    Jx[0] = Jx[1] - ls[0];
    Jy[0] = Jy[1] + lc[0];
    const HP vx = Jx[0] * HP(dq[0]) + Jx[1] * HP(dq[1]) + Jx[2] * HP(dq[2]);
    const HP vy = Jy[0] * HP(dq[0]) + Jy[1] * HP(dq[1]) + Jy[2] * HP(dq[2]);
    T bx, by;
    if (P.bias_mode == BIAS_JDOT_QDOT) {
      bx = -T(lc[0] * w[0] * w[0] + lc[1] * w[1] * w[1] + lc[2] * w[2] * w[2]);
      by = -T(ls[0] * w[0] * w[0] + ls[1] * w[1] * w[1] + ls[2] * w[2] * w[2]);
    } else {  // pinocchio classical acceleration with zero spatial acceleration: omega x v
      bx = -T(om * vy);
      by = T(om * vx);
    }
    const HP hl = HP(P.env[8]), hw = HP(P.env[9]);
    R.put_c(0, -x - hl);
    R.put_c(1, -y - hw);
    R.put_c(2, y - hw);
    R.put_Jdq(0, -vx);
    R.put_Jdq(1, -vy);
    R.put_Jdq(2, vy);
    R.put_b(0, -bx);
    R.put_b(1, -by);
    R.put_b(2, by);
    ATACOM_UNROLL
    for (int j = 0; j < 3; ++j) {
      R.put_J(0, j, -JT(Jx[j]));
      R.put_J(1, j, -JT(Jy[j]));
      R.put_J(2, j, JT(Jy[j]));
      const HP qm = HP(P.env[5 + j]), qj = q[j];
      R.put_c(3 + j, qj * qj - qm * qm);
      R.put_Jdq(3 + j, HP(2) * qj * HP(dq[j]));
      R.put_b(3 + j, T(2) * dq[j] * dq[j]);
      ATACOM_UNROLL
      for (int i = 0; i < 3; ++i) R.put_J(3 + j, i, (i == j) ? JT(2) * JT(q[j]) : JT(0));
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
template <typename T, typename U> ATACOM_HD Vec3<T> vcast(const Vec3<U>& a) { return {cvt<T>(a.x), cvt<T>(a.y), cvt<T>(a.z)}; }
template <typename T> ATACOM_HD T dot(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

template <int NJ>
struct IiwaEnv {
  static_assert(NJ == 6 || NJ == 7, "iiwa: 6 (isolated joint 7) or 7 controlled joints");
  using D = Dims<NJ, 1, 5 + NJ>;
  static constexpr int NDIAG = NJ;  // joint-limit rows q_j^2 - qmax_j^2

  template <typename T, typename HP, class Out>
  static ATACOM_HD void eval(const ParamsT<T>& P, const T* q, const T* dq, Out& R) {
    using JT = typename Out::JT;
    // joint placement in the parent frame: translation along z or along y
    constexpr double len[7] = {0.1575, 0.2025, 0.2045, 0.2155, 0.1845, 0.2155, 0.081};
    constexpr int along_y[7] = {0, 0, 1, 0, 1, 0, 1};
    constexpr int rot_kind[7] = {0, 1, 1, 2, 1, 2, 1};  // 0: I, 1: A, 2: Bm (see header comment)
    constexpr double tip_len = 0.585;                   // env_base.py:148-149

    // ---- one forward pass over the chain.  Geometry in HP: frame axes ex, ey, ez and origin o in the
    // robot base frame.  Riding along in T (velocity-product terms are not amplified by K_c): the
    // zero-joint-acceleration recursion for the acceleration of the link-4 / link-7 origins and the tip.
    Vec3<HP> ex{HP(1), HP(0), HP(0)}, ey{HP(0), HP(1), HP(0)}, ez{HP(0), HP(0), HP(1)}, o{HP(0), HP(0), HP(0)};
    Vec3<HP> zax[7], org[7];
    Vec3<T> w{T(0), T(0), T(0)}, al = w, ao = w, oprev = w;
    Vec3<T> acc4 = w, acc7 = w, w4 = w, w7 = w;
    ATACOM_UNROLL
    for (int i = 0; i < 7; ++i) {
      o = o + HP(len[i]) * (along_y[i] ? ey : ez);
      Vec3<HP> nx, ny, nz;  // R <- R * Rc
      if (rot_kind[i] == 0) { nx = ex; ny = ey; nz = ez; }
      else if (rot_kind[i] == 1) { nx = HP(-1) * ex; ny = ez; nz = ey; }
      else { nx = ex; ny = ez; nz = HP(-1) * ey; }
      if (i < NJ) {         // R <- R * Rz(q_i)
        double sn, cs;
        sincos_hp(cvt<double>(q[i < NJ ? i : 0]), &sn, &cs);
        ex = HP(cs) * nx + HP(sn) * ny;
        ey = HP(cs) * ny - HP(sn) * nx;
      } else {
        ex = nx;
        ey = ny;
      }
      ez = nz;
      zax[i] = ez;
      org[i] = o;
      const Vec3<T> oi = vcast<T>(o);
      const Vec3<T> rr = oi - oprev;
      ao = ao + cross(al, rr) + cross(w, cross(w, rr));
      if (i == 3) { acc4 = ao; w4 = w; }
      if (i == 6) { acc7 = ao; w7 = w; }
      if (i < NJ) {
        const Vec3<T> zq = dq[i < NJ ? i : 0] * vcast<T>(ez);
        al = al + cross(w, zq);
        w = w + zq;
      }
      oprev = oi;
    }
    const Vec3<HP> tip = o + HP(tip_len) * ez;
    const Vec3<T> rt = vcast<T>(tip) - oprev;
    Vec3<T> acct = ao + cross(al, rt) + cross(w, cross(w, rt));

    // ---- world-aligned linear Jacobians, column j = z_j x (p - o_j), handed on as they are formed; J dq
    // accumulated in HP.  Only the z rows of the link-4 / link-7 Jacobians enter the constraints.
    HP dqh[NJ];
    ATACOM_UNROLL
    for (int j = 0; j < NJ; ++j) dqh[j] = cvt<HP>(dq[j]);
    Vec3<HP> vt{HP(0), HP(0), HP(0)};
    HP v4z = HP(0), v7z = HP(0);
    ATACOM_UNROLL
    for (int j = 0; j < NJ; ++j) {
      // the tip lies on the axis of joint 7: its Jacobian column is exactly zero (env_base.py:147-151; "isolated
      // joint 7"), not the rounding noise of the cross product — the sign of such noise would decide the sign of a
      // Householder reflector of the LAPACK basis
      const Vec3<HP> ct = (j == 6) ? Vec3<HP>{HP(0), HP(0), HP(0)} : cross(zax[j], tip - org[j]);
      vt = vt + dqh[j] * ct;
      HP c4z = HP(0), c7z = HP(0);
      if (j < 3) c4z = zax[j].x * (org[3].y - org[j].y) - zax[j].y * (org[3].x - org[j].x);
      if (j < 6) c7z = zax[j].x * (org[6].y - org[j].y) - zax[j].y * (org[6].x - org[j].x);
      v4z += dqh[j] * c4z;
      v7z += dqh[j] * c7z;
      R.put_J(0, j, cvt<JT>(ct.z));
      R.put_J(1, j, cvt<JT>(-ct.x));
      R.put_J(2, j, cvt<JT>(-ct.y));
      R.put_J(3, j, cvt<JT>(ct.y));
      R.put_J(4, j, cvt<JT>(-c4z));
      R.put_J(5, j, cvt<JT>(-c7z));
      const HP qm = HP(P.env[6 + j]), qj = cvt<HP>(q[j]);
      R.put_c(6 + j, qj * qj - qm * qm);
      R.put_Jdq(6 + j, HP(2) * qj * dqh[j]);
      R.put_b(6 + j, T(2) * dq[j] * dq[j]);
      ATACOM_UNROLL
      for (int i = 0; i < NJ; ++i) R.put_J(6 + j, i, (i == j) ? cvt<JT>(HP(2) * qj) : JT(0));
    }

    if (P.bias_mode == BIAS_OMEGA_X_V) {
      // pinocchio classical acceleration with zero spatial acceleration: omega x v.  The link-4 / link-7
      // frames are the child frames of joints 4 / 7: their angular velocity includes that joint's own rate
      Vec3<HP> v4{HP(0), HP(0), HP(0)}, v7 = v4;
      ATACOM_UNROLL
      for (int j = 0; j < NJ; ++j) {
        if (j < 3) v4 = v4 + dqh[j] * cross(zax[j], org[3] - org[j]);
        if (j < 6) v7 = v7 + dqh[j] * cross(zax[j], org[6] - org[j]);
      }
      const Vec3<T> w4f = w4 + dq[3] * vcast<T>(zax[3]);
      const Vec3<T> w7f = (NJ == 7) ? w7 + dq[NJ == 7 ? 6 : 0] * vcast<T>(zax[6]) : w7;
      acct = cross(w, vcast<T>(vt));
      acc4 = cross(w4f, vcast<T>(v4));
      acc7 = cross(w7f, vcast<T>(v7));
    }

    const HP xw = tip.x + HP(P.env[0]);
    const HP hl = HP(P.env[1]), hw = HP(P.env[2]);
    R.put_c(0, tip.z - HP(P.env[3]));     R.put_Jdq(0, vt.z);    R.put_b(0, acct.z);
    R.put_c(1, -xw - hl);                 R.put_Jdq(1, -vt.x);   R.put_b(1, -acct.x);
    R.put_c(2, -tip.y - hw);              R.put_Jdq(2, -vt.y);   R.put_b(2, -acct.y);
    R.put_c(3, tip.y - hw);               R.put_Jdq(3, vt.y);    R.put_b(3, acct.y);
    R.put_c(4, HP(P.env[4]) - org[3].z);  R.put_Jdq(4, -v4z);    R.put_b(4, -acc4.z);
    R.put_c(5, HP(P.env[5]) - org[6].z);  R.put_Jdq(5, -v7z);    R.put_b(5, -acc7.z);
  }
};

// ----------------------------------------------------------------------------- projection dispatch
// Dense Householder path, kept out of line: it is the rarely taken fallback of the structured path and
// must not set the register budget of the kernels that inline the fast path.
template <typename T, class D>
ATACOM_NOINLINE uint8_t project_dense_outlined(const T* Af, const T* Ag, const T* s, const T* r, const T* alpha,
                                               T tol, bool want_null, T* w_mn, T* w_null) {
  return project_dense<T, D>(Af, Ag, s, r, alpha, tol, want_null, w_mn, w_null);
}

constexpr int STRUCTURED_TMAX = 2;      // stiff inequality rows handled by the fast path
#define ATACOM_STIFF_TAU 400            // row i is stiff when |a_i|^2 > tau s_i^2

// A: C x n row-major (equality rows first, all K-scaled), the last NDIAG inequality rows diagonal.
template <typename T, class D, int NDIAG, bool SYNC = false>
ATACOM_HD uint8_t project(const T* A, const T* s, const T* r, const T* alpha, T tol, bool want_null, T* w_mn,
                          T* w_null) {
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, GD = G - NDIAG;
  T dg[at_least_1<NDIAG>::value];
  ATACOM_UNROLL
  for (int j = 0; j < NDIAG; ++j) dg[j] = A[(F + GD + j) * n + j];
  const T* Ag = A + (F < C ? F : 0) * n;
  uint8_t st = Structured<T, D, NDIAG, STRUCTURED_TMAX, SYNC>::project(A, Ag, dg, s, r, alpha, tol, T(ATACOM_STIFF_TAU),
                                                                 want_null, w_mn, w_null);
  if (st & ST_DENSE_PATH) {
    // private copies: only these (not the register-resident operands of the fast path) are addressable
    constexpr int N = D::N, k = D::k;
    T A2[at_least_1<C * n>::value], s2[at_least_1<G>::value], r2[at_least_1<C>::value];
    T al2[at_least_1<k>::value], wm2[N], wn2[N];
    ATACOM_UNROLL
    for (int i = 0; i < C * n; ++i) A2[i] = A[i];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) s2[i] = s[i];
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) r2[i] = r[i];
    ATACOM_UNROLL
    for (int i = 0; i < k; ++i) al2[i] = alpha[i];
    st = ST_DENSE_PATH | project_dense_outlined<T, D>(A2, A2 + (F < C ? F : 0) * n, s2, r2, al2, tol, want_null,
                                                      wm2, wn2);
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      w_mn[i] = wm2[i];
      w_null[i] = wn2[i];
    }
  }
  return st;
}

// ----------------------------------------------------------------------------- shared tail
// constraints.py:33-43 (fun / K_J / b), atacom.py:151-196 (Jc, psi, c), :130-137 (assembly,
// slack integration, acceleration truncation); error_correction_wrapper.py:117-134 for VARIANT_EC.
template <typename T, typename HP, class D, int NDIAG, bool SYNC = false>
ATACOM_HD uint8_t step_from_raw(const ParamsT<T>& P, const RawConstraints<T, HP, D>& R, const T* dq, const T* s,
                                const T* alpha, T* ddq, T* s_out, T* w_dbg) {
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  constexpr int C1 = at_least_1<C>::value;
  T A[C1][n], r[C1];
  const bool ec = P.variant == VARIANT_EC;
  ATACOM_UNROLL
  for (int i = 0; i < C; ++i) {
    const T K = i < F ? P.K_f[i < F ? i : 0] : P.K_g[i >= F ? i - F : 0];
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) A[i][j] = K * R.J[i][j];        // constraints.py:39-40
    HP ct = R.c[i] + HP(K) * R.Jdq[i];                          // constraints.py:37
    if (i >= F) {
      const HP si = s[i >= F ? i - F : 0];
      ct += HP(0.5) * si * si;                                  // atacom.py:195
    }
    const HP psi = R.Jdq[i] + HP(K) * HP(R.b[i]);               // constraints.py:43
    r[i] = T((ec ? HP(0) : psi) + HP(P.K_c[i]) * ct);           // atacom.py:130,132,181
  }
  T w_mn[N], w_null[N];
  phase_sync<SYNC>();
  uint8_t st = project<T, D, NDIAG, SYNC>(&A[0][0], s, r, alpha, P.rref_tol, !ec, w_mn, w_null);
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

#ifndef ATACOM_COOP_MAX_LANES      // more asking lanes than this in one warp: each runs the serial general routine
#define ATACOM_COOP_MAX_LANES 6
#endif

// Gains the dual path multiplies in HP, converted and combined once on the host side of the launch:
//   r_i = psi_i + K_c,i c_i  with  psi_i = (J dq)_i + K_i b_i,  c_i = c_i(q) + K_i (J dq)_i [+ s_i^2 / 2]
//       = K_c,i c_i(q) + wJ_i (J dq)_i + wb_i b_i [+ K_c,i s_i^2 / 2]      (constraints.py:33-43, atacom.py:181,195)
// variant E drops psi (error_correction_wrapper.py:117-134).
template <typename HP>
struct DualConsts {
  HP K[20];     // K_f then K_g (constraints.py:26-29)
  HP K_c[20];   // atacom.py:42-48
  HP wJ[20];    // 1 + K_c K   (variant E: K_c K)
  HP wb[20];    // K           (variant E: 0)
  HP dt, tol;
};

template <typename T, typename HP>
inline DualConsts<HP> make_dual_consts(const ParamsT<T>& P, int F, int G) {
  DualConsts<HP> c;
  const bool ec = P.variant == VARIANT_EC;
  for (int i = 0; i < 20; ++i) {
    c.K[i] = i < F ? HP(P.K_f[i]) : (i - F < G && i - F < 16 ? HP(P.K_g[i - F]) : HP(0));
    c.K_c[i] = HP(P.K_c[i]);
    c.wJ[i] = (ec ? HP(0) : HP(1)) + c.K_c[i] * c.K[i];
    c.wb[i] = ec ? HP(0) : c.K[i];
  }
  c.dt = HP(P.dt);
  c.tol = HP(P.rref_tol);
  return c;
}

// ParamsT<float> (the C ABI struct) widened to another scalar type: the roll-out kernels run in double
// throughout, so that a T-step roll-out tracks the float64 reference instead of accumulating fp32 rounding.
template <typename To, typename From>
inline ParamsT<To> widen_params(const ParamsT<From>& P) {
  ParamsT<To> Q;
  for (int i = 0; i < 4; ++i) Q.K_f[i] = To(P.K_f[i]);
  for (int i = 0; i < 16; ++i) Q.K_g[i] = To(P.K_g[i]);
  for (int i = 0; i < 20; ++i) Q.K_c[i] = To(P.K_c[i]);
  for (int i = 0; i < 8; ++i) {
    Q.K_q[i] = To(P.K_q[i]);
    Q.vel_max[i] = To(P.vel_max[i]);
    Q.acc_max[i] = To(P.acc_max[i]);
  }
  Q.dt = To(P.dt);
  Q.rref_tol = To(P.rref_tol);
  Q.variant = P.variant;
  Q.bias_mode = P.bias_mode;
  Q.clip_acc = P.clip_acc;
  Q.basis_mode = P.basis_mode;
  for (int i = 0; i < 24; ++i) Q.env[i] = P.env[i];
  return Q;
}

// Sink of the dual path: K-scaled dense Jacobian rows go straight to the scratch Y, the diagonal rows
// to dg, and the three scalar streams are folded into r as they arrive.
template <typename T, typename HP, class D, int NDIAG, class YS>
struct DualSink {
  using JT = HP;
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, m = D::F + D::G - NDIAG;
  const DualConsts<HP>& Kd;
  YS& Y;
  HP r[at_least_1<C>::value], dg[at_least_1<NDIAG>::value];
  ATACOM_HD DualSink(const DualConsts<HP>& Kd_, YS& Y_) : Kd(Kd_), Y(Y_) {
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) r[i] = HP(0);
  }
  // the slack term of c (atacom.py:195), added once the slacks are at hand
  ATACOM_HD void add_slack_terms(const HP* sh) {
    ATACOM_UNROLL
    for (int i = F; i < C; ++i) r[i] += HP(0.5) * Kd.K_c[i] * sh[i - F] * sh[i - F];
  }
  ATACOM_HD void put_c(int i, HP v) { r[i] += Kd.K_c[i] * v; }
  ATACOM_HD void put_Jdq(int i, HP v) { r[i] += Kd.wJ[i] * v; }
  ATACOM_HD void put_b(int i, T v) { r[i] += Kd.wb[i] * cvt<HP>(v); }
  ATACOM_HD void put_J(int i, int j, HP v) {
    if (i < m) Y.set(i * n + j, Kd.K[i] * v);                 // constraints.py:39-40
    else if (j == i - m) dg[j] = Kd.K[i] * v;
  }
};

// Assembly shared by the dual and the LAPACK-basis paths: ddq_ds = w_mn + w_null (variant E: alpha in place of the
// null part), slack integration in HP (atacom.py:135), acceleration truncation (atacom.py:117-121).
template <class D, typename T, typename HP>
ATACOM_HD uint8_t finish_step(const ParamsT<T>& P, const DualConsts<HP>& Kd, const HP* w_mn, HP* w_null, const HP* sh,
                              const T* alpha, const T* dq, T* ddq, HP* s_new, T* w_dbg, uint8_t st) {
  constexpr int n = D::n, G = D::G, N = D::N;
  const bool ec = P.variant == VARIANT_EC;
  if (ec) {
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) w_null[j] = cvt<HP>(alpha[j]);    // error_correction_wrapper.py:127
  }
  bool finite = true;
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) {
    const HP wz = cvt<HP>(w_mn[n + i]) + w_null[n + i];
    s_new[i] = sh[i] + wz * Kd.dt;                                // atacom.py:135
    finite = finite && (s_new[i] - s_new[i] == HP(0));
  }
  ATACOM_UNROLL
  for (int j = 0; j < n; ++j) {
    T a = cvt<T>(cvt<HP>(w_mn[j]) + w_null[j]);
    finite = finite && (a - a == T(0));
    if (P.clip_acc) {                                             // atacom.py:117-121
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
      w_dbg[i] = cvt<T>(w_mn[i]);
      w_dbg[N + i] = cvt<T>(w_null[i]);
    }
  }
  if (!finite) st |= ST_NONFINITE;
  return st;
}

// The whole step around the dual projection (atacom_dual.cuh): everything between the fp32 inputs and
// the fp32 outputs is carried in HP.
// `fetch(s, alpha)` is called AFTER the constraint functor has run and must fill the slack row (G values) and
// the action row (n values, zero-padded): the kinematics need neither, so a kernel can have them copied in
// the background (bulk copy into shared memory) while the kinematics run on q and dq.
// Projection + assembly + slack integration + acceleration truncation on operands that are already in place:
// Y holds the K-scaled dense Jacobian rows, dg the diagonal rows, r the right-hand side (slack terms included),
// sh the current slacks.  s_new = s + dt w_z in HP (atacom.py:135); ddq clipped (atacom.py:117-121).
// Where the minimum-norm part of an environment that is left to the LAPACK-basis routine goes: `row()` returns the N
// HP values of this environment or null.  A functor, so that a kernel can work the address out only when it is needed
// (nothing stays live in registers across the projection).
template <typename HP>
struct WmnRow {
  HP* p;
  ATACOM_HD HP* row() const { return p; }
};

template <class Env, typename T, typename HP, bool BAND_ONLY = false, class YS, class LS, class WO = WmnRow<HP>>
ATACOM_HD uint8_t dual_tail(const ParamsT<T>& P, const DualConsts<HP>& Kd, YS& Y, LS& Ls, const HP* dg, const HP* r,
                            const HP* sh, const T* alpha, const T* dq, T* ddq, HP* s_new, T* w_dbg,
                            HP* coop = nullptr, int coop_slots = 1, int coop_stride = 0,
                            const WO& wmn_out = WO{nullptr}) {
  // coop: the first of the block's coop_slots scratch slots (coop_stride doubles apart) of the warp-cooperative
  // general routine, on the device when Y and L live in shared memory
  using D = typename Env::D;
  constexpr int NDIAG = Env::NDIAG;
  constexpr int n = D::n, G = D::G, N = D::N, k = D::k;
  const bool ec = P.variant == VARIANT_EC;
  HP ah[at_least_1<k>::value];
  ATACOM_UNROLL
  for (int l = 0; l < k; ++l) ah[l] = cvt<HP>(alpha[l]);
  HP w_mn[N], w_null[N];      // (w_mn in HP: kept in fp32 it cost two conversions per entry and bought no registers)
#if defined(__CUDA_ARCH__)
  // Three or more constraints active at once (the general null-space routine): with Y and L in shared memory the
  // thread only asks for it, and the warp then decides.  Few askers (the realistic case: a handful of environments
  // of a 65 536 batch): all 32 lanes run the routine together for each asking lane in turn, a few microseconds
  // each — run serially by its thread it takes ~70 us, and a single such environment would set the duration of the
  // whole launch.  Many askers: each runs the serial routine, all at once.  The warp is converged here (threads
  // past the end of the batch recompute the last environment); a scratch slot is shared by the warps that map to
  // it under the lock stored behind it.
  constexpr bool COOP = is_shared_store<YS>::value && N >= 2 * n && !BAND_ONLY;
  const bool band_defer = P.basis_mode == BASIS_LAPACK && k > 1;     // (k = 1: the null basis is unique up to its sign)
  uint8_t st = Dual<HP, D, NDIAG>::template project<COOP, BAND_ONLY>(Y, Ls, dg, sh, r, ah, Kd.tol, !ec, w_mn, w_null,
                                                                     band_defer);
  if constexpr (COOP) {
    using DU = Dual<HP, D, NDIAG>;
    unsigned pend = __ballot_sync(0xffffffffu, (st & ST_DENSE_PATH) != 0);
    if (__builtin_expect(pend != 0u, 0) && (coop == nullptr || __popc(pend) > ATACOM_COOP_MAX_LANES)) {
      if (st & ST_DENSE_PATH) st = (st & ST_RANK_DEFICIENT) | DU::general_from_deferred(Y, Ls, sh, ah, Kd.tol, w_null);
    } else if (__builtin_expect(pend != 0u, 0)) {
      const int lane = static_cast<int>(threadIdx.x & 31u);
      HP* slot = coop + static_cast<int>((threadIdx.x >> 5) % static_cast<unsigned>(coop_slots)) * coop_stride;
      volatile HP* sc = slot;
      unsigned* lock = reinterpret_cast<unsigned*>(slot + DU::COOP_DOUBLES);
      if (lane == 0) {
        while (atomicCAS(lock, 0u, 1u) != 0u) __nanosleep(200);   // holders are warps of this block and always release
      }
      __syncwarp();
      __threadfence_block();
      while (pend != 0u) {
        const int owner = __ffs(pend) - 1;
        pend &= pend - 1u;
        if (lane == owner) {
          ATACOM_UNROLL
          for (int j = 0; j < 2 * n; ++j) sc[DU::COOP_SG0 + j] = w_null[j < N ? j : 0];
          ATACOM_UNROLL
          for (int i = 0; i < G; ++i) sc[DU::COOP_S0 + i] = sh[i];
          ATACOM_UNROLL
          for (int l = 0; l < k; ++l) sc[DU::COOP_A0 + l] = ah[l];
        }
        __syncwarp();
        const uint8_t st2 = DU::null_part_general_warp(shared_base(Y) - lane + owner, shared_base(Ls) - lane + owner,
                                                       Kd.tol, sc);
        if (lane == owner) {
          st = (st & ST_RANK_DEFICIENT) | st2;
          ATACOM_UNROLL
          for (int i = 0; i < N; ++i) w_null[i] = sc[DU::COOP_WN0 + i];
        }
        __syncwarp();
      }
      __threadfence_block();
      if (lane == 0) atomicExch(lock, 0u);
    }
  }
#else
  const bool band_defer = P.basis_mode == BASIS_LAPACK && k > 1;
  uint8_t st = Dual<HP, D, NDIAG>::project(Y, Ls, dg, sh, r, ah, Kd.tol, !ec, w_mn, w_null, band_defer);
#endif
  if (st & ST_LAPACK_PATH) {
    // left to the LAPACK-basis routine: the minimum-norm part does not depend on the basis and is handed on
    HP* out = wmn_out.row();
    if (out != nullptr) {
      ATACOM_UNROLL
      for (int i = 0; i < N; ++i) out[i] = w_mn[i];
    }
    return st;
  }
  if (st & ST_DENSE_PATH) return st;     // (only after the warp-wide ballot above)
  return finish_step<D, T, HP>(P, Kd, w_mn, w_null, sh, alpha, dq, ddq, s_new, w_dbg, st);
}

template <class Env, typename T, typename HP, bool BAND_ONLY = false, class YS, class LS, class Fetch,
          class WO = WmnRow<HP>>
ATACOM_HD uint8_t step_dual_lazy(const ParamsT<T>& P, const DualConsts<HP>& Kd, YS& Y, LS& Ls, const T* q,
                                 const T* dq, Fetch&& fetch, T* ddq, T* s_out, T* w_dbg, HP* coop = nullptr,
                                 int coop_slots = 1, int coop_stride = 0, const WO& wmn_out = WO{nullptr}) {
  using D = typename Env::D;
  constexpr int NDIAG = Env::NDIAG;
  constexpr int n = D::n, G = D::G;
  DualSink<T, HP, D, NDIAG, YS> sink(Kd, Y);
  Env::template eval<T, HP>(P, q, dq, sink);
  T s[at_least_1<G>::value], alpha[n];
  fetch(s, alpha);
  HP sh[at_least_1<G>::value], sn[at_least_1<G>::value];
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) sh[i] = cvt<HP>(s[i]);
  sink.add_slack_terms(sh);
  const uint8_t st = dual_tail<Env, T, HP, BAND_ONLY>(P, Kd, Y, Ls, sink.dg, sink.r, sh, alpha, dq, ddq, sn, w_dbg, coop,
                                                      coop_slots, coop_stride, wmn_out);
  if (st & (ST_DENSE_PATH | ST_LAPACK_PATH)) return st;
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) s_out[i] = cvt<T>(sn[i]);
  return st;
}

template <class Env, typename T, typename HP, class YS, class LS>
ATACOM_HD uint8_t step_dual(const ParamsT<T>& P, const DualConsts<HP>& Kd, YS& Y, LS& Ls, const T* q, const T* dq,
                            const T* s, const T* alpha, T* ddq, T* s_out, T* w_dbg, HP* wmn_out = nullptr) {
  using D = typename Env::D;
  auto fetch = [&](T* s_row, T* a_row) {
    ATACOM_UNROLL
    for (int i = 0; i < D::G; ++i) s_row[i] = s[i];
    ATACOM_UNROLL
    for (int j = 0; j < D::n; ++j) a_row[j] = alpha[j];
  };
  return step_dual_lazy<Env, T, HP>(P, Kd, Y, Ls, q, dq, fetch, ddq, s_out, w_dbg, nullptr, 1, 0, WmnRow<HP>{wmn_out});
}

// The whole step on the LAPACK-basis routine (atacom_lapack.cuh): what the step kernels run for the environments
// the dual path flagged (Dual::project, band_defer), and the generic kernel for every environment.  S: a store of
// Lapack<HP, D>::SIZE entries (a column of a shared-memory array on the device).
// LPE lanes of a warp may share one environment (`Grp`, see Lapack::project): every lane evaluates the constraints,
// the rows of Jc are written round robin, the outputs are valid in lane 0 of the group.
template <class D, typename T, typename HP, int LPE = 1, class ST, class GRP = SoloGroup>
ATACOM_HD uint8_t step_lapack_from_raw(const ParamsT<T>& P, const DualConsts<HP>& Kd, ST& S,
                                       const RawConstraints<T, HP, D, HP>& R, const T* dq, const T* s, const T* alpha,
                                       T* ddq, T* s_out, T* w_dbg, const GRP& Grp = GRP()) {
  using LP = Lapack<HP, D>;
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  const bool ec = P.variant == VARIANT_EC;
  HP r[at_least_1<C>::value], sh[at_least_1<G>::value], sn[at_least_1<G>::value];
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) sh[i] = cvt<HP>(s[i]);
  ATACOM_UNROLL
  for (int i = 0; i < C; ++i) {
    if (LPE == 1 || Grp.write_lane() < 0 || (i % LPE) == Grp.write_lane()) {
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) S.set(LP::a(i, j), Kd.K[i] * R.J[i][j]);               // constraints.py:39-40
      ATACOM_UNROLL
      for (int j = n; j < N; ++j) S.set(LP::a(i, j), (i >= F && j - n == i - F) ? sh[i >= F ? i - F : 0] : HP(0));   // atacom.py:151-165
    }
    HP ri = Kd.K_c[i] * R.c[i] + Kd.wJ[i] * R.Jdq[i] + Kd.wb[i] * cvt<HP>(R.b[i]);       // see DualConsts
    if (i >= F) ri += HP(0.5) * Kd.K_c[i] * sh[i >= F ? i - F : 0] * sh[i >= F ? i - F : 0];
    r[i] = ri;
  }
  HP ah[at_least_1<k>::value], w_mn[N], w_null[N];
  ATACOM_UNROLL
  for (int l = 0; l < k; ++l) ah[l] = cvt<HP>(alpha[l]);
  uint8_t st = LP::template project<LPE>(S, r, ah, Kd.tol, !ec, w_mn, w_null, Grp) | ST_LAPACK_PATH;
  st = finish_step<D, T, HP>(P, Kd, w_mn, w_null, sh, alpha, dq, ddq, sn, w_dbg, st);
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) s_out[i] = cvt<T>(sn[i]);
  return st;
}

// Sink of the LAPACK path: the K-scaled Jacobian entries go straight to their cells of Jc (row i is written by lane
// i mod LPE of the group), the three scalar streams are folded into the right-hand side as they arrive (DualSink's
// rule), so no C x n array is ever live in registers.
template <typename T, typename HP, class D, int LPE, class ST>
struct LapackSink {
  using JT = HP;
  using LP = Lapack<HP, D>;
  const DualConsts<HP>& Kd;
  ST& S;
  int lane;
  HP r[at_least_1<D::C>::value];
  ATACOM_HD LapackSink(const DualConsts<HP>& Kd_, ST& S_, int lane_) : Kd(Kd_), S(S_), lane(lane_) {
    ATACOM_UNROLL
    for (int i = 0; i < D::C; ++i) r[i] = HP(0);
  }
  ATACOM_HD void put_c(int i, HP v) { r[i] += Kd.K_c[i] * v; }
  ATACOM_HD void put_Jdq(int i, HP v) { r[i] += Kd.wJ[i] * v; }
  ATACOM_HD void put_b(int i, T v) { r[i] += Kd.wb[i] * cvt<HP>(v); }
  ATACOM_HD void put_J(int i, int j, HP v) {
    if (LPE == 1 || lane < 0 || (i % LPE) == lane) S.set(LP::a(i, j), Kd.K[i] * v);                  // constraints.py:39-40
  }
};

template <class Env, typename T, typename HP, int LPE = 1, class ST, class GRP = SoloGroup>
ATACOM_HD uint8_t step_lapack(const ParamsT<T>& P, const DualConsts<HP>& Kd, ST& S, const T* q, const T* dq, const T* s,
                              const T* alpha, T* ddq, T* s_out, T* w_dbg, const GRP& Grp = GRP()) {
  using D = typename Env::D;
  using LP = Lapack<HP, D>;
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  const bool ec = P.variant == VARIANT_EC;
  LapackSink<T, HP, D, LPE, ST> sink(Kd, S, Grp.write_lane());
  Env::template eval<T, HP>(P, q, dq, sink);
  HP sh[at_least_1<G>::value], sn[at_least_1<G>::value];
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) sh[i] = cvt<HP>(s[i]);
  ATACOM_UNROLL
  for (int i = 0; i < C; ++i) {
    if (LPE == 1 || Grp.write_lane() < 0 || (i % LPE) == Grp.write_lane()) {
      ATACOM_UNROLL
      for (int j = n; j < N; ++j) S.set(LP::a(i, j), (i >= F && j - n == i - F) ? sh[i >= F ? i - F : 0] : HP(0));   // atacom.py:151-165
    }
    if (i >= F) sink.r[i] += HP(0.5) * Kd.K_c[i] * sh[i >= F ? i - F : 0] * sh[i >= F ? i - F : 0];      // atacom.py:195
  }
  HP ah[at_least_1<k>::value], w_mn[N], w_null[N];
  ATACOM_UNROLL
  for (int l = 0; l < k; ++l) ah[l] = cvt<HP>(alpha[l]);
  uint8_t st = LP::template project<LPE>(S, sink.r, ah, Kd.tol, !ec, w_mn, w_null, Grp) | ST_LAPACK_PATH;
  st = finish_step<D, T, HP>(P, Kd, w_mn, w_null, sh, alpha, dq, ddq, sn, w_dbg, st);
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) s_out[i] = cvt<T>(sn[i]);
  return st;
}

// The fix-up kernel's step: only the NULL part is redone with the LAPACK basis.  The minimum-norm part -Jc^+ r is the
// same for every basis; the dual path of the step kernel has computed it already and hands it on (`w_mn`, N values in
// HP), so the right-hand side is not even assembled here: the sink keeps the K-scaled Jacobian entries and drops the
// three scalar streams (what only they depend on — the bias terms, the constraint values — is dead code).
template <typename T, typename HP, class D, int LPE, class ST>
struct NullOnlySink {
  using JT = HP;
  using LP = Lapack<HP, D>;
  const DualConsts<HP>& Kd;
  ST& S;
  int lane;
  ATACOM_HD NullOnlySink(const DualConsts<HP>& Kd_, ST& S_, int lane_) : Kd(Kd_), S(S_), lane(lane_) {}
  ATACOM_HD void put_c(int, HP) {}
  ATACOM_HD void put_Jdq(int, HP) {}
  ATACOM_HD void put_b(int, T) {}
  ATACOM_HD void put_J(int i, int j, HP v) {
    if (LPE == 1 || lane < 0 || (i % LPE) == lane) S.set(LP::a(i, j), Kd.K[i] * v);                  // constraints.py:39-40
  }
};

template <class Env, typename T, typename HP, int LPE = 1, class ST, class GRP = SoloGroup>
ATACOM_HD uint8_t step_lapack_null(const ParamsT<T>& P, const DualConsts<HP>& Kd, ST& S, const T* q, const T* dq,
                                   const T* s, const T* alpha, const HP* w_mn, T* ddq, T* s_out, T* w_dbg,
                                   const GRP& Grp = GRP()) {
  using D = typename Env::D;
  using LP = Lapack<HP, D>;
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  NullOnlySink<T, HP, D, LPE, ST> sink(Kd, S, Grp.write_lane());
  Env::template eval<T, HP>(P, q, dq, sink);
  HP sh[at_least_1<G>::value], sn[at_least_1<G>::value];
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) sh[i] = cvt<HP>(s[i]);
  ATACOM_UNROLL
  for (int i = 0; i < C; ++i) {
    if (LPE == 1 || Grp.write_lane() < 0 || (i % LPE) == Grp.write_lane()) {
      ATACOM_UNROLL
      for (int j = n; j < N; ++j) S.set(LP::a(i, j), (i >= F && j - n == i - F) ? sh[i >= F ? i - F : 0] : HP(0));   // atacom.py:151-165
    }
  }
  HP ah[at_least_1<k>::value], w_null[N];
  ATACOM_UNROLL
  for (int l = 0; l < k; ++l) ah[l] = cvt<HP>(alpha[l]);
  uint8_t st = LP::template project<LPE, false>(S, nullptr, ah, Kd.tol, true, nullptr, w_null, Grp) | ST_LAPACK_PATH;
  st = finish_step<D, T, HP>(P, Kd, w_mn, w_null, sh, alpha, dq, ddq, sn, w_dbg, st);
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) s_out[i] = cvt<T>(sn[i]);
  return st;
}

// atacom.py:145-149
template <typename T, typename HP, class D, typename JT>
ATACOM_HD void slack_from_raw(const ParamsT<T>& P, const RawConstraints<T, HP, D, JT>& R, T* s) {
  constexpr int F = D::F, G = D::G;
  ATACOM_UNROLL
  for (int i = 0; i < G; ++i) {
    const HP v = HP(-2) * (R.c[F + i] + HP(P.K_g[i]) * R.Jdq[F + i]);
    s[i] = v > HP(0) ? T(::sqrt(static_cast<double>(v))) : T(0);
  }
}

// ----------------------------------------------------------------------------- point reach (env C)
// collision_avoidance_atacom.py:29-47,72-127.  env[] = radius^2, K, K_c.  No K_g factor in Jc,
// no alpha_max scaling, no acceleration truncation; bp + bq as written (positions, not velocities).
template <int G_>
struct PointReachEnv {
  using D = Dims<2, 0, G_>;
  template <typename T, typename HP>
  static ATACOM_HD uint8_t step(const ParamsT<T>& P, const T* q, const T* dq, const T* p, const T* dp,
                                const T* s, const T* action, T* w_out, T* s_out, T* w_dbg) {
    constexpr int G = G_, N = 2 + G_;
    const HP rad2 = HP(P.env[0]), K = HP(P.env[1]), Kc = HP(P.env[2]);
    const HP qx = q[0], qy = q[1];
    T A[G][2], r[G];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      const HP px = p[2 * i], py = p[2 * i + 1];
      const HP dx = qx - px, dy = qy - py;
      const HP c_o = rad2 - (dx * dx + dy * dy);
      A[i][0] = T(HP(-2) * dx);
      A[i][1] = T(HP(-2) * dy);
      const HP dc = HP(2) * (dx * HP(dp[2 * i]) + dy * HP(dp[2 * i + 1])) - HP(2) * (dx * HP(dq[0]) + dy * HP(dq[1]));
      const HP bp = (HP(-2) * px + HP(2) * qx) * px + (HP(-2) * py + HP(2) * qy) * py;
      const HP bq = (HP(-2) * qx + HP(2) * px) * qx + (HP(-2) * qy + HP(2) * py) * qy;
      const HP si = s[i];
      const HP c = c_o + HP(0.5) * si * si + K * dc;
      r[i] = T(dc + K * (bp + bq) + Kc * c);
    }
    T w_mn[N], w_null[N];
    uint8_t st = project<T, D, 0>(&A[0][0], s, r, action, P.rref_tol, true, w_mn, w_null);
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
  template <typename T, typename HP>
  static ATACOM_HD void slack_init(const ParamsT<T>& P, const T* q, const T* p, T* s) {
    ATACOM_UNROLL
    for (int i = 0; i < G_; ++i) {
      const HP dx = HP(q[0]) - HP(p[2 * i]), dy = HP(q[1]) - HP(p[2 * i + 1]);
      const HP v = HP(-2) * (HP(P.env[0]) - (dx * dx + dy * dy));
      s[i] = v > HP(0) ? T(::sqrt(static_cast<double>(v))) : T(0);
    }
  }
};

}  // namespace atacom
