// Dual (covariance-form) projection: the hot path of the step kernels.
//
// Same result as project_dense (atacom_core.cuh) — w_mn = -Jc^+ r and w_null = Nc alpha with
// the canonical column-ordered null basis and the reference's tolerance-RREF
// (atacom/atacom.py:127-133, atacom/utils/null_space_coordinate.py:8-26,40-79) — computed
// through the block structure  Jc = [[A_f, 0], [A_g, diag(s)]]  on the DUAL side, where nothing
// is ever divided by a slack variable, so soft rows, stiff rows (s_i -> 0) and active
// constraints (s_i = 0) are one uniform, branch-free code path:
//
//  (1) a trailing inequality row with a diagonal Jacobian (joint limit: d_j x_j + s_j z_j = rho_j)
//      is absorbed exactly by a change of coordinates in its (x_j, z_j) plane:
//          x_j = mu_j + sigma_j t_j,  z_j = zeta_j - gamma_j t_j,
//          (sigma_j, gamma_j) = (s_j, d_j) / sqrt(d_j^2 + s_j^2),
//      an isometry from the free coordinate t_j onto the row's solution line;
//  (2) the m = F + GD dense rows become  B t + S_d z_d = rho~  with  B = A_d diag(sigma); their
//      Gram matrix  H = B B^T + S_d^2  (m x m, SPD whenever Jc has full row rank) is factored
//      H = L L^T and the minimum-norm solution is  t = B^T lambda, z_d = S_d lambda,
//      lambda = H^-1 rho~;
//  (3) the x-x block of the projector onto null(Jc) is  Gx = Sigma^1/2 (I - Y^T Y) Sigma^1/2
//      with  Y = L^-1 B;
//  (4) the canonical basis (oracle/nullspace.py:canonical_null_basis) is the Cholesky
//      factorisation of the projector in column order, skipping a column whose pivot is
//      <= tol: on the x columns that is a skip-Cholesky of Gx.  The RREF multipliers beta
//      (forward substitution) give the x entries of w_null directly — a dropped column keeps
//      only the rows pivoted before it, which is what the upper-triangular storage holds;
//  (5,6) the slack entries are those of the null vector  v* = Pi u,  u = sum_r theta_r e_{c_r},
//      theta = V_p^-1 beta, again through the dual:  lambda* = H^-1 B g,  z_d = -S_d lambda*,
//      t* = g - B^T lambda*,  z_diag = -gamma t*;
//  (7) if one pivot is still missing after the x columns (an active constraint took a
//      direction away from x), the remaining null direction is  (nu, -A_g nu / s)  with nu
//      supported on the two (F = 1) / one (F = 0) unpivoted x columns; the first slack column
//      whose entry exceeds tol becomes the coordinate; with two pivots missing the same is done in the
//      remaining plane (two_slack_pivots); three or more missing pivots, or F > 1, take the general
//      routine, the oracle's own procedure in the reduced coordinates: null_part_general (out of line,
//      one thread) or, where Y and L live in shared memory, null_part_general_warp (the 32 lanes of the
//      warp on one environment at a time; the caller of project<true>() arranges it).
//
// All arithmetic is in R (double on the device: B200 runs FP64 at half the FP32 rate and the
// dual Gram matrix squares cond(Jc)); ~1.0 kFLOP per iiwa environment, no data-dependent
// branch before (7).
#pragma once

#include "atacom_core.cuh"
#include "atacom_lapack.cuh"

namespace atacom {

// 1 / sqrt(x): MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-20, the whole double range — no conversion to fp32 and
// back, which cost two F2F per call and 1.8 % of the step kernel) and one third-order correction in R: 6
// instructions, no slow path; ::rsqrt(double) costs ~27 with a call.  Arguments are floored at DUAL_TINY by the callers.
template <typename R>
ATACOM_HD R dual_rsqrt(R x) {
#if defined(__CUDA_ARCH__)
  double yd;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yd) : "d"(static_cast<double>(x)));
  const R y = static_cast<R>(yd);
  const R e = ::fma(-x * y, y, R(1));               // 1 - x y^2 ~ 1e-6
  return ::fma(y, ::fma(R(0.375), e, R(0.5)) * e, y);   // y (1 + e/2 + 3 e^2 / 8): error O(e^3)
#else
  return R(1) / ::sqrt(x);
#endif
}
// The same with a second (first-order) correction, for the slack-pivot phases: they form 1 / |s_i| ~ 1e12 for an
// exactly active constraint and Gram determinants of such entries (1e48 and beyond — the fp32-seeded rsqrt this
// replaced overflowed there), and what they compute is compared against a tolerance.
template <typename R>
ATACOM_HD R dual_rsqrt_wide(R x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(static_cast<double>(x)));
  const double e = ::fma(-static_cast<double>(x) * y, y, 1.0);      // ~ 1e-6
  y = ::fma(y, ::fma(0.375, e, 0.5) * e, y);                        // error O(e^3)
  const double e2 = ::fma(-static_cast<double>(x) * y, y, 1.0);
  return static_cast<R>(::fma(y, 0.5 * e2, y));
#else
  return R(1) / ::sqrt(x);
#endif
}
constexpr double DUAL_TINY = 1e-24;   // floor of every rsqrt argument: squares of fp32 data below it are zero here

// Per-environment scratch of the dual projection.  LocalStore: a plain array (registers / local
// memory; the host build and the small environments).  SharedStore: the environment's column of a
// [SIZE][STRIDE] shared-memory array, one environment per thread — conflict-free, and it takes the
// two big operands (Y: m x n, L: m(m+1)/2) out of the register budget of the iiwa kernels.
template <typename R, int SIZE>
struct LocalStore {
  R v[SIZE > 0 ? SIZE : 1];
  ATACOM_HD R get(int i) const { return v[i]; }
  ATACOM_HD void set(int i, R x) { v[i] = x; }
  // entry at a run-time index (a shared-memory address on the device; here a select chain so that the
  // array can stay in registers)
  ATACOM_HD R get_dyn(int i) const {
    R x = R(0);
    ATACOM_UNROLL
    for (int c = 0; c < SIZE; ++c) x = (c == i) ? v[c] : x;
    return x;
  }
};
template <typename R, int STRIDE>
struct SharedStore {
  R* base;   // &array[0][threadIdx.x]
  // volatile: a plain access lets the compiler forward the stored value to the later load through a
  // register, i.e. keep the operand in the register file after all (and spill it to local memory)
  ATACOM_HD R get(int i) const { return *static_cast<const volatile R*>(base + i * STRIDE); }
  ATACOM_HD void set(int i, R x) { *static_cast<volatile R*>(base + i * STRIDE) = x; }
  ATACOM_HD R get_dyn(int i) const { return *static_cast<const volatile R*>(base + i * STRIDE); }
};

// The same column of a [SIZE][STRIDE] shared-memory array with plain accesses: for code that WANTS the compiler to
// schedule its loads early and keep values in registers (the LAPACK-basis routine: rows and reflector vectors).
template <typename R, int STRIDE>
struct PlainSharedStore {
  R* base;
  ATACOM_HD R get(int i) const { return base[i * STRIDE]; }
  ATACOM_HD void set(int i, R x) { base[i * STRIDE] = x; }
};

template <class S> struct is_shared_store { static constexpr bool value = false; };
template <typename R, int STRIDE> struct is_shared_store<SharedStore<R, STRIDE>> { static constexpr bool value = true; };
template <typename R, int STRIDE>
ATACOM_HD R* shared_base(const SharedStore<R, STRIDE>& s) { return s.base; }
template <typename R, int SIZE>
ATACOM_HD R* shared_base(LocalStore<R, SIZE>&) { return nullptr; }

template <typename R, class D, int NDIAG>
struct Dual {
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  static constexpr int GD = G - NDIAG;  // dense inequality rows
  static constexpr int m = F + GD;      // dense rows (equalities first)
  static constexpr int M1 = at_least_1<m>::value, K1 = at_least_1<k>::value, G1 = at_least_1<G>::value;
  static constexpr int Y_SIZE = m * n, L_SIZE = m * (m + 1) / 2;
  static_assert(NDIAG <= n && NDIAG <= G, "diagonal rows: row GD + j has its single entry in column j");
  static constexpr ATACOM_HD int lidx(int i, int l) { return i * (i - 1) / 2 + l; }   // strictly lower part of L
  static constexpr int LI0 = m * (m - 1) / 2;                                        // then 1 / diag(L)

  // (7c) Any number of slack pivots (and any F): the null part once more, literally the way the oracle does it
  // (oracle/nullspace.py:canonical_null_basis + tol_rref), but in the reduced coordinates (t, z_d) where the
  // projector is I - W^T W with W = L^-1 [B, S_d] — still nothing divided by a slack.  Rolled loops over a
  // (n + GD)^2 matrix in local memory: ~2 k FLOP, small code; taken when three or more constraints are active at
  // once.  Column c of Jc corresponds to the reduced vector sc * e_idx: x_c -> sigma_c e_{t_c}, dense slack i ->
  // e_{z_i}, diagonal slack j -> -gamma_j e_{t_j}.
  template <class YS, class LS>
  static ATACOM_NOINLINE uint8_t null_part_general(YS& Y, LS& Ls, const R* s, const R* sig, const R* gam,
                                                   const R* alpha, R tol, R* w_null) {
    constexpr int NR = n + GD;
    uint8_t status = 0;
    R W[M1][NR], Pm[NR][NR], V[K1][NR], beta[K1];
    ATACOM_ROLLED
    for (int l = 0; l < m; ++l) {
      ATACOM_ROLLED
      for (int j = 0; j < n; ++j) W[l][j] = Y.get_dyn(l * n + j);
      ATACOM_ROLLED
      for (int i = 0; i < GD; ++i) W[l][n + i] = R(0);
    }
    ATACOM_ROLLED
    for (int i = 0; i < GD; ++i) {   // column F + i of L^-1 times s_i (forward substitution)
      const int c0 = F + i;
      ATACOM_ROLLED
      for (int l = c0; l < m; ++l) {
        R acc = (l == c0) ? s[i] : R(0);
        ATACOM_ROLLED
        for (int p = c0; p < l; ++p) acc -= Ls.get_dyn(lidx(l, p)) * W[p][n + i];
        W[l][n + i] = acc * Ls.get_dyn(LI0 + l);
      }
    }
    ATACOM_ROLLED
    for (int a = 0; a < NR; ++a) {
      R wa[M1];
      ATACOM_UNROLL
      for (int l = 0; l < m; ++l) wa[l] = W[l][a];
      ATACOM_ROLLED
      for (int b = a; b < NR; ++b) {
        R acc = (a == b) ? R(1) : R(0);
        ATACOM_UNROLL
        for (int l = 0; l < m; ++l) acc -= wa[l] * W[l][b];
        Pm[a][b] = acc;
        Pm[b][a] = acc;
      }
    }
    int npiv = 0;
    ATACOM_ROLLED
    for (int c = 0; c < N; ++c) {
      int idx;
      R sc;
      if (c < n) {
        idx = c;
        sc = sig[c];
      } else if (c - n < GD) {
        idx = c;          // n + i
        sc = R(1);
      } else {
        idx = c - n - GD;
        sc = -gam[idx];
      }
      R part = R(0);     // sum over the rows pivoted so far of beta_r V[r][c]
      ATACOM_ROLLED
      for (int r = 0; r < npiv; ++r) part += beta[r] * sc * V[r][idx];
      const R d = sc * sc * Pm[idx][idx];
      if (npiv < k && d > tol * tol && d > R(DUAL_TINY)) {
        const R inv = dual_rsqrt(d);
        ATACOM_ROLLED
        for (int a = 0; a < NR; ++a) V[npiv][a] = Pm[a][idx] * sc * inv;
        ATACOM_ROLLED
        for (int a = 0; a < NR; ++a) {
          const R va = V[npiv][a];
          ATACOM_UNROLL
          for (int b = 0; b < NR; ++b) Pm[a][b] -= va * V[npiv][b];
        }
        beta[npiv] = (alpha[npiv] - part) * inv;
        w_null[c] = alpha[npiv];
        if (c >= n) status |= ST_SLACK_PIVOT;
        ++npiv;
      } else {
        if (npiv < k) status |= ST_COLUMN_DROPPED;
        w_null[c] = part;   // dropped: the rows pivoted before it; untested (npiv == k): every row
      }
    }
    if (npiv < k) status |= ST_RANK_DEFICIENT;
    return status;
  }

#if defined(__CUDACC__)
  // (7c'), device only: the same procedure run by all 32 lanes of the warp for ONE of its environments (the
  // "owner"), with the matrices in a small scratch in shared memory — a few microseconds instead of ~70 us for
  // the serial routine, whose local-memory accesses are one dependent chain (a single such environment would set
  // the duration of the whole launch).  Every lane of the warp must call it (converged).  Yo / Lo point at the
  // owner's columns of the [entry][lane] scratch arrays (stride 32); sc holds, filled in by the owner beforehand:
  // sigma and gamma (COOP_SG0: n + n), the slacks (COOP_S0: G) and alpha (COOP_A0: k).  Returns the status bits
  // and leaves the owner's w_null (N values) in sc[COOP_WN0 ...].
  static constexpr int COOP_NR = n + GD;
  static constexpr int COOP_W0 = 0;                                   // W (m x NR), later V (k x NR)
  static constexpr int COOP_P0 = (m > k ? m : k) * (n + GD);          // Pm (NR x NR)
  static constexpr int COOP_WN0 = COOP_P0 + (n + GD) * (n + GD);      // w_null (N)
  static constexpr int COOP_SG0 = COOP_WN0 + N;                       // sigma (n), gamma (n)
  static constexpr int COOP_S0 = COOP_SG0 + 2 * n;                    // slacks (G)
  static constexpr int COOP_A0 = COOP_S0 + G;                         // alpha (k)
  static constexpr int COOP_DOUBLES = COOP_A0 + k;
  static_assert(COOP_NR <= 32, "one lane per reduced coordinate");
  static __device__ __noinline__ uint8_t null_part_general_warp(const volatile R* Yo, const volatile R* Lo, R tol,
                                                                volatile R* sc) {
    constexpr int NR = COOP_NR;
    const int lane = threadIdx.x & 31;
    uint8_t status = 0;
    // W = [Y, L^-1 S_d]: the x part element-wise, one dense slack column per lane (forward substitution)
    for (int e = lane; e < m * n; e += 32) sc[COOP_W0 + (e / n) * NR + (e % n)] = Yo[e * 32];
    if (lane < GD) {
      const R si = sc[COOP_S0 + lane];
      const int c0 = F + lane;
      for (int l = 0; l < m; ++l) {
        R acc = R(0);
        if (l >= c0) {
          acc = (l == c0) ? si : R(0);
          for (int p = c0; p < l; ++p) acc -= Lo[lidx(l, p) * 32] * sc[COOP_W0 + p * NR + n + lane];
          acc *= Lo[(LI0 + l) * 32];
        }
        sc[COOP_W0 + l * NR + n + lane] = acc;
      }
    }
    __syncwarp();
    for (int e = lane; e < NR * NR; e += 32) {          // projector of the reduced coordinates: I - W^T W
      const int a = e / NR, b = e % NR;
      R acc = (a == b) ? R(1) : R(0);
      for (int l = 0; l < m; ++l) acc -= sc[COOP_W0 + l * NR + a] * sc[COOP_W0 + l * NR + b];
      sc[COOP_P0 + e] = acc;
    }
    __syncwarp();
    R beta[K1];
    ATACOM_UNROLL
    for (int l = 0; l < K1; ++l) beta[l] = R(0);
    int npiv = 0;
    for (int c = 0; c < N; ++c) {      // uniform control flow: every lane reads the same values, takes the same decisions
      int idx;
      R scl;
      if (c < n) {
        idx = c;
        scl = sc[COOP_SG0 + c];
      } else if (c - n < GD) {
        idx = c;
        scl = R(1);
      } else {
        idx = c - n - GD;
        scl = -sc[COOP_SG0 + n + idx];
      }
      R part = R(0);
      ATACOM_UNROLL
      for (int r = 0; r < k; ++r) part += (r < npiv) ? beta[r] * scl * sc[COOP_W0 + r * NR + idx] : R(0);
      const R d = scl * scl * sc[COOP_P0 + idx * NR + idx];
      if (npiv < k && d > tol * tol && d > R(DUAL_TINY)) {
        const R inv = dual_rsqrt(d);
        const R al = sc[COOP_A0 + npiv];
        __syncwarp();                  // everyone has read row npiv of V's storage and the diagonal
        if (lane < NR) sc[COOP_W0 + npiv * NR + lane] = sc[COOP_P0 + idx * NR + lane] * scl * inv;   // row npiv of V
        __syncwarp();
        for (int e = lane; e < NR * NR; e += 32)
          sc[COOP_P0 + e] -= sc[COOP_W0 + npiv * NR + e / NR] * sc[COOP_W0 + npiv * NR + e % NR];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) beta[l] = (l == npiv) ? (al - part) * inv : beta[l];
        if (lane == 0) sc[COOP_WN0 + c] = al;
        if (c >= n) status |= ST_SLACK_PIVOT;
        ++npiv;
        __syncwarp();
      } else {
        if (npiv < k) status |= ST_COLUMN_DROPPED;
        if (lane == 0) sc[COOP_WN0 + c] = part;   // dropped: the rows pivoted before it; untested: every row
      }
    }
    __syncwarp();
    if (npiv < k) status |= ST_RANK_DEFICIENT;
    return status;
  }

#endif

  // The serial general routine for a thread whose project<true>() call returned the request (sigma and gamma in
  // w_null[0 .. 2n)): what the caller runs when too many lanes of its warp ask at once for the warp-cooperative
  // routine to pay off.
  template <class YS, class LS>
  static ATACOM_HD uint8_t general_from_deferred(YS& Y, LS& Ls, const R* s, const R* alpha, R tol, R* w_null) {
    R s2[G1], sg2[n], gm2[n], al2[K1], wn2[N];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) s2[i] = s[i];
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) {
      sg2[j] = w_null[j < N ? j : 0];
      gm2[j] = w_null[n + j < N ? n + j : 0];
    }
    ATACOM_UNROLL
    for (int l = 0; l < k; ++l) al2[l] = alpha[l];
    const uint8_t st2 = null_part_general(Y, Ls, s2, sg2, gm2, al2, tol, wn2);
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) w_null[i] = wn2[i];
    return st2;
  }

  // (7b) Two pivots missing after the x columns (two active constraints took two directions away from x; k >= 2).
  // The remaining null space is two-dimensional: tau supported on the F + 2 unpivoted columns with B_f tau = 0,
  // spanned by tau1, tau2; the full vectors are n_j = (sigma tau_j, -gamma tau_j, -(B tau_j) / s) with Gram matrix
  // g.  The component of the projected unit vector of slack column i in that plane has norm^2
  // [a_i b_i] g^-1 [a_i b_i]^T, (a_i, b_i) the z_i entries of n_1, n_2: the first column above tol^2 is the first
  // pivot (row q1), the direction of the plane with zero z_i1 entry is what is left for the second (row q2).
  template <class YS, class LS>
  static ATACOM_HD uint8_t two_slack_pivots(YS& Y, LS& Ls, const R* s, const R* alpha, R tol, const R* gam,
                                            const bool* take, int npiv, uint8_t status, R* w_null) {
    int u1 = 0, u2 = 0, u3 = 0;
    {
      int cnt = 0;
      ATACOM_UNROLL
      for (int c = 0; c < n; ++c) {
        u1 = (!take[c] && cnt == 0) ? c : u1;
        u2 = (!take[c] && cnt == 1) ? c : u2;
        u3 = (!take[c] && cnt == 2) ? c : u3;
        cnt += take[c] ? 0 : 1;
      }
    }
    R t1[3] = {R(1), R(0), R(0)}, t2[3] = {R(0), R(1), R(0)};   // coordinates at u1, u2, u3 (F = 0: unit vectors)
    if (F == 1) {
      const R v1 = Y.get_dyn(u1), v2 = Y.get_dyn(u2), v3 = Y.get_dyn(u3);   // equality row of Y ~ B_f
      const R m1 = num<R>::abs(v1), m2 = num<R>::abs(v2), m3 = num<R>::abs(v3);
      const bool p3 = m3 >= m1 && m3 >= m2, p2 = !p3 && m2 >= m1;
      // the largest entry is the one solved for
      t1[0] = p3 ? v3 : (p2 ? v2 : -v2);
      t1[1] = p3 ? R(0) : (p2 ? -v1 : v1);
      t1[2] = p3 ? -v1 : R(0);
      t2[0] = p3 ? R(0) : (p2 ? R(0) : -v3);
      t2[1] = p3 ? v3 : (p2 ? -v3 : R(0));
      t2[2] = p3 ? -v2 : (p2 ? v2 : v1);
    }
    // n_1, n_2 as vectors over PC components: the 3 tau coordinates (the x and z_diag entries of a joint have
    // joint norm |tau_j|, sigma^2 + gamma^2 = 1) and the GD dense slack entries.  A slack that is (nearly) zero
    // makes its entry huge, so the Gram determinant and everything derived from it is formed from 2 x 2
    // minors and differences (Lagrange identity) — never from g11 g22 - g12^2, which would cancel.
    constexpr int PC = 3 + GD;
    R U[PC], V[PC], ya[M1], yb[M1], ea[G1], eb[G1];
    ATACOM_UNROLL
    for (int c = 0; c < 3; ++c) {
      U[c] = t1[c];
      V[c] = t2[c];
    }
    ATACOM_UNROLL
    for (int l = 0; l < m; ++l) {   // B = L Y  =>  b_i . tau = sum_{l <= i} L[i][l] (Y[l] . tau)
      const R y1 = Y.get_dyn(l * n + u1), y2 = Y.get_dyn(l * n + u2), y3 = (F == 1) ? Y.get_dyn(l * n + u3) : R(0);
      ya[l] = y1 * t1[0] + y2 * t1[1] + y3 * t1[2];
      yb[l] = y1 * t2[0] + y2 * t2[1] + y3 * t2[2];
    }
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      if (i < GD) {
        const int ri = F + (i < GD ? i : 0);
        const R lii = Ls.get(LI0 + ri);
        const R dii = (lii > R(0)) ? dual_rsqrt_wide(lii * lii) : R(0);   // L[ri][ri] = 1 / li
        R an = ya[ri] * dii, bn = yb[ri] * dii;
        ATACOM_UNROLL
        for (int l = 0; l < ri; ++l) {
          const R lv = Ls.get(lidx(ri, l));
          an += lv * ya[l];
          bn += lv * yb[l];
        }
        const R s2 = s[i] * s[i];
        const R rs = dual_rsqrt(s2 > R(DUAL_TINY) ? s2 : R(DUAL_TINY));   // 1 / |s_i|, s_i = 0 -> the limit
        const R sg = s[i] < R(0) ? rs : -rs;
        ea[i] = an * sg;
        eb[i] = bn * sg;
        U[3 + (i < GD ? i : 0)] = ea[i];
        V[3 + (i < GD ? i : 0)] = eb[i];
      } else {
        const int j = i >= GD ? i - GD : 0;
        const R ta = (j == u1) ? t1[0] : ((j == u2) ? t1[1] : ((F == 1 && j == u3) ? t1[2] : R(0)));
        const R tb = (j == u1) ? t2[0] : ((j == u2) ? t2[1] : ((F == 1 && j == u3) ? t2[2] : R(0)));
        ea[i] = -gam[j] * ta;
        eb[i] = -gam[j] * tb;
      }
    }
    R det = R(0);
    ATACOM_UNROLL
    for (int p = 0; p < PC; ++p) {
      ATACOM_UNROLL
      for (int q = p + 1; q < PC; ++q) {
        const R mn = U[p] * V[q] - U[q] * V[p];
        det += mn * mn;
      }
    }
    if (!(det > R(DUAL_TINY))) return status | ST_DENSE_PATH;
    const R rd = dual_rsqrt_wide(det);
    const R idet = rd * rd;
    // first pivot: the first slack column whose component in the plane exceeds tol:
    // |a_i n_2 - b_i n_1|^2 / det > tol^2
    int i1 = G;
    R a1 = R(0), b1 = R(0), z1 = R(0);
    ATACOM_UNROLL
    for (int i = G - 1; i >= 0; --i) {
      R nn = R(0);
      ATACOM_UNROLL
      for (int p = 0; p < PC; ++p) {
        const R d = ea[i] * V[p] - eb[i] * U[p];
        nn += d * d;
      }
      const bool hit = nn * idet > tol * tol;
      i1 = hit ? i : i1;
      a1 = hit ? ea[i] : a1;
      b1 = hit ? eb[i] : b1;
      z1 = hit ? w_null[n + i] : z1;
    }
    if (i1 == G) return status | ST_RANK_DEFICIENT;
    // d = a1 n_2 - b1 n_1: the direction of the plane with zero z_i1 entry (second row, up to sign and norm);
    // the first row is the unit vector of the plane orthogonal to it: q1 = (k1 n_1 + k2 n_2) with
    // (k1, k2) = g^-1 (a1, b1) / p1 = ((n_2 . d), -(n_1 . d)) / (det p1),  p1^2 = |d|^2 / det
    R nrm2 = R(0), vd = R(0), ud = R(0);
    ATACOM_UNROLL
    for (int p = 0; p < PC; ++p) {
      const R d = a1 * V[p] - b1 * U[p];
      nrm2 += d * d;
      vd += V[p] * d;
      ud += U[p] * d;
    }
    const R rn = dual_rsqrt_wide(nrm2 > R(DUAL_TINY) ? nrm2 : R(DUAL_TINY));
    const R ip1 = rn * (det * rd);                 // 1 / p1 = sqrt(det) / |d|
    const R k1 = vd * idet * ip1, k2 = -ud * idet * ip1;
    int i2 = G;
    R f2 = R(0), z2 = R(0), q12 = R(0);
    ATACOM_UNROLL
    for (int i = G - 1; i >= 0; --i) {
      const R fi = (b1 * ea[i] - a1 * eb[i]) * rn;     // entry of the second row (up to its sign)
      const R qi = k1 * ea[i] + k2 * eb[i];            // entry of the first row
      const bool hit = (i > i1) && (num<R>::abs(fi) > tol);
      i2 = hit ? i : i2;
      f2 = hit ? fi : f2;
      z2 = hit ? w_null[n + i] : z2;
      q12 = hit ? qi : q12;
      ea[i] = qi;     // from here on: first-row entries in ea, second-row entries in eb
      eb[i] = fi;
    }
    if (i2 == G) return status | ST_RANK_DEFICIENT;
    R al1 = R(0), al2 = R(0);
    ATACOM_UNROLL
    for (int l = 0; l < k; ++l) {
      al1 = (l == npiv) ? alpha[l] : al1;
      al2 = (l == npiv + 1) ? alpha[l] : al2;
    }
    const R bt1 = (al1 - z1) * ip1;
    const R rf = dual_rsqrt_wide(f2 * f2);
    const R bt2 = (al2 - z2 - bt1 * q12) * (f2 * rf * rf);   // / f2: beta of the second row times the sign of its pivot
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      const R v = w_null[n + i];
      w_null[n + i] = (i < i1) ? v : ((i == i1) ? al1 : ((i < i2) ? v + bt1 * ea[i] : ((i == i2) ? al2 : v + bt1 * ea[i] + bt2 * eb[i])));
    }
    return status | ST_SLACK_PIVOT | ST_COLUMN_DROPPED;
  }

  // Y: on entry the m x n dense rows of A (row-major, equalities first), overwritten; Ls: scratch for L.
  // dg: NDIAG diagonal entries, s: G, r: C (rows ordered equality, dense inequality, diagonal
  // inequality), alpha: k.  w_mn (type W), w_null (type R): N each.
  // defer_general: do not run the general routine here but return ST_DENSE_PATH as a request, with sigma and gamma
  // in w_null[0 .. 2n) — the caller then runs it with the whole warp (null_part_general_warp) and overwrites w_null.
  // band_defer: leave every environment whose result may depend on the null BASIS to the LAPACK-basis routine
  // (atacom_lapack.cuh) and return ST_LAPACK_PATH for it, w_null unfinished.  The reference's rref can only drop
  // column j (r pivots found so far) if the canonical pivot p_j — the norm of what is left of P e_j, computed here —
  // is at most sqrt(k - r) tol: its candidate is the largest of k - r entries y_i = <w_i, P e_j> of rows w_i that
  // are an elimination-transformed basis of the remaining null space with Gram matrix >= I, so max |y_i| >=
  // p_j / sqrt(k - r).  Outside that band no column is dropped by either procedure, the pivots are the first k
  // columns, and the result is the basis-free closed form computed here.
  // BAND_ONLY: the same decided at compile time (the step kernels of ATACOM_BASIS_LAPACK): everything from (7) on —
  // slack pivots, the general routine — is then unreachable and not even compiled into the kernel.
  template <bool defer_general = false, bool BAND_ONLY = false, class YS, class LS, typename W>
  static ATACOM_HD uint8_t project(YS& Y, LS& Ls, const R* dg, const R* s, const R* r, const R* alpha, R tol,
                                   bool want_null, W* w_mn, R* w_null, bool band_defer = false) {
    uint8_t status = 0;

    // ---- (1) diagonal rows -> coordinates t_j
    R sig[n], gam[n], mu[n], zeta[n];
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) {
      sig[j] = R(1);
      gam[j] = R(0);
      mu[j] = R(0);
      zeta[j] = R(0);
      if (j < NDIAG) {
        const R d = dg[j < NDIAG ? j : 0], sj = s[GD + (j < NDIAG ? j : 0)];
        const R den = d * d + sj * sj;
        if (den > R(DUAL_TINY)) {
          const R rs = dual_rsqrt(den);
          const R rho = -r[F + GD + (j < NDIAG ? j : 0)] * rs;
          sig[j] = sj * rs;
          gam[j] = d * rs;
          mu[j] = rho * gam[j];
          zeta[j] = rho * sig[j];
        } else {
          status |= ST_RANK_DEFICIENT;  // all-zero row of Jc
        }
      }
    }

    // ---- (2) dense rows, one column of A at a time: rho~ = -r - A mu, B = A diag(sigma) (stored back),
    // H = B B^T + S_d^2 accumulated as rank-1 updates
    R L[M1][M1], li[M1], u[M1];
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      u[i] = -r[i];
      ATACOM_UNROLL
      for (int l = 0; l <= i; ++l) L[i][l] = (l == i && i >= F) ? s[i >= F ? i - F : 0] * s[i >= F ? i - F : 0] : R(0);
    }
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) {
      R bj[M1];
      ATACOM_UNROLL
      for (int i = 0; i < m; ++i) {
        const R a = Y.get(i * n + j);
        u[i] -= a * mu[j];
        bj[i] = (j < NDIAG) ? a * sig[j] : a;
        if (j < NDIAG) Y.set(i * n + j, bj[i]);
      }
      ATACOM_UNROLL
      for (int i = 0; i < m; ++i) {
        ATACOM_UNROLL
        for (int l = 0; l <= i; ++l) L[i][l] += bj[i] * bj[l];
      }
    }
    // H = L L^T (in place), li = 1 / diag(L)
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      ATACOM_UNROLL
      for (int l = 0; l < i; ++l) {
        R v = L[i][l];
        ATACOM_UNROLL
        for (int p = 0; p < l; ++p) v -= L[i][p] * L[l][p];
        L[i][l] = v * li[l];
      }
      R d = L[i][i];
      const R d0 = d;
      ATACOM_UNROLL
      for (int p = 0; p < i; ++p) d -= L[i][p] * L[i][p];
      if (d > R(512) * num<R>::eps() * d0 && d > R(DUAL_TINY)) {
        li[i] = dual_rsqrt(d);
      } else {  // dependent row of Jc: dropped from the solve, flagged
        status |= ST_RANK_DEFICIENT;
        li[i] = R(0);
      }
    }
    // u = L^-1 rho~, lambda = L^-T u;  B^T lambda = Y^T u is picked up while Y is formed below
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      R acc = u[i];
      ATACOM_UNROLL
      for (int l = 0; l < i; ++l) acc -= L[i][l] * u[l];
      u[i] = acc * li[i];
    }
    {
      R lam[M1];
      ATACOM_UNROLL
      for (int i = m - 1; i >= 0; --i) {
        R acc = u[i];
        ATACOM_UNROLL
        for (int l = i + 1; l < m; ++l) acc -= L[l][i] * lam[l];
        lam[i] = acc * li[i];
      }
      ATACOM_UNROLL
      for (int i = 0; i < G; ++i) {
        if (i < GD) w_mn[n + i] = cvt<W>(s[i] * lam[F + (i < GD ? i : 0)]);
        w_null[n + i] = R(0);
      }
    }

    // ---- (3a) Y = L^-1 B, one column at a time (stored back); t = Y^T u
    const bool null_part = want_null && k > 0;
    {
      R cur[M1], nxt[M1];
      ATACOM_UNROLL
      for (int i = 0; i < m; ++i) cur[i] = Y.get(i * n);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) {   // the next column is on its way while this one is solved
          ATACOM_UNROLL
          for (int i = 0; i < m; ++i) nxt[i] = Y.get(i * n + (j + 1 < n ? j + 1 : 0));
        }
        R yj[M1];
        R t = R(0);
        ATACOM_UNROLL
        for (int i = 0; i < m; ++i) {
          R v = cur[i];
          ATACOM_UNROLL
          for (int l = 0; l < i; ++l) v -= L[i][l] * yj[l];
          yj[i] = v * li[i];
          t += yj[i] * u[i];
        }
        if (null_part) {
          ATACOM_UNROLL
          for (int i = 0; i < m; ++i) Y.set(i * n + j, yj[i]);
        }
        w_mn[j] = cvt<W>((j < NDIAG) ? mu[j] + sig[j] * t : t);
        if (j < NDIAG) w_mn[n + GD + j] = cvt<W>(zeta[j] - gam[j] * t);
        w_null[j] = R(0);
        ATACOM_UNROLL
        for (int i = 0; i < m; ++i) cur[i] = nxt[i];
      }
    }
    if (!null_part) return status;
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      Ls.set(LI0 + i, li[i]);
      ATACOM_UNROLL
      for (int l = 0; l < i; ++l) Ls.set(lidx(i, l), L[i][l]);
    }

    // ---- (3b) Gx = Sigma^1/2 (I - Y^T Y) Sigma^1/2 (upper triangle), one row of Y at a time
    R Gx[n][n];
    ATACOM_UNROLL
    for (int a = 0; a < n; ++a) {
      ATACOM_UNROLL
      for (int b = a; b < n; ++b) Gx[a][b] = (a == b) ? R(1) : R(0);
    }
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      R yi[n];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) yi[j] = Y.get(i * n + j);
      ATACOM_UNROLL
      for (int a = 0; a < n; ++a) {
        ATACOM_UNROLL
        for (int b = a; b < n; ++b) Gx[a][b] -= yi[a] * yi[b];
      }
    }
    ATACOM_UNROLL
    for (int a = 0; a < n; ++a) {
      ATACOM_UNROLL
      for (int b = a; b < n; ++b) {
        if (a < NDIAG) Gx[a][b] *= sig[a];
        if (b < NDIAG) Gx[a][b] *= sig[b];
      }
    }

    // ---- (4) skip-Cholesky of Gx in column order with the rref tolerance, in place; row c of the basis is
    // stored by pivot COLUMN (zero row for a skipped column), so every loop range is static.
    R (&Vc)[n][n] = Gx;
    R vinv[n], bcol[n];
    bool take[n];
    int npiv = 0;
    bool band = false;
    ATACOM_UNROLL
    for (int c = 0; c < n; ++c) {
      const R d = Gx[c][c];
      const bool tested = npiv < k;
      // (a numerically zero pivot — a column that is structurally dependent, e.g. the last joint of iiwa-7 — is dropped
      // by every basis alike and what is zeroed there is rounding noise: not a reason to defer)
      band = band || (tested && !(d > tol * tol * R(k - npiv)) && d > R(1e-14));
      const bool tk = tested && (d > tol * tol) && (d > R(DUAL_TINY));
      const R inv = tk ? dual_rsqrt(d) : R(0);
      take[c] = tk;
      vinv[c] = inv;
      if (tested && !tk) status |= ST_COLUMN_DROPPED;
      ATACOM_UNROLL
      for (int b = c; b < n; ++b) Vc[c][b] = Gx[c][b] * inv;
      ATACOM_UNROLL
      for (int a = c + 1; a < n; ++a) {
        ATACOM_UNROLL
        for (int b = a; b < n; ++b) Gx[a][b] -= Vc[c][a] * Vc[c][b];
      }
      R part = R(0);
      ATACOM_UNROLL
      for (int cp = 0; cp < c; ++cp) part += bcol[cp] * Vc[cp][c];
      R al = R(0);
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) al = (l == npiv) ? alpha[l] : al;
      bcol[c] = (al - part) * inv;        // 0 for a skipped column
      w_null[c] = tk ? al : part;         // pivot: alpha; dropped / untested: rows pivoted before it
      npiv += tk ? 1 : 0;
    }

    // ---- (5) theta = V_p^-1 beta (back substitution), g = Sigma^1/2 theta
    R gt[n];
    ATACOM_UNROLL
    for (int c = n - 1; c >= 0; --c) {
      R acc = bcol[c];
      ATACOM_UNROLL
      for (int b = c + 1; b < n; ++b) acc -= Vc[c][b] * gt[b];
      gt[c] = acc * vinv[c];
    }
    ATACOM_UNROLL
    for (int j = 0; j < NDIAG; ++j) gt[j] *= sig[j];

    // ---- (6) slack entries of v* = Pi u:  y = Y g,  t* = g - Y^T y,  lambda* = L^-T y
    R yv[M1], ts[n];
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) ts[j] = gt[j];
    ATACOM_UNROLL
    for (int i = 0; i < m; ++i) {
      R yi[n];
      R acc = R(0);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) {
        yi[j] = Y.get(i * n + j);
        acc += yi[j] * gt[j];
      }
      yv[i] = acc;
      ATACOM_UNROLL
      for (int j = 0; j < NDIAG; ++j) ts[j] -= yi[j] * acc;
    }
    ATACOM_UNROLL
    for (int j = 0; j < NDIAG; ++j) w_null[n + GD + j] = -gam[j] * ts[j];
    ATACOM_UNROLL
    for (int i = m - 1; i >= 0; --i) {
      R acc = yv[i];
      ATACOM_UNROLL
      for (int l = i + 1; l < m; ++l) acc -= Ls.get(lidx(l, i)) * yv[l];
      yv[i] = acc * Ls.get(LI0 + i);
    }
    ATACOM_UNROLL
    for (int i = 0; i < GD; ++i) w_null[n + i] = -s[i] * yv[F + i];
    if constexpr (BAND_ONLY) {
      return (band || npiv < k) ? (status | ST_LAPACK_PATH) : status;      // (a rank-deficient Jc is flagged there too)
    }
    if (band_defer && (band || npiv < k) && !(status & ST_RANK_DEFICIENT)) return status | ST_LAPACK_PATH;
    if (npiv == k) return status;

    // ---- (7) slack columns become tangent-space coordinates
    if (k - npiv == 2 && F <= 1) {
      const uint8_t st2 = two_slack_pivots(Y, Ls, s, alpha, tol, gam, take, npiv, status, w_null);
      if (!(st2 & ST_DENSE_PATH)) return st2;     // (a degenerate plane falls through to the general routine)
    }
    if (k - npiv > 1 || F > 1) {
      if constexpr (defer_general) {   // hand sigma and gamma to the caller in w_null (N >= 2 n is the caller's condition)
        if (N >= 2 * n) {
          ATACOM_UNROLL
          for (int j = 0; j < n; ++j) {
            w_null[j < N ? j : 0] = sig[j];
            w_null[n + j < N ? n + j : 0] = gam[j];
          }
        }
        return (status & ST_RANK_DEFICIENT) | ST_DENSE_PATH;
      } else {
      // private copies: only these (not the register-resident operands of the hot path) get their address taken
      R s2[G1], sg2[n], gm2[n], al2[K1], wn2[N];
      ATACOM_UNROLL
      for (int i = 0; i < G; ++i) s2[i] = s[i];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) {
        sg2[j] = sig[j];
        gm2[j] = gam[j];
      }
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) al2[l] = alpha[l];
      const uint8_t st2 = null_part_general(Y, Ls, s2, sg2, gm2, al2, tol, wn2);
      ATACOM_UNROLL
      for (int i = 0; i < N; ++i) w_null[i] = wn2[i];
      return (status & ST_RANK_DEFICIENT) | st2;
      }
    }
    // remaining null direction in the free coordinates: tau supported on the unpivoted columns with
    // B_f tau = 0; its entries are x_j = sigma_j tau_j, z_diag,j = -gamma_j tau_j, z_dense,i = -(b_i . tau) / s_i.
    // The equality row of Y is B_f / L_00: same direction.
    int u1 = 0, u2 = 0;
    {
      int cnt = 0;
      ATACOM_UNROLL
      for (int c = 0; c < n; ++c) {
        u1 = (!take[c] && cnt == 0) ? c : u1;
        u2 = (!take[c] && cnt == 1) ? c : u2;
        cnt += take[c] ? 0 : 1;
      }
    }
    // tau = (t1 at column u1, t2 at column u2) with B_f tau = 0 (F = 1), or the unit vector of u1 (F = 0)
    const R t1 = (F == 1) ? Y.get_dyn(u2) : R(1);
    const R t2 = (F == 1) ? -Y.get_dyn(u1) : R(0);
    R e[G1], yt[M1];
    R nrm2 = t1 * t1 + t2 * t2;
    ATACOM_UNROLL
    for (int l = 0; l < m; ++l) {   // B = L Y  =>  b_i . tau = sum_{l <= i} L[i][l] (Y[l] . tau)
      R acc = Y.get_dyn(l * n + u1) * t1;
      if (F == 1) acc += Y.get_dyn(l * n + u2) * t2;
      yt[l] = acc;
    }
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      if (i < GD) {
        const int ri = F + (i < GD ? i : 0);
        const R lii = Ls.get(LI0 + ri);
        R an = (lii > R(0)) ? yt[ri] * dual_rsqrt_wide(lii * lii) : R(0);   // L[ri][ri] = 1 / li
        ATACOM_UNROLL
        for (int l = 0; l < ri; ++l) an += Ls.get(lidx(ri, l)) * yt[l];
        const R s2 = s[i] * s[i];
        const R rs = dual_rsqrt(s2 > R(DUAL_TINY) ? s2 : R(DUAL_TINY));   // 1 / |s_i|, s_i = 0 -> the limit
        e[i] = (s[i] < R(0) ? an : -an) * rs;
        nrm2 += e[i] * e[i];
      } else {
        const int j = i >= GD ? i - GD : 0;
        e[i] = -gam[j] * ((j == u1) ? t1 : ((F == 1 && j == u2) ? t2 : R(0)));   // sigma^2 + gamma^2 = 1: already in nrm2
      }
    }
    const R rn = dual_rsqrt_wide(nrm2);
    int ip = G;
    ATACOM_UNROLL
    for (int i = G - 1; i >= 0; --i) {
      e[i] *= rn;
      ip = (num<R>::abs(e[i]) > tol) ? i : ip;
    }
    if (ip == G) return status | ST_RANK_DEFICIENT;
    status |= ST_SLACK_PIVOT;
    if (ip > 0) status |= ST_COLUMN_DROPPED;
    R ep = R(0), zp = R(0), al = R(0);
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      ep = (i == ip) ? e[i] : ep;
      zp = (i == ip) ? w_null[n + i] : zp;
    }
    ATACOM_UNROLL
    for (int l = 0; l < k; ++l) al = (l == npiv) ? alpha[l] : al;
    const R ri = dual_rsqrt_wide(ep * ep);
    const R bl = (al - zp) * (ep * ri * ri);   // (al - zp) / ep: beta of the new row times the sign of its pivot entry
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      const R v = w_null[n + i];
      w_null[n + i] = (i == ip) ? al : (i > ip ? v + bl * e[i] : v);
    }
    return status;
  }
};

}  // namespace atacom
