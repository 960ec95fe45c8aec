// The reference's null-space basis, restated: what scipy.linalg.svd (LAPACK gesdd) hands to
// atacom/utils/null_space_coordinate.py:8-26, followed by the reference's tolerance-RREF (:40-79): its decisions and
// its result, evaluated lazily (see the comment at the rref below).
//
// For a full-row-rank C x N matrix (C < N) gesdd never touches the null vectors with its SVD iteration: with
// JOBZ = 'A' the trailing N - C rows of VT are [0 I] P^T, where P = G_0 G_1 ... G_{C-1} is the product of the RIGHT
// Householder reflectors of
//   * the bidiagonalisation dgebd2 (lower bidiagonal form: G_i annihilates row i right of the diagonal, then the left
//     reflector H_i annihilates column i below the subdiagonal) — LAPACK's path 5t, taken while N < int(11 C / 6);
//   * the LQ factorisation dgelq2 (no left reflectors) — path 4t, N >= int(11 C / 6).
// So the "implementation-defined" SVD null basis is a deterministic function of Jc: the last k = N - C columns of P,
// with dlarfg's sign convention beta = -sign(alpha) |[alpha; x]|.  oracle/nullspace.py:lapack_null_basis is the NumPy
// restatement; tests/test_oracle.py checks it against SciPy's vectors (1e-12, signs included).  With this basis the
// reference's output is reproduced on BOTH strata — also where the tolerance branch of rref fires and the result
// depends on the basis.
//
// Used for the environments the fast dual path (atacom_dual.cuh) flags: a column can only fire in the reference if
// the canonical pivot is at most sqrt(k - r) tol (r pivots so far), see Dual::project; everything outside that
// band is basis-free and stays on the fast path.  Also the whole projection of the generic / point-reach kernels.
//
// Storage: ONE C x N array (`S`, any store with get / set at a run-time index: a column of a shared-memory array on
// the device).  It holds Jc, then the reflector vectors (v_i right of the diagonal of row i, u_i below the
// subdiagonal of column i, the bidiagonal in between), then — in cells that have died by then — the null basis Z
// (N x k), which the rref only reads, and the N entries of the result.  Cost ~ 4 C^2 (N - C/3) + 4 N C k flops:
// ~10 kFLOP for 12 x 17.
#pragma once

#include "atacom_core.cuh"

#if defined(__CUDA_ARCH__) && ATACOM_LAPACK_SWEEP_UNROLL == 2
#define ATACOM_SWEEP_LOOP _Pragma("unroll 2")
#else
#define ATACOM_SWEEP_LOOP ATACOM_ROLLED
#endif

namespace atacom {

constexpr uint8_t ST_LAPACK_PATH = 32;   // the fast path left this environment to the LAPACK-basis routine

// Plain array store with run-time indexing (host build, small shapes in local memory).
template <typename R, int SIZE>
struct ArrayStore {
  R v[SIZE > 0 ? SIZE : 1];
  ATACOM_HD R get(int i) const { return v[i]; }
  ATACOM_HD void set(int i, R x) { v[i] = x; }
};

// The lanes working on one environment: the host build and the generic kernels use one (SoloGroup); the fix-up kernel
// of the step families uses LPE lanes of a warp (WarpLanes, atacom_kernels.cu).
struct SoloGroup {
  ATACOM_HD int sub() const { return 0; }
  ATACOM_HD void sync() const {}
  ATACOM_HD int write_lane() const { return 0; }      // which rows of Jc this lane fills in: i % LPE == write_lane(); < 0: all
  template <int LPE, typename R> ATACOM_HD R sum(R x) const { return x; }     // over the lanes of the group
};

template <typename R, class D>
struct Lapack {
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  static constexpr int C1 = at_least_1<C>::value, K1 = at_least_1<k>::value;
  static constexpr bool LQ_PATH = N >= (C * 11) / 6;            // gesdd: MNTHR = int(minmn * 11 / 6)
#ifndef ATACOM_LAPACK_NACC
#define ATACOM_LAPACK_NACC 2      // (4 measured equal: 52.1 against 52.0 us per step)
#endif
#ifndef ATACOM_LAPACK_SWEEP_UNROLL
#define ATACOM_LAPACK_SWEEP_UNROLL 1
#endif
  static constexpr int NACC = ATACOM_LAPACK_NACC;               // partial sums per dot product
  // Z (N x k) lives in dead cells of the C x N array when the shape allows it (see zcell), else behind it
  static constexpr bool Z_OVERLAY = (C >= 2 * k - 1) && (N >= 2 * k) && (k > 0);
  // w_null is produced column by column at run-time indices: N more cells, in columns k and k + 1 of the array (dead
  // by then, and clear of Z) when the shape allows it, else behind everything
  static constexpr bool W_OVERLAY = Z_OVERLAY && (N >= 2 * k + 2) && (N - C <= C);
  static constexpr int SIZE = C * N + (Z_OVERLAY ? 0 : N * k) + (W_OVERLAY ? 0 : N);
  static constexpr ATACOM_HD int wcell(int j) {
    return W_OVERLAY ? (j < C ? j * N + k : (j - C) * N + k + 1) : C * N + (Z_OVERLAY ? 0 : N * k) + j;
  }
  static constexpr ATACOM_HD int a(int i, int j) { return i * N + j; }
  // row j of Z.  Overlay: rows j < C take the last k cells of row j of the array (v_j lives there until reflector j
  // has been applied, which is when row j of Z comes into being); the identity rows j >= C take the first k cells
  // of row j - k, below the diagonal (left-reflector / L entries, dead after the minimum-norm solve).
  static constexpr ATACOM_HD int zcell(int j, int c) {
    return Z_OVERLAY ? (j < C ? j * N + (N - k) + c : (j - k) * N + c) : C * N + j * k + c;
  }

  // dlarfg: reflector I - tau [1; v][1; v]^T mapping [alpha; x] to [beta; 0].  x is given by its squared norm;
  // returns beta and tau, `scale` = 1 / (alpha - beta) turns x into v (0 when x = 0: the identity, as LAPACK).
  // With beta = -sign(alpha) nrm:  tau = (beta - alpha) / beta = 1 + |alpha| / nrm  and  1 / (alpha - beta) =
  // sign(alpha) / (|alpha| + nrm) — on the device one reciprocal square root and one reciprocal, each a MUFU seed plus
  // Newton steps in double (~25 instructions for the whole reflector against ~70 with sqrt and two divisions, on the
  // critical path of every one of the 2 C - 1 reflectors); results agree with the division form to 2 ulp.
  // 1 / x: MUFU.RCP64H seed (~2^-23) and two Newton steps on the device (the argument is a pivot: |x| > tol)
  static ATACOM_HD R rcp(R x) {
#if defined(__CUDA_ARCH__)
    double z;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(static_cast<double>(x)));
    z = ::fma(z, ::fma(-static_cast<double>(x), z, 1.0), z);
    z = ::fma(z, ::fma(-static_cast<double>(x), z, 1.0), z);
    return static_cast<R>(z);
#else
    return R(1) / x;
#endif
  }

  // (the division form, as LAPACK writes it: the host build, and out of line on the device for arguments outside the
  // range of the MUFU seeds — never met with this data)
  struct Reflector { R beta, tau, scale; };
#if defined(__CUDA_ARCH__)
  static __device__ __noinline__ Reflector larfg_exact(R alpha, R ss) {
#else
  static inline Reflector larfg_exact(R alpha, R ss) {
#endif
    const R nrm = num<R>::sqrt(ss);
    const R b = -::copysign(nrm, alpha);      // Fortran SIGN: the sign of alpha, negative zero included
    return {b, (b - alpha) / b, R(1) / (alpha - b)};
  }
  static ATACOM_HD void larfg(R alpha, R xn2, R* beta, R* tau, R* scale) {
    if (!(xn2 > R(0))) {
      *beta = alpha;
      *tau = R(0);
      *scale = R(0);
      return;
    }
    const R ss = alpha * alpha + xn2;
#if defined(__CUDA_ARCH__)
    if (ss > R(1e-280) && ss < R(1e280)) {       // (always, for this data; the seeds flush denormals)
      double y;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(static_cast<double>(ss)));
      const double e = ::fma(-static_cast<double>(ss) * y, y, 1.0);         // ~ 1e-6
      y = ::fma(y, ::fma(0.375, e, 0.5) * e, y);                            // error O(e^3)
      const double e2 = ::fma(-static_cast<double>(ss) * y, y, 1.0);
      const double rs = ::fma(y, 0.5 * e2, y);                              // 1 / sqrt(ss)
      double nrm = static_cast<double>(ss) * rs;
      nrm = ::fma(::fma(-nrm, nrm, static_cast<double>(ss)), 0.5 * rs, nrm);
      const double aa = ::fabs(static_cast<double>(alpha));
      const double den = aa + nrm;
      double z;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(den));
      z = ::fma(z, ::fma(-den, z, 1.0), z);                                 // seed ~2^-23: two Newton steps
      z = ::fma(z, ::fma(-den, z, 1.0), z);
      *beta = static_cast<R>(-::copysign(nrm, static_cast<double>(alpha)));   // Fortran SIGN: negative zero included
      *tau = static_cast<R>(::fma(aa, rs, 1.0));
      *scale = static_cast<R>(::copysign(z, static_cast<double>(alpha)));
      return;
    }
#endif
    const Reflector h = larfg_exact(alpha, ss);
    *beta = h.beta;
    *tau = h.tau;
    *scale = h.scale;
  }

  // S: on entry Jc (C x N, row-major, S.get(i * N + j)); destroyed.  r: C (right-hand side psi + K_c c; destroyed),
  // alpha: k.  w_mn = -Jc^+ r, w_null = Nc alpha (N each; valid in lane 0 of the group).  Returns status bits.
  // WITH_MN = false: the null part only — the minimum-norm part -Jc^+ r does not depend on the basis and the caller
  // has it already (the fix-up kernel takes it from the dual path of the step kernel): r and w_mn are not touched, the
  // left reflectors are not applied to a right-hand side, nothing of the bidiagonal factor is kept (tau_i takes the
  // diagonal cell), and one vector less is swept back.
  //
  // LPE lanes of one warp work on one environment (LPE = 1: one thread, the host build).  The rows the right
  // reflector G_i updates are independent of each other, and so are the columns the left reflector H_i updates: the
  // lanes of a group take them round robin (lane `sub` of LPE), every lane forms the reflector itself from the row /
  // column it needs, and a __syncwarp separates the phases (`G.sync()`, over the lanes of the warp that take part).
  // The null basis Z is formed column-wise in registers, one or two columns per lane, and the elimination steps of the
  // rref are shared column-wise as well.  What bounds the routine on the device is the length of one environment's
  // dependent instruction chain.
  template <int LPE = 1, bool WITH_MN = true, class ST, class GRP = SoloGroup>
  static ATACOM_HD uint8_t project(ST& S, R* r, const R* alpha, R tol, bool want_null, R* w_mn, R* w_null,
                                   const GRP& Grp = GRP()) {
    const int sub = Grp.sub();
    uint8_t status = 0;
    R tau_s[WITH_MN ? C1 : 1];
    R bmax = R(0), bmin = R(1e300);                // largest and smallest |diagonal entry| of the bidiagonal / L factor

    // ---- reflectors (dgebd2 / dgelq2).
    // (a) Null part only, bidiagonal path: ONE pass over the trailing rows per reflector pair.  What bounds the sweeps
    // on the device is shared-memory bandwidth — every trailing entry read and written once per reflector — so the
    // left reflector H_i is not applied in a column sweep of its own: a lane that has a row in registers applies the
    // PENDING left reflector H_{i-1} to it (row_l -= u_l y^T, u_l = A[l][i-1] scq the row's own entry of column i-1),
    // then G_i, and adds the row's share to what H_i needs — P_j = sum_l A[l][i] A[l][j] over the rows below the
    // subdiagonal and the squared norm of column i; the column dot products of H_i are y_j = tauq (A[i+1][j] +
    // scq P_j), a scaling away.  The lanes' partial sums meet in a butterfly (Grp.sum), row i+1 is read by every lane,
    // H_i and, from that row with H_i applied, G_{i+1} are formed by every lane alike.  Half the shared-memory
    // traffic of (b), the same arithmetic up to the order of two roundings.
    constexpr bool FUSED_PAIRS = !WITH_MN && !LQ_PATH && (LPE == 1 || LPE == 2 || LPE == 4) && C >= 2;
    if constexpr (FUSED_PAIRS) {
      R v[N], y[N];
      R tau, scq_prev = R(0);
      {
        Grp.sync();
        R xs[NACC];
        ATACOM_UNROLL
        for (int t = 0; t < NACC; ++t) xs[t] = R(0);
        const R x0 = S.get(a(0, 0));
        ATACOM_UNROLL
        for (int j = 1; j < N; ++j) {
          v[j] = S.get(a(0, j));
          xs[(j - 1) % NACC] += v[j] * v[j];
        }
        R xn2 = xs[0];
        ATACOM_UNROLL
        for (int t = 1; t < NACC; ++t) xn2 += xs[t];
        R beta, sc;
        larfg(x0, xn2, &beta, &tau, &sc);
        const R ab = num<R>::abs(beta);
        bmax = ab > bmax ? ab : bmax;
        bmin = ab < bmin ? ab : bmin;
        ATACOM_UNROLL
        for (int j = 1; j < N; ++j) v[j] *= sc;
        Grp.sync();                                // every lane has read row 0
        if (sub == 0) {
          S.set(a(0, 0), tau);
          ATACOM_UNROLL
          for (int j = 1; j < N; ++j) S.set(a(0, j), v[j]);
        }
      }
      ATACOM_UNROLL
      for (int i = 0; i < C - 1; ++i) {
        R Pj[N];
        R cn2 = R(0);
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) Pj[j] = R(0);
        ATACOM_ROLLED
        for (int l = i + 1 + sub; l < C; l += LPE) {
          R row[N];
          ATACOM_UNROLL
          for (int j = i; j < N; ++j) row[j] = S.get(a(l, j));
          if (i >= 1) {                            // H_{i-1}, pending from the step before
            const R ul = S.get(a(l, i >= 1 ? i - 1 : 0)) * scq_prev;
            ATACOM_UNROLL
            for (int j = i; j < N; ++j) row[j] -= ul * y[j];
          }
          R ws[NACC];                              // G_i
          ATACOM_UNROLL
          for (int t = 0; t < NACC; ++t) ws[t] = (t == 0) ? row[i] : R(0);
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) ws[(j - i) % NACC] += row[j] * v[j];
          R wsum = ws[0];
          ATACOM_UNROLL
          for (int t = 1; t < NACC; ++t) wsum += ws[t];
          const R w = wsum * tau;
          row[i] -= w;
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) row[j] -= w * v[j];
          if (l != i + 1) {                        // below the subdiagonal: the row's share of H_i
            cn2 += row[i] * row[i];
            ATACOM_UNROLL
            for (int j = i + 1; j < N; ++j) Pj[j] += row[i] * row[j];
          }
          ATACOM_UNROLL
          for (int j = i; j < N; ++j) S.set(a(l, j), row[j]);
        }
        Grp.sync();                                // the rows are final up to H_i
        cn2 = Grp.template sum<LPE>(cn2);
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) Pj[j] = Grp.template sum<LPE>(Pj[j]);
        R pr[N];                                   // row i + 1: the row H_i pivots on, then the one G_{i+1} is made of
        ATACOM_UNROLL
        for (int j = i; j < N; ++j) pr[j] = S.get(a(i + 1, j));
        R betaq, tauq, scq;
        larfg(pr[i], cn2, &betaq, &tauq, &scq);
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) {
          y[j] = (pr[j] + scq * Pj[j]) * tauq;
          pr[j] -= y[j];
        }
        R xs[NACC];
        ATACOM_UNROLL
        for (int t = 0; t < NACC; ++t) xs[t] = R(0);
        ATACOM_UNROLL
        for (int j = i + 2; j < N; ++j) xs[(j - i) % NACC] += pr[j] * pr[j];
        R xn2 = xs[0];
        ATACOM_UNROLL
        for (int t = 1; t < NACC; ++t) xn2 += xs[t];
        R beta, sc;
        larfg(pr[i + 1], xn2, &beta, &tau, &sc);
        const R ab = num<R>::abs(beta);
        bmax = ab > bmax ? ab : bmax;
        bmin = ab < bmin ? ab : bmin;
        ATACOM_UNROLL
        for (int j = i + 2; j < N; ++j) v[j] = pr[j] * sc;
        scq_prev = scq;
        Grp.sync();                                // every lane has read row i + 1
        if (sub == 0) {
          S.set(a(i + 1, i + 1), tau);
          ATACOM_UNROLL
          for (int j = i + 2; j < N; ++j) S.set(a(i + 1, j), v[j]);
        }
      }
    }
    // (b) Everything else: left reflectors applied in a sweep over the columns, and to the right-hand side as they are
    // formed.  The loop over i is unrolled: every range below is static, nothing is predicated.
    ATACOM_UNROLL
    for (int i = 0; i < (FUSED_PAIRS ? 0 : C); ++i) {
      Grp.sync();                                  // row i is final: the previous reflectors have been applied to it
      R v[N];                                      // row i right of the diagonal, then the reflector vector
      R xs[NACC];                                  // NACC partial sums: the dependent chains are what bounds a warp here
      ATACOM_UNROLL
      for (int t = 0; t < NACC; ++t) xs[t] = R(0);
      const R x0 = S.get(a(i, i));
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) {
        v[j] = S.get(a(i, j));
        xs[(j - i - 1) % NACC] += v[j] * v[j];
      }
      R xn2 = xs[0];
      ATACOM_UNROLL
      for (int t = 1; t < NACC; ++t) xn2 += xs[t];
      R beta, tau, sc;
      larfg(x0, xn2, &beta, &tau, &sc);
      {
        const R ab = num<R>::abs(beta);
        bmax = ab > bmax ? ab : bmax;
        bmin = ab < bmin ? ab : bmin;
      }
      if (WITH_MN) tau_s[WITH_MN ? i : 0] = tau;
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) v[j] *= sc;
      if (sub == 0) {                              // (tau = 0, the identity: x = 0 = v, every update below is a no-op)
        S.set(a(i, i), WITH_MN ? beta : tau);
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) S.set(a(i, j), v[j]);
      }
      {
        ATACOM_SWEEP_LOOP
        for (int l = i + 1 + sub; l < C; l += LPE) {      // rows below: A <- A G_i, one load and one store per entry
          R row[N];
          ATACOM_UNROLL
          for (int j = i; j < N; ++j) row[j] = S.get(a(l, j));
          R ws[NACC];
          ATACOM_UNROLL
          for (int t = 0; t < NACC; ++t) ws[t] = (t == 0) ? row[i] : R(0);
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) ws[(j - i) % NACC] += row[j] * v[j];
          R wsum = ws[0];
          ATACOM_UNROLL
          for (int t = 1; t < NACC; ++t) wsum += ws[t];
          const R w = wsum * tau;
          S.set(a(l, i), row[i] - w);
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) S.set(a(l, j), row[j] - w * v[j]);
        }
      }
      if (!LQ_PATH && i + 1 < C) {
        Grp.sync();                                // column i is final
        R u[C1];                                   // column i below the subdiagonal, then the reflector vector
        R us[NACC];
        ATACOM_UNROLL
        for (int t = 0; t < NACC; ++t) us[t] = R(0);
        const R c0 = S.get(a(i + 1 < C ? i + 1 : 0, i));
        ATACOM_UNROLL
        for (int l = i + 2; l < C; ++l) {
          u[l] = S.get(a(l, i));
          us[(l - i) % NACC] += u[l] * u[l];
        }
        R un2 = us[0];
        ATACOM_UNROLL
        for (int t = 1; t < NACC; ++t) un2 += us[t];
        R betaq, tauq, scq;
        larfg(c0, un2, &betaq, &tauq, &scq);
        if (WITH_MN && sub == 0) S.set(a(i + 1 < C ? i + 1 : 0, i), betaq);     // (u itself is not needed again: it is applied to r here)
        {
          ATACOM_UNROLL
          for (int l = i + 2; l < C; ++l) u[l] *= scq;
          ATACOM_SWEEP_LOOP
          for (int j = i + 1 + sub; j < N; j += LPE) {    // columns to the right: A <- H_i A
            R col[C1];
            ATACOM_UNROLL
            for (int l = i + 1; l < C; ++l) col[l] = S.get(a(l, j));
            R ws[NACC];
            ATACOM_UNROLL
            for (int t = 0; t < NACC; ++t) ws[t] = (t == 0) ? col[i + 1 < C ? i + 1 : 0] : R(0);
            ATACOM_UNROLL
            for (int l = i + 2; l < C; ++l) ws[(l - i - 1) % NACC] += col[l] * u[l];
            R wsum = ws[0];
            ATACOM_UNROLL
            for (int t = 1; t < NACC; ++t) wsum += ws[t];
            const R w = wsum * tauq;
            S.set(a(i + 1 < C ? i + 1 : 0, j), col[i + 1 < C ? i + 1 : 0] - w);
            ATACOM_UNROLL
            for (int l = i + 2; l < C; ++l) S.set(a(l, j), col[l] - w * u[l]);
          }
          if (WITH_MN) {
            R w = r[i + 1 < C ? i + 1 : 0];        // and the right-hand side (every lane keeps its own copy)
            ATACOM_UNROLL
            for (int l = i + 2; l < C; ++l) w += u[l] * r[l];
            w *= tauq;
            r[i + 1 < C ? i + 1 : 0] -= w;
            ATACOM_UNROLL
            for (int l = i + 2; l < C; ++l) r[l] -= w * u[l];
          }
        }
      }
    }
    Grp.sync();
    // a dependent row of Jc shows as a (numerically) zero diagonal entry of the factor: |beta_i| is the norm of what
    // the earlier reflectors left of row i
    const R rank_floor = R(64) * num<R>::eps() * bmax;
    if (!WITH_MN && !(bmin > rank_floor)) status |= ST_RANK_DEFICIENT;

    // ---- minimum-norm part: Jc = U B P^T (B lower bidiagonal) or L Q;  x = P [y; 0],  B y = -U^T r  /  L y = -r
    R t[WITH_MN ? N : 1];
    if (WITH_MN) {
      ATACOM_UNROLL
      for (int i = 0; i < N; ++i) t[WITH_MN ? i : 0] = R(0);
      ATACOM_UNROLL
      for (int i = 0; i < C; ++i) {
        R acc = -r[i];
        if (LQ_PATH) {
          ATACOM_UNROLL
          for (int j = 0; j < i; ++j) acc -= S.get(a(i, j)) * t[WITH_MN ? j : 0];
        } else if (i > 0) {
          acc -= S.get(a(i, i > 0 ? i - 1 : 0)) * t[WITH_MN && i > 0 ? i - 1 : 0];
        }
        const R d = S.get(a(i, i));
        if (num<R>::abs(d) > rank_floor) {
          t[WITH_MN ? i : 0] = acc / d;
        } else {                                   // dependent row of Jc: dropped from the solve, flagged
          status |= ST_RANK_DEFICIENT;
          t[WITH_MN ? i : 0] = R(0);
        }
      }
    }
    // P = G_0 ... G_{C-1} applied to the vectors in one sweep over the reflectors (every index static).  WITH_MN:
    // vector 0 is [y; 0], which becomes the minimum-norm solution.  The others are the unit vectors e_C .. e_{N-1},
    // which become the columns of Z — the null basis as gesdd returns it (rows C.. of VT).  The lanes take the vectors
    // round robin.
    const bool null_part = want_null && k > 0;
    constexpr int MN = WITH_MN ? 1 : 0;            // index of the first null vector
    constexpr int CPL = (k + MN + LPE - 1) / LPE;  // vectors per lane
    constexpr int CPL1 = CPL > 0 ? CPL : 1;
    R X[CPL1][N];
    ATACOM_UNROLL
    for (int cc = 0; cc < CPL; ++cc) {
      ATACOM_UNROLL
      for (int j = 0; j < N; ++j) {
        const R unit = (j - C == sub + cc * LPE - MN) ? R(1) : R(0);
        X[cc][j] = (WITH_MN && sub + cc * LPE == 0) ? t[WITH_MN ? j : 0] : unit;
      }
    }
    ATACOM_UNROLL
    for (int i = C - 1; i >= 0; --i) {
      R vi[N];
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) vi[j] = S.get(a(i, j));
      const R taui = WITH_MN ? tau_s[WITH_MN ? i : 0] : S.get(a(i, i));
      ATACOM_UNROLL
      for (int cc = 0; cc < CPL; ++cc) {
        R wa = X[cc][i], wb = R(0);
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) {
          if ((j - i) & 1) wb += vi[j] * X[cc][j];
          else wa += vi[j] * X[cc][j];
        }
        R w = (wa + wb) * taui;
        X[cc][i] -= w;
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) X[cc][j] -= w * vi[j];
      }
    }
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      if (WITH_MN) w_mn[i] = X[0][i];              // (the minimum-norm part: lane 0's first vector)
      w_null[i] = R(0);
    }
    if (!null_part) return status;
    Grp.sync();                                    // every lane has read the last reflector: its cells may be reused
    ATACOM_UNROLL
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = sub + cc * LPE - MN;
      if (c >= 0 && c < k) {
        ATACOM_UNROLL
        for (int j = 0; j < N; ++j) S.set(zcell(j, c), X[cc][j]);
      }
    }
    Grp.sync();

    // ---- the reference's rref on V = Z^T (k x N), null_space_coordinate.py:40-79, evaluated lazily.  The procedure
    // walks the columns; the pivot candidate is the first largest |.| among the rows not used yet; <= tol: zero those
    // entries and move on; else swap, scale the pivot row, eliminate the column from every other row.  All of that are
    // row operations, so the current column j is E z_j — z_j the ORIGINAL column (row j of Z), E the k x k product of
    // the row operations so far — and the matrix itself is never rewritten: a pivot step updates E alone (k x k
    // against an elimination sweep over every remaining column in shared memory).  What the finished matrix holds in
    // column j: the unit vector of its pivot row (a pivot column); its state when it was dropped, the rows not yet
    // used zeroed (later operations only combine rows that are zero there); E_final z_j for the columns behind the
    // last pivot.  So w_null[j] = sum_l alpha_l R[l][j] is alpha_r for the r-th pivot column and g . z_j otherwise,
    // g = sum over the rows used so far of alpha_l E[l].
    // One stage per pivot, unrolled: in stage r only the rows r.. are candidates, so a column costs (k - r) dot
    // products, and the stages that walk through many dropped columns (the last pivot of an environment with an
    // active constraint is a slack column) are the cheap ones.  Every lane of a group runs this redundantly; lane 0
    // notes the per-column results in shared memory (w_null is indexed at run time), one synchronisation at the end.
    R E[K1][K1], g[K1];
    ATACOM_UNROLL
    for (int i = 0; i < k; ++i) {
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) E[i][m] = (i == m) ? R(1) : R(0);
      g[i] = R(0);
    }
    // Within a stage the walk over the columns only LOOKS for the pivot (a dropped column costs its dot products and
    // one store); the pivot step itself comes after the loop, where the groups of a warp — which have dropped
    // different numbers of columns on the way — are together again and execute it once.
    int j = 0;
    ATACOM_UNROLL
    for (int rr = 0; rr < k; ++rr) {
      bool found = false;
      R z[K1], c[K1];
      R gz = R(0);
      int kk = rr;
      ATACOM_ROLLED
      while (j < N) {
        ATACOM_UNROLL
        for (int m = 0; m < k; ++m) z[m] = S.get(zcell(j, m));
        gz = g[0] * z[0];
        ATACOM_UNROLL
        for (int m = 1; m < k; ++m) gz += g[m] * z[m];
        R p = R(-1);
        kk = rr;
        ATACOM_UNROLL
        for (int i = rr; i < k; ++i) {
          R acc = E[i][0] * z[0];
          ATACOM_UNROLL
          for (int m = 1; m < k; ++m) acc += E[i][m] * z[m];
          c[i] = acc;
          const R av = num<R>::abs(acc);
          if (av > p) {                            // (strict: the first of equal candidates, as np.argmax)
            p = av;
            kk = i;
          }
        }
        if (p > tol) {
          found = true;
          break;
        }
        status |= ST_COLUMN_DROPPED;
        if (sub == 0) S.set(wcell(j), gz);
        ++j;
      }
      if (!found) {
        status |= ST_RANK_DEFICIENT;               // ran out of columns (the later stages find j == N as well)
        continue;
      }
      if (j >= n) status |= ST_SLACK_PIVOT;
      if (sub == 0) S.set(wcell(j), alpha[rr]);
      // rows rr and kk change places; the pivot row is scaled
      R ckk = c[rr];
      ATACOM_UNROLL
      for (int i = rr + 1; i < k; ++i) {
        const R ci = c[i];
        c[i] = (i == kk) ? c[rr] : ci;
        ckk = (i == kk) ? ci : ckk;
      }
      const R inv = rcp(ckk);
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) {
        R e = E[rr][m];
        ATACOM_UNROLL
        for (int i = rr + 1; i < k; ++i) {
          const R ei = E[i][m];
          E[i][m] = (i == kk) ? E[rr][m] : ei;
          e = (i == kk) ? ei : e;
        }
        E[rr][m] = e * inv;
      }
      // ... and is eliminated from the rows still to come and from the rows used before
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        if (i == rr) continue;
        R ci;
        if (i > rr) {
          ci = c[i];
        } else {
          ci = E[i][0] * z[0];
          ATACOM_UNROLL
          for (int m = 1; m < k; ++m) ci += E[i][m] * z[m];
        }
        ATACOM_UNROLL
        for (int m = 0; m < k; ++m) E[i][m] -= ci * E[rr][m];
      }
      const R dg = alpha[rr] - gz;                 // g = sum_{l <= rr} alpha_l E[l] of the updated rows
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) g[m] += dg * E[rr][m];
      ++j;
    }
    Grp.sync();                                    // lane 0's per-column results are visible to the group
    ATACOM_UNROLL
    for (int jj = 0; jj < N; ++jj) {
      if (jj < j) {
        w_null[jj] = S.get(wcell(jj));
      } else {                                     // behind the last pivot
        R acc = R(0);
        ATACOM_UNROLL
        for (int m = 0; m < k; ++m) acc += g[m] * S.get(zcell(jj, m));
        w_null[jj] = acc;
      }
    }
    return status;
  }
};

}  // namespace atacom
