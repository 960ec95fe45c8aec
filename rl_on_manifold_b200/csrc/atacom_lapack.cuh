// The reference's null-space basis, restated: what scipy.linalg.svd (LAPACK gesdd) hands to
// atacom/utils/null_space_coordinate.py:8-26, followed by the reference's tolerance-RREF (:40-79) exactly as written.
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
// (N x k) on which the RREF runs in place.  Cost ~ 4 C^2 (N - C/3) + 4 N C k flops: ~10 kFLOP for 12 x 17.
#pragma once

#include "atacom_core.cuh"

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
  ATACOM_HD bool all(bool p) const { return p; }      // true when p holds for every environment that syncs along
};

template <typename R, class D>
struct Lapack {
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  static constexpr int C1 = at_least_1<C>::value, K1 = at_least_1<k>::value;
  static constexpr bool LQ_PATH = N >= (C * 11) / 6;            // gesdd: MNTHR = int(minmn * 11 / 6)
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

    // ---- reflectors (dgebd2 / dgelq2), left reflectors applied to the right-hand side as they are formed.
    // The loop over i is unrolled: every range below is static, nothing is predicated.
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) {
      Grp.sync();                                  // row i is final: the previous reflectors have been applied to it
      R v[N];                                      // row i right of the diagonal, then the reflector vector
      R xn2 = R(0), xn2b = R(0);
      const R x0 = S.get(a(i, i));
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) {
        v[j] = S.get(a(i, j));
        if ((j - i) & 1) xn2 += v[j] * v[j];
        else xn2b += v[j] * v[j];
      }
      xn2 += xn2b;
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
        ATACOM_ROLLED
        for (int l = i + 1 + sub; l < C; l += LPE) {      // rows below: A <- A G_i, one load and one store per entry
          R row[N];
          ATACOM_UNROLL
          for (int j = i; j < N; ++j) row[j] = S.get(a(l, j));
          R w0 = row[i], w1 = R(0);
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) {
            if ((j - i) & 1) w1 += row[j] * v[j];
            else w0 += row[j] * v[j];
          }
          const R w = (w0 + w1) * tau;
          S.set(a(l, i), row[i] - w);
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) S.set(a(l, j), row[j] - w * v[j]);
        }
      }
      if (!LQ_PATH && i + 1 < C) {
        Grp.sync();                                // column i is final
        R u[C1];                                   // column i below the subdiagonal, then the reflector vector
        R un2 = R(0);
        const R c0 = S.get(a(i + 1 < C ? i + 1 : 0, i));
        ATACOM_UNROLL
        for (int l = i + 2; l < C; ++l) {
          u[l] = S.get(a(l, i));
          un2 += u[l] * u[l];
        }
        R betaq, tauq, scq;
        larfg(c0, un2, &betaq, &tauq, &scq);
        if (WITH_MN && sub == 0) S.set(a(i + 1 < C ? i + 1 : 0, i), betaq);     // (u itself is not needed again: it is applied to r here)
        {
          ATACOM_UNROLL
          for (int l = i + 2; l < C; ++l) u[l] *= scq;
          ATACOM_ROLLED
          for (int j = i + 1 + sub; j < N; j += LPE) {    // columns to the right: A <- H_i A
            R col[C1];
            ATACOM_UNROLL
            for (int l = i + 1; l < C; ++l) col[l] = S.get(a(l, j));
            R w0 = col[i + 1 < C ? i + 1 : 0], w1 = R(0);
            ATACOM_UNROLL
            for (int l = i + 2; l < C; ++l) {
              if ((l - i) & 1) w1 += col[l] * u[l];
              else w0 += col[l] * u[l];
            }
            const R w = (w0 + w1) * tauq;
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
        R w = X[cc][i];
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) w += vi[j] * X[cc][j];
        w *= taui;
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
    // the row operations so far — and the matrix itself is never rewritten: a pivot step is a rank-one update of E
    // (25 fused multiply-adds for k = 5 against an elimination sweep over every remaining column in shared memory).
    // What the finished matrix holds in column j: the unit vector of its pivot row (a pivot column); its state when
    // it was dropped, the rows not yet used zeroed (later operations only combine rows that are zero there);
    // E_final z_j for the columns behind the last pivot.  w_null[j] = sum_l alpha_l R[l][j] follows column by
    // column.  Rows are tracked physically: `pos[i]` is the position the reference's swaps have given row i (it
    // decides which of two equal candidates is "first"), `coef[i]` the alpha of the pivot the row carries (0: unused).
    // Every lane of a group runs this redundantly — no shared-memory writes but its own results, no synchronisation.
    R E[K1][K1], coef[K1];
    int pos[K1];
    ATACOM_UNROLL
    for (int i = 0; i < k; ++i) {
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) E[i][m] = (i == m) ? R(1) : R(0);
      coef[i] = R(0);
      pos[i] = i;
    }
    unsigned used = 0u;
    int rr = 0, j = 0;
    ATACOM_ROLLED
    for (; j < N && rr < k; ++j) {
      R z[K1], c[K1];
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) z[m] = S.get(zcell(j, m));
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        R acc = E[i][0] * z[0];
        ATACOM_UNROLL
        for (int m = 1; m < k; ++m) acc += E[i][m] * z[m];
        c[i] = acc;
      }
      R p = R(-1), piv = R(1), wdrop = R(0);
      int pk = 0, ppos = k;
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        const bool un = ((used >> i) & 1u) == 0u;
        const R av = num<R>::abs(c[i]);
        const bool better = un && (av > p || (av == p && pos[i] < ppos));
        p = better ? av : p;
        piv = better ? c[i] : piv;
        pk = better ? i : pk;
        ppos = better ? pos[i] : ppos;
        wdrop += coef[i] * c[i];                   // (what a dropped column keeps: the rows used so far)
      }
      const bool pivot = p > tol;
      status |= pivot ? (j >= n ? ST_SLACK_PIVOT : 0) : ST_COLUMN_DROPPED;
      R arr = R(0);                                // alpha of this pivot
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) arr = (l == rr) ? alpha[l] : arr;
      S.set(wcell(j), pivot ? arr : wdrop);
      // pivot: row pk becomes E[pk] / piv and every other row i loses c_i times that — for all rows at once
      // E -= d (x) prow with d_i = c_i - [i == pk]; no pivot: d = 0, nothing changes
      const R inv = pivot ? rcp(piv) : R(0);
      R prow[K1], d[K1];
      ATACOM_UNROLL
      for (int m = 0; m < k; ++m) {
        R e = E[0][m];
        ATACOM_UNROLL
        for (int i = 1; i < k; ++i) e = (i == pk) ? E[i][m] : e;
        prow[m] = e * inv;
      }
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        d[i] = pivot ? ((i == pk) ? c[i] - R(1) : c[i]) : R(0);
        ATACOM_UNROLL
        for (int m = 0; m < k; ++m) E[i][m] -= d[i] * prow[m];
        // the swap: the row that sat at position rr takes the pivot row's old position
        const bool un = ((used >> i) & 1u) == 0u;
        if (pivot && un && pos[i] == rr) pos[i] = ppos;
        if (pivot && i == pk) {
          pos[i] = rr;
          coef[i] = arr;
        }
      }
      used |= pivot ? (1u << pk) : 0u;
      rr += pivot ? 1 : 0;
    }
    if (rr < k) status |= ST_RANK_DEFICIENT;
    // columns behind the last pivot: w_null[j] = sum_i coef_i (E z_j)_i = g . z_j
    R g[K1];
    ATACOM_UNROLL
    for (int m = 0; m < k; ++m) {
      R acc = R(0);
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) acc += coef[i] * E[i][m];
      g[m] = acc;
    }
    ATACOM_UNROLL
    for (int jj = 0; jj < N; ++jj) {
      if (jj < j) {
        w_null[jj] = S.get(wcell(jj));
      } else {
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
