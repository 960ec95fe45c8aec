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
};

template <typename R, class D>
struct Lapack {
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  static constexpr int C1 = at_least_1<C>::value, K1 = at_least_1<k>::value;
  static constexpr bool LQ_PATH = N >= (C * 11) / 6;            // gesdd: MNTHR = int(minmn * 11 / 6)
  // Z (N x k) lives in dead cells of the C x N array when the shape allows it (see zcell), else behind it
  static constexpr bool Z_OVERLAY = (C >= 2 * k - 1) && (N >= 2 * k) && (k > 0);
  static constexpr int SIZE = C * N + (Z_OVERLAY ? 0 : N * k);
  static constexpr ATACOM_HD int a(int i, int j) { return i * N + j; }
  // row j of Z.  Overlay: rows j < C take the last k cells of row j of the array (v_j lives there until reflector j
  // has been applied, which is when row j of Z comes into being); the identity rows j >= C take the first k cells
  // of row j - k, below the diagonal (left-reflector / L entries, dead after the minimum-norm solve).
  static constexpr ATACOM_HD int zcell(int j, int c) {
    return Z_OVERLAY ? (j < C ? j * N + (N - k) + c : (j - k) * N + c) : C * N + j * k + c;
  }

  // dlarfg: reflector I - tau [1; v][1; v]^T mapping [alpha; x] to [beta; 0].  x is given by its squared norm;
  // returns beta and tau, `scale` = 1 / (alpha - beta) turns x into v (0 when x = 0: the identity, as LAPACK).
  static ATACOM_HD void larfg(R alpha, R xn2, R* beta, R* tau, R* scale) {
    if (!(xn2 > R(0))) {
      *beta = alpha;
      *tau = R(0);
      *scale = R(0);
      return;
    }
    const R nrm = num<R>::sqrt(alpha * alpha + xn2);
    const R b = -::copysign(nrm, alpha);      // Fortran SIGN: the sign of alpha, negative zero included
    *beta = b;
    *tau = (b - alpha) / b;
    *scale = R(1) / (alpha - b);
  }

  // S: on entry Jc (C x N, row-major, S.get(i * N + j)); destroyed.  r: C (right-hand side psi + K_c c; destroyed),
  // alpha: k.  w_mn = -Jc^+ r, w_null = Nc alpha (N each; valid in lane 0 of the group).  Returns status bits.
  //
  // LPE lanes of one warp work on one environment (LPE = 1: one thread, the host build).  The rows the right
  // reflector G_i updates are independent of each other, and so are the columns the left reflector H_i updates: the
  // lanes of a group take them round robin (lane `sub` of LPE), every lane forms the reflector itself from the row /
  // column it needs, and a __syncwarp separates the phases (`G.sync()`, over the lanes of the warp that take part).
  // The null basis Z is formed column-wise in registers, one or two columns per lane, and the elimination steps of the
  // rref are shared column-wise as well.  What bounds the routine on the device is shared-memory traffic — every
  // trailing entry is loaded and stored once per reflector — and with LPE = 4 there are enough warps to keep it busy.
  template <int LPE = 1, class ST, class GRP = SoloGroup>
  static ATACOM_HD uint8_t project(ST& S, R* r, const R* alpha, R tol, bool want_null, R* w_mn, R* w_null,
                                   const GRP& Grp = GRP()) {
    const int sub = Grp.sub();
    uint8_t status = 0;
    R tau_s[C1];
    R amax = R(0);

    // ---- reflectors (dgebd2 / dgelq2), left reflectors applied to the right-hand side as they are formed.
    // The loop over i is unrolled: every range below is static, nothing is predicated.
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) {
      Grp.sync();                                  // row i is final: the previous reflectors have been applied to it
      R v[N];                                      // row i right of the diagonal, then the reflector vector
      R xn2 = R(0);
      const R x0 = S.get(a(i, i));
      amax = num<R>::abs(x0) > amax ? num<R>::abs(x0) : amax;
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) {
        v[j] = S.get(a(i, j));
        const R ax = num<R>::abs(v[j]);
        amax = ax > amax ? ax : amax;
        xn2 += v[j] * v[j];
      }
      R beta, tau, sc;
      larfg(x0, xn2, &beta, &tau, &sc);
      tau_s[i] = tau;
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) v[j] *= sc;
      if (sub == 0) {                              // (tau = 0, the identity: x = 0 = v, every update below is a no-op)
        S.set(a(i, i), beta);
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
        if (sub == 0) S.set(a(i + 1 < C ? i + 1 : 0, i), betaq);     // (u itself is not needed again: it is applied to r here)
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
          R w = r[i + 1 < C ? i + 1 : 0];          // and the right-hand side (every lane keeps its own copy)
          ATACOM_UNROLL
          for (int l = i + 2; l < C; ++l) w += u[l] * r[l];
          w *= tauq;
          r[i + 1 < C ? i + 1 : 0] -= w;
          ATACOM_UNROLL
          for (int l = i + 2; l < C; ++l) r[l] -= w * u[l];
        }
      }
    }
    Grp.sync();
    const R rank_floor = R(64) * num<R>::eps() * amax;     // (amax: the largest entry met while walking the rows)

    // ---- minimum-norm part: Jc = U B P^T (B lower bidiagonal) or L Q;  x = P [y; 0],  B y = -U^T r  /  L y = -r
    R t[N];
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) t[i] = R(0);
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) {
      R acc = -r[i];
      if (LQ_PATH) {
        ATACOM_UNROLL
        for (int j = 0; j < i; ++j) acc -= S.get(a(i, j)) * t[j];
      } else if (i > 0) {
        acc -= S.get(a(i, i > 0 ? i - 1 : 0)) * t[i > 0 ? i - 1 : 0];
      }
      const R d = S.get(a(i, i));
      if (num<R>::abs(d) > rank_floor) {
        t[i] = acc / d;
      } else {                                     // dependent row of Jc: dropped from the solve, flagged
        status |= ST_RANK_DEFICIENT;
        t[i] = R(0);
      }
    }
    // P = G_0 ... G_{C-1} applied to k + 1 vectors in one sweep over the reflectors (every index static): vector 0 is
    // [y; 0], which becomes the minimum-norm solution; vectors 1..k are the unit vectors e_C .. e_{N-1}, which become
    // the columns of Z — the null basis as gesdd returns it (rows C.. of VT).  The lanes take the vectors round robin.
    const bool null_part = want_null && k > 0;
    constexpr int CPL = (k + 1 + LPE - 1) / LPE;   // vectors per lane
    R X[CPL][N];
    ATACOM_UNROLL
    for (int cc = 0; cc < CPL; ++cc) {
      ATACOM_UNROLL
      for (int j = 0; j < N; ++j) X[cc][j] = (sub + cc * LPE == 0) ? t[j] : ((j - C == sub + cc * LPE - 1) ? R(1) : R(0));
    }
    ATACOM_UNROLL
    for (int i = C - 1; i >= 0; --i) {
      R vi[N];
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) vi[j] = S.get(a(i, j));
      ATACOM_UNROLL
      for (int cc = 0; cc < CPL; ++cc) {
        R w = X[cc][i];
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) w += vi[j] * X[cc][j];
        w *= tau_s[i];
        X[cc][i] -= w;
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) X[cc][j] -= w * vi[j];
      }
    }
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      w_mn[i] = X[0][i];                           // (the minimum-norm part: lane 0's first vector)
      w_null[i] = R(0);
    }
    if (!null_part) return status;
    Grp.sync();                                    // every lane has read the last reflector: its cells may be reused
    ATACOM_UNROLL
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = sub + cc * LPE - 1;
      if (c >= 0 && c < k) {
        ATACOM_UNROLL
        for (int j = 0; j < N; ++j) S.set(zcell(j, c), X[cc][j]);
      }
    }
    Grp.sync();

    // ---- the reference's rref on V = Z^T (k x N), null_space_coordinate.py:40-79 as written: walk the columns; the
    // pivot candidate is the first largest |.| among the rows not used yet; <= tol: zero those entries and move on;
    // else swap (columns j.. only), scale the pivot row, eliminate the column from every other row.  Every lane takes
    // the same decisions; the columns an elimination step touches are shared out round robin.
    int rr = 0;
    ATACOM_ROLLED
    for (int j = 0; j < N && rr < k; ++j) {
      R colj[K1];
      int kk = rr;
      R p = R(-1);
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        colj[i] = S.get(zcell(j, i));
        const R av = num<R>::abs(colj[i]);
        if (i >= rr && av > p) {
          p = av;
          kk = i;
        }
      }
      if (!(p > tol)) {
        status |= ST_COLUMN_DROPPED;
        Grp.sync();                                // everyone has read column j
        if (sub == 0) {
          ATACOM_UNROLL
          for (int i = 0; i < k; ++i) {
            if (i >= rr) S.set(zcell(j, i), R(0));
          }
        }
        continue;
      }
      if (j >= n) status |= ST_SLACK_PIVOT;
      // entries of column j in rows kk (the pivot) and rr: the row index is just an address here
      const R piv = S.get(zcell(j, kk)), other = S.get(zcell(j, rr));
      const R inv = R(1) / piv;
      // multipliers of the elimination: row i loses f_i times the scaled pivot row; after the swap row kk holds what
      // was row rr (its multiplier is `fk`; rows rr and kk are rewritten after the generic update below)
      R f[K1];
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) f[i] = colj[i];
      const R fk = (kk == rr) ? R(0) : other;
      Grp.sync();                                  // everyone has read column j before it is rewritten
      ATACOM_ROLLED
      for (int jj = j + sub; jj < N; jj += LPE) {
        R e[K1];
        ATACOM_UNROLL
        for (int i = 0; i < k; ++i) e[i] = S.get(zcell(jj, i));
        const R ek = S.get(zcell(jj, kk)), er = S.get(zcell(jj, rr));
        const R lead = ek * inv;
        ATACOM_UNROLL
        for (int i = 0; i < k; ++i) S.set(zcell(jj, i), e[i] - f[i] * lead);
        S.set(zcell(jj, kk), er - fk * lead);      // swap rr <-> kk: row kk takes over what was row rr ...
        S.set(zcell(jj, rr), lead);                // ... and row rr becomes the scaled pivot row (also when kk == rr)
      }
      Grp.sync();
      ++rr;
    }
    Grp.sync();
    if (rr < k) status |= ST_RANK_DEFICIENT;
    ATACOM_UNROLL
    for (int j = 0; j < N; ++j) {
      R acc = R(0);
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) acc += alpha[i] * S.get(zcell(j, i));
      w_null[j] = acc;
    }
    return status;
  }
};

}  // namespace atacom
