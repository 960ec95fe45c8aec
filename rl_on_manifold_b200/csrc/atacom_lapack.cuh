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
  // alpha: k.  w_mn = -Jc^+ r, w_null = Nc alpha (N each).  Returns status bits.
  //
  // Memory traffic is what bounds this routine on the device (one thread per environment, the array in shared
  // memory): the current reflector vector lives in registers, a row / column of the trailing matrix is loaded
  // once, updated and stored once per reflector (inner loops unrolled over the static bound with predicates, the
  // outer ones rolled), and the null basis Z is formed entirely in registers.
  template <class ST>
  static ATACOM_HD uint8_t project(ST& S, R* r, const R* alpha, R tol, bool want_null, R* w_mn, R* w_null) {
    uint8_t status = 0;
    R taup[C1];
    R amax = R(0);

    // ---- reflectors (dgebd2 / dgelq2), left reflectors applied to the right-hand side as they are formed
    ATACOM_ROLLED
    for (int i = 0; i < C; ++i) {
      R v[N];                                      // row i from the diagonal on, then the reflector vector
      R xn2 = R(0), x0 = R(0);
      ATACOM_UNROLL
      for (int j = 0; j < N; ++j) {
        v[j] = R(0);
        if (j >= i) {
          const R x = S.get(a(i, j));
          const R ax = num<R>::abs(x);
          amax = ax > amax ? ax : amax;
          if (j == i) x0 = x;
          else {
            v[j] = x;
            xn2 += x * x;
          }
        }
      }
      R beta, tau, sc;
      larfg(x0, xn2, &beta, &tau, &sc);
      S.set(a(i, i), beta);
      taup[i] = tau;
      if (tau != R(0)) {
        ATACOM_UNROLL
        for (int j = 0; j < N; ++j) {
          v[j] *= sc;
          if (j > i) S.set(a(i, j), v[j]);
        }
        ATACOM_ROLLED
        for (int l = i + 1; l < C; ++l) {          // rows below: A <- A G_i, one load and one store per entry
          R row[N];
          R w0 = R(0), w1 = R(0), w2 = R(0), w3 = R(0);
          ATACOM_UNROLL
          for (int j = 0; j < N; ++j) {
            row[j] = R(0);
            if (j >= i) {
              row[j] = S.get(a(l, j));
              const R pr = (j == i) ? row[j] : row[j] * v[j];
              if ((j & 3) == 0) w0 += pr;
              else if ((j & 3) == 1) w1 += pr;
              else if ((j & 3) == 2) w2 += pr;
              else w3 += pr;
            }
          }
          const R w = ((w0 + w1) + (w2 + w3)) * tau;
          ATACOM_UNROLL
          for (int j = 0; j < N; ++j) {
            if (j == i) S.set(a(l, j), row[j] - w);
            else if (j > i) S.set(a(l, j), row[j] - w * v[j]);
          }
        }
      }
      if (!LQ_PATH && i + 1 < C) {
        R u[C1];                                   // column i below the diagonal, then the reflector vector
        R un2 = R(0), c0 = R(0);
        ATACOM_UNROLL
        for (int l = 0; l < C; ++l) {
          u[l] = R(0);
          if (l >= i + 1) {
            const R x = S.get(a(l, i));
            if (l == i + 1) c0 = x;
            else {
              u[l] = x;
              un2 += x * x;
            }
          }
        }
        R betaq, tauq, scq;
        larfg(c0, un2, &betaq, &tauq, &scq);
        S.set(a(i + 1, i), betaq);                 // (u itself is not needed again: it is applied to r right here)
        if (tauq != R(0)) {
          ATACOM_UNROLL
          for (int l = 0; l < C; ++l) u[l] = (l == i + 1) ? R(1) : u[l] * scq;
          ATACOM_ROLLED
          for (int j = i + 1; j < N; ++j) {        // columns to the right: A <- H_i A
            R col[C1];
            R w0 = R(0), w1 = R(0);
            ATACOM_UNROLL
            for (int l = 0; l < C; ++l) {
              col[l] = R(0);
              if (l >= i + 1) {
                col[l] = S.get(a(l, j));
                if (l & 1) w1 += col[l] * u[l];
                else w0 += col[l] * u[l];
              }
            }
            const R w = (w0 + w1) * tauq;
            ATACOM_UNROLL
            for (int l = 0; l < C; ++l) {
              if (l >= i + 1) S.set(a(l, j), col[l] - w * u[l]);
            }
          }
          R w = R(0);                              // and the right-hand side: r <- H_i r   (u is zero above row i + 1)
          ATACOM_UNROLL
          for (int l = 0; l < C; ++l) w += u[l] * r[l];
          w *= tauq;
          ATACOM_UNROLL
          for (int l = 0; l < C; ++l) r[l] -= w * u[l];
        }
      }
    }
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
    R tau_s[C1];                                   // (static copies: the loops below are unrolled)
    ATACOM_UNROLL
    for (int i = 0; i < C; ++i) tau_s[i] = taup[i];
    ATACOM_UNROLL
    for (int i = C - 1; i >= 0; --i) {
      R w = t[i];
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) w += S.get(a(i, j)) * t[j];
      w *= tau_s[i];
      t[i] -= w;
      ATACOM_UNROLL
      for (int j = i + 1; j < N; ++j) t[j] -= w * S.get(a(i, j));
    }
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) {
      w_mn[i] = t[i];
      w_null[i] = R(0);
    }
    if (!want_null || k == 0) return status;

    // ---- Z = G_0 ... G_{C-1} [0; I]: the null basis as gesdd returns it (columns of Z = rows C.. of VT), formed in
    // registers (every index static); it goes to its cells of the array only when the last reflector has been read
    {
      R Z[N][K1];
      ATACOM_UNROLL
      for (int j = 0; j < N; ++j) {
        ATACOM_UNROLL
        for (int c = 0; c < k; ++c) Z[j][c] = (j - C == c) ? R(1) : R(0);
      }
      ATACOM_UNROLL
      for (int i = C - 1; i >= 0; --i) {
        R vi[N];
        ATACOM_UNROLL
        for (int j = i + 1; j < N; ++j) vi[j] = S.get(a(i, j));
        ATACOM_UNROLL
        for (int c = 0; c < k; ++c) {
          R w = R(0);                              // (row i of Z is zero before reflector i)
          ATACOM_UNROLL
          for (int j = (i + 1 > C ? i + 1 : C); j < N; ++j) w += vi[j] * Z[j][c];      // rows < C above i are zero too
          ATACOM_UNROLL
          for (int j = i + 1; j < C; ++j) w += vi[j] * Z[j][c];
          w *= tau_s[i];
          Z[i][c] = -w;
          ATACOM_UNROLL
          for (int j = i + 1; j < N; ++j) Z[j][c] -= w * vi[j];
        }
      }
      ATACOM_UNROLL
      for (int j = 0; j < N; ++j) {
        ATACOM_UNROLL
        for (int c = 0; c < k; ++c) S.set(zcell(j, c), Z[j][c]);
      }
    }

    // ---- the reference's rref on V = Z^T (k x N), null_space_coordinate.py:40-79 as written: walk the columns; the
    // pivot candidate is the first largest |.| among the rows not used yet; <= tol: zero those entries and move on;
    // else swap (columns j.. only), scale the pivot row, eliminate the column from every other row.
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
        ATACOM_UNROLL
        for (int i = 0; i < k; ++i) {
          if (i >= rr) S.set(zcell(j, i), R(0));
        }
        continue;
      }
      if (j >= n) status |= ST_SLACK_PIVOT;
      R piv = R(0), other = R(0);                  // entries of column j in rows kk (the pivot) and rr
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) {
        piv = (i == kk) ? colj[i] : piv;
        other = (i == rr) ? colj[i] : other;
      }
      const R inv = R(1) / piv;
      // multipliers of the elimination: row i loses f_i times the scaled pivot row; after the swap row kk holds what
      // was row rr
      R f[K1];
      ATACOM_UNROLL
      for (int i = 0; i < k; ++i) f[i] = (i == rr) ? R(0) : ((i == kk) ? other : colj[i]);
      ATACOM_UNROLL
      for (int jj = 0; jj < N; ++jj) {
        if (jj >= j) {
          R e[K1];
          ATACOM_UNROLL
          for (int i = 0; i < k; ++i) e[i] = S.get(zcell(jj, i));
          R ek = R(0), er = R(0);
          ATACOM_UNROLL
          for (int i = 0; i < k; ++i) {
            ek = (i == kk) ? e[i] : ek;
            er = (i == rr) ? e[i] : er;
          }
          const R lead = ek * inv;
          ATACOM_UNROLL
          for (int i = 0; i < k; ++i) {
            const R cur = (i == kk) ? er : e[i];                 // swap rr <-> kk (a no-op when kk == rr)
            const R out = (i == rr) ? lead : cur - f[i] * lead;
            S.set(zcell(jj, i), out);
          }
        }
      }
      ++rr;
    }
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
