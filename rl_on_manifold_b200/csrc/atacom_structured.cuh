// Structured projection: the fast path of the kernel family.
//
// Same result as project_dense (atacom_core.cuh) — w_mn = -Jc^+ r and w_null = Nc alpha with the
// canonical column-ordered null basis and the reference's tolerance-RREF — computed through the
// block structure  Jc = [[A_f, 0], [A_g, diag(s)]]  instead of a dense N x C factorisation:
//
//  * an inequality row whose slack is not small ("soft": |a_i|^2 <= tau s_i^2) is eliminated
//    analytically, z_i = -(r_i + a_i x)/s_i, which turns its contribution into the n x n SPD
//    metric  M = I + sum_i a_i^T a_i / s_i^2  and the linear term  h = sum_i r_i a_i^T / s_i^2;
//    with M = U U^T (U upper triangular) and xi = U^T x the null space becomes Euclidean;
//  * equality rows and "stiff" inequality rows (small slack: an active constraint, s_i -> 0 is
//    allowed) stay explicit constraints, each with its own slack coordinate zeta_t, and are
//    triangularised by Householder reflectors of support {pivot} + free xi coordinates — no
//    division by a small slack anywhere;
//  * trailing rows of A_g that are diagonal (joint limits: row j = d_j e_j) only touch the
//    diagonal of M.
//
// Cost for the iiwa shape (n=6, F=1, G=11): ~1.5 kFLOP against ~15 kFLOP for the dense path.
// At most TMAX stiff rows are handled here; an environment with more returns ST_DENSE_PATH and
// the caller runs project_dense on it.
#pragma once

#include "atacom_core.cuh"

namespace atacom {

template <typename T, class D, int NDIAG, int TMAX, bool SYNC = false>
struct Structured {
  static constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  static constexpr int GD = G - NDIAG;  // dense inequality rows
  static constexpr int K1 = at_least_1<k>::value, F1 = at_least_1<F>::value, G1 = at_least_1<G>::value;
  static constexpr int GD1 = at_least_1<GD>::value, ND1 = at_least_1<NDIAG>::value, T1 = at_least_1<TMAX>::value;
  static_assert(NDIAG <= n && NDIAG <= G, "diagonal rows: row GD + j has its single entry in column j");

  // row i of A_g dotted with an n-vector
  static ATACOM_HD T gdot(const T* Ad, const T* dg, int i, const T* x) {
    if (i < GD) {
      T d = T(0);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) d += Ad[i * n + j] * x[j];
      return d;
    }
    return dg[i - GD] * x[i - GD];
  }

  // U upper triangular with inverse diagonal ud:  y = U^{-1} y (back substitution)
  static ATACOM_HD void solve_U(const T (*U)[n], const T* ud, T* y) {
    ATACOM_UNROLL
    for (int i = n - 1; i >= 0; --i) {
      T acc = y[i];
      ATACOM_UNROLL
      for (int l = i + 1; l < n; ++l) acc -= U[i][l] * y[l];
      y[i] = acc * ud[i];
    }
  }
  // y = U^{-T} y (forward substitution)
  static ATACOM_HD void solve_Ut(const T (*U)[n], const T* ud, T* y) {
    ATACOM_UNROLL
    for (int i = 0; i < n; ++i) {
      T acc = y[i];
      ATACOM_UNROLL
      for (int l = 0; l < i; ++l) acc -= U[l][i] * y[l];
      y[i] = acc * ud[i];
    }
  }

  // Af: F x n, Ad: GD x n dense rows of A_g, dg: NDIAG diagonal entries, s: G, r: C, alpha: k
  static ATACOM_HD uint8_t project(const T* Af, const T* Ad, const T* dg, const T* s, const T* r,
                                   const T* alpha, T tol, T tau, bool want_null, T* w_mn, T* w_null) {
    uint8_t status = 0;

    // ---- S1: classify inequality rows, accumulate M (upper triangle) and h, collect stiff slots
    T M[n][n], h[n];
    ATACOM_UNROLL
    for (int i = 0; i < n; ++i) {
      h[i] = T(0);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) M[i][j] = (i == j) ? T(1) : T(0);
    }
    T sa[T1][n], ss[T1], sr[T1];  // stiff slots; unused slot = (a = 0, s = 1, r = 0), a no-op
    int srow[T1];
    ATACOM_UNROLL
    for (int t = 0; t < TMAX; ++t) {
      ss[t] = T(1);
      sr[t] = T(0);
      srow[t] = -1;
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) sa[t][j] = T(0);
    }
    // Row i is stiff when |a_i|^2 > tau s_i^2 (tau = 400: |a_i| / |s_i| > 20, the regime in which the rref
    // tolerance 0.05 starts dropping x columns).  Soft rows then keep cond(M) <= 1 + tau G, measured fp32
    // error <= 1e-5; a stiff row that finds no free slot overflows to the dense path.
    int nst = 0;
    bool soft[G1];
    T tinv[G1];
    bool overflow = false;
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      T nrm2 = T(0);
      if (i < GD) {
        ATACOM_UNROLL
        for (int j = 0; j < n; ++j) nrm2 += Ad[(i < GD ? i : 0) * n + j] * Ad[(i < GD ? i : 0) * n + j];
      } else {
        nrm2 = dg[i >= GD ? i - GD : 0] * dg[i >= GD ? i - GD : 0];
      }
      const T s2 = s[i] * s[i];
      const bool zero_row = !(nrm2 > T(0)) && !(s2 > T(0));
      if (zero_row) status |= ST_RANK_DEFICIENT;
      const bool stiff = !(nrm2 <= tau * s2) && !zero_row;
      const bool takes = stiff && nst < TMAX;
      overflow = overflow || (stiff && !takes);
      soft[i] = !takes;
      ATACOM_UNROLL
      for (int t2 = 0; t2 < TMAX; ++t2) srow[t2] = (takes && t2 == nst) ? i : srow[t2];
      nst += takes ? 1 : 0;
      tinv[i] = (soft[i] && s2 > T(0)) ? num<T>::div(T(1), s2) : T(0);
    }
    // copy the chosen rows into their slots (zeros of a diagonal row are compile-time constants)
    ATACOM_UNROLL
    for (int t = 0; t < TMAX; ++t) {
      ATACOM_UNROLL
      for (int i = 0; i < G; ++i) {
        if (srow[t] == i) {
          ss[t] = s[i];
          sr[t] = r[F + i];
          if (i < GD) {
            ATACOM_UNROLL
            for (int j = 0; j < n; ++j) sa[t][j] = Ad[(i < GD ? i : 0) * n + j];
          } else {
            sa[t][i >= GD ? i - GD : 0] = dg[i >= GD ? i - GD : 0];
          }
        }
      }
    }
    // no early return on overflow (block barriers follow): the caller reruns such an environment densely
    if (overflow) status |= ST_DENSE_PATH;
    phase_sync<SYNC>();
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      const T t = tinv[i];
      const T tr = t * r[F + i];
      if (i < GD) {
        ATACOM_UNROLL
        for (int a = 0; a < n; ++a) {
          const T ta = t * Ad[(i < GD ? i : 0) * n + a];
          h[a] += tr * Ad[(i < GD ? i : 0) * n + a];
          ATACOM_UNROLL
          for (int b = a; b < n; ++b) M[a][b] += ta * Ad[(i < GD ? i : 0) * n + b];
        }
      } else {
        const int j = i >= GD ? i - GD : 0;
        M[j][j] += t * dg[j] * dg[j];
        h[j] += tr * dg[j];
      }
    }

    // ---- S2: M = U U^T, U upper triangular (elimination from the last variable up)
    T U[n][n], ud[n];
    ATACOM_UNROLL
    for (int j = n - 1; j >= 0; --j) {
      T d = M[j][j];
      ATACOM_UNROLL
      for (int l = j + 1; l < n; ++l) d -= U[j][l] * U[j][l];
      const T inv = num<T>::rsqrt(d);
      ud[j] = inv;
      U[j][j] = d * inv;
      ATACOM_UNROLL
      for (int i = 0; i < j; ++i) {
        T v = M[i][j];
        ATACOM_UNROLL
        for (int l = j + 1; l < n; ++l) v -= U[i][l] * U[j][l];
        U[i][j] = v * inv;
      }
    }

    phase_sync<SYNC>();
    // ---- S3: explicit constraints in (zeta, xi): equality rows (pivot xi_f), stiff slots (pivot zeta_t)
    T fv[F1][n], fb[F1], tf[F1];  // equality reflectors (entries < f unused), beta, solved multipliers
    ATACOM_UNROLL
    for (int f = 0; f < F; ++f) {
      T v[n];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) v[j] = Af[f * n + j];
      solve_U(U, ud, v);
      T acc = -r[f];
      ATACOM_UNROLL
      for (int g = 0; g < f; ++g) {  // earlier equality reflectors
        T d = T(0);
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) d += fv[g][j] * v[j];
        d *= fb[g];
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) v[j] -= d * fv[g][j];
        acc -= v[g] * tf[g];
      }
      T s2 = T(0);
      ATACOM_UNROLL
      for (int j = f; j < n; ++j) s2 += v[j] * v[j];
      const T sigma = num<T>::sqrt(s2);
      if (!(sigma > T(0))) {
        status |= ST_RANK_DEFICIENT;
        fb[f] = T(0);
        tf[f] = T(0);
      } else {
        const T x0 = v[f];
        const T sg = x0 >= T(0) ? T(1) : T(-1);
        v[f] = x0 + sg * sigma;
        fb[f] = T(1) / (sigma * (sigma + num<T>::abs(x0)));
        tf[f] = acc / (-sg * sigma);
      }
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) fv[f][j] = j >= f ? v[j] : T(0);
    }

    T sz[T1], sx[T1][K1], sb[T1], ts[T1];  // slot reflectors: zeta entry, free-xi entries, beta, multipliers
    ATACOM_UNROLL
    for (int t = 0; t < TMAX; ++t) {
      T v[n];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) v[j] = sa[t][j];
      solve_U(U, ud, v);
      T acc = -sr[t];
      ATACOM_UNROLL
      for (int g = 0; g < F; ++g) {
        T d = T(0);
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) d += fv[g][j] * v[j];
        d *= fb[g];
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) v[j] -= d * fv[g][j];
        acc -= v[g] * tf[g];
      }
      ATACOM_UNROLL
      for (int u = 0; u < t; ++u) {  // earlier slots: this vector has no zeta_u entry yet
        T d = T(0);
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) d += sx[u][l] * v[F + l];
        d *= sb[u];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) v[F + l] -= d * sx[u][l];
        acc -= (-d * sz[u]) * ts[u];
      }
      T s2 = ss[t] * ss[t];
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) s2 += v[F + l] * v[F + l];
      const T sigma = num<T>::sqrt(s2);
      if (!(sigma > T(0))) {
        status |= ST_RANK_DEFICIENT;
        sb[t] = T(0);
        ts[t] = T(0);
        sz[t] = T(0);
      } else {
        const T x0 = ss[t];
        const T sg = x0 >= T(0) ? T(1) : T(-1);
        sz[t] = x0 + sg * sigma;
        sb[t] = T(1) / (sigma * (sigma + num<T>::abs(x0)));
        ts[t] = acc / (-sg * sigma);
      }
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) sx[t][l] = v[F + l];
    }

    // forward (Q^T) and backward (Q) application of all constraint reflectors to (zeta, xi)
    auto apply_Qt = [&](T* ze, T* xi) {
      ATACOM_UNROLL
      for (int g = 0; g < F; ++g) {
        T d = T(0);
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) d += fv[g][j] * xi[j];
        d *= fb[g];
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) xi[j] -= d * fv[g][j];
      }
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) {
        T d = sz[t] * ze[t];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) d += sx[t][l] * xi[F + l];
        d *= sb[t];
        ze[t] -= d * sz[t];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) xi[F + l] -= d * sx[t][l];
      }
    };
    auto apply_Q = [&](T* ze, T* xi) {
      ATACOM_UNROLL
      for (int t = TMAX - 1; t >= 0; --t) {
        T d = sz[t] * ze[t];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) d += sx[t][l] * xi[F + l];
        d *= sb[t];
        ze[t] -= d * sz[t];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) xi[F + l] -= d * sx[t][l];
      }
      ATACOM_UNROLL
      for (int g = F - 1; g >= 0; --g) {
        T d = T(0);
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) d += fv[g][j] * xi[j];
        d *= fb[g];
        ATACOM_UNROLL
        for (int j = g; j < n; ++j) xi[j] -= d * fv[g][j];
      }
    };

    phase_sync<SYNC>();
    // ---- S4: minimum-norm part
    T x_mn[n], z_mn[T1];
    {
      T xi[n], ze[T1];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) xi[j] = h[j];
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) ze[t] = T(0);
      solve_U(U, ud, xi);
      apply_Qt(ze, xi);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) xi[j] = j < F ? tf[j < F ? j : 0] : -xi[j];
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) ze[t] = ts[t];
      apply_Q(ze, xi);
      solve_Ut(U, ud, xi);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) x_mn[j] = xi[j];
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) z_mn[t] = ze[t];
    }

    // slack entries of the min-norm part; stiff rows overwrite theirs below
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) {
      w_mn[j] = x_mn[j];
      w_null[j] = T(0);
    }
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      w_null[n + i] = T(0);
      w_mn[n + i] = soft[i] ? (-r[F + i] - gdot(Ad, dg, i, x_mn)) / s[i] : T(0);
      if (soft[i] && !(s[i] * s[i] > T(0))) w_mn[n + i] = T(0);  // all-zero row
      if (!soft[i]) {
        ATACOM_UNROLL
        for (int t = 0; t < TMAX; ++t)
          if (srow[t] == i) w_mn[n + i] = z_mn[t];
      }
    }
    phase_sync<SYNC>();
    if (!(want_null && k > 0)) {
      // error-correction variant: same barrier count as the null-space branch below
      ATACOM_UNROLL
      for (int c = 0; c < (n + 1) / 2 + 3; ++c) phase_sync<SYNC>();
      return status;
    }

    // ---- S5: orthonormal null basis, x parts Gx[l][:] and stiff-slack parts Bz[l][:]
    T Gx[K1][n], Bz[K1][T1];
    ATACOM_UNROLL
    for (int l = 0; l < k; ++l) {
      T xi[n], ze[T1];
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) xi[j] = (j == F + l) ? T(1) : T(0);
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) ze[t] = T(0);
      apply_Q(ze, xi);
      solve_Ut(U, ud, xi);
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) Gx[l][j] = xi[j];
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) Bz[l][t] = ze[t];
    }

    phase_sync<SYNC>();
    // ---- S6/S7: column-ordered Gram-Schmidt of the columns of Gx (coordinates of P e_c in the basis) with
    // the rref tolerance.  Directions are stored per COLUMN (zero for a dropped / untested column), so
    // every loop range is static.  bcol[c] is the rref multiplier beta of the row pivoted on column c;
    // it is solved on the fly (forward substitution of V~_piv^T beta = alpha).
    T Q[n][K1], bcol[n], xpart[n];
    bool xdrop[n];
    int npiv = 0;
    ATACOM_UNROLL
    for (int c = 0; c < n; ++c) {
      if ((c & 1) == 0) phase_sync<SYNC>();
      T res[K1];
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) res[l] = Gx[l][c];
      T part = T(0);
      ATACOM_UNROLL
      for (int cp = 0; cp < c; ++cp) {
        T rr = T(0);
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) rr += Q[cp][l] * res[l];
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) res[l] -= rr * Q[cp][l];
        part += rr * bcol[cp];
      }
      T s2 = T(0);
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) s2 += res[l] * res[l];
      const T sigma = num<T>::sqrt(s2);
      const bool tested = npiv < k;
      const bool take = tested && (sigma > tol);
      T al = T(0);
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) al = (l == npiv) ? alpha[l] : al;
      const T inv = take ? num<T>::div(T(1), sigma) : T(0);
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) Q[c][l] = res[l] * inv;
      bcol[c] = (al - part) * inv;
      xdrop[c] = tested && !take;
      xpart[c] = part;
      if (xdrop[c]) status |= ST_COLUMN_DROPPED;
      npiv += take ? 1 : 0;
    }

    // exact null vector sum_l beta_l v_l of the rows pivoted so far: x part and stiff-slack part
    T xl[n], zl[T1];
    {
      T gam[K1];
      ATACOM_UNROLL
      for (int l = 0; l < k; ++l) {
        T acc = T(0);
        ATACOM_UNROLL
        for (int c = 0; c < n; ++c) acc += bcol[c] * Q[c][l];
        gam[l] = acc;
      }
      ATACOM_UNROLL
      for (int j = 0; j < n; ++j) {
        T acc = T(0);
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) acc += gam[l] * Gx[l][j];
        xl[j] = acc;
      }
      ATACOM_UNROLL
      for (int t = 0; t < TMAX; ++t) {
        T acc = T(0);
        ATACOM_UNROLL
        for (int l = 0; l < k; ++l) acc += gam[l] * Bz[l][t];
        zl[t] = acc;
      }
    }

    phase_sync<SYNC>();
    // ---- S8: slack columns, only while pivots are missing (a constraint becomes a coordinate).  Work in the
    // remaining subspace (dimension rem = k - npiv <= RMAX here), kept as null vectors in x-functional form:
    // Xr[j] its x part, Zr[j] its stiff-slack part; a soft slack entry is then -(a_i . Xr[j]) / s_i.
    constexpr int RMAX = 2;
    bool ztested[G1];
    T zval[G1];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      ztested[i] = false;
      zval[i] = T(0);
    }
    if (npiv < k) {
      int rem = k - npiv;
      if (rem > RMAX) {
        status |= ST_DENSE_PATH;
      } else {
        T Xr[RMAX][n], Zr[RMAX][T1];
        T Qr[RMAX][K1];
        ATACOM_UNROLL
        for (int j = 0; j < RMAX; ++j) {
          // complete the basis: the unit coordinate vector with the largest remainder, orthogonalised
          T best = T(-1);
          int lb = 0;
          ATACOM_UNROLL
          for (int l = 0; l < k; ++l) {
            T rho = T(1);
            ATACOM_UNROLL
            for (int c = 0; c < n; ++c) rho -= Q[c][l] * Q[c][l];
            ATACOM_UNROLL
            for (int jp = 0; jp < j; ++jp) rho -= Qr[jp][l] * Qr[jp][l];
            if (rho > best) {
              best = rho;
              lb = l;
            }
          }
          T res[K1];
          ATACOM_UNROLL
          for (int l = 0; l < k; ++l) res[l] = (l == lb) ? T(1) : T(0);
          ATACOM_UNROLL
          for (int c = 0; c < n; ++c) {
            T rr = T(0);
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) rr += Q[c][l] * res[l];
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) res[l] -= rr * Q[c][l];
          }
          ATACOM_UNROLL
          for (int jp = 0; jp < j; ++jp) {
            T rr = T(0);
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) rr += Qr[jp][l] * res[l];
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) res[l] -= rr * Qr[jp][l];
          }
          T s2 = T(0);
          ATACOM_UNROLL
          for (int l = 0; l < k; ++l) s2 += res[l] * res[l];
          const T inv = (j < rem && s2 > T(0)) ? num<T>::rsqrt(s2) : T(0);
          ATACOM_UNROLL
          for (int l = 0; l < k; ++l) Qr[j][l] = res[l] * inv;
          ATACOM_UNROLL
          for (int c = 0; c < n; ++c) {
            T acc = T(0);
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) acc += Qr[j][l] * Gx[l][c];
            Xr[j][c] = acc;
          }
          ATACOM_UNROLL
          for (int t = 0; t < TMAX; ++t) {
            T acc = T(0);
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) acc += Qr[j][l] * Bz[l][t];
            Zr[j][t] = acc;
          }
        }
        // Slack columns in order.  Straight-line code (no per-column branches): each of the <= RMAX passes
        // evaluates every slack column on the remaining directions, takes the first one past the previous
        // pivot whose remainder exceeds the tolerance, and marks the columns it skipped as dropped.
        unsigned testmask = 0u;
        T zloc[G1], ninv[G1];
        int slot_of[G1];
        ATACOM_UNROLL
        for (int i = 0; i < G; ++i) {
          zloc[i] = T(0);
          ninv[i] = (soft[i] && s[i] * s[i] > T(0)) ? num<T>::div(T(-1), s[i]) : T(0);
          slot_of[i] = -1;
          ATACOM_UNROLL
          for (int t = 0; t < TMAX; ++t) slot_of[i] = (srow[t] == i) ? t : slot_of[i];
        }
        const T tol2 = tol * tol;
        int last = -1;
        ATACOM_UNROLL
        for (int it = 0; it < RMAX; ++it) {
          if (rem > 0) {
            T e0[G1], e1[G1], part[G1];
            int piv = G;
            ATACOM_UNROLL
            for (int i = G - 1; i >= 0; --i) {
              T a0, a1, ap;
              if (i < GD) {
                a0 = gdot(Ad, dg, i, Xr[0]);
                a1 = gdot(Ad, dg, i, Xr[1]);
                ap = gdot(Ad, dg, i, xl);
              } else {
                const T d = dg[i >= GD ? i - GD : 0];
                a0 = d * Xr[0][i >= GD ? i - GD : 0];
                a1 = d * Xr[1][i >= GD ? i - GD : 0];
                ap = d * xl[i >= GD ? i - GD : 0];
              }
              T z0 = T(0), z1 = T(0), zp = T(0);
              ATACOM_UNROLL
              for (int t = 0; t < TMAX; ++t) {
                z0 = (t == slot_of[i]) ? Zr[0][t] : z0;
                z1 = (t == slot_of[i]) ? Zr[1][t] : z1;
                zp = (t == slot_of[i]) ? zl[t] : zp;
              }
              e0[i] = soft[i] ? a0 * ninv[i] : z0;
              e1[i] = soft[i] ? a1 * ninv[i] : z1;
              part[i] = soft[i] ? ap * ninv[i] : zp;
              const bool ok = (i > last) && (e0[i] * e0[i] + e1[i] * e1[i] > tol2);
              piv = ok ? i : piv;
            }
            T pe0 = T(0), pe1 = T(0), pp = T(0);
            ATACOM_UNROLL
            for (int i = 0; i < G; ++i) {
              const bool skipped = (i > last) && (i < piv);
              zloc[i] = skipped ? part[i] : zloc[i];
              testmask |= (skipped || i == piv) ? (1u << i) : 0u;
              if (skipped) status |= ST_COLUMN_DROPPED;
              pe0 = (i == piv) ? e0[i] : pe0;
              pe1 = (i == piv) ? e1[i] : pe1;
              pp = (i == piv) ? part[i] : pp;
            }
            const bool take = piv < G;
            const T sigma = num<T>::sqrt(pe0 * pe0 + pe1 * pe1);
            T al = T(0);
            ATACOM_UNROLL
            for (int l = 0; l < k; ++l) al = (l == npiv) ? alpha[l] : al;
            ATACOM_UNROLL
            for (int i = 0; i < G; ++i) zloc[i] = (i == piv) ? al : zloc[i];
            if (take) status |= ST_SLACK_PIVOT;
            const T inv = take ? num<T>::div(T(1), sigma) : T(0);
            const T bnew = (al - pp) * inv;
            // direction of the new row inside the remaining subspace; what is left is its in-plane normal
            const T c0 = pe0 * inv, c1 = pe1 * inv;
            ATACOM_UNROLL
            for (int j = 0; j < n; ++j) {
              const T x0 = Xr[0][j], x1 = Xr[1][j];
              xl[j] += bnew * (c0 * x0 + c1 * x1);
              Xr[0][j] = take ? (c0 * x1 - c1 * x0) : x0;
              Xr[1][j] = take ? T(0) : x1;
            }
            ATACOM_UNROLL
            for (int t = 0; t < TMAX; ++t) {
              const T z0 = Zr[0][t], z1 = Zr[1][t];
              zl[t] += bnew * (c0 * z0 + c1 * z1);
              Zr[0][t] = take ? (c0 * z1 - c1 * z0) : z0;
              Zr[1][t] = take ? T(0) : z1;
            }
            npiv += take ? 1 : 0;
            rem -= take ? 1 : 0;
            last = piv;
          }
        }
        ATACOM_UNROLL
        for (int i = 0; i < G; ++i) {
          ztested[i] = (testmask >> i) & 1u;
          zval[i] = zloc[i];
        }
      }
    }
    if (npiv < k) status |= ST_RANK_DEFICIENT;

    phase_sync<SYNC>();
    // ---- S9: w_null = sum_l beta_l V~[l][:]; a dropped column keeps only the rows pivoted before it
    ATACOM_UNROLL
    for (int j = 0; j < n; ++j) w_null[j] = xdrop[j] ? xpart[j] : xl[j];
    ATACOM_UNROLL
    for (int i = 0; i < G; ++i) {
      T v;
      if (ztested[i]) {
        v = zval[i];
      } else if (soft[i]) {
        v = (s[i] * s[i] > T(0)) ? -num<T>::div(gdot(Ad, dg, i, xl), s[i]) : T(0);
      } else {
        v = T(0);
        ATACOM_UNROLL
        for (int t = 0; t < TMAX; ++t) v = (srow[t] == i) ? zl[t] : v;
      }
      w_null[n + i] = v;
    }
    return status;
  }
};

}  // namespace atacom
