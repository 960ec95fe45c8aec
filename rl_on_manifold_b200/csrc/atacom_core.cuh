// ATACOM tangent-space projection: per-environment algebra (sm_100a device code).
//
// Replaces, for one environment, the body of AtacomEnvWrapper.step_action_function
// (reference atacom/atacom.py:126-133): given the stacked, K-scaled constraint
// Jacobian  Jc = [[A_f, 0], [A_g, diag(s)]]  (atacom.py:151-165) and the right-hand
// side  r = psi + K_c*c  it returns
//     w_mn   = -Jc^+ r                     (act_a + act_err, atacom.py:130,132)
//     w_null =  Nc alpha                   (act_b, atacom.py:131)
// where Nc = rref(orthonormal null basis of Jc, tol)  (atacom.py:127-128,
// atacom/utils/null_space_coordinate.py:8-26,40-79).
//
// The reference takes its null basis from LAPACK gesdd, which is implementation
// defined for a null space of dimension > 1; its tolerance-RREF output then
// depends on that choice whenever the tolerance branch fires.  This code uses
// the one basis that depends on null(Jc) only: Gram-Schmidt of the projected
// unit vectors P e_0, P e_1, ... in column order, skipping a column when what is
// left of it has norm <= tol (oracle/nullspace.py:canonical_null_basis).  The
// reference's own rref applied to that basis gives exactly what is computed
// here; when the tolerance branch does not fire the result is basis-free and
// equals the reference's.
//
// Everything is __host__ __device__ so tests/host_harness can compile the same
// source with g++ (float and double) and check it against the NumPy oracle on a
// box without a GPU.  The product library only ever runs it on the device.
#pragma once

#include <math.h>
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define ATACOM_HD __host__ __device__ __forceinline__
#define ATACOM_UNROLL _Pragma("unroll")
#else
#define ATACOM_HD inline
#define ATACOM_UNROLL
#endif

// rarely taken paths are kept out of line so that they do not set the register budget of the hot path
#if defined(__CUDACC__)
#define ATACOM_NOINLINE __host__ __device__ __noinline__
#define ATACOM_ROLLED _Pragma("unroll 1")
#else
#define ATACOM_NOINLINE __attribute__((noinline))
#define ATACOM_ROLLED
#endif

namespace atacom {

// The per-environment code is one long fully unrolled instruction stream (> 100 KB of SASS); warps
// that drift apart each stream it through the instruction cache on their own.  Kernels in which every
// thread of the block runs the projection place block barriers between its phases (SYNC = true) so
// the resident warps stay in the same code region and share instruction-cache lines.
template <bool SYNC>
ATACOM_HD void phase_sync() {
#if defined(__CUDA_ARCH__)
  if (SYNC) __syncthreads();
#endif
}

// status bits written per environment
enum : uint8_t {
  ST_RANK_DEFICIENT = 1,   // Jc lost row rank (pinv_null would cut a singular value)
  ST_COLUMN_DROPPED = 2,   // the tolerance branch of rref fired on some column
  ST_SLACK_PIVOT = 4,      // a slack column became a tangent-space coordinate
  ST_NONFINITE = 8,        // a non-finite value reached the output
  ST_DENSE_PATH = 16,      // the structured fast path deferred to the dense path
};

template <typename T> struct num;
template <> struct num<float> {
  static ATACOM_HD float eps() { return 1.1920929e-07f; }
  static ATACOM_HD float sqrt(float x) { return ::sqrtf(x); }
  static ATACOM_HD float abs(float x) { return ::fabsf(x); }
  static ATACOM_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);      // MUFU.RCP + FMUL, <= 2 ulp: far below eps32 * cond(Jc)
#else
    return a / b;
#endif
  }
  static ATACOM_HD float rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return ::rsqrtf(x);
#else
    return 1.0f / ::sqrtf(x);
#endif
  }
};
template <> struct num<double> {
  static ATACOM_HD double eps() { return 2.220446049250313e-16; }
  static ATACOM_HD double sqrt(double x) { return ::sqrt(x); }
  static ATACOM_HD double abs(double x) { return ::fabs(x); }
  static ATACOM_HD double div(double a, double b) { return a / b; }
  static ATACOM_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
};

// float <-> double conversion of a run-time value.  The library is compiled with --ftz=true, under which a
// plain cast becomes a flush (FMUL / DSETP) plus the conversion; the exact single-instruction cvt is what the
// mixed-precision code wants (fp32 denormals do not occur in this data).
template <typename To, typename From>
ATACOM_HD To cvt(From x) {
#if defined(__CUDA_ARCH__)
  if constexpr (std::is_same<To, double>::value && std::is_same<From, float>::value) {
    double y;
    asm("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(x));
    return y;
  } else if constexpr (std::is_same<To, float>::value && std::is_same<From, double>::value) {
    float y;
    asm("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x));
    return y;
  } else {
    return static_cast<To>(x);
  }
#else
  return static_cast<To>(x);
#endif
}

template <int N_, int F_, int G_>
struct Dims {
  static constexpr int n = N_;       // dim_q
  static constexpr int F = F_;       // equality rows
  static constexpr int G = G_;       // inequality rows = slack variables
  static constexpr int C = F_ + G_;  // rows of Jc
  static constexpr int N = N_ + G_;  // columns of Jc
  static constexpr int k = N_ - F_;  // tangent-space (action) dimension
};

template <int V> struct at_least_1 { static constexpr int value = V > 0 ? V : 1; };

// ---------------------------------------------------------------------------------------
// Dense path: Householder QR of Jc^T (N x C), exactly the orthogonal route the reference's
// pinv_null takes through LAPACK, followed by the column-ordered triangularisation of the
// null basis and the tolerance-RREF.  No division by a slack variable anywhere, so it is
// valid for s_i = 0 (active constraint) and any mix of stiff and soft rows.  Cost is
// O(C^2 N): it is the robust reference path of the kernel family and the fallback of the
// structured path below.
//
// Af: F x n, Ag: G x n (row-major, leading dimension n), s: G, r: C, alpha: k.
// w_mn, w_null: N each.  Returns status bits.
template <typename T, class D>
ATACOM_HD uint8_t project_dense(const T* Af, const T* Ag, const T* s, const T* r, const T* alpha,
                                T tol, bool want_null, T* w_mn, T* w_null) {
  constexpr int n = D::n, F = D::F, G = D::G, C = D::C, N = D::N, k = D::k;
  constexpr int C1 = at_least_1<C>::value, K1 = at_least_1<k>::value;
  uint8_t status = 0;

  T Tm[N][C1];  // Jc^T; becomes reflectors (on and below the diagonal) and R (above)
  ATACOM_UNROLL
  for (int i = 0; i < N; ++i) {
    ATACOM_UNROLL
    for (int j = 0; j < C; ++j) {
      T v = T(0);
      if (i < n) v = (j < F) ? Af[(j < F ? j : 0) * n + i] : Ag[(j >= F ? j - F : 0) * n + i];
      else if (i - n == j - F) v = s[i - n];
      Tm[i][j] = v;
    }
  }

  T beta[C1], rdiag[C1], t[C1];
  T colmax = T(0);
  ATACOM_UNROLL
  for (int j = 0; j < C; ++j) {
    T s2 = T(0);
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) s2 += Tm[i][j] * Tm[i][j];
    colmax = s2 > colmax ? s2 : colmax;
  }
  const T rank_floor = T(64) * num<T>::eps() * num<T>::sqrt(colmax);

  ATACOM_UNROLL
  for (int j = 0; j < C; ++j) {
    T s2 = T(0);
    ATACOM_UNROLL
    for (int i = j; i < N; ++i) s2 += Tm[i][j] * Tm[i][j];
    const T sigma = num<T>::sqrt(s2);
    if (!(sigma > rank_floor)) {  // dependent row of Jc: identity reflector, flagged
      status |= ST_RANK_DEFICIENT;
      beta[j] = T(0);
      rdiag[j] = T(0);
      continue;
    }
    const T x0 = Tm[j][j];
    const T sg = x0 >= T(0) ? T(1) : T(-1);
    Tm[j][j] = x0 + sg * sigma;
    beta[j] = T(1) / (sigma * (sigma + num<T>::abs(x0)));
    rdiag[j] = -sg * sigma;
    ATACOM_UNROLL
    for (int jj = j + 1; jj < C; ++jj) {
      T d = T(0);
      ATACOM_UNROLL
      for (int i = j; i < N; ++i) d += Tm[i][j] * Tm[i][jj];
      d *= beta[j];
      ATACOM_UNROLL
      for (int i = j; i < N; ++i) Tm[i][jj] -= d * Tm[i][j];
    }
  }

  // minimum-norm part: R^T t = -r, w_mn = Q [t; 0]
  ATACOM_UNROLL
  for (int j = 0; j < C; ++j) {
    T acc = -r[j];
    ATACOM_UNROLL
    for (int i = 0; i < j; ++i) acc -= Tm[i][j] * t[i];
    t[j] = rdiag[j] != T(0) ? acc / rdiag[j] : T(0);
  }
  ATACOM_UNROLL
  for (int i = 0; i < N; ++i) w_mn[i] = i < C ? t[i < C ? i : 0] : T(0);
  ATACOM_UNROLL
  for (int j = C - 1; j >= 0; --j) {
    T d = T(0);
    ATACOM_UNROLL
    for (int i = j; i < N; ++i) d += Tm[i][j] * w_mn[i];
    d *= beta[j];
    ATACOM_UNROLL
    for (int i = j; i < N; ++i) w_mn[i] -= d * Tm[i][j];
  }

  ATACOM_UNROLL
  for (int i = 0; i < N; ++i) w_null[i] = T(0);
  if (!want_null || k == 0) return status;

  // orthonormal null basis: E[:, l] = Q e_{C+l}
  T E[N][K1];
  ATACOM_UNROLL
  for (int l = 0; l < k; ++l) {
    T e[N];
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) e[i] = (i == C + l) ? T(1) : T(0);
    ATACOM_UNROLL
    for (int j = C - 1; j >= 0; --j) {
      T d = T(0);
      ATACOM_UNROLL
      for (int i = j; i < N; ++i) d += Tm[i][j] * e[i];
      d *= beta[j];
      ATACOM_UNROLL
      for (int i = j; i < N; ++i) e[i] -= d * Tm[i][j];
    }
    ATACOM_UNROLL
    for (int i = 0; i < N; ++i) E[i][l] = e[i];
  }

  // column-ordered triangularisation with the rref tolerance.  Row c of E holds the
  // coordinates of P e_c in the basis; a reflector in R^k per accepted column.
  int npiv = 0;
  int pivcol[K1];
  int dropat[N];
  ATACOM_UNROLL
  for (int c = 0; c < N; ++c) dropat[c] = k;
  for (int c = 0; c < N && npiv < k; ++c) {
    T s2 = T(0);
    for (int l = npiv; l < k; ++l) s2 += E[c][l] * E[c][l];
    const T p = num<T>::sqrt(s2);
    if (!(p > tol)) {
      dropat[c] = npiv;
      status |= ST_COLUMN_DROPPED;
      continue;
    }
    if (c >= n) status |= ST_SLACK_PIVOT;
    T v[K1];
    for (int l = 0; l < k; ++l) v[l] = l >= npiv ? E[c][l] : T(0);
    const T x0 = v[npiv];
    const T sg = x0 >= T(0) ? T(1) : T(-1);
    v[npiv] = x0 + sg * p;
    const T bt = T(1) / (p * (p + num<T>::abs(x0)));
    for (int cc = 0; cc < N; ++cc) {
      T d = T(0);
      for (int l = npiv; l < k; ++l) d += v[l] * E[cc][l];
      d *= bt;
      for (int l = npiv; l < k; ++l) E[cc][l] -= d * v[l];
    }
    pivcol[npiv] = c;
    ++npiv;
  }
  if (npiv < k) status |= ST_RANK_DEFICIENT;

  // rref: E~_piv^T b = alpha (forward substitution), w_null = sum_l b_l E~[:, l] with the
  // reference's truncation (entries of a dropped column are zero in rows pivoted later)
  T b[K1];
  for (int i = 0; i < k; ++i) b[i] = T(0);
  for (int i = 0; i < npiv; ++i) {
    const int pc = pivcol[i];
    T acc = alpha[i];
    for (int l = 0; l < i; ++l) acc -= E[pc][l] * b[l];
    b[i] = acc / E[pc][i];
  }
  for (int c = 0; c < N; ++c) {
    T acc = T(0);
    const int lim = dropat[c] < npiv ? dropat[c] : npiv;
    for (int l = 0; l < lim; ++l) acc += b[l] * E[c][l];
    w_null[c] = acc;
  }
  return status;
}

}  // namespace atacom
