// Host-side test harness: compiles the DEVICE algebra headers of the product
// (rl_on_manifold_b200/csrc/*.cuh, all __host__ __device__) with g++ so the
// `-m "not gpu"` tests can check the very same source against the NumPy oracle on
// a box without a GPU.  Test infrastructure only: never linked into the product
// library, never a fallback.
#include <cstdint>
#include "../../rl_on_manifold_b200/csrc/atacom_envs.cuh"
#include "../../rl_on_manifold_b200/csrc/atacom_structured.cuh"
#include "../../rl_on_manifold_b200/csrc/atacom_lapack.cuh"

using namespace atacom;

template <typename T, int n, int F, int G>
static void run_dense(int64_t B, const T* Af, const T* Ag, const T* s, const T* r, const T* alpha, T tol,
                      T* w_mn, T* w_null, uint8_t* status) {
  using D = Dims<n, F, G>;
  for (int64_t b = 0; b < B; ++b) {
    status[b] = project_dense<T, D>(Af + b * F * n, Ag + b * G * n, s + b * G, r + b * D::C,
                                    alpha + b * D::k, tol, true, w_mn + b * D::N, w_null + b * D::N);
  }
}

// the LAPACK-basis routine (atacom_lapack.cuh) on the same interface: Jc assembled from A_f, A_g, s
template <typename T, int n, int F, int G>
static void run_lapack(int64_t B, const T* Af, const T* Ag, const T* s, const T* r, const T* alpha, T tol,
                       T* w_mn, T* w_null, uint8_t* status) {
  using D = Dims<n, F, G>;
  using LP = Lapack<T, D>;
  for (int64_t b = 0; b < B; ++b) {
    ArrayStore<T, LP::SIZE> S;
    for (int i = 0; i < LP::SIZE; ++i) S.set(i, T(0));
    for (int i = 0; i < D::C; ++i)
      for (int j = 0; j < n; ++j) S.set(LP::a(i, j), i < F ? Af[(b * F + i) * n + j] : Ag[(b * G + i - F) * n + j]);
    for (int i = 0; i < G; ++i) S.set(LP::a(F + i, n + i), s[b * G + i]);
    T rr[D::C > 0 ? D::C : 1];
    for (int i = 0; i < D::C; ++i) rr[i] = r[b * D::C + i];
    status[b] = LP::project(S, rr, alpha + b * D::k, tol, true, w_mn + b * D::N, w_null + b * D::N);
  }
}

#define DISPATCH(T, FN)                                                                      \
  if (n == 2 && F == 1 && G == 1) { FN<T, 2, 1, 1>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 0 && G == 6) { FN<T, 3, 0, 6>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 6 && F == 1 && G == 11) { FN<T, 6, 1, 11>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 7 && F == 1 && G == 12) { FN<T, 7, 1, 12>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 2 && F == 0 && G == 4) { FN<T, 2, 0, 4>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 4 && F == 2 && G == 3) { FN<T, 4, 2, 3>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 1 && G == 3) { FN<T, 3, 1, 3>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 1 && G == 0) { FN<T, 3, 1, 0>(B, Af, Ag, s, r, alpha, tol, w_mn, w_null, status); return 0; } \
  return -1;

extern "C" {
int harness_dense_f32(int n, int F, int G, int64_t B, const float* Af, const float* Ag, const float* s,
                      const float* r, const float* alpha, float tol, float* w_mn, float* w_null,
                      uint8_t* status) {
  DISPATCH(float, run_dense)
}
int harness_dense_f64(int n, int F, int G, int64_t B, const double* Af, const double* Ag, const double* s,
                      const double* r, const double* alpha, double tol, double* w_mn, double* w_null,
                      uint8_t* status) {
  DISPATCH(double, run_dense)
}
int harness_lapack_f64(int n, int F, int G, int64_t B, const double* Af, const double* Ag, const double* s,
                       const double* r, const double* alpha, double tol, double* w_mn, double* w_null,
                       uint8_t* status) {
  DISPATCH(double, run_lapack)
}
}

// ---- full step: env functor + viability terms + projection + slack integration + clipping
// params: the ParamsT<T> fields flattened in declaration order with the four int32 fields as T.
template <typename T>
static ParamsT<T> unpack(const double* f) {
  ParamsT<T> P;
  const double* p = f;
  for (int i = 0; i < 4; ++i) P.K_f[i] = static_cast<T>(*p++);
  for (int i = 0; i < 16; ++i) P.K_g[i] = static_cast<T>(*p++);
  for (int i = 0; i < 20; ++i) P.K_c[i] = static_cast<T>(*p++);
  for (int i = 0; i < 8; ++i) P.K_q[i] = static_cast<T>(*p++);
  for (int i = 0; i < 8; ++i) P.vel_max[i] = static_cast<T>(*p++);
  for (int i = 0; i < 8; ++i) P.acc_max[i] = static_cast<T>(*p++);
  P.dt = static_cast<T>(*p++);
  P.rref_tol = static_cast<T>(*p++);
  P.variant = (int32_t)*p++;
  P.bias_mode = (int32_t)*p++;
  P.clip_acc = (int32_t)*p++;
  P.basis_mode = (int32_t)*p++;
  for (int i = 0; i < 24; ++i) P.env[i] = *p++;
  return P;
}

static int g_general_path = 0;
extern "C" void harness_set_general_path(int on) { g_general_path = on; }
static int g_fix_full = 0;
extern "C" void harness_set_fix_full(int on) { g_fix_full = on; }

template <typename T, class Env>
static void run_step(int64_t B, const double* params, const T* q, const T* dq, const T* s, const T* alpha, T* ddq,
                     T* s_out, T* w_dbg, uint8_t* status, int init_only) {
  using D = typename Env::D;
  const ParamsT<T> P = unpack<T>(params);
  const int na = P.variant == VARIANT_EC ? D::n : D::k;
  for (int64_t b = 0; b < B; ++b) {
    T al[D::n];
    for (int j = 0; j < D::n; ++j) al[j] = j < na ? alpha[b * na + j] : T(0);
    if (g_general_path) {   // the fp32 structured / dense path (generic kernels, fallback of the dual path)
      RawConstraints<T, double, D> R;
      Env::template eval<T, double>(P, q + b * D::n, dq + b * D::n, R);
      if (init_only) {
        slack_from_raw<T, double, D>(P, R, s_out + b * D::G);
        continue;
      }
      status[b] = step_from_raw<T, double, D, Env::NDIAG>(P, R, dq + b * D::n, s + b * D::G, al, ddq + b * D::n,
                                                          s_out + b * D::G, w_dbg + b * 2 * D::N);
    } else {                // what the step kernels run: dual projection in double
      if (init_only) {
        RawConstraints<T, double, D> R;
        Env::template eval<T, double>(P, q + b * D::n, dq + b * D::n, R);
        slack_from_raw<T, double, D>(P, R, s_out + b * D::G);
        continue;
      }
      const DualConsts<double> Kd = make_dual_consts<T, double>(P, D::F, D::G);
      using DU = Dual<double, D, Env::NDIAG>;
      LocalStore<double, DU::Y_SIZE> Ys;
      LocalStore<double, DU::L_SIZE> Ls;
      double wmn[D::N];          // the minimum-norm part the dual path hands on for a deferred environment
      uint8_t st = step_dual<Env, T, double>(P, Kd, Ys, Ls, q + b * D::n, dq + b * D::n, s + b * D::G, al,
                                             ddq + b * D::n, s_out + b * D::G, w_dbg + b * 2 * D::N, wmn);
      if (st & ST_LAPACK_PATH) {   // what the fix-up kernel does for the environments the dual path flagged
        ArrayStore<double, Lapack<double, D>::SIZE> S;
        if (g_fix_full)          // (A/B: the LAPACK routine redoes the minimum-norm part too)
          st = step_lapack<Env, T, double>(P, Kd, S, q + b * D::n, dq + b * D::n, s + b * D::G, al, ddq + b * D::n,
                                           s_out + b * D::G, w_dbg + b * 2 * D::N);
        else
          st = step_lapack_null<Env, T, double>(P, Kd, S, q + b * D::n, dq + b * D::n, s + b * D::G, al, wmn,
                                                ddq + b * D::n, s_out + b * D::G, w_dbg + b * 2 * D::N);
      }
      status[b] = st;
    }
  }
}

template <typename T>
static int step_dispatch(int env, int64_t B, const double* params, const T* q, const T* dq, const T* s, const T* alpha,
                         T* ddq, T* s_out, T* w_dbg, uint8_t* status, int init_only) {
  switch (env) {
    case 0: run_step<T, CircleEnv>(B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only); return 0;
    case 1: run_step<T, PlanarEnv>(B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only); return 0;
    case 2: run_step<T, IiwaEnv<6>>(B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only); return 0;
    case 3: run_step<T, IiwaEnv<7>>(B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only); return 0;
  }
  return -1;
}

template <typename T, int G>
static void run_point(int64_t B, const double* params, const T* q, const T* dq, const T* p, const T* dp, const T* s,
                      const T* act, T* w, T* s_out, T* w_dbg, uint8_t* status, int init_only) {
  const ParamsT<T> P = unpack<T>(params);
  for (int64_t b = 0; b < B; ++b) {
    if (init_only) {
      PointReachEnv<G>::template slack_init<T, double>(P, q + 2 * b, p + 2 * G * b, s_out + G * b);
      continue;
    }
    status[b] = PointReachEnv<G>::template step<T, double>(P, q + 2 * b, dq + 2 * b, p + 2 * G * b, dp + 2 * G * b,
                                                  s + G * b, act + 2 * b, w + 2 * b, s_out + G * b,
                                                  w_dbg + b * 2 * (2 + G));
  }
}

template <typename T>
static int point_dispatch(int G, int64_t B, const double* params, const T* q, const T* dq, const T* p, const T* dp,
                          const T* s, const T* act, T* w, T* s_out, T* w_dbg, uint8_t* status, int init_only) {
  switch (G) {
    case 1: run_point<T, 1>(B, params, q, dq, p, dp, s, act, w, s_out, w_dbg, status, init_only); return 0;
    case 2: run_point<T, 2>(B, params, q, dq, p, dp, s, act, w, s_out, w_dbg, status, init_only); return 0;
    case 4: run_point<T, 4>(B, params, q, dq, p, dp, s, act, w, s_out, w_dbg, status, init_only); return 0;
  }
  return -1;
}

extern "C" {
int harness_step_f32(int env, int64_t B, const double* params, const float* q, const float* dq, const float* s,
                     const float* alpha, float* ddq, float* s_out, float* w_dbg, uint8_t* status, int init_only) {
  return step_dispatch<float>(env, B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only);
}
int harness_step_f64(int env, int64_t B, const double* params, const double* q, const double* dq, const double* s,
                     const double* alpha, double* ddq, double* s_out, double* w_dbg, uint8_t* status,
                     int init_only) {
  return step_dispatch<double>(env, B, params, q, dq, s, alpha, ddq, s_out, w_dbg, status, init_only);
}
int harness_point_f32(int G, int64_t B, const double* params, const float* q, const float* dq, const float* p,
                      const float* dp, const float* s, const float* act, float* w, float* s_out, float* w_dbg,
                      uint8_t* status, int init_only) {
  return point_dispatch<float>(G, B, params, q, dq, p, dp, s, act, w, s_out, w_dbg, status, init_only);
}
int harness_point_f64(int G, int64_t B, const double* params, const double* q, const double* dq, const double* p,
                      const double* dp, const double* s, const double* act, double* w, double* s_out,
                      double* w_dbg, uint8_t* status, int init_only) {
  return point_dispatch<double>(G, B, params, q, dq, p, dp, s, act, w, s_out, w_dbg, status, init_only);
}
}

// ---- structured path (with the dense fallback the kernels use when it defers)
template <typename T, int n, int F, int G, int NDIAG, int TMAX>
static void run_structured(int64_t B, const T* Af, const T* Ag, const T* s, const T* r, const T* alpha, T tol, T tau,
                           T* w_mn, T* w_null, uint8_t* status) {
  using D = Dims<n, F, G>;
  constexpr int GD = G - NDIAG;
  for (int64_t b = 0; b < B; ++b) {
    T dg[NDIAG > 0 ? NDIAG : 1];
    for (int j = 0; j < NDIAG; ++j) dg[j] = Ag[(b * G + GD + j) * n + j];
    uint8_t st = Structured<T, D, NDIAG, TMAX>::project(Af + b * F * n, Ag + b * G * n, dg, s + b * G, r + b * D::C,
                                                        alpha + b * D::k, tol, tau, true, w_mn + b * D::N,
                                                        w_null + b * D::N);
    if (st & ST_DENSE_PATH)
      st = ST_DENSE_PATH | project_dense<T, D>(Af + b * F * n, Ag + b * G * n, s + b * G, r + b * D::C,
                                               alpha + b * D::k, tol, true, w_mn + b * D::N, w_null + b * D::N);
    status[b] = st;
  }
}

#define SDISPATCH(T)                                                                                                 \
  if (n == 2 && F == 1 && G == 1) { run_structured<T, 2, 1, 1, 0, 1>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 0 && G == 6) { run_structured<T, 3, 0, 6, 3, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 6 && F == 1 && G == 11) { run_structured<T, 6, 1, 11, 6, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 7 && F == 1 && G == 12) { run_structured<T, 7, 1, 12, 7, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 2 && F == 0 && G == 4) { run_structured<T, 2, 0, 4, 0, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 4 && F == 2 && G == 3) { run_structured<T, 4, 2, 3, 0, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 1 && G == 3) { run_structured<T, 3, 1, 3, 0, 2>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  if (n == 3 && F == 1 && G == 0) { run_structured<T, 3, 1, 0, 0, 0>(B, Af, Ag, s, r, alpha, tol, tau, w_mn, w_null, status); return 0; } \
  return -1;

extern "C" {
int harness_structured_f32(int n, int F, int G, int64_t B, const float* Af, const float* Ag, const float* s,
                           const float* r, const float* alpha, float tol, float tau, float* w_mn, float* w_null,
                           uint8_t* status) {
  SDISPATCH(float)
}
int harness_structured_f64(int n, int F, int G, int64_t B, const double* Af, const double* Ag, const double* s,
                           const double* r, const double* alpha, double tol, double tau, double* w_mn,
                           double* w_null, uint8_t* status) {
  SDISPATCH(double)
}
}
