/* atacom_b200.h — C ABI of libatacom_b200.so: batched ATACOM tangent-space projection on B200 (sm_100a).
 *
 * The reference (PuzeLiu/rl_on_manifold) has no FFI layer: its boundary is the Python method
 * AtacomEnvWrapper.step_action_function(sim_state, alpha) (atacom/atacom.py:123-139) that a base
 * environment calls back once per simulator sub-step, one environment at a time, on NumPy float64
 * vectors.  Each entry point below replaces that call (and the helpers it reaches) for a BATCH of B
 * independent environments of one family; INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - all array pointers are caller-owned DEVICE memory, fp32, contiguous row-major [B, dim];
 *    arithmetic: constraint values / kinematics and the whole projection in FP64 (the reference
 *    multiplies the residual c + K J dq + s^2/2, a cancellation of O(1) terms, by K_c ~ 100-240, and the
 *    dual Gram matrix squares cond(Jc)); fp32 only for the I/O, the acceleration limits and (in the
 *    JDOT_QDOT mode) the velocity-product recursion;
 *    `*_host` entry points take HOST pointers instead and do the copies themselves;
 *  - `stream` is a cudaStream_t (NULL = default stream); calls are asynchronous on it; there is no
 *    global mutable state, so calls on different streams are independent;
 *  - s_in and s_out may alias (the reference integrates the slack in place, atacom.py:135);
 *  - `status` (optional, may be NULL) receives one ATACOM_ST_* bit set per environment;
 *  - `w_dbg` (optional, may be NULL) receives [B, 2*(n+G)]: the min-norm part -Jc^+(psi + K_c c)
 *    (= _act_a + _act_err of the reference) followed by the null-space part Nc alpha (= _act_b);
 *  - return value: ATACOM_OK or a negative ATACOM_ERR_* code; never throws, never aborts.
 */
#ifndef ATACOM_B200_H_
#define ATACOM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATACOM_MAX_Q 8   /* dim_q                      */
#define ATACOM_MAX_F 4   /* equality rows              */
#define ATACOM_MAX_G 16  /* inequality rows / slacks   */
#define ATACOM_MAX_C 20  /* F + G                      */
#define ATACOM_ENV_PARAMS 24
#define ATACOM_MAX_PEERS 8 /* GPUs of one NVSwitch domain */

enum {
  ATACOM_OK = 0,
  ATACOM_ERR_NULL_POINTER = -1,
  ATACOM_ERR_BAD_DIMS = -2,
  ATACOM_ERR_BAD_PARAM = -3,
  ATACOM_ERR_CUDA = -4,
  ATACOM_ERR_NO_DEVICE = -5,
  ATACOM_ERR_ALIGNMENT = -6
};

/* per-environment status bits */
enum {
  ATACOM_ST_RANK_DEFICIENT = 1, /* Jc lost row rank: pinv_null would cut a singular value           */
  ATACOM_ST_COLUMN_DROPPED = 2, /* the tolerance branch of rref fired (null_space_coordinate.py:60) */
  ATACOM_ST_SLACK_PIVOT = 4,    /* a slack column became a tangent coordinate                       */
  ATACOM_ST_NONFINITE = 8,
  ATACOM_ST_DENSE_PATH = 16,    /* generic / point-reach kernels: the structured fast path deferred to the dense
                                   Householder path (the step kernels of the four families never set it)       */
  ATACOM_ST_LAPACK_PATH = 32    /* inside the tolerance band of rref: projected with the reference's LAPACK null
                                   basis (ATACOM_BASIS_LAPACK) instead of the basis-free fast path            */
};

enum { ATACOM_VARIANT_ATACOM = 0, ATACOM_VARIANT_ERROR_CORRECTION = 1 };
/* Null basis handed to rref(tol) (atacom.py:127-128).  Where the tolerance branch of rref does not fire the result
 * does not depend on it.  Where it fires (an active constraint: 15-25 % of the synthetic benchmark batches) it does:
 *   LAPACK     (default, 0): the reference's own basis — SciPy svd / LAPACK gesdd returns, for a full-row-rank Jc,
 *              the trailing columns of the product of the right Householder reflectors of the bidiagonalisation
 *              (dgebd2; dgelq2 when N >= 11 C / 6): a deterministic function of Jc, reproduced here
 *              (csrc/atacom_lapack.cuh), so the output equals the reference's on both strata.  The step kernels
 *              run the basis-free fast path everywhere and redo the environments inside the tolerance band
 *              (flagged ATACOM_ST_LAPACK_PATH) with this basis in a second, compacted launch;
 *   CANONICAL  (1): Gram-Schmidt of the projected unit vectors in column order — a function of null(Jc) alone;
 *              one launch, ~40 % faster, equal to the reference's output only where the tolerance branch stays
 *              silent. */
enum { ATACOM_BASIS_LAPACK = 0, ATACOM_BASIS_CANONICAL = 1 };
/* b(q, dq) of the Cartesian rows.  OMEGA_X_V (default of the *_default_params): what the reference executes —
 * pinocchio's classical frame acceleration after a FIRST-order forwardKinematics, data.a = 0, i.e. omega x v
 * (iiwa_hit_atacom.py:87-89,122-128; atacom_air_hockey.py:94-96).  JDOT_QDOT: the full dJ/dt dq the paper
 * defines (constraints.py:19-20), opt-in. */
enum { ATACOM_BIAS_JDOT_QDOT = 0, ATACOM_BIAS_OMEGA_X_V = 1 };

/* Constructor arguments of AtacomEnvWrapper (atacom/atacom.py:10-71) resolved to arrays, plus the
 * per-family constants the reference reads from its simulator / URDF. */
typedef struct AtacomParams {
  float K_f[ATACOM_MAX_F];     /* ViabilityConstraint.K of the equality rows   (constraints.py:26-29) */
  float K_g[ATACOM_MAX_G];     /* ... of the inequality rows                                          */
  float K_c[ATACOM_MAX_C];     /* error-correction gain                        (atacom.py:42-45)      */
  float K_q[ATACOM_MAX_Q];     /* viability acceleration-bound gain            (atacom.py:65-69)      */
  float vel_max[ATACOM_MAX_Q]; /* (atacom.py:53-57) */
  float acc_max[ATACOM_MAX_Q]; /* (atacom.py:59-63) */
  float dt;                    /* time_step                                    (atacom.py:28)         */
  float rref_tol;              /* 0.05 in the wrapper                          (atacom.py:128)        */
  int32_t variant;             /* ATACOM_VARIANT_*                                                    */
  int32_t bias_mode;           /* ATACOM_BIAS_* for the Cartesian rows of planar / iiwa               */
  int32_t clip_acc;            /* 1: apply acc_truncation (atacom.py:117-121)                         */
  int32_t basis_mode;          /* ATACOM_BASIS_*: which null basis the tolerance-RREF sees (0 = the reference's) */
  double env[ATACOM_ENV_PARAMS]; /* geometry constants, double because K_c amplifies their rounding:
                                    planar: l1 l2 l3 base_x base_y qmax[3] half_len half_wid;
                                    iiwa: base_x half_len half_wid height z4_min z7_min qmax[7];
                                    circle: action_scale base_dt goal_x goal_y;
                                    point_reach: radius^2 K K_c action_scale base_dt goal_x goal_y wall_lo
                                    wall_hi obst_lo obst_hi obst_action_scale obst_vel_max obst_circle_radius */
} AtacomParams;

const char* atacom_version(void);
const char* atacom_error_string(int code);
/* number of kernels this library has launched so far in this process (for benchmarks' bookkeeping) */
int64_t atacom_launch_count(void);

/* Fill `p` with the constructor defaults of the reference's env classes:
 *  circle       CircleEnvAtacom            circle_atacom.py:8-18
 *  planar       AirHockeyPlanarAtacom      atacom_air_hockey.py:28-43   (URDF constants unverified)
 *  iiwa         AirHockeyIiwaAtacom        iiwa_hit_atacom.py:23-40, env_base.py:50,147-159, urdf/iiwa_1.urdf
 *  point_reach  PointReachAtacom           collision_avoidance_atacom.py:9-16 */
int atacom_circle_default_params(AtacomParams* p);
int atacom_planar_default_params(AtacomParams* p);
int atacom_iiwa_default_params(AtacomParams* p, int n_ctrl_joints /* 6 or 7 */);
int atacom_point_reach_default_params(AtacomParams* p);

/* ---- AtacomEnvWrapper.step_action_function (atacom.py:123-139) /
 *      ErrorCorrectionEnvWrapper.step_action_function (error_correction_wrapper.py:117-134) ----
 * q, dq: [B, n]; s_in, s_out: [B, G]; alpha: [B, k] (variant ATACOM; already scaled by alpha_max,
 * atacom.py:107-108) or [B, n] (variant ERROR_CORRECTION); ddq: [B, n] clipped joint acceleration
 * (the argument the reference hands to acc_to_ctrl_action, atacom.py:137-138).
 *  circle: n=2 F=1 G=1 k=1     planar: n=3 F=0 G=6 k=3     iiwa: n=6|7 F=1 G=5+n k=n-1 */
int atacom_circle_step(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq,
                       float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                       void* stream);
int atacom_planar_step(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq,
                       float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                       void* stream);
int atacom_iiwa_step(int n_ctrl_joints, const float* q, const float* dq, const float* s_in, const float* alpha,
                     float* ddq, float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                     void* stream);

/* ---- the same step with the per-step gather fused into its epilogue (multi-GPU) ----
 * The batch is sharded by environment; every rank needs all projected accelerations after each step.
 * Instead of a separate all-gather, the kernel stores each environment's ddq row straight into every
 * rank's gather buffer: peer_ddq[w] is rank w's [world * B, n] buffer mapped into this process (NVLink
 * peer / symmetric memory; peer_ddq[rank] is the local one), this rank's rows start at row_offset.
 * ddq (the local [B, n] copy) may be NULL.  The caller synchronises the ranks afterwards (one barrier). */
int atacom_iiwa_step_gather(int n_ctrl_joints, const float* q, const float* dq, const float* s_in,
                            const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B,
                            const AtacomParams* p, void* stream, float* const* peer_ddq, int world,
                            int64_t row_offset);

/* The same with the cross-rank synchronisation inside the kernel as well: when this launch has completed on a
 * rank, the rows of EVERY rank are in that rank's gather buffer.  peer_flags[w] is rank w's flag array
 * (uint32 [world], zero-initialised once, in peer-mapped memory like the gather buffers); local_sync points to
 * 2 zero-initialised uint32 in this rank's own device memory (block counter, step sequence number).  The last
 * block of the launch publishes the step to every rank (release, system scope) and waits for every rank's
 * flag of the same step (acquire).  Every rank must issue the same sequence of launches; safe to capture in
 * a CUDA graph (the sequence number lives in device memory). */
int atacom_iiwa_step_gather_sync(int n_ctrl_joints, const float* q, const float* dq, const float* s_in,
                                 const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B,
                                 const AtacomParams* p, void* stream, float* const* peer_ddq, int world,
                                 int64_t row_offset, uint32_t* const* peer_flags, uint32_t* local_sync, int rank);

/* ---- K simulator sub-steps of one agent step in one launch (PyBullet-style environments) ----
 * The base environment calls the hook once per simulator sub-step (env_base.py:161-165; MushroomRL's
 * n_intermediate_steps = 4) while q, dq stay what the last step() set (atacom.py:111-112,123-126): K projections
 * on the same (q, dq, alpha) that differ only through the slacks each call integrates (atacom.py:135).
 * ddq: [K, B, n], row block k is the acceleration of sub-step k; s_out: [B, G] the slacks after the K-th call.
 * workspace: device memory for atacom_iiwa_substeps_workspace_doubles(n) * B doubles (the kinematics' products
 * are parked there after the first sub-step), or NULL (the kinematics are recomputed every sub-step). */
int atacom_iiwa_substeps_workspace_doubles(int n_ctrl_joints);
int atacom_iiwa_step_substeps(int n_ctrl_joints, int K, const float* q, const float* dq, const float* s_in,
                              const float* alpha, float* ddq, float* s_out, uint8_t* status, double* workspace,
                              int64_t B, const AtacomParams* p, void* stream);

/* ---- AtacomEnvWrapper._compute_slack_variables (atacom.py:145-149): s = sqrt(max(-2 g~, 0)) ----
 * mask (optional, may be NULL): uint8 [B]; only environments with mask != 0 are re-initialised. */
int atacom_circle_slack_init(const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                             const AtacomParams* p, void* stream);
int atacom_planar_slack_init(const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                             const AtacomParams* p, void* stream);
int atacom_iiwa_slack_init(int n_ctrl_joints, const float* q, const float* dq, float* s, const uint8_t* mask,
                           int64_t B, const AtacomParams* p, void* stream);

/* ---- PointReachAtacom.step / reset (collision_avoidance_atacom.py:18-48) ----
 * q, dq: [B, 2]; obs_p, obs_dp: [B, 2*n_objects]; s: [B, n_objects]; action: [B, 2] (unscaled,
 * collision_avoidance_atacom.py:46); w: [B, 2] the projected action handed to PointGoalReach.step
 * (before its clip and x10, collision_avoidance_base.py:44-45).  n_objects in {1,2,3,4,6,8}. */
int atacom_point_reach_step(int n_objects, const float* q, const float* dq, const float* obs_p,
                            const float* obs_dp, const float* s_in, const float* action, float* w,
                            float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                            void* stream);
int atacom_point_reach_slack_init(int n_objects, const float* q, const float* obs_p, float* s,
                                  const uint8_t* mask, int64_t B, const AtacomParams* p, void* stream);

/* ---- constraint statistics (AtacomEnvWrapper._update_constraint_stats / get_constraints_logs, atacom.py:201-216) ----
 * Per environment c_i = max(|c_f(q)|, c_g(q)) (origin constraints) and c_dq_i = max_j(|dq_j| - vel_max_j).
 * per_env: [B, 2] = (c_i, c_dq_i) or NULL; stats: 4 doubles in device memory or NULL, updated atomically:
 * { sum of c_i, max of c_i, max of c_dq_i, number of samples } — initialise to { 0, -inf, -inf, 0 }; after an epoch
 * (c_avg, c_max, c_dq_max) = (stats[0] / stats[3], stats[1], stats[2]). */
int atacom_circle_constraint_stats(const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                                   const AtacomParams* p, void* stream);
int atacom_planar_constraint_stats(const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                                   const AtacomParams* p, void* stream);
int atacom_iiwa_constraint_stats(int n_ctrl_joints, const float* q, const float* dq, float* per_env, double* stats,
                                 int64_t B, const AtacomParams* p, void* stream);

/* ---- fused roll-outs of the two environments whose base dynamics are plain arithmetic ----
 * T agent steps per launch with the state in registers: action scaling, projection, base-env integration,
 * reward and the constraint log, in double precision throughout (state, slack and rewards are stored fp32).
 *
 * atacom_circle_rollout: CircleEnvAtacom / CircleEnvErrorCorrection.step for T steps
 *   (atacom.py:106-115, circle_base.py:53-67,86-115, circle_atacom.py:26-27).  state: [B, 4] (x, y, dx, dy),
 *   s: [B, 1], actions: [T, B, 1] (variant ATACOM) or [T, B, 2] (ERROR_CORRECTION), raw agent actions (clipped to
 *   [-1, 1] and scaled here); rewards: [T, B] or NULL; stats as above, with the circle's own log
 *   (c = max(|x^2+y^2-1|, -y-0.5), c_dq = max(|dx|, |dy|) - 1, taken BEFORE each step).
 *
 * atacom_point_reach_rollout: PointReachAtacom.step for T steps (collision_avoidance_atacom.py:29-48,
 *   collision_avoidance_base.py:41-76).  state: [B, 4 + 4 G] (agent x y dx dy, then per obstacle x y dx dy),
 *   s: [B, G], actions: [T, B, 2]; obstacle_draws: [T, B, 2 G] U(-1, 1) draws of the obstacles' random walk
 *   (the reference draws them with np.random.uniform inside step), or NULL with obstacle_centers [B, 2 G] for the
 *   circling obstacles starting at time0, or both NULL for static obstacles; stats[2] stays untouched
 *   (the reference logs c_dq = 0). */
int atacom_circle_rollout(float* state, float* s, const float* actions, float* rewards, double* stats,
                          uint8_t* status, int64_t B, int T, const AtacomParams* p, void* stream);
int atacom_point_reach_rollout(int n_objects, float* state, float* s, const float* actions,
                               const float* obstacle_draws, const float* obstacle_centers, double time0,
                               float* rewards, double* stats, uint8_t* status, int64_t B, int T,
                               const AtacomParams* p, void* stream);

/* ---- generic ConstraintsSet (atacom/constraints.py:46-82) ----
 * For a user-defined ConstraintsSet whose callbacks were evaluated batched by the caller:
 * c: [B, C] = fun(q) (origin), J: [B, C, n], b: [B, C] = b_state(q, dq); equality rows first.
 * Supported (n, F, G): see atacom_generic_supported(). */
int atacom_generic_supported(int n, int F, int G);
int atacom_generic_step(int n, int F, int G, const float* c, const float* J, const float* b, const float* dq,
                        const float* s_in, const float* alpha, float* ddq, float* s_out, uint8_t* status,
                        float* w_dbg, int64_t B, const AtacomParams* p, void* stream);

/* ---- host-buffer entry points (what a NumPy caller of the reference binds) ----
 * Same semantics as the device entry points but every array pointer is HOST memory (pinned memory makes
 * the copies asynchronous).  The context owns the device staging buffers and streams; the batch is
 * cut into `chunks` pieces whose H2D copy, kernel and D2H copy overlap.
 *
 * Data paths (atacom_host_ctx_set_mode); the pipeline of a call is captured in a CUDA graph and replayed while
 * the caller keeps passing the same buffers:
 *   ATACOM_HOST_STAGED     device staging buffers, copy engines in both directions;
 *   ATACOM_HOST_HYBRID     inputs through the copy engines, outputs stored by the kernels straight into the
 *                          caller's buffers over PCIe (the OUTPUT buffers must be page-locked and mapped:
 *                          cudaHostAlloc / cudaHostRegister / torch pin_memory());
 *   ATACOM_HOST_ZERO_COPY  one kernel reads and writes the caller's buffers directly (all of them mapped); its
 *                          bulk loads are admitted a window of warps at a time in block-start order, so the first
 *                          blocks compute and store while the last still wait for PCIe (environment overrides for
 *                          experiments: ATACOM_ZC_WINDOW warps, 0 = all at once; ATACOM_ZC_TPB threads per block);
 *   ATACOM_HOST_AUTO       (default) zero-copy when every buffer of the call is mapped, else staged.
 * Measured on PCIe gen5 x16, 65 536 iiwa environments per call: zero-copy 188 us (231 us with all loads issued at
 * once), staged 262 us, hybrid 265 us. */
typedef struct AtacomHostCtx AtacomHostCtx;
enum { ATACOM_HOST_AUTO = 0, ATACOM_HOST_STAGED = 1, ATACOM_HOST_ZERO_COPY = 2, ATACOM_HOST_HYBRID = 3 };
int atacom_host_ctx_create(AtacomHostCtx** ctx, int64_t max_B, int chunks);
int atacom_host_ctx_set_mode(AtacomHostCtx* ctx, int mode);
int atacom_host_ctx_destroy(AtacomHostCtx* ctx);
int atacom_iiwa_step_host(AtacomHostCtx* ctx, int n_ctrl_joints, const float* q, const float* dq,
                          const float* s_in, const float* alpha, float* ddq, float* s_out, uint8_t* status,
                          int64_t B, const AtacomParams* p);
/* The other families, same context, same argument meaning as their device entry points (w_dbg omitted):
 *   circle / planar: what a NumPy caller of CircleEnvAtacom (circle_atacom.py) / AirHockeyPlanarAtacom
 *   (atacom_air_hockey.py) binds — all four data paths;
 *   point_reach: PointReachAtacom.step (collision_avoidance_atacom.py:29-48); generic: any ConstraintsSet whose
 *   callbacks the caller evaluated batched in NumPy (constraints.py:46-82) — staged data path only
 *   (ATACOM_HOST_AUTO or ATACOM_HOST_STAGED; the other modes return ATACOM_ERR_BAD_PARAM). */
int atacom_circle_step_host(AtacomHostCtx* ctx, const float* q, const float* dq, const float* s_in,
                            const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B,
                            const AtacomParams* p);
int atacom_planar_step_host(AtacomHostCtx* ctx, const float* q, const float* dq, const float* s_in,
                            const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B,
                            const AtacomParams* p);
int atacom_point_reach_step_host(AtacomHostCtx* ctx, int n_objects, const float* q, const float* dq,
                                 const float* obs_p, const float* obs_dp, const float* s_in, const float* action,
                                 float* w, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p);
int atacom_generic_step_host(AtacomHostCtx* ctx, int n, int F, int G, const float* c, const float* J,
                             const float* b, const float* dq, const float* s_in, const float* alpha, float* ddq,
                             float* s_out, uint8_t* status, int64_t B, const AtacomParams* p);

/* Device-side waits (ordered admission of the zero-copy path, in-kernel barrier of atacom_iiwa_step_gather_sync)
 * are bounded by a clock budget (~2 s): a wait that exceeds it gives up and is counted.  *count = number of
 * such time-outs on the current device since the library was loaded (0 in a healthy run).  Synchronises. */
int atacom_spin_timeouts(unsigned* count);

#ifdef __cplusplus
}
#endif
#endif /* ATACOM_B200_H_ */
