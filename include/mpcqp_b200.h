/* mpcqp_b200.h — C ABI of the B200-native batched MPC QP engine.
 *
 * This is the FFI line a maintainer of kotakondo/Intent-MPC binds instead of OSQP's C API
 * (reference: trajectory_planner/include/trajectory_planner/third_party/osqp/osqp.h:41-400, used through
 * OsqpEigen::Solver at trajectory_planner/include/trajectory_planner/mpcPlanner.cpp:436-527).
 * Plain C types only; no torch, no Eigen.  All functions return 0 on success, a negative mpcqp_error
 * otherwise; mpcqp_engine_last_error() gives the message.  There is NO CPU fallback: every solve runs the
 * sm_100a kernels, and engine creation fails if no CUDA device is usable.
 *
 * Three groups:
 *   (1) engine lifecycle, one engine per GPU / host thread;
 *   (2) the batched entry point: B independent mpcPlanner control-step QPs are ASSEMBLED ON THE DEVICE from
 *       the planner's own inputs (current state, reference window, per-stage obstacle ellipsoids,
 *       linearisation point, warm start) and solved, one CTA per QP (csrc/mpcqp_core.cuh);
 *   (3) an OSQP-shaped single-problem set (setup / warm_start / solve / info / solution / cleanup) taking
 *       explicit CSC matrices, which is what the OsqpEigen::Solver-shaped C++ facade
 *       (intent-mpc_b200/host/OsqpEigenB200.hpp) calls.
 */
#ifndef MPCQP_B200_H
#define MPCQP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpcqp_engine mpcqp_engine;
typedef struct mpcqp_problem mpcqp_problem;

enum mpcqp_error {
  MPCQP_OK = 0,
  MPCQP_ERR_CUDA = -1,          /* a CUDA call failed (no device, out of memory, launch failure) */
  MPCQP_ERR_ARG = -2,           /* null pointer / bad size */
  MPCQP_ERR_DATA = -3,          /* OSQP_DATA_VALIDATION_ERROR analogue (osqp constants.h:41-49): l > u, bad dims */
  MPCQP_ERR_SETTINGS = -4,      /* OSQP_SETTINGS_VALIDATION_ERROR analogue, or a setting the engine does not implement */
  MPCQP_ERR_STRUCTURE = -5,     /* CSC problem has no mpcPlanner stage structure AND is too large for the dense generic kernel (n + m > 4096) */
  MPCQP_ERR_NOT_INIT = -7       /* OSQP_WORKSPACE_NOT_INIT_ERROR analogue */
};

/* Solver status values — identical to OSQP's (third_party/osqp/constants.h:18-30). */
enum mpcqp_status {
  MPCQP_DUAL_INFEASIBLE_INACCURATE = 4, MPCQP_PRIMAL_INFEASIBLE_INACCURATE = 3, MPCQP_SOLVED_INACCURATE = 2,
  MPCQP_SOLVED = 1, MPCQP_MAX_ITER_REACHED = -2, MPCQP_PRIMAL_INFEASIBLE = -3, MPCQP_DUAL_INFEASIBLE = -4,
  MPCQP_NON_CVX = -7, MPCQP_UNSOLVED = -10
};

/* Mirrors OSQPSettings (third_party/osqp/types.h:139-176).  Fields the engine does not implement must keep
 * their default or setup fails with MPCQP_ERR_SETTINGS: polish (0), scaled_termination (0), time_limit (0).
 * Defaults are OSQP's (constants.h:59-118) except the two determinism pins of SURVEY.md §8(c):
 * adaptive_rho_interval = 25 (OSQP's 0 derives it from wall-clock time) and time_limit = 0. */
typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance, adaptive_rho_fraction, delta, time_limit;
  int64_t max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination, warm_start;
  int64_t scaled_termination, polish, polish_refine_iter, verbose;
} mpcqp_settings;

/* Replaces osqp_set_default_settings (osqp.h:41). */
void mpcqp_set_default_settings(mpcqp_settings* s);

/* Planner parameters = mpcPlanner::initParam keys (mpcPlanner.cpp:19-173) that shape the QP. */
typedef struct {
  int32_t horizon;                 /* stages; mpcWindow N = horizon-1 (mpcPlanner.cpp:382) */
  double ts;
  double max_vel, max_acc;         /* updateMaxVel / updateMaxAcc */
  double y_min, y_max, z_min, z_max;
  double static_safety_dist, dynamic_safety_dist;
  double static_slack, dynamic_slack;
  double position_weight, velocity_weight, acceleration_weight;
} mpcqp_mpc_params;

/* intent_mpc_demo defaults (autonomous_flight/cfg/mpc_navigation/planner_param.yaml:25-39). */
void mpcqp_default_mpc_params(mpcqp_mpc_params* p);

/* ---- (1) engine lifecycle ------------------------------------------------------------------------- */
/* Number of usable CUDA devices (0 when there is none: every other call then fails; there is no CPU path). */
int mpcqp_device_count(void);
int mpcqp_engine_create(int device, mpcqp_engine** out);
int mpcqp_engine_destroy(mpcqp_engine* e);
const char* mpcqp_engine_last_error(const mpcqp_engine* e);
/* Device time (ms, CUDA events on the engine's stream) of the kernels of the last batch call, and how many
 * kernels it launched. */
double mpcqp_engine_last_kernel_ms(const mpcqp_engine* e);
/* Same, solve kernel alone (the dominant kernel; excludes the device-side builder). */
double mpcqp_engine_last_solve_kernel_ms(const mpcqp_engine* e);
int64_t mpcqp_engine_last_launches(const mpcqp_engine* e);
/* Which solve kernel the last batch ran: 2 = CTA kernel (one 4-warp CTA per QP, parallel-cyclic-reduction solve,
 * horizon 30, num_obs <= 8; horizons 20 and 25 without the assistant variant), 1 = one-warp-per-QP register-resident kernel
 * (horizon 30), 0 = generic one-warp
 * shared-memory kernel (any horizon/num_obs that fits); for problems WITHOUT the mpcPlanner stage structure (polyTrajSolver's
 * minimum-snap QPs): 5 = sparse generic kernel (banded L D L' of the reverse-Cuthill-McKee-ordered KKT matrix, one warp per QP;
 * taken when the half-bandwidth is <= 31), 4 = dense generic kernel (anything else up to n + m = 4096).  force_generic(4) pins the
 * dense kernel for unstructured problems (A/B tests).  force_generic(1) pins the generic kernel, (2) the
 * one-warp register kernel, (3) the CTA kernel without its assistant warps (launches that give a CTA an SM to itself
 * normally carry three more warps that hold the PCR matrices of levels 1..3 in registers), (0) restores the default
 * dispatch (used by the tests to cover all of them). */
int mpcqp_engine_last_path(const mpcqp_engine* e);
int mpcqp_engine_force_generic(mpcqp_engine* e, int on);
/* Scheduling hint (default on): the batched entry point remembers how many iterations each batch slot took and, when
 * the next call has the same batch size and num_obs (a receding-horizon loop: slot b is the same scenario one control
 * step later), starts the slots that ran long first, one per SM.  Results never depend on it.  0 switches it off. */
int mpcqp_engine_use_history(mpcqp_engine* e, int on);
/* Migration (default on): in batches larger than the SM count, an instance that is still iterating after 300 iterations
 * parks its state and is resumed — bit-identically — by a follow-up launch (on an SM of its own for small batches), so
 * that instances nobody could predict to be long do not form the tail of the batch.  Results never depend on it. */
int mpcqp_engine_use_migration(mpcqp_engine* e, int on);
/* Layout of the obs_dyn argument of the batched entry point: 0 (default) = one [N][R] pattern shared by the batch (the
 * candidates of one control step), 1 = [B][N][R], one pattern per instance (Monte-Carlo sweeps where every instance has
 * its own mix of dynamic and static obstacles; updateObstacleParam's flags, mpcPlanner.cpp:1148-1197). */
int mpcqp_engine_obs_dyn_per_instance(mpcqp_engine* e, int on);
/* Per-instance obstacle counts for the batched entry point (Monte-Carlo sweeps: every instance sees its own number of
 * obstacles).  nobs = host array [B] with 0 <= nobs[b] <= num_obs, or NULL to switch back.  While set, num_obs is the
 * STRIDE of the obstacle arrays (instance b uses rows 0 .. nobs[b]-1 of every stage, the rest is padding that is never
 * read), its QP has m_b = 16*horizon + 5*N + nobs[b]*N constraints laid out compactly at y + b*m (m from num_obs).
 * Needs horizon 30 and 1 <= num_obs <= 32.  The pointer is read at every call until it is replaced. */
int mpcqp_engine_num_obs_per_instance(mpcqp_engine* e, const int32_t* nobs);
/* Per-instance velocity / acceleration limits: limits = host array [B][2] = (max_vel, max_acc) of each instance, replacing
 * p->max_vel / p->max_acc in the state and input boxes (updateMaxVel / updateMaxAcc per scenario, mpcPlanner.cpp:243-255), or
 * NULL to switch back.  Needs horizon 30 and num_obs <= 32 (the CTA kernels).  Read at every call until replaced. */
int mpcqp_engine_limits_per_instance(mpcqp_engine* e, const double* limits);
/* Wait for the engine's stream (needed after a *_device call before reading results / last_kernel_ms). */
int mpcqp_engine_sync(mpcqp_engine* e);
/* FP64 FMA-pipe microbenchmark (all SMs, 8 independent DFMA chains per thread): the measured roofline
 * denominator for the solve kernel, in TFLOP/s. */
int mpcqp_fp64_fma_peak(mpcqp_engine* e, double* tflops);
/* The engine's CUDA stream (cudaStream_t) so callers can order their own work against it. */
void* mpcqp_engine_stream(const mpcqp_engine* e);

/* ---- (2) batched entry point ---------------------------------------------------------------------- *
 * Replaces, for B instances at once, mpcPlanner::solveTraj's assemble + initSolver + setWarmStart +
 * solveProblem + getSolution sequence (mpcPlanner.cpp:375-541).  Array shapes (row-major, N = horizon-1,
 * R = num_obs, n = 8*horizon + 5*N, m = 16*horizon + 5*N + R*N):
 *   x0       [B][6]        current position, velocity            (updateCurrStates, mpcPlanner.cpp:257-263)
 *   xref     [B][N+1][3]   reference positions                   (getXRef, mpcPlanner.cpp:968-981)
 *   obs_c    [B][N][R][3]  obstacle centre per stage             (updateObstacleParam, mpcPlanner.cpp:1148-1197)
 *   obs_semi [B][N][R][3]  semi-axes = size/2 + safety distance
 *   obs_yaw  [B][N][R]
 *   obs_dyn  [N][R] int32  1: row is softened by slack input 3 (dynamic), 0: by slack input 4 (static); HOST pointer in
 *                          both forms; [B][N][R] after mpcqp_engine_obs_dyn_per_instance(e, 1)
 *   lin_pt   [B][N][3]     linearisation point (previous plan, unshifted, or current position; :1042-1051)
 *   warm_x   [B][n] or NULL  primal warm start (previous plan; dual warm start is always 0, :487)
 * Outputs: x [B][n]; y [B][m] or NULL; status/iter/rho_updates [B] int32; obj/pri_res/dua_res [B].
 * The *_host form takes host pointers (copies in and out on the engine's stream, synchronous on return);
 * the *_device form takes device pointers and is asynchronous on the engine's stream. */
int mpcqp_solve_mpc_batch_host(mpcqp_engine* e, const mpcqp_mpc_params* p, const mpcqp_settings* s, int32_t B,
                               int32_t num_obs, const double* x0, const double* xref, const double* obs_c,
                               const double* obs_semi, const double* obs_yaw, const int32_t* obs_dyn,
                               const double* lin_pt, const double* warm_x, double* x, double* y, int32_t* status,
                               int32_t* iter, int32_t* rho_updates, double* obj, double* pri_res, double* dua_res);
int mpcqp_solve_mpc_batch_device(mpcqp_engine* e, const mpcqp_mpc_params* p, const mpcqp_settings* s, int32_t B,
                                 int32_t num_obs, const double* x0, const double* xref, const double* obs_c,
                                 const double* obs_semi, const double* obs_yaw, const int32_t* obs_dyn,
                                 const double* lin_pt, const double* warm_x, double* x, double* y, int32_t* status,
                                 int32_t* iter, int32_t* rho_updates, double* obj, double* pri_res,
                                 double* dua_res);

/* ---- (2b) candidate scoring and selection ----------------------------------------------------------- *
 * Device pointers, asynchronous on the engine's stream.  Replace, for the candidates of many scenarios at once,
 * getTrajectoryScore (consistency / detour / safety, mpcPlanner.cpp:771-852) and evaluateTraj (:854-887).
 *   x         [B][n]        candidate solutions from the batched entry point
 *   prev_plan [B][n] / NULL the plan each candidate is compared with (currentStatesSol_; the warm start array); NULL on
 *                           the first control step (consistency score 0, :783-785)
 *   xref, obs_c, obs_semi   as given to the solve; the first n_dynamic rows of a stage are dynamic obstacles (full-size
 *                           xy diagonal + dynamicSafetyDist_), the rest static (half-size diagonal + staticSafetyDist_)
 *   obs_c_last, obs_semi_last [B][R][3]  the same for stage N: getSafetyScore walks every state 0..N (mpcPlanner.cpp:818-826)
 *                           while the QP only has obstacle rows for stages 0..N-1; required when num_obs > 0
 *   score     [B][3]        (consistency, detour, safety)
 * select: scenario s has C candidates, rows cand[s][c] of `score` / `x_all`, with weights weight[s][c] (the intent
 * probabilities in the order evaluateTraj indexes them).  best[s] = argmax_c weight * (avg_c/c_c + avg_d/d_c + s_c/avg_s),
 * first maximum, NaN never wins; weighted [S][C] (optional) receives the values; plan [S][n] (optional) the chosen x. */
/* Candidate enumeration (getIntentComb + findClosestObstacle, mpcPlanner.cpp:663-769) for S scenarios with D predicted
 * obstacles each:  pred_pos / pred_size [S][D][4 intents][NP >= N][3] (intent order FORWARD, LEFT, RIGHT, STOP of
 * dynamic_predictor/utils.h:15-20), prob [S][D][4], prev_plan [S][n] or NULL on the first step (then pos [S][3] is used).
 * The six hypotheses of a scenario, sorted by descending weight, become rows of two solve batches: scen_a [4S] / obs_c_a,
 * obs_semi_a [4S][N][D][3] (one intent for the closest obstacle) and scen_b [2S] / obs_c_b, obs_semi_b [2S][N][D+1][3] (two
 * intents: the closest obstacle twice); obs_c_last_a, obs_semi_last_a [4S][D][3] and obs_c_last_b, obs_semi_last_b
 * [2S][D+1][3] (all four or none; need NP >= horizon) receive the predictions of stage N for the scoring; cand [S][6] = row of sorted candidate c in the concatenation (batch a, then b),
 * weight [S][6] as evaluateTraj indexes it.  mpcqp_gather_rows_device replicates scenario-level arrays (x0, xref, lin_pt,
 * warm_x) per row: dst[b] = src[idx[b]]. */
int mpcqp_intent_candidates_device(mpcqp_engine* e, const mpcqp_mpc_params* p, int32_t S, int32_t D, int32_t NP,
                                   const double* pred_pos, const double* pred_size, const double* prob, const double* prev_plan,
                                   const double* pos, int32_t* scen_a, int32_t* scen_b, double* obs_c_a, double* obs_semi_a,
                                   double* obs_c_b, double* obs_semi_b, double* obs_c_last_a, double* obs_semi_last_a,
                                   double* obs_c_last_b, double* obs_semi_last_b, double* weight, int32_t* cand);
int mpcqp_gather_rows_device(mpcqp_engine* e, int64_t B, int32_t width, const int32_t* idx, const double* src, double* dst);
int mpcqp_score_candidates_device(mpcqp_engine* e, const mpcqp_mpc_params* p, int32_t B, int32_t num_obs, int32_t n_dynamic,
                                  const double* x, const double* prev_plan, const double* xref, const double* obs_c,
                                  const double* obs_semi, const double* obs_c_last, const double* obs_semi_last, double* score);
int mpcqp_select_candidates_device(mpcqp_engine* e, int32_t S, int32_t C, int32_t n, const int32_t* cand,
                                   const double* weight, const double* score, const double* x_all, int32_t* best,
                                   double* weighted, double* plan);

/* ---- (2c) predictor rollouts --------------------------------------------------------------------------- *
 * The step immediately upstream of the planner: dynamicPredictor::predictor::predict (dynamic_predictor/include/
 * dynamic_predictor/dynamicPredictor.cpp:162-195) = intentProb (:197-281) + predTraj (:283-541: forward / turning / stop
 * sampling, mean path, box inflated by 2 sqrt(var) z) for num_obstacles obstacles at once; its outputs are exactly the
 * arguments of updatePredObstacles (mpcPlanner.cpp:343-373) / mpcqp_intent_candidates_device.  Device pointers,
 * asynchronous on the engine's stream.
 *   pos_hist, vel_hist [num_obstacles][num_hist][3]   history, index 0 = newest (getDynamicObstaclesHist)
 *   size               [num_obstacles][3]             current box size (robot size already added, fakeDetector.cpp:540)
 *   pred_pos, pred_size [num_obstacles][4][prediction_size + 1][3]   intent order FORWARD, LEFT, RIGHT, STOP
 *   intent_prob        [num_obstacles][4]
 * The occupancy map the reference consults while sampling (isInflatedOccupied) is taken as free space. */
typedef struct {
  int32_t prediction_size;                /* predictor_param.yaml keys, defaults of intent_mpc_demo */
  double prediction_time_step, min_turning_time, max_turning_time, prediction_z_score;
  double max_front_prob, front_angle_deg, stop_velocity_threshold, prob_scale_param;
} mpcqp_predictor_params;
void mpcqp_default_predictor_params(mpcqp_predictor_params* p);
int mpcqp_predict_device(mpcqp_engine* e, const mpcqp_predictor_params* pp, int32_t num_obstacles, int32_t num_hist,
                         const double* pos_hist, const double* vel_hist, const double* size, double* pred_pos,
                         double* pred_size, double* intent_prob);

/* ---- (3) OSQP-shaped single problem (explicit CSC) -------------------------------------------------- *
 * mpcqp_setup replaces osqp_setup (osqp.h:58): data is copied, the caller's arrays may die afterwards.
 * P is upper-triangular CSC (n x n), A is CSC (m x n), indices int64 like OSQP's c_int (glob_opts.h:80).
 * A problem with the mpcPlanner stage structure (diagonal P; dynamics / box / obstacle row blocks as
 * mpcPlanner.cpp:989-1071 lays them out) runs on the stage-structured kernels.  Anything else — the reference's second
 * consumer of this boundary is polyTrajSolver (polyTrajSolver.cpp:14-37, 162-239, 848-900: block-diagonal minimum-snap
 * Hessian, continuity equalities, corridor boxes) — runs on the dense generic kernel (csrc/mpcqp_dense.cuh, one CTA per
 * QP, same OSQP iterate sequence) when n + m <= 4096, and returns MPCQP_ERR_STRUCTURE beyond that.  There is no CPU path. */
int mpcqp_setup(mpcqp_engine* e, mpcqp_problem** out, int64_t n, int64_t m, const int64_t* P_colptr,
                const int64_t* P_rowidx, const double* P_val, const double* q, const int64_t* A_colptr,
                const int64_t* A_rowidx, const double* A_val, const double* l, const double* u,
                const mpcqp_settings* s);
int mpcqp_warm_start(mpcqp_problem* pr, const double* x, const double* y);   /* osqp_warm_start, osqp.h:157 */
int mpcqp_warm_start_x(mpcqp_problem* pr, const double* x);                  /* osqp_warm_start_x */
int mpcqp_update_lin_cost(mpcqp_problem* pr, const double* q_new);           /* osqp_update_lin_cost */
int mpcqp_update_bounds(mpcqp_problem* pr, const double* l_new, const double* u_new); /* osqp_update_bounds */
int mpcqp_solve(mpcqp_problem* pr);                                          /* osqp_solve, osqp.h:78 */
typedef struct {                                                             /* OSQPInfo subset, types.h:66-91 */
  int64_t iter, status_val, rho_updates;
  double obj_val, pri_res, dua_res, setup_time, solve_time;
} mpcqp_info;
int mpcqp_get_info(const mpcqp_problem* pr, mpcqp_info* info);
int mpcqp_get_solution(const mpcqp_problem* pr, double* x, double* y);       /* work->solution->x / ->y */
int mpcqp_cleanup(mpcqp_problem* pr);                                        /* osqp_cleanup */

/* ---- (3b) batch of unstructured QPs sharing one CSC pattern ------------------------------------------ *
 * What polyTrajSolver does with three OsqpEigen::Solver objects (x, y, z: same P and A, own bounds;
 * polyTrajSolver.cpp:162-239, solved one after the other at :848-900) as ONE launch of the dense generic kernel, one CTA
 * per QP at a time; also the entry point for scoring many candidate paths at once.  Values are [B][nnz] / [B][n] / [B][m]
 * row-major; warm_x, warm_y, y, iter, rho_updates, obj, pri_res, dua_res may be NULL.  Host arrays in and out; synchronous.
 * Same validation and error codes as mpcqp_setup (n + m <= 4096, else MPCQP_ERR_STRUCTURE). */
int mpcqp_solve_qp_batch_host(mpcqp_engine* e, const mpcqp_settings* s, int32_t B, int64_t n, int64_t m,
                              const int64_t* P_colptr, const int64_t* P_rowidx, const double* P_val, const double* q,
                              const int64_t* A_colptr, const int64_t* A_rowidx, const double* A_val, const double* l,
                              const double* u, const double* warm_x, const double* warm_y, double* x, double* y,
                              int32_t* status, int32_t* iter, int32_t* rho_updates, double* obj, double* pri_res,
                              double* dua_res);

#ifdef __cplusplus
}
#endif
#endif
