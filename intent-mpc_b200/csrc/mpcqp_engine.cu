// mpcqp_engine.cu — sm_100a kernels and the C ABI (include/mpcqp_b200.h) of the batched MPC QP engine.
//
// Kernels:
//   mpc_assemble_kernel  device-side builder: planner inputs -> structured QP data (q, x0, obstacle-row
//                        gradients and lower bounds), restating mpcPlanner.cpp:952-966, 1040-1071, 1114-1139.
//   solve kernels        mpcqp_kernels.cuh / mpcqp_kernels.cu (bodies in mpcqp_core.cuh), reached through getters.
// There is no host solve path in this library.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/mpcqp_b200.h"
#include "mpcqp_core.cuh"
#include "mpcqp_kernels.cuh"
#include "mpcqp_dense.cuh"
#include "mpcqp_band.cuh"
#include "mpcqp_band_host.hpp"

namespace mpcqp {

// ------------------------------------------------------------------------------------------------
// device-side builder
// ------------------------------------------------------------------------------------------------
struct AsmArgs {
  int B, NS, R, n;
  double Qp[3];                 // position weights (full precision, MP.cpp:952-966 uses Q without the float cast)
  const double* x0;             // [B][6]
  const double* xref;           // [B][NS][3]
  const double* obs_c; const double* obs_semi; const double* obs_yaw;   // [B][N][R][3], [B][N][R][3], [B][N][R]
  const double* lin_pt;         // [B][N][3]
  double* q; double* x0s; double* g; double* low;
  int* hard;                    // [B] set when a stage-0 obstacle row is violated by the (fixed) current position
  const int* hist; int hist_thresh;   // iterations each slot took in the previous call (or null): >= thresh -> hard
  const int* nobs;              // [B] obstacle rows per stage of each instance (<= R, the array stride), or null: R everywhere
};

__global__ void mpc_assemble_kernel(const __grid_constant__ AsmArgs a) {
  const int N = a.NS - 1;
  const long long nq = (long long)a.B * a.n, n0 = (long long)a.B * 8, no = (long long)a.B * N * a.R;
  const long long total = nq + n0 + no;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    if (t < nq) {                                   // gradient: q = [-Q xRef_k ; 0]   (castMPCToQPGradient)
      const int b = (int)(t / a.n), v = (int)(t - (long long)b * a.n);
      double val = 0.0;
      if (v < 8 * a.NS) { const int k = v >> 3, j = v & 7; if (j < 3) val = -a.Qp[j] * a.xref[((long long)b * a.NS + k) * 3 + j]; }
      a.q[t] = val;
    } else if (t < nq + n0) {                       // x0 = [pos, vel, 0, 0]           (MP.cpp:401-408)
      const long long u = t - nq; const int b = (int)(u >> 3), j = (int)(u & 7);
      a.x0s[u] = j < 6 ? a.x0[(long long)b * 6 + j] : 0.0;
      // Scheduling hint only: in receding-horizon use slot b of consecutive calls is the same scenario one control
      // step later, so a slot that ran long last time very likely runs long again.
      if (j == 0 && a.hist && a.hist[b] >= a.hist_thresh) a.hard[b] = 1;
    } else {                                        // obstacle rows                   (MP.cpp:1040-1071, 1114-1139)
      const long long u = t - nq - n0;
      const long long bk = u / a.R;                 // b*N + k
      if (a.nobs && (int)(u - bk * a.R) >= a.nobs[bk / N]) {      // padding beyond this instance's rows: never read
        a.g[u * 3] = 0.0; a.g[u * 3 + 1] = 0.0; a.g[u * 3 + 2] = 0.0; a.low[u] = 0.0;
        continue;
      }
      const double cx = a.lin_pt[bk * 3], cy = a.lin_pt[bk * 3 + 1], cz = a.lin_pt[bk * 3 + 2];
      const double ox = a.obs_c[u * 3], oy = a.obs_c[u * 3 + 1], oz = a.obs_c[u * 3 + 2];
      const double sx = a.obs_semi[u * 3], sy = a.obs_semi[u * 3 + 1], sz = a.obs_semi[u * 3 + 2];
      double sn, cs; sincos(a.obs_yaw[u], &sn, &cs);
      const double xi = (cx - ox) * cs + (cy - oy) * sn;
      const double eta = -(cx - ox) * sn + (cy - oy) * cs;
      const double fxyz = xi * xi / (sx * sx) + eta * eta / (sy * sy) + (cz - oz) * (cz - oz) / (sz * sz);
      const double fxx = 2 * xi / (sx * sx) * cs + 2 * eta / (sy * sy) * (-sn);
      const double fyy = 2 * xi / (sx * sx) * sn + 2 * eta / (sy * sy) * cs;
      const double fzz = 2 * (cz - oz) / (sz * sz);
      a.g[u * 3] = fxx; a.g[u * 3 + 1] = fyy; a.g[u * 3 + 2] = fzz;
      const double lw = 1 - fxyz + fxx * cx + fyy * cy + fzz * cz;
      a.low[u] = lw;
      // Scheduling hint only (results do not depend on it): x_0 is pinned by the stage-0 equality, so a stage-0 obstacle
      // row that the current position violates makes the QP infeasible; OSQP then runs to max_iter.  Such
      // instances are started first so that they do not form the tail of the batch.
      const int N_ = a.NS - 1;
      if (bk % N_ == 0) {
        const long long b = bk / N_;
        if (fxx * a.x0[b * 6] + fyy * a.x0[b * 6 + 1] + fzz * a.x0[b * 6 + 2] < lw) a.hard[b] = 1;
      }
    }
  }
}

// Parked instances by decreasing key (rank sort: one thread per slot counts the slots ahead of it; n is a few hundred to a few
// thousand).  Ties and equal keys keep slot order.
__global__ void mpc_rank_parked_kernel(const int* count, int cap, const double* key, int* order) {
  __shared__ double tile[1024];
  int n = *count; if (n > cap) n = cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double ki = i < n ? key[i] : 0.0;
  int r = 0;
  for (int j0 = 0; j0 < n; j0 += 1024) {                   // every block walks all keys, 1,024 at a time through shared memory
    const int m = n - j0 < 1024 ? n - j0 : 1024;
    __syncthreads();
    for (int t = threadIdx.x; t < m; t += blockDim.x) tile[t] = key[j0 + t];
    __syncthreads();
#pragma unroll 8
    for (int t = 0; t < m; ++t) { const double kj = tile[t]; r += (kj > ki || (kj == ki && j0 + t < i)) ? 1 : 0; }
  }
  if (i < n) order[r] = i;
}

// Compacts the indices of the instances flagged `hard` into order[0 .. cnt[0]).
__global__ void mpc_order_kernel(int B, const int* hard, int* order, int* cnt) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    if (hard[b]) order[atomicAdd(&cnt[0], 1)] = b;
  }
}

// CTA kernels for horizons other than 30, registered by their translation units at load time (mpcqp_kernels.cuh)
static SolveKernel g_alt_kernels[kAltNsMax + 1][kAltRMax + 1];
AltKernelRegistration::AltKernelRegistration(int ns, int r, SolveKernel k) {
  if (ns >= 0 && ns <= kAltNsMax && r >= 0 && r <= kAltRMax) g_alt_kernels[ns][r] = k;
}
SolveKernel mpcqp_kernel_cta_alt(int ns, int r) {
  return (ns >= 0 && ns <= kAltNsMax && r >= 0 && r <= kAltRMax) ? g_alt_kernels[ns][r] : nullptr;
}

// want: 2 = CTA kernel if available, 1 = warp-fast kernel if available, 0 = generic.  *mode returns what was picked.
static SolveKernel pick_kernel(int NS, int R, int want, int* mode, bool assist = false, bool* wide = nullptr, bool force_wide = false) {
  if (wide) *wide = false;
  if (NS != 30 && want == kModeCta) {                      // other horizons: the plain 4-warp CTA kernel where it was built
    if (SolveKernel k = mpcqp_kernel_cta_alt(NS, R)) { *mode = kModeCta; return k; }
  }
  if (NS == 30 && want == kModeCta && wide && (R > 8 || (force_wide && R > 0)) && R <= kWideMax) {   // run-time obstacle count, rows in shared memory
    *mode = kModeCta; *wide = true;
    return mpcqp_kernel_cta_wide();
  }
  if (NS == 30 && want == kModeCta) {
    *mode = kModeCta;
    switch (R) {
#define X(r) case r: return mpcqp_kernel_cta_##r(assist);
      MPCQP_FAST_R_LIST
#undef X
      default: break;
    }
  }
  if (NS == 30 && want >= kModeWarp) {
    *mode = kModeWarp;
    switch (R) {
#define X(r) case r: return mpcqp_kernel_warp_##r();
      MPCQP_FAST_R_LIST
#undef X
      default: break;
    }
  }
  *mode = kModeGeneric;
  return mpcqp_kernel_generic();
}
static SolveKernel pick_setup_kernel(int R) {
  switch (R) {
#define X(r) case r: return mpcqp_kernel_setup_##r();
    MPCQP_FAST_R_LIST
#undef X
    default: return nullptr;
  }
}

}  // namespace mpcqp

// ------------------------------------------------------------------------------------------------
// host side: engine object + C ABI
// ------------------------------------------------------------------------------------------------
using namespace mpcqp;

struct DevBuf;
// Every DevBuf constructed while an engine is being built registers itself here, so that mpcqp_engine_destroy releases
// all of them without a hand-kept list (new members cannot be forgotten).
static thread_local std::vector<DevBuf*>* g_devbuf_registry = nullptr;
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  DevBuf() { if (g_devbuf_registry) g_devbuf_registry->push_back(this); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};
struct DevBufScope {   // first member of the engine: opens the registry before the DevBuf members are constructed
  std::vector<DevBuf*> all;
  DevBufScope() { g_devbuf_registry = &all; }
};
struct DevBufScopeEnd { DevBufScopeEnd() { g_devbuf_registry = nullptr; } };   // last member: closes it

struct DenseDev {   // device buffers of one call of the generic (unstructured) path
  DevBuf Pc, Pi, Px, Ac, Ai, Ax, q, l, u, wx, wy, x, y, ii, dd, ws, pat;
};

struct mpcqp_engine {
  DevBufScope bufs;                                          // must stay the first member (see DevBufScope)
  DenseDev dense;
  int device = 0, num_sms = 0, max_smem_optin = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;          // main stream; side stream for the second solve launch
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs = nullptr;   // batch start / end, solve-kernel start
  cudaEvent_t evf = nullptr, evj = nullptr;                  // fork / join of the side stream
  std::string err;
  int live_problems = 0; bool destroy_pending = false;       // mpcqp_problem handles keep their engine alive (see mpcqp_engine_destroy)
  double last_ms = 0.0, last_solve_ms = 0.0; long long last_launches = 0; int last_fast = 0; int force_generic = 0; int no_assist = 0; int dyn_per_instance = 0; int large_batch_factor = 0; int split_setup = 1; int migrate = 1, suspend_at = 300, suspend_at_small = 25, hist_active = 0;
  const int32_t* nobs_host = nullptr; const double* limits_host = nullptr;
  // structured-problem buffers (device)
  DevBuf pd, slack, q, x0s, g, low, ws, counter, hard, order, hist, dbg, nobs, limits, susp_cold, susp_scal, susp_list, susp_key, susp_order, susp_ctr, cand_tab;
  int hist_B = 0, hist_R = -1, use_history = 1;       // iteration counts of the previous batch call (same B, R) as a scheduling hint
  // staging for the *_host entry point
  DevBuf in_x0, in_xref, in_c, in_semi, in_yaw, in_lin, in_warm, out_x, out_y, out_i, out_d;
  DevBufScopeEnd bufs_end;                                   // must stay the last member
};

static float f32(double v) { return (float)v; }

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string(#call) + ": " + cudaGetErrorString(_e); return MPCQP_ERR_CUDA; } } while (0)

extern "C" void mpcqp_set_default_settings(mpcqp_settings* s) {
  if (!s) return;
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->adaptive_rho_tolerance = 5.0; s->adaptive_rho_fraction = 0.4;
  s->delta = 1e-6; s->time_limit = 0.0;
  s->max_iter = 4000; s->scaling = 10; s->adaptive_rho = 1; s->adaptive_rho_interval = 25; s->check_termination = 25;
  s->warm_start = 1; s->scaled_termination = 0; s->polish = 0; s->polish_refine_iter = 3; s->verbose = 0;
}

extern "C" void mpcqp_default_mpc_params(mpcqp_mpc_params* p) {
  if (!p) return;
  p->horizon = 30; p->ts = 0.1; p->max_vel = 5.0; p->max_acc = 20.0; p->y_min = -5.0; p->y_max = 5.0;
  p->z_min = 0.5; p->z_max = 4.5; p->static_safety_dist = 0.8; p->dynamic_safety_dist = 1.5;
  p->static_slack = 0.01; p->dynamic_slack = 0.2; p->position_weight = 1000.0; p->velocity_weight = 0.0;
  p->acceleration_weight = 10.0;
}

extern "C" int mpcqp_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}

extern "C" int mpcqp_engine_create(int device, mpcqp_engine** out) {
  if (!out) return MPCQP_ERR_ARG;
  *out = nullptr;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0 || device < 0 || device >= cnt) return MPCQP_ERR_CUDA;
  mpcqp_engine* e = new mpcqp_engine();
  e->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete e; return MPCQP_ERR_CUDA; }
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { delete e; return MPCQP_ERR_CUDA; }
  e->num_sms = pr.multiProcessorCount;
  e->max_smem_optin = (int)pr.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->evf, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->evj, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess || cudaEventCreate(&e->evs) != cudaSuccess) { delete e; return MPCQP_ERR_CUDA; }
  // development knobs (scheduling only; results never depend on them)
  if (const char* v = getenv("MPCQP_LARGE_BATCH_FACTOR")) { const int f = atoi(v); if (f >= 0) e->large_batch_factor = f; }
  if (const char* v = getenv("MPCQP_SPLIT_SETUP")) e->split_setup = atoi(v) != 0;
  if (const char* v = getenv("MPCQP_SUSPEND_AT")) { const int f = atoi(v); if (f >= 0) e->suspend_at = e->suspend_at_small = f; }
  *out = e;
  return MPCQP_OK;
}

// Problems created by mpcqp_setup hold a pointer to their engine.  Destroying an engine that still has live problems
// (an OsqpEigen::Solver that outlives the thread-local engine of its thread, OsqpEigenB200.hpp) only marks it; the last
// mpcqp_cleanup then releases it.
extern "C" int mpcqp_engine_destroy(mpcqp_engine* e) {
  if (!e) return MPCQP_ERR_ARG;
  if (e->live_problems > 0) { e->destroy_pending = true; return MPCQP_OK; }
  cudaSetDevice(e->device);
  for (DevBuf* b : e->bufs.all) b->release();
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->evs) cudaEventDestroy(e->evs);
  if (e->evf) cudaEventDestroy(e->evf);
  if (e->evj) cudaEventDestroy(e->evj);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return MPCQP_OK;
}

extern "C" const char* mpcqp_engine_last_error(const mpcqp_engine* e) { return e ? e->err.c_str() : "null engine"; }
extern "C" double mpcqp_engine_last_kernel_ms(const mpcqp_engine* e) { return e ? e->last_ms : 0.0; }
extern "C" double mpcqp_engine_last_solve_kernel_ms(const mpcqp_engine* e) { return e ? e->last_solve_ms : 0.0; }
extern "C" int64_t mpcqp_engine_last_launches(const mpcqp_engine* e) { return e ? e->last_launches : 0; }
extern "C" int mpcqp_engine_last_path(const mpcqp_engine* e) { return e ? e->last_fast : -1; }
extern "C" int mpcqp_engine_force_generic(mpcqp_engine* e, int on) { if (!e) return MPCQP_ERR_ARG; e->force_generic = on; return MPCQP_OK; }
extern "C" int mpcqp_engine_obs_dyn_per_instance(mpcqp_engine* e, int on) { if (!e) return MPCQP_ERR_ARG; e->dyn_per_instance = on ? 1 : 0; return MPCQP_OK; }
extern "C" int mpcqp_engine_num_obs_per_instance(mpcqp_engine* e, const int32_t* nobs) { if (!e) return MPCQP_ERR_ARG; e->nobs_host = nobs; return MPCQP_OK; }
extern "C" int mpcqp_engine_limits_per_instance(mpcqp_engine* e, const double* limits) { if (!e) return MPCQP_ERR_ARG; e->limits_host = limits; return MPCQP_OK; }
extern "C" int mpcqp_engine_use_migration(mpcqp_engine* e, int on) { if (!e) return MPCQP_ERR_ARG; e->migrate = on ? 1 : 0; return MPCQP_OK; }
extern "C" int mpcqp_engine_use_history(mpcqp_engine* e, int on) { if (!e) return MPCQP_ERR_ARG; e->use_history = on ? 1 : 0; if (!on) e->hist_B = 0; return MPCQP_OK; }
#ifdef MPCQP_PHASE_TIMING
extern "C" int mpcqp_debug_phase_clocks(mpcqp_engine* e, long long* out, int B) {   // development builds only
  if (!e || !out) return MPCQP_ERR_ARG;
  CK(cudaMemcpy(out, e->dbg.p, (size_t)B * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return MPCQP_OK;
}
#endif
extern "C" void* mpcqp_engine_stream(const mpcqp_engine* e) { return e ? (void*)e->stream : nullptr; }

static int check_settings(mpcqp_engine* e, const mpcqp_settings* s, Settings* o) {
  if (!s) { e->err = "null settings"; return MPCQP_ERR_ARG; }
  // validate_settings analogue (OSQP rejects these too) + features this engine does not implement
  if (!(s->rho > 0) || !(s->sigma > 0) || !(s->alpha > 0 && s->alpha < 2) || s->eps_abs < 0 || s->eps_rel < 0 ||
      (s->eps_abs == 0 && s->eps_rel == 0) || !(s->eps_prim_inf > 0) || !(s->eps_dual_inf > 0) || s->max_iter <= 0 ||
      s->scaling < 0 || s->check_termination < 0 || s->adaptive_rho_interval < 0 || !(s->adaptive_rho_tolerance >= 1.0)) {
    e->err = "invalid settings"; return MPCQP_ERR_SETTINGS;
  }
  if (s->polish != 0 || s->scaled_termination != 0 || s->time_limit != 0.0) {
    e->err = "settings polish / scaled_termination / time_limit are not implemented by this engine (must be 0)";
    return MPCQP_ERR_SETTINGS;
  }
  if (s->adaptive_rho && s->adaptive_rho_interval == 0) {
    e->err = "adaptive_rho_interval = 0 (wall-clock derived in OSQP) is not reproducible; set it explicitly (default 25)";
    return MPCQP_ERR_SETTINGS;
  }
  o->rho = s->rho; o->sigma = s->sigma; o->alpha = s->alpha; o->eps_abs = s->eps_abs; o->eps_rel = s->eps_rel;
  o->eps_prim_inf = s->eps_prim_inf; o->eps_dual_inf = s->eps_dual_inf; o->adaptive_rho_tolerance = s->adaptive_rho_tolerance;
  o->max_iter = (int)s->max_iter; o->scaling = (int)s->scaling; o->adaptive_rho = (int)s->adaptive_rho;
  o->adaptive_rho_interval = (int)s->adaptive_rho_interval; o->check_termination = (int)s->check_termination;
  o->warm_start = (int)s->warm_start;
  return MPCQP_OK;
}

// Shape + P diagonal from planner parameters (MP.cpp:891-951).
static int shape_from_params(mpcqp_engine* e, const mpcqp_mpc_params* p, int R, Shape* sh, std::vector<double>* pd) {
  if (!p || p->horizon < 3 || R < 0) { e->err = "bad mpc params (horizon >= 3, num_obs >= 0)"; return MPCQP_ERR_ARG; }
  const int NS = p->horizon, N = NS - 1;
  sh->NS = NS; sh->R = R; sh->n = 8 * NS + 5 * N; sh->m = 16 * NS + 5 * N + R * N;
  sh->obs_hi = INFINITY;                                    // upperBound = +inf on the obstacle rows (MP.cpp:1139)
  // Ad/Bd entries pass through `float value` (MP.cpp:1003,1014)
  sh->a_pv = (double)f32(p->ts); sh->b_pa = (double)f32(1.0 / 2 * (p->ts * p->ts)); sh->b_va = (double)f32(p->ts);
  const double sks = 1.0 - (1 - p->static_slack) * (1 - p->static_slack);
  const double skd = 1.0 - (1 - p->dynamic_slack) * (1 - p->dynamic_slack);
  const double lo[13] = { -INFINITY, p->y_min, p->z_min, -p->max_vel, -p->max_vel, -p->max_vel, -INFINITY, -INFINITY,
                          -p->max_acc, -p->max_acc, -p->max_acc, 0.0, 0.0 };
  const double hi[13] = { INFINITY, p->y_max, p->z_max, p->max_vel, p->max_vel, p->max_vel, INFINITY, INFINITY,
                          p->max_acc, p->max_acc, p->max_acc, skd, sks };
  for (int j = 0; j < 13; ++j) { if (lo[j] > hi[j]) { e->err = "box bounds: lower > upper"; return MPCQP_ERR_DATA; } sh->blo[j] = lo[j]; sh->bhi[j] = hi[j]; }
  // castMPCToQPHessian (MP.cpp:932-951): float-rounded weights; R indexed by GLOBAL variable index % 5
  const double Q[8] = { p->position_weight, p->position_weight, p->position_weight, p->velocity_weight, p->velocity_weight,
                        p->velocity_weight, 100.0, 1000.0 };
  const double Rw[5] = { p->acceleration_weight, p->acceleration_weight, p->acceleration_weight, 1.0, 1.0 };
  pd->assign((size_t)NS * 13, 0.0);
  for (int k = 0; k < NS; ++k) {
    for (int j = 0; j < 8; ++j) (*pd)[k * 13 + j] = (double)f32(Q[j]);
    if (k < N) for (int j = 0; j < 5; ++j) (*pd)[k * 13 + 8 + j] = (double)f32(Rw[(8 * NS + 5 * k + j) % 5]);
  }
  return MPCQP_OK;
}

// Launch the solve kernel on structured data already on the device.
static int launch_solve(mpcqp_engine* e, const Shape& sh, const Settings& st, Batch bt) {
  int mode = kModeGeneric;
  // force_generic: 0 = best available (CTA kernel), 1 = generic kernel, 2 = warp-fast kernel
  const int want = e->force_generic == 1 ? kModeGeneric : (e->force_generic == 2 ? kModeWarp : kModeCta);
  e->no_assist = e->force_generic == 3;           // 3 = CTA kernel without the assistant warps (A/B tests)
  bool wide = false;
  SolveKernel kern = pick_kernel(sh.NS, sh.R, want, &mode, false, &wide, bt.nobs != nullptr);
  if (bt.limits && mode != kModeCta) { e->err = "per-instance limits need a CTA kernel (horizon 30, num_obs <= 32)"; return MPCQP_ERR_ARG; }
  if (bt.nobs && !wide) { e->err = "per-instance obstacle counts need horizon 30 and 1 <= num_obs <= 32"; return MPCQP_ERR_ARG; }
  const int threads = mode == kModeCta ? (wide ? 224 : 128) : 32;
  const size_t smem = (size_t)smem_doubles(sh.NS, sh.R, mode, wide) * sizeof(double);
  if ((long long)smem > (long long)e->max_smem_optin) {
    e->err = "problem does not fit shared memory: horizon/num_obs too large (" + std::to_string(smem) + " B needed)";
    return MPCQP_ERR_ARG;
  }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) { e->err = "solve kernel cannot be resident"; return MPCQP_ERR_CUDA; }
  long long grid = (long long)e->num_sms * occ;
  if (grid > bt.B) grid = bt.B;
  const int wsd = (ws_doubles(sh.NS, sh.R, mode) + 1) & ~1;     // even: every block's scratch starts 16-byte aligned (bulk copies)
  CK(e->counter.need(4 * sizeof(int)));
  CK(cudaMemsetAsync(e->counter.p, 0, sizeof(int), e->stream));
  e->last_fast = mode;
  bt.queue = 0; bt.nhard = nullptr;
  if (mode != kModeCta) { if (bt.order) bt.nhard = e->counter.as<int>() + 1; }      // one-warp kernels: hard list first, then the rest
  if (mode == kModeCta && wide) {
    // one CTA (4 solver warps + 3 PCR assistants) per SM; instances flagged hard first, then the rest
    CK(e->ws.need((size_t)grid * wsd * sizeof(double)));
    bt.ws = e->ws.as<double>();
    if (bt.order) { bt.queue = 3; bt.nhard = e->counter.as<int>() + 1; }
    CK(cudaEventRecord(e->evs, e->stream));
    kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
    CK(cudaGetLastError());
    e->last_launches += 1;
    return MPCQP_OK;
  }
  if (mode == kModeCta) {
    // A CTA that has an SM to itself iterates ~1.5x faster than two sharing one.  Small batches therefore run one CTA
    // per SM (the launch asks for the whole shared memory of the SM).  Larger batches run as two concurrent launches:
    // the instances flagged hard (they run to max_iter and would otherwise form the tail of the batch) one per SM on
    // the main stream, everything else two per SM on the side stream, on whatever SMs the first launch leaves free.
    const size_t smem_solo = (size_t)e->max_smem_optin - 2048;   // more than half an SM: nothing else fits beside it
    // one-per-SM launches use the variant with PCR assistant warps (224 threads, matrices of levels 1..3 in registers) and, from
    // four obstacle rows per stage on, the row helper (256 threads)
    int mode_a = 0;
    const bool plain_solo = e->no_assist || sh.NS != 30;        // (the assistant / helper variants exist for horizon 30 only)
    SolveKernel kern_solo = plain_solo ? kern : pick_kernel(sh.NS, sh.R, want, &mode_a, true);
    const int threads_solo = plain_solo ? threads : cta_threads(sh.R, true, false);
    CK(cudaFuncSetAttribute(kern_solo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solo));
    if (bt.B <= e->num_sms || !bt.order) {
      const bool solo = bt.B <= e->num_sms;
      CK(e->ws.need((size_t)grid * wsd * sizeof(double)));
      bt.ws = e->ws.as<double>();
      CK(cudaEventRecord(e->evs, e->stream));
      if (solo) kern_solo<<<(unsigned)grid, threads_solo, smem_solo, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
      else kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
      CK(cudaGetLastError());
      e->last_launches += 1;
      return MPCQP_OK;
    }
    // Migration of long-running instances: whoever is still iterating after suspend_at iterations in a two-per-SM launch
    // parks its state; a follow-up launch resumes these instances (bit-identically) — on SMs of their own for small
    // batches, evenly spread for large ones — so that an instance nobody could predict to be long does not form the tail.
    // In the two-launch regime (below) the two-per-SM launch is a probe: setup, factorisation and ONE check interval, then
    // everybody still running is parked and the one-per-SM launch resumes them by decreasing primal residual — a step waits
    // for its longest instance anyway, and the residual after 25 iterations tells which ones those are (measured, parked
    // after 25 / 100 / 300 iterations: 1,024 QPs 5.97 / 6.38 / 6.53 ms, 2,048 QPs 6.56 / 7.68 / 7.85 ms, 4,096 QPs 10.5 / 11.4 /
    // 10.6 ms).  Large batches without the setup kernel keep 300.
    const int lb_factor = e->large_batch_factor > 0 ? e->large_batch_factor : (e->hist_active ? 13 : (e->split_setup ? 26 : 40));
    const int suspend_at = bt.B < (long long)lb_factor * grid ? e->suspend_at_small : e->suspend_at;
    const bool migrate = e->migrate && suspend_at > 0 && suspend_at < st.max_iter;
    if (migrate) {
      const int cap = bt.B < 8192 ? bt.B : 8192;
      const int stride = cold_slots(sh.R) * sh.NS;
      CK(e->susp_cold.need((size_t)cap * stride * sizeof(double)));
      CK(e->susp_scal.need((size_t)cap * 8 * sizeof(double)));
      CK(e->susp_list.need((size_t)cap * sizeof(int)));
      CK(e->susp_ctr.need(sizeof(int)));
      CK(cudaMemsetAsync(e->susp_ctr.p, 0, sizeof(int), e->stream));
      bt.suspend_at = suspend_at; bt.susp_cap = cap; bt.susp_stride = stride;
      bt.susp_cold = e->susp_cold.as<double>(); bt.susp_scal = e->susp_scal.as<double>();
      bt.susp_list = e->susp_list.as<int>(); bt.susp_count = e->susp_ctr.as<int>();
      CK(e->susp_key.need((size_t)cap * sizeof(double)));
      CK(e->susp_order.need((size_t)cap * sizeof(int)));
      bt.susp_key = e->susp_key.as<double>();
    }
    // measured crossovers (configs[1]-shaped batches): with a history the single launch (setup split off into its own kernel)
    // wins from ~13 x 296 instances, without one (nothing known about the instances) the probe + ordered resume win up to ~26 x 296
    // (6,144 QPs: 16.0 against 17.0 ms; 8,192: 21.5 either way)
    if (bt.B >= (long long)lb_factor * grid) {
      // (migration: the follow-up launch only starts when this one has drained, so it pays only where nothing is known
      // about the instances — no iteration history — and the batch is long enough to amortise the second launch)
      if (e->hist_active) bt.suspend_at = 0;
      // Large batch: every SM stays busy with two CTAs to the end anyway, and a one-per-SM launch would only halve the
      // occupancy of the SMs it takes.  One launch, two CTAs per SM, the hard list first.
      SolveKernel kset = (e->split_setup && sh.NS == 30) ? pick_setup_kernel(sh.R) : nullptr;
      if (kset) {
        // Setup (Ruiz scaling, rho vector, warm start) in its own kernel, three CTAs per SM with only the cold block in shared
        // memory; it leaves every instance parked at iteration 0 and the solve launch resumes them all (hard list first).
        const int stride = cold_slots(sh.R) * sh.NS;
        const size_t smem_set = ((size_t)stride + 8) * sizeof(double);
        CK(e->susp_cold.need((size_t)bt.B * stride * sizeof(double)));
        CK(e->susp_scal.need((size_t)bt.B * 8 * sizeof(double)));
        CK(e->susp_list.need((size_t)bt.B * sizeof(int)));
        CK(e->susp_ctr.need(sizeof(int)));
        CK(cudaMemsetAsync(e->susp_ctr.p, 0, sizeof(int), e->stream));
        CK(cudaFuncSetAttribute(kset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_set));
        int occ_s = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, kset, 128, smem_set));
        if (occ_s < 1) { e->err = "setup kernel cannot be resident"; return MPCQP_ERR_CUDA; }
        const long long gs = (long long)e->num_sms * occ_s;
        CK(e->ws.need((size_t)(grid > gs ? grid : gs) * wsd * sizeof(double)));
        bt.ws = e->ws.as<double>();
        bt.suspend_at = 0; bt.susp_cap = bt.B; bt.susp_stride = stride;
        bt.susp_cold = e->susp_cold.as<double>(); bt.susp_scal = e->susp_scal.as<double>();
        bt.susp_list = e->susp_list.as<int>(); bt.susp_count = e->susp_ctr.as<int>();
        bt.nhard = e->counter.as<int>() + 1;
        CK(cudaEventRecord(e->evs, e->stream));
        kset<<<(unsigned)gs, 128, smem_set, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
        CK(cudaGetLastError());
        Batch br = bt; br.queue = 4;
        kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, br, wsd, e->counter.as<int>());
        CK(cudaGetLastError());
        e->last_launches += 2;
        return MPCQP_OK;
      }
      CK(e->ws.need((size_t)grid * wsd * sizeof(double)));
      bt.ws = e->ws.as<double>();
      bt.queue = 3; bt.nhard = e->counter.as<int>() + 1;
      CK(cudaEventRecord(e->evs, e->stream));
      kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
      CK(cudaGetLastError());
      e->last_launches += 1;
      if (bt.suspend_at > 0) {                             // the parked instances, again two per SM over the whole GPU
        Batch br = bt; br.queue = 4; br.suspend_at = 0;
        kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, br, wsd, e->counter.as<int>());
        CK(cudaGetLastError());
        e->last_launches += 1;
      }
      return MPCQP_OK;
    }
    const long long gh = e->num_sms < bt.B ? e->num_sms : bt.B;
    CK(e->ws.need((size_t)(grid + 2 * gh) * wsd * sizeof(double)));
    bt.nhard = e->counter.as<int>() + 1;
    CK(cudaEventRecord(e->evs, e->stream));
    CK(cudaEventRecord(e->evf, e->stream));
    CK(cudaStreamWaitEvent(e->stream2, e->evf, 0));
    Batch bh = bt; bh.queue = 1; bh.ws = e->ws.as<double>(); bh.suspend_at = 0;
    kern_solo<<<(unsigned)gh, threads_solo, smem_solo, e->stream>>>(sh, st, bh, wsd, e->counter.as<int>());
    CK(cudaGetLastError());
    Batch bn = bt; bn.queue = 2; bn.ws = e->ws.as<double>() + (size_t)gh * wsd;
    kern<<<(unsigned)grid, threads, smem, e->stream2>>>(sh, st, bn, wsd, e->counter.as<int>());
    CK(cudaGetLastError());
    if (migrate) {                                         // the parked instances, one per SM (with assistants), behind the second launch
      // ... those with the largest primal residual first: after a few dozen iterations it separates the instances that run to
      // max_iter from the rest almost perfectly (8,192 configs[1] instances: all long ones among the first 148 of every 1,024)
      mpc_rank_parked_kernel<<<(unsigned)((bt.susp_cap + 127) / 128), 128, 0, e->stream2>>>(bt.susp_count, bt.susp_cap, bt.susp_key, e->susp_order.as<int>());
      CK(cudaGetLastError());
      Batch br = bt; br.queue = 4; br.suspend_at = 0; br.ws = e->ws.as<double>() + (size_t)(gh + grid) * wsd;
      br.susp_order = e->susp_order.as<int>();
      kern_solo<<<(unsigned)gh, threads_solo, smem_solo, e->stream2>>>(sh, st, br, wsd, e->counter.as<int>());
      CK(cudaGetLastError());
      e->last_launches += 1;
    }
    CK(cudaEventRecord(e->evj, e->stream2));
    CK(cudaStreamWaitEvent(e->stream, e->evj, 0));
    e->last_launches += 2;
    return MPCQP_OK;
  }
  CK(e->ws.need((size_t)grid * wsd * sizeof(double)));
  bt.ws = e->ws.as<double>();
  CK(cudaEventRecord(e->evs, e->stream));
  kern<<<(unsigned)grid, threads, smem, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
  CK(cudaGetLastError());
  e->last_launches += 1;
  return MPCQP_OK;
}

static int solve_mpc_device(mpcqp_engine* e, const mpcqp_mpc_params* p, const mpcqp_settings* s, int B, int R,
                            const double* x0, const double* xref, const double* obs_c, const double* obs_semi,
                            const double* obs_yaw, const int32_t* obs_dyn_host, const double* lin_pt, const double* warm_x,
                            double* x, double* y, int32_t* status, int32_t* iter, int32_t* rho_updates, double* obj,
                            double* pri_res, double* dua_res) {
  Settings st; Shape sh; std::vector<double> pd;
  int rc = check_settings(e, s, &st); if (rc) return rc;
  rc = shape_from_params(e, p, R, &sh, &pd); if (rc) return rc;
  if (B <= 0) { e->err = "B must be positive"; return MPCQP_ERR_ARG; }
  if (!x0 || !xref || !x || !status || !iter || !rho_updates || !obj || !pri_res || !dua_res || (R > 0 && (!obs_c || !obs_semi || !obs_yaw || !obs_dyn_host || !lin_pt))) {
    e->err = "null array"; return MPCQP_ERR_ARG;
  }
  const int NS = sh.NS, N = NS - 1;
  CK(cudaSetDevice(e->device));
  CK(e->pd.need(pd.size() * sizeof(double)));
  CK(cudaMemcpyAsync(e->pd.p, pd.data(), pd.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  const size_t nflag = (size_t)N * R * (e->dyn_per_instance ? B : 1);
  std::vector<unsigned char> slk(nflag ? nflag : 1, 0);
  for (size_t i = 0; i < nflag; ++i) slk[i] = obs_dyn_host[i] ? 0 : 1;   // MP.cpp:1064-1069
  CK(e->slack.need(slk.size()));
  CK(cudaMemcpyAsync(e->slack.p, slk.data(), slk.size(), cudaMemcpyHostToDevice, e->stream));
  CK(e->q.need((size_t)B * sh.n * sizeof(double)));
  CK(e->x0s.need((size_t)B * 8 * sizeof(double)));
  CK(e->g.need((size_t)B * N * (R > 0 ? R : 1) * 3 * sizeof(double)));
  CK(e->low.need((size_t)B * N * (R > 0 ? R : 1) * sizeof(double)));
  CK(e->hard.need((size_t)B * sizeof(int)));
  CK(e->order.need((size_t)B * sizeof(int)));
  CK(e->counter.need(4 * sizeof(int)));
  e->last_launches = 0;
  CK(cudaEventRecord(e->ev0, e->stream));
  AsmArgs a;
  a.B = B; a.NS = NS; a.R = R; a.n = sh.n;
  a.Qp[0] = a.Qp[1] = a.Qp[2] = p->position_weight;
  a.x0 = x0; a.xref = xref; a.obs_c = obs_c; a.obs_semi = obs_semi; a.obs_yaw = obs_yaw; a.lin_pt = lin_pt;
  a.q = e->q.as<double>(); a.x0s = e->x0s.as<double>(); a.g = e->g.as<double>(); a.low = e->low.as<double>();
  a.hard = e->hard.as<int>();
  a.nobs = nullptr;
  if (e->nobs_host) {                                  // per-instance obstacle counts: R is the array stride
    if (R < 1) { e->err = "per-instance obstacle counts need num_obs >= 1 (the array stride)"; return MPCQP_ERR_ARG; }
    for (int b = 0; b < B; ++b) if (e->nobs_host[b] < 0 || e->nobs_host[b] > R) { e->err = "num_obs per instance out of [0, num_obs]"; return MPCQP_ERR_ARG; }
    CK(e->nobs.need((size_t)B * sizeof(int)));
    CK(cudaMemcpyAsync(e->nobs.p, e->nobs_host, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    a.nobs = e->nobs.as<int>();
  }
  a.hist = (e->use_history && e->hist_B == B && e->hist_R == R) ? e->hist.as<int>() : nullptr;
  e->hist_active = a.hist != nullptr;
  a.hist_thresh = 500;
  CK(cudaMemsetAsync(e->hard.p, 0, (size_t)B * sizeof(int), e->stream));
  CK(cudaMemsetAsync(e->counter.p, 0, 4 * sizeof(int), e->stream));
  {
    long long total = (long long)B * sh.n + (long long)B * 8 + (long long)B * N * R;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)e->num_sms * 8;
    if (blocks > cap) blocks = cap;
    mpc_assemble_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(a);
    CK(cudaGetLastError());
    mpc_order_kernel<<<(unsigned)((B + 255) / 256 < cap ? (B + 255) / 256 : cap), 256, 0, e->stream>>>(B, a.hard, e->order.as<int>(), e->counter.as<int>() + 1);
    CK(cudaGetLastError());
    e->last_launches += 2;
  }
  Batch bt; memset(&bt, 0, sizeof bt);
  bt.nobs = a.nobs;
  if (e->limits_host) {
    for (int b = 0; b < 2 * B; ++b) if (!(e->limits_host[b] > 0.0)) { e->err = "per-instance limits must be positive"; return MPCQP_ERR_ARG; }
    CK(e->limits.need((size_t)B * 2 * sizeof(double)));
    CK(cudaMemcpyAsync(e->limits.p, e->limits_host, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    bt.limits = e->limits.as<double>();
  }
  bt.pd = e->pd.as<double>(); bt.slack = e->slack.as<unsigned char>(); bt.slack_stride = e->dyn_per_instance ? N * R : 0; bt.q = a.q; bt.x0 = a.x0s; bt.g = a.g; bt.low = a.low;
  bt.warm_x = warm_x; bt.x = x; bt.y = y; bt.status = status; bt.iter = iter; bt.rho_updates = rho_updates;
  bt.obj = obj; bt.pri_res = pri_res; bt.dua_res = dua_res; bt.B = B; bt.order = e->order.as<int>(); bt.hard = e->hard.as<int>();
#ifdef MPCQP_PHASE_TIMING
  CK(e->dbg.need((size_t)B * 16 * sizeof(long long))); bt.dbg = e->dbg.as<long long>();
#endif
  rc = launch_solve(e, sh, st, bt); if (rc) return rc;
  CK(cudaEventRecord(e->ev1, e->stream));
  if (e->use_history) {
    CK(e->hist.need((size_t)B * sizeof(int)));
    CK(cudaMemcpyAsync(e->hist.p, iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, e->stream));
    e->hist_B = B; e->hist_R = R;
  }
  return MPCQP_OK;
}

extern "C" int mpcqp_solve_mpc_batch_device(mpcqp_engine* e, const mpcqp_mpc_params* p, const mpcqp_settings* s, int32_t B,
                                            int32_t num_obs, const double* x0, const double* xref, const double* obs_c,
                                            const double* obs_semi, const double* obs_yaw, const int32_t* obs_dyn,
                                            const double* lin_pt, const double* warm_x, double* x, double* y,
                                            int32_t* status, int32_t* iter, int32_t* rho_updates, double* obj,
                                            double* pri_res, double* dua_res) {
  if (!e) return MPCQP_ERR_ARG;
  return solve_mpc_device(e, p, s, B, num_obs, x0, xref, obs_c, obs_semi, obs_yaw, obs_dyn, lin_pt, warm_x, x, y, status,
                          iter, rho_updates, obj, pri_res, dua_res);
}

extern "C" int mpcqp_engine_sync(mpcqp_engine* e) {
  if (!e) return MPCQP_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_ms = ms;
  if (cudaEventElapsedTime(&ms, e->evs, e->ev1) == cudaSuccess) e->last_solve_ms = ms;
  (void)cudaGetLastError();          // events not recorded yet (no solve so far) leave an error behind: not ours to report later
  return MPCQP_OK;
}

extern "C" int mpcqp_solve_mpc_batch_host(mpcqp_engine* e, const mpcqp_mpc_params* p, const mpcqp_settings* s, int32_t B,
                                          int32_t R, const double* x0, const double* xref, const double* obs_c,
                                          const double* obs_semi, const double* obs_yaw, const int32_t* obs_dyn,
                                          const double* lin_pt, const double* warm_x, double* x, double* y, int32_t* status,
                                          int32_t* iter, int32_t* rho_updates, double* obj, double* pri_res, double* dua_res) {
  if (!e) return MPCQP_ERR_ARG;
  if (!p || p->horizon < 3 || B <= 0 || R < 0) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  const int NS = p->horizon, N = NS - 1, n = 8 * NS + 5 * N, m = 16 * NS + 5 * N + R * N;
  CK(cudaSetDevice(e->device));
  const size_t d = sizeof(double);
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    if (!src || !bytes) return cudaSuccess;
    cudaError_t r = b.need(bytes); if (r != cudaSuccess) return r;
    return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream);
  };
  CK(up(e->in_x0, x0, (size_t)B * 6 * d));
  CK(up(e->in_xref, xref, (size_t)B * NS * 3 * d));
  CK(up(e->in_c, obs_c, (size_t)B * N * R * 3 * d));
  CK(up(e->in_semi, obs_semi, (size_t)B * N * R * 3 * d));
  CK(up(e->in_yaw, obs_yaw, (size_t)B * N * R * d));
  CK(up(e->in_lin, lin_pt, (size_t)B * N * 3 * d));
  CK(up(e->in_warm, warm_x, (size_t)B * n * d));
  CK(e->out_x.need((size_t)B * n * d));
  if (y) CK(e->out_y.need((size_t)B * m * d));
  CK(e->out_i.need((size_t)B * 3 * sizeof(int32_t)));
  CK(e->out_d.need((size_t)B * 3 * d));
  int32_t* oi = e->out_i.as<int32_t>(); double* od = e->out_d.as<double>();
  int rc = solve_mpc_device(e, p, s, B, R, x0 ? e->in_x0.as<double>() : nullptr, xref ? e->in_xref.as<double>() : nullptr,
                            (R && obs_c) ? e->in_c.as<double>() : nullptr, (R && obs_semi) ? e->in_semi.as<double>() : nullptr,
                            (R && obs_yaw) ? e->in_yaw.as<double>() : nullptr, obs_dyn, (R && lin_pt) ? e->in_lin.as<double>() : nullptr,
                            warm_x ? e->in_warm.as<double>() : nullptr, e->out_x.as<double>(), y ? e->out_y.as<double>() : nullptr,
                            oi, oi + B, oi + 2 * B, od, od + B, od + 2 * B);
  if (rc) return rc;
  CK(cudaMemcpyAsync(x, e->out_x.p, (size_t)B * n * d, cudaMemcpyDeviceToHost, e->stream));
  if (y) CK(cudaMemcpyAsync(y, e->out_y.p, (size_t)B * m * d, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(status, oi, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(iter, oi + B, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(rho_updates, oi + 2 * B, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(obj, od, (size_t)B * d, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(pri_res, od + B, (size_t)B * d, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(dua_res, od + 2 * B, (size_t)B * d, cudaMemcpyDeviceToHost, e->stream));
  return mpcqp_engine_sync(e);
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA-pipe microbenchmark: the roofline denominator for the solve kernel (MEASURED_PEAKS.json has no
// FP64 figure).  8 independent DFMA chains per thread, all SMs, timed with CUDA events on the engine stream.
// ------------------------------------------------------------------------------------------------
namespace mpcqp {
__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace mpcqp

extern "C" int mpcqp_fp64_fma_peak(mpcqp_engine* e, double* tflops) {
  if (!e || !tflops) return MPCQP_ERR_ARG;
  CK(cudaSetDevice(e->device));
  const int blocks = e->num_sms * 8, threads = 256, iters = 4096;
  CK(e->out_d.need((size_t)blocks * threads * sizeof(double)));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e->ev0, e->stream));
    fp64_fma_peak_kernel<<<blocks, threads, 0, e->stream>>>(e->out_d.as<double>(), iters, 0.999999, 1e-9);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return MPCQP_OK;
}

// ------------------------------------------------------------------------------------------------
// (2b) candidate scoring and selection on the device: getTrajectoryScore / getConsistencyScore / getDetourScore /
// getSafetyScore / evaluateTraj (mpcPlanner.cpp:771-887) for the candidates of many scenarios at once, so that the six
// solves of makePlanWithPred (mpcPlanner.cpp:609-644), their scoring and the choice of the plan need no host round trip.
// ------------------------------------------------------------------------------------------------
namespace mpcqp {
// one warp per candidate, lane = stage (in chunks of 32 for longer horizons).  The per-stage terms are computed in parallel
// and then added by lane 0 in stage order, which is the reference's own summation order (mpcPlanner.cpp:789-795, 804-810,
// 817-848): the scores differ from the reference's only by the rounding of tanh (CUDA's vs glibc's).
__device__ __forceinline__ double ordered_warp_sum(double acc, double term, int count) {
  for (int k = 0; k < count; ++k) acc += __shfl_sync(0xffffffffu, term, k);
  return acc;
}
__global__ void __launch_bounds__(128) mpc_score_kernel(int B, int NS, int R, int n_dynamic, int n, double dyn_safety, double stat_safety,
                                                         const double* __restrict__ x, const double* __restrict__ prev,
                                                         const double* __restrict__ xref, const double* __restrict__ obs_c,
                                                         const double* __restrict__ obs_semi, const double* __restrict__ obs_c_last,
                                                         const double* __restrict__ obs_semi_last, double* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int N = NS - 1;
  const double k05 = 0.54930614433405484570;             // atanh(0.5), mpcPlanner.cpp:830,840
  for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += (long long)gridDim.x * (blockDim.x >> 5)) {
    const double* xb = x + b * n;
    double cons = 0.0, det = 0.0, saf = 0.0;
    for (int k0 = 0; k0 < NS; k0 += 32) {
      const int k = k0 + lane, cnt = NS - k0 < 32 ? NS - k0 : 32;
      double tc = 0.0, td = 0.0, tsf = 0.0;
      if (k < NS) {
        const double px = xb[8 * k], py = xb[8 * k + 1], pz = xb[8 * k + 2];
        if (prev && k < 10) {                              // numConsistencyStep = 10 (mpcPlanner.cpp:781)
          const double* pv = prev + b * n + 8 * k;
          const double dx = pv[0] - px, dy = pv[1] - py, dz = pv[2] - pz;
          tc = sqrt(dx * dx + dy * dy + dz * dz);
        }
        {
          const double* rf = xref + (b * NS + k) * 3;
          const double dx = rf[0] - px, dy = rf[1] - py, dz = rf[2] - pz;
          td = sqrt(dx * dx + dy * dy + dz * dz);
        }
        if (R > 0) {
          // obstaclePos[j][i] for EVERY state i = 0..N (mpcPlanner.cpp:818-826): the solve holds stages 0..N-1, the
          // prediction of stage N comes in obs_*_last
          const double* oc = k < N ? obs_c + ((b * N + k) * R) * 3 : obs_c_last + (b * R) * 3;
          const double* om = k < N ? obs_semi + ((b * N + k) * R) * 3 : obs_semi_last + (b * R) * 3;
          double dist = 0.0, tw = 0.0;
          for (int o = 0; o < R; ++o) {
            const bool dynamic = o < n_dynamic;
            const double sd = dynamic ? dyn_safety : stat_safety;
            // dynamic: maxSize = |full size (x, y)|; static: |half size (x, y)|   (mpcPlanner.cpp:828,838); semi = size/2 + safety
            const double hx = om[3 * o] - sd, hy = om[3 * o + 1] - sd;
            const double sx = dynamic ? 2.0 * hx : hx, sy = dynamic ? 2.0 * hy : hy;
            const double ms = sqrt(sx * sx + sy * sy);
            const double ex = px - oc[3 * o], ey = py - oc[3 * o + 1];
            const double d = sqrt(ex * ex + ey * ey + 0.0);
            const double w = 1.0 - tanh(k05 / (sd + ms) * d);
            dist += d * w; tw += w;
          }
          tsf = dist / tw;
        }
      }
      cons = ordered_warp_sum(cons, tc, cnt < 10 - k0 ? cnt : (10 - k0 > 0 ? 10 - k0 : 0));
      det = ordered_warp_sum(det, td, cnt);
      saf = ordered_warp_sum(saf, tsf, cnt);
    }
    if (lane == 0) {
      const int steps = NS < 10 ? NS : 10;
      score[b * 3] = prev ? fmax(cons / steps, 0.1) : 0.0;
      score[b * 3 + 1] = fmax(det / NS, 0.1);
      score[b * 3 + 2] = R > 0 ? saf / NS : nan("");       // no obstacles: 0/0 in the reference
    }
  }
}

// one thread per scenario: evaluateTraj (mpcPlanner.cpp:854-887) over its C candidates, best plan copied out
__global__ void mpc_select_kernel(int S, int C, int n, const int* __restrict__ cand, const double* __restrict__ weight,
                                  const double* __restrict__ score, const double* __restrict__ x_all, int* __restrict__ best,
                                  double* __restrict__ weighted, double* __restrict__ plan) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    double avg[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < C; ++c) { const double* sc = score + (long long)cand[s * C + c] * 3; avg[0] += sc[0]; avg[1] += sc[1]; avg[2] += sc[2]; }
    avg[0] /= C; avg[1] /= C; avg[2] /= C;
    // weightedScore.maxCoeff(&bestTrajIdx) (mpcPlanner.cpp:883), Eigen 3.3's visitor: start from element 0, replace on a
    // strict `>`; a NaN elsewhere never wins, a NaN in element 0 is never replaced
    int bi = 0; double bv = 0.0;
    for (int c = 0; c < C; ++c) {
      const double* sc = score + (long long)cand[s * C + c] * 3;
      const double w = weight[s * C + c] * (1.0 * (avg[0] / sc[0]) + 1.0 * (avg[1] / sc[1]) + 1.0 * (sc[2] / avg[2]));
      if (weighted) weighted[s * C + c] = w;
      if (c == 0) bv = w; else if (w > bv) { bv = w; bi = c; }
    }
    best[s] = bi;
  }
  (void)x_all; (void)plan; (void)n;
}
__global__ void mpc_gather_plan_kernel(int S, int C, int n, const int* __restrict__ cand, const int* __restrict__ best,
                                       const double* __restrict__ x_all, double* __restrict__ plan) {
  const long long total = (long long)S * n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t / n), j = (int)(t - (long long)s * n);
    plan[t] = x_all[(long long)cand[s * C + best[s]] * n + j];
  }
}
}  // namespace mpcqp

namespace mpcqp {
// getIntentComb / findClosestObstacle (mpcPlanner.cpp:663-769), one thread per scenario: the closest obstacle (direction-
// weighted distance on the previous plan, plain distance on the first step), the six intent hypotheses for it sorted by
// descending weight (std::sort on (weight, index) pairs taken from the back, :728, :753-756), every other obstacle with its
// most likely intent.  Output: for sorted candidate c of scenario s the obstacle / intent of each of its rows, its row in
// the 4S-row batch (one intent: D rows per stage) or the 2S-row batch (two intents: the closest obstacle twice, D+1 rows).
__global__ void mpc_intent_enumerate_kernel(int S, int D, int NP, int n, const double* __restrict__ pp, const double* __restrict__ prob,
                                            const double* __restrict__ prev_plan, const double* __restrict__ pos,
                                            int* __restrict__ row_ob, int* __restrict__ row_it, int* __restrict__ scen_a, int* __restrict__ scen_b,
                                            double* __restrict__ weight, int* __restrict__ cand) {
  constexpr int FORWARD = 0, LEFT = 1, RIGHT = 2, STOP = 3;          // dynamic_predictor/utils.h:15-20
  const int combo_it[6][2] = {{STOP, -1}, {LEFT, -1}, {RIGHT, -1}, {FORWARD, -1}, {LEFT, FORWARD}, {RIGHT, FORWARD}};
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    // closest obstacle: positions "now" are the FORWARD predictions at step 0
    int ob = 0; double best = INFINITY;
    double s0[3], ta = 0.0;
    if (prev_plan) {
      const double* pl = prev_plan + (long long)s * n;
      s0[0] = pl[0]; s0[1] = pl[1]; s0[2] = pl[2];
      ta = atan2(pl[8 + 1] - pl[1], pl[8] - pl[0]);
    } else { s0[0] = pos[s * 3]; s0[1] = pos[s * 3 + 1]; s0[2] = pos[s * 3 + 2]; }
    const int nterm = ((n + 5) / 13) / 3;                 // currentStatesSol_.size() / 3 (mpcPlanner.cpp:689)
    for (int j = 0; j < D; ++j) {
      const double* o = pp + ((((long long)s * D + j) * 4 + FORWARD) * NP) * 3;
      const double dx = s0[0] - o[0], dy = s0[1] - o[1], dz = s0[2] - o[2];
      const double d = sqrt(dx * dx + dy * dy + dz * dz);
      double w = d;
      if (prev_plan) {
        // the reference adds exp(-t) * d * (3 - cos(.)) for t = 0 .. size/3 - 1 with the SAME state every term and stops
        // once the sum passes the best so far (mpcPlanner.cpp:689-702); same order of additions here
        const double c3 = 3.0 - cos(ta - atan2(o[1] - s0[1], o[0] - s0[0]));
        w = 0.0;
        for (int t = 0; t < nterm; ++t) { w += exp(-(double)t) * d * c3; if (w > best) break; }
      }
      if (w < best) { best = w; ob = j; }
    }
    const double* pr = prob + ((long long)s * D + ob) * 4;
    double w6[6] = {pr[STOP], pr[LEFT], pr[RIGHT], pr[FORWARD], fmax(pr[LEFT], pr[FORWARD]), fmax(pr[RIGHT], pr[FORWARD])};
    int ord[6] = {0, 1, 2, 3, 4, 5};
    for (int a = 1; a < 6; ++a) {                      // descending by (weight, index)
      const int v = ord[a]; int b = a - 1;
      while (b >= 0 && (w6[ord[b]] < w6[v] || (w6[ord[b]] == w6[v] && ord[b] < v))) { ord[b + 1] = ord[b]; --b; }
      ord[b + 1] = v;
    }
    int na = 0, nb = 0;
    for (int c = 0; c < 6; ++c) {
      weight[s * 6 + c] = w6[c];                       // ORIGINAL combo order: evaluateTraj indexes it with the sorted position (:866-880)
      const int id = ord[c], two = id >= 4;
      const int row = two ? 4 * S + 2 * s + nb : 4 * s + na;
      if (two) { scen_b[2 * s + nb] = s; ++nb; } else { scen_a[4 * s + na] = s; ++na; }
      cand[s * 6 + c] = row;
      int* ro = row_ob + (long long)row * (D + 1); int* ri = row_it + (long long)row * (D + 1);
      int r = 0;
      ro[r] = ob; ri[r] = combo_it[id][0]; ++r;
      if (two) { ro[r] = ob; ri[r] = combo_it[id][1]; ++r; }
      for (int j = 0; j < D; ++j) if (j != ob) {
        const double* pj = prob + ((long long)s * D + j) * 4;
        int mi = 0; for (int t = 1; t < 4; ++t) if (pj[t] > pj[mi]) mi = t;
        ro[r] = j; ri[r] = mi; ++r;
      }
    }
  }
}
// the obstacle rows of every candidate: centre = predicted position, semi-axes = predicted size / 2 + dynamicSafetyDist_
// (updateObstacleParam, mpcPlanner.cpp:1160-1172) for the N stages the QP constrains, plus — in the *_last arrays — the
// prediction of stage N, which only getSafetyScore reads (mpcPlanner.cpp:818-826)
__global__ void mpc_intent_fill_kernel(int S, int D, int NP, int N, double safety, const double* __restrict__ pp, const double* __restrict__ ps,
                                       const int* __restrict__ row_ob, const int* __restrict__ row_it, const int* __restrict__ scen_a,
                                       const int* __restrict__ scen_b, double* __restrict__ ca, double* __restrict__ sa,
                                       double* __restrict__ cb, double* __restrict__ sb, double* __restrict__ la, double* __restrict__ lsa,
                                       double* __restrict__ lb, double* __restrict__ lsb) {
  const int NK = la ? N + 1 : N;                          // stage N goes to the *_last arrays
  const long long na = (long long)4 * S * NK * D, nb = (long long)2 * S * NK * (D + 1);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < na + nb; t += (long long)gridDim.x * blockDim.x) {
    const bool two = t >= na;
    const long long u = two ? t - na : t;
    const int R = two ? D + 1 : D;
    const int o = (int)(u % R); const long long bk = u / R; const int k = (int)(bk % NK); const int b = (int)(bk / NK);
    const int row = two ? 4 * S + b : b;
    const int s = two ? scen_b[b] : scen_a[b];
    const int j = row_ob[(long long)row * (D + 1) + o], it = row_it[(long long)row * (D + 1) + o];
    const long long src = ((((long long)s * D + j) * 4 + it) * NP + k) * 3;
    double* c; double* m;
    if (k < N) { const long long v = (((long long)b * N + k) * R + o) * 3; c = (two ? cb : ca) + v; m = (two ? sb : sa) + v; }
    else { const long long v = ((long long)b * R + o) * 3; c = (two ? lb : la) + v; m = (two ? lsb : lsa) + v; }
    c[0] = pp[src]; c[1] = pp[src + 1]; c[2] = pp[src + 2];
    m[0] = ps[src] / 2 + safety; m[1] = ps[src + 1] / 2 + safety; m[2] = ps[src + 2] / 2 + safety;
  }
}
// dst[b][:] = src[idx[b]][:]  (scenario-level inputs replicated per candidate: x0, xref, lin_pt, warm_x)
__global__ void mpc_gather_rows_kernel(long long B, int w, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < B * w; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / w; dst[t] = src[(long long)idx[b] * w + (t - b * w)];
  }
}
}  // namespace mpcqp

extern "C" int mpcqp_intent_candidates_device(mpcqp_engine* e, const mpcqp_mpc_params* p, int32_t S, int32_t D, int32_t NP,
                                              const double* pred_pos, const double* pred_size, const double* prob, const double* prev_plan,
                                              const double* pos, int32_t* scen_a, int32_t* scen_b, double* obs_c_a, double* obs_semi_a,
                                              double* obs_c_b, double* obs_semi_b, double* obs_c_last_a, double* obs_semi_last_a,
                                              double* obs_c_last_b, double* obs_semi_last_b, double* weight, int32_t* cand) {
  if (!e) return MPCQP_ERR_ARG;
  const bool want_last = obs_c_last_a || obs_semi_last_a || obs_c_last_b || obs_semi_last_b;
  if (want_last && (!obs_c_last_a || !obs_semi_last_a || !obs_c_last_b || !obs_semi_last_b || !p || NP < p->horizon)) {
    e->err = "the stage-N obstacle arrays need all four pointers and predictions of at least `horizon` steps"; return MPCQP_ERR_ARG;
  }
  if (!p || S <= 0 || D <= 0 || D > 31 || NP < p->horizon - 1 || !pred_pos || !pred_size || !prob || (!prev_plan && !pos) || !scen_a || !scen_b ||
      !obs_c_a || !obs_semi_a || !obs_c_b || !obs_semi_b || !weight || !cand) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  CK(cudaSetDevice(e->device));
  const int NS = p->horizon, N = NS - 1, n = 8 * NS + 5 * N;
  CK(e->cand_tab.need((size_t)6 * S * (D + 1) * 2 * sizeof(int)));
  int* row_ob = e->cand_tab.as<int>(); int* row_it = row_ob + (size_t)6 * S * (D + 1);
  mpc_intent_enumerate_kernel<<<(unsigned)((S + 127) / 128), 128, 0, e->stream>>>(S, D, NP, n, pred_pos, prob, prev_plan, pos, row_ob, row_it,
                                                                                    scen_a, scen_b, weight, cand);
  CK(cudaGetLastError());
  const int NK = want_last ? N + 1 : N;
  const long long total = (long long)4 * S * NK * D + (long long)2 * S * NK * (D + 1);
  long long blocks = (total + 255) / 256; const long long cap = (long long)e->num_sms * 16;
  if (blocks > cap) blocks = cap;
  mpc_intent_fill_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(S, D, NP, N, p->dynamic_safety_dist, pred_pos, pred_size, row_ob, row_it,
                                                                   scen_a, scen_b, obs_c_a, obs_semi_a, obs_c_b, obs_semi_b,
                                                                   obs_c_last_a, obs_semi_last_a, obs_c_last_b, obs_semi_last_b);
  CK(cudaGetLastError());
  return MPCQP_OK;
}

extern "C" int mpcqp_gather_rows_device(mpcqp_engine* e, int64_t B, int32_t width, const int32_t* idx, const double* src, double* dst) {
  if (!e) return MPCQP_ERR_ARG;
  if (B <= 0 || width <= 0 || !idx || !src || !dst) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  CK(cudaSetDevice(e->device));
  long long blocks = (B * width + 255) / 256; const long long cap = (long long)e->num_sms * 16;
  if (blocks > cap) blocks = cap;
  mpc_gather_rows_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(B, width, idx, src, dst);
  CK(cudaGetLastError());
  return MPCQP_OK;
}

extern "C" int mpcqp_score_candidates_device(mpcqp_engine* e, const mpcqp_mpc_params* p, int32_t B, int32_t R, int32_t n_dynamic,
                                             const double* x, const double* prev_plan, const double* xref, const double* obs_c,
                                             const double* obs_semi, const double* obs_c_last, const double* obs_semi_last, double* score) {
  if (!e) return MPCQP_ERR_ARG;
  if (!p || B <= 0 || R < 0 || n_dynamic < 0 || n_dynamic > R || !x || !xref || !score || (R > 0 && (!obs_c || !obs_semi))) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  if (R > 0 && (!obs_c_last || !obs_semi_last)) {
    e->err = "getSafetyScore reads the obstacle prediction of stage N as well (mpcPlanner.cpp:818-826): obs_c_last / obs_semi_last are required";
    return MPCQP_ERR_ARG;
  }
  CK(cudaSetDevice(e->device));
  const int NS = p->horizon, n = 8 * NS + 5 * (NS - 1);
  long long blocks = ((long long)B + 3) / 4; const long long cap = (long long)e->num_sms * 16;
  if (blocks > cap) blocks = cap;
  mpc_score_kernel<<<(unsigned)blocks, 128, 0, e->stream>>>(B, NS, R, n_dynamic, n, p->dynamic_safety_dist, p->static_safety_dist, x, prev_plan,
                                                              xref, obs_c, obs_semi, obs_c_last, obs_semi_last, score);
  CK(cudaGetLastError());
  return MPCQP_OK;
}

extern "C" int mpcqp_select_candidates_device(mpcqp_engine* e, int32_t S, int32_t C, int32_t n, const int32_t* cand, const double* weight,
                                              const double* score, const double* x_all, int32_t* best, double* weighted, double* plan) {
  if (!e) return MPCQP_ERR_ARG;
  if (S <= 0 || C <= 0 || !cand || !weight || !score || !best || (plan && (!x_all || n <= 0))) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  CK(cudaSetDevice(e->device));
  mpc_select_kernel<<<(unsigned)((S + 127) / 128), 128, 0, e->stream>>>(S, C, n, cand, weight, score, x_all, best, weighted, nullptr);
  CK(cudaGetLastError());
  if (plan) {
    long long blocks = ((long long)S * n + 255) / 256; const long long cap = (long long)e->num_sms * 8;
    if (blocks > cap) blocks = cap;
    mpc_gather_plan_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(S, C, n, cand, best, x_all, plan);
    CK(cudaGetLastError());
  }
  return MPCQP_OK;
}

// ------------------------------------------------------------------------------------------------
// (2c) predictor rollouts on the device: dynamicPredictor::predictor's intentProb and predTraj
// (dynamic_predictor/include/dynamic_predictor/dynamicPredictor.cpp:197-541) for many obstacles at once — the producer of
// updatePredObstacles' arguments.  One warp per (obstacle, intent).  The sampling loops of modelForward / modelTurning run on
// accumulating DOUBLE counters exactly as the reference writes them (every lane walks the loop nest, so all lanes see the
// same counter values and the same sample count; sample s is rolled out by lane s mod 32); the mean and the two-pass
// variance of genTraj are warp reductions.  The occupancy map is free space (isInflatedOccupied == false): perception is
// outside SURVEY.md section 8.
// ------------------------------------------------------------------------------------------------
namespace mpcqp {
struct PredConsts { int numPred; double dt, zScore, minTurn, maxTurn, frontAngle, stopVel, pscale, paramf, paraml, paramr, params; };
constexpr int kPredMax = 64;                    // prediction_size + 1 <= kPredMax

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PASS 0: accumulate sum x, sum y per step; PASS 1: accumulate squared deviations from (mx, my)
template <int PASS>
__device__ __forceinline__ void pred_rollout(const PredConsts& c, int intent, const double* pos, double speed, double angleInit, double angVel,
                                             double endAngle, bool turning, double* ax, double* ay, const double* mx, const double* my) {
  double angle = angleInit;
  double x = pos[0], y = pos[1], vx = speed * cos(angle), vy = speed * sin(angle);
  if (PASS == 0) { ax[0] += x; ay[0] += y; } else { ax[0] += (x - mx[0]) * (x - mx[0]); ay[0] += (y - my[0]) * (y - my[0]); }
  for (int k = 1; k <= c.numPred; ++k) {
    x = x + c.dt * vx; y = y + c.dt * vy;                      // model * currState (dynamicPredictor.cpp:372-376, 446-450)
    if (PASS == 0) { ax[k] += x; ay[k] += y; } else { ax[k] += (x - mx[k]) * (x - mx[k]); ay[k] += (y - my[k]) * (y - my[k]); }
    if (turning) {                                           // :461-471
      angle += angVel * c.dt;
      angle = intent == 1 ? fmin(angle, endAngle) : fmax(angle, endAngle);
      const double v = sqrt(vx * vx + vy * vy);
      vx = v * cos(angle); vy = v * sin(angle);
    }
  }
}

template <int PASS>
__device__ __forceinline__ int pred_samples(const PredConsts& c, int intent, const double* pos, const double* vel, int lane, double* ax, double* ay,
                                            const double* mx, const double* my) {
  const double v = sqrt(vel[0] * vel[0] + vel[1] * vel[1]);
  const double angleInit = atan2(vel[1], vel[0]);
  const double minVel = v - v, maxVel = v + v;
  int s = 0;
  if (intent == 0) {                                         // modelForward, :353-404
    const double minAngle = angleInit - c.frontAngle, maxAngle = angleInit + c.frontAngle;
    for (double i = minAngle; i < maxAngle; i += 0.1)
      for (double j = minVel; j < maxVel; j += 0.1) {
        if ((s & 31) == lane) pred_rollout<PASS>(c, intent, pos, j, i, 0.0, 0.0, false, ax, ay, mx, my);
        ++s;
      }
  } else {                                                   // modelTurning, :406-491
    const double kPi = 3.14159265358979323846;
    double endMin, endMax, minAngVel, maxAngVel;
    if (intent == 1) { endMin = c.frontAngle + angleInit; endMax = (kPi - c.frontAngle) + angleInit; minAngVel = (kPi / 2) / c.maxTurn; maxAngVel = (kPi / 2) / c.minTurn; }
    else { endMin = -(kPi - c.frontAngle) + angleInit; endMax = -c.frontAngle + angleInit; minAngVel = (-kPi / 2) / c.minTurn; maxAngVel = (-kPi / 2) / c.maxTurn; }
    for (double i = minVel; i < maxVel; i += 0.2)
      for (double j = minAngVel; j < maxAngVel; j += 0.2)
        for (double endAngle = endMin; endAngle < endMax; endAngle += 0.2) {
          if ((s & 31) == lane) pred_rollout<PASS>(c, intent, pos, i, angleInit, j, endAngle, true, ax, ay, mx, my);
          ++s;
        }
  }
  return s;
}

__global__ void __launch_bounds__(128) mpc_predict_kernel(const __grid_constant__ PredConsts c, int NOB, int H, const double* __restrict__ pos_hist,
                                                           const double* __restrict__ vel_hist, const double* __restrict__ size,
                                                           double* __restrict__ pred_pos, double* __restrict__ pred_size, double* __restrict__ prob) {
  const int lane = threadIdx.x & 31;
  const int T = c.numPred + 1;
  __shared__ double acc[4][4][kPredMax];                     // per warp: sum x, sum y -> mean x, mean y, then var x, var y
  double* const sx = acc[threadIdx.x >> 5][0]; double* const sy = acc[threadIdx.x >> 5][1];
  double* const vx_ = acc[threadIdx.x >> 5][2]; double* const vy_ = acc[threadIdx.x >> 5][3];
  for (long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); w < (long long)NOB * 4; w += (long long)gridDim.x * 4) {
    const int ob = (int)(w >> 2), intent = (int)(w & 3);
    const double* ph = pos_hist + (long long)ob * H * 3; const double* vh = vel_hist + (long long)ob * H * 3;
    const double pos[3] = {ph[0], ph[1], ph[2]}, vel[3] = {vh[0], vh[1], vh[2]};
    const double sz[3] = {size[ob * 3], size[ob * 3 + 1], size[ob * 3 + 2]};
    double* op = pred_pos + ((long long)ob * 4 + intent) * T * 3; double* os = pred_size + ((long long)ob * 4 + intent) * T * 3;
    const double v = sqrt(vel[0] * vel[0] + vel[1] * vel[1]);
    if (intent == 0 && lane == 0) {                          // intentProb, :197-226 (+ genTransitionMatrix / Vector, :229-281)
      double P[4] = {0.25, 0.25, 0.25, 0.25};
      const double kPi = 3.14159265358979323846;
      for (int j = 2; j < H - 1; ++j) {                        // the reference's last pass (j = H - 1) reads index -1: left out
        const double* prevPos = ph + (H - j - 1) * 3; const double* currPos = ph + (H - j - 2) * 3; const double* older = ph + (H - j) * 3;
        const double* currVel = vh + (H - j - 2) * 3;
        const double prevAngle = atan2(prevPos[1] - older[1], prevPos[0] - older[0]);
        const double currAngle = atan2(currPos[1] - prevPos[1], currPos[0] - prevPos[0]);
        double theta = currAngle - prevAngle;
        if (theta > kPi) theta = theta - 2 * kPi; else if (theta <= -kPi) theta = theta + 2 * kPi;
        const double r = sqrt(currVel[0] * currVel[0] + currVel[1] * currVel[1]);
        double nP[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < 4; ++i) {
          double sc[4] = {1.0, 1.0, 1.0, 1.0}; sc[i] = c.pscale;
          double pf = sc[0] * (exp(-0.5 * (theta / c.paramf) * (theta / c.paramf)) + c.paraml);
          double pl = sc[1] * (c.paraml * (1 + sin(theta)));
          double pr = sc[2] * (c.paramr * (1 - sin(theta)));
          const double ps = (1 - tanh(c.params / sc[3] * r));
          const double sum = pr + pl + pf;
          pr = (1 - ps) * pr / sum; pl = (1 - ps) * pl / sum; pf = (1 - ps) * pf / sum;
          nP[0] += pf * P[i]; nP[1] += pl * P[i]; nP[2] += pr * P[i]; nP[3] += ps * P[i];
        }
        P[0] = nP[0]; P[1] = nP[1]; P[2] = nP[2]; P[3] = nP[3];
      }
      prob[ob * 4] = P[0]; prob[ob * 4 + 1] = P[1]; prob[ob * 4 + 2] = P[2]; prob[ob * 4 + 3] = P[3];
    }
    if (intent == 3 || v <= c.stopVel) {                     // modelStop, :493-505 (genTraj adds 2 sqrt(0) z = 0)
      if (lane == 0) {
        double s0 = sz[0], s1 = sz[1];
        const double g = 2 * fmin(v, c.stopVel) * c.dt;
        for (int k = 0; k < T; ++k) {
          op[k * 3] = pos[0]; op[k * 3 + 1] = pos[1]; op[k * 3 + 2] = pos[2];
          os[k * 3] = s0 + 2 * sqrt(0.0) * c.zScore; os[k * 3 + 1] = s1 + 2 * sqrt(0.0) * c.zScore; os[k * 3 + 2] = sz[2];
          s0 += g; s1 += g;
        }
      }
      continue;
    }
    double ax[kPredMax], ay[kPredMax];
    for (int k = 0; k < T; ++k) { ax[k] = 0.0; ay[k] = 0.0; }
    const int cnt = pred_samples<0>(c, intent, pos, vel, lane, ax, ay, nullptr, nullptr);
    for (int k = 0; k < T; ++k) { const double a = warp_sum_all(ax[k]), b = warp_sum_all(ay[k]); if (lane == 0) { sx[k] = a / cnt; sy[k] = b / cnt; } }
    __syncwarp();
    for (int k = 0; k < T; ++k) { ax[k] = 0.0; ay[k] = 0.0; }
    pred_samples<1>(c, intent, pos, vel, lane, ax, ay, sx, sy);
    for (int k = 0; k < T; ++k) { const double a = warp_sum_all(ax[k]), b = warp_sum_all(ay[k]); if (lane == 0) { vx_[k] = a / cnt; vy_[k] = b / cnt; } }
    __syncwarp();
    for (int k = lane; k < T; k += 32) {                      // genTraj, :507-541
      op[k * 3] = sx[k]; op[k * 3 + 1] = sy[k]; op[k * 3 + 2] = pos[2];
      os[k * 3] = sz[0] + 2 * sqrt(vx_[k]) * c.zScore; os[k * 3 + 1] = sz[1] + 2 * sqrt(vy_[k]) * c.zScore; os[k * 3 + 2] = sz[2];
    }
    __syncwarp();
  }
}
}  // namespace mpcqp

extern "C" void mpcqp_default_predictor_params(mpcqp_predictor_params* p) {
  if (!p) return;
  p->prediction_size = 30; p->prediction_time_step = 0.1; p->min_turning_time = 2.0; p->max_turning_time = 3.0; p->prediction_z_score = 0.674;
  p->max_front_prob = 0.5; p->front_angle_deg = 25.0; p->stop_velocity_threshold = 0.1; p->prob_scale_param = 5.0;
}

extern "C" int mpcqp_predict_device(mpcqp_engine* e, const mpcqp_predictor_params* pp, int32_t num_obstacles, int32_t num_hist, const double* pos_hist,
                                    const double* vel_hist, const double* size, double* pred_pos, double* pred_size, double* intent_prob) {
  if (!e) return MPCQP_ERR_ARG;
  if (!pp || num_obstacles <= 0 || num_hist < 1 || !pos_hist || !vel_hist || !size || !pred_pos || !pred_size || !intent_prob) { e->err = "bad arguments"; return MPCQP_ERR_ARG; }
  if (pp->prediction_size < 1 || pp->prediction_size + 1 > kPredMax || !(pp->prediction_time_step > 0) || !(pp->stop_velocity_threshold > 0) ||
      !(pp->min_turning_time > 0) || !(pp->max_turning_time > 0) || !(3 * pp->max_front_prob - 1 > 0)) { e->err = "bad predictor parameters"; return MPCQP_ERR_ARG; }
  CK(cudaSetDevice(e->device));
  PredConsts c;                                               // initParam, dynamicPredictor.cpp:14-117
  c.numPred = pp->prediction_size; c.dt = pp->prediction_time_step; c.zScore = pp->prediction_z_score; c.minTurn = pp->min_turning_time;
  c.maxTurn = pp->max_turning_time; c.stopVel = pp->stop_velocity_threshold; c.pscale = pp->prob_scale_param;
  c.paraml = (1 - pp->max_front_prob) / (3 * pp->max_front_prob - 1); c.paramr = c.paraml;
  c.frontAngle = pp->front_angle_deg * M_PI / 180;
  c.paramf = sqrt(c.frontAngle * c.frontAngle / (-2 * log(c.paraml * (1 + sin(c.frontAngle)) - c.paraml)));
  c.params = atanh(0.5) / c.stopVel;
  long long blocks = ((long long)num_obstacles * 4 + 3) / 4; const long long cap = (long long)e->num_sms * 8;
  if (blocks > cap) blocks = cap;
  mpc_predict_kernel<<<(unsigned)blocks, 128, 0, e->stream>>>(c, num_obstacles, num_hist, pos_hist, vel_hist, size, pred_pos, pred_size, intent_prob);
  CK(cudaGetLastError());
  return MPCQP_OK;
}

// ------------------------------------------------------------------------------------------------
// (3) OSQP-shaped single problem with explicit CSC data (what OsqpEigen::Solver hands to osqp_setup,
// OsqpEigen/Data.tpp:38-39,77; osqp.h:58).  A problem with the mpcPlanner stage structure
// (mpcPlanner.cpp:932-1146) is parsed on the host into the structured form the stage kernels take and solved by them as a
// batch of one; anything else goes to the dense generic kernel (csrc/mpcqp_dense.cuh).  There is no host solve.
// ------------------------------------------------------------------------------------------------
struct mpcqp_problem {
  DevBufScope bufs;                                          // must stay the first member (see DevBufScope)
  mpcqp_engine* e = nullptr;
  Shape sh; Settings st; mpcqp_settings user;
  std::vector<double> pd, q, x0, g, low, warm_x, warm_y, sol_x, sol_y;
  std::vector<unsigned char> slack;
  bool has_wx = false, has_wy = false, solved = false;
  mpcqp_info info;
  DevBuf d_pd, d_slack, d_q, d_x0, d_g, d_low, d_wx, d_wy, d_x, d_y, d_i, d_d;
  // generic (unstructured) path, csrc/mpcqp_dense.cuh: dense copies of the user's P (mirrored), A, A' and the bounds
  bool dense = false;
  std::vector<int64_t> dPc, dPi, dAc, dAi;
  std::vector<double> dPx, dAx, dl, du;
  DenseDev dd;
  // the caller's CSC data of a structured problem, kept so that mpcqp_update_bounds can move it to the generic path when the
  // new bounds leave the planner's pattern (osqp_update_bounds accepts any l <= u)
  std::vector<int64_t> cPc, cPi, cAc, cAi;
  std::vector<double> cPx, cAx;
  DevBufScopeEnd bufs_end;                                   // must stay the last member
};

namespace mpcqp_dense {
int grid_size(int B, int n, int m, int device);
int launch(const Batch& bt, int grid, const Settings& st, cudaStream_t stream);
}
namespace mpcqp_band {
void pattern_bind(Pattern* pt, int n, int m, int w, int nnzP, int nnzA, const int* dev_flat, const int* off);
int grid_size(int B, int N, int w, int device);
int launch(const Batch& bt, int grid, const Settings& st, cudaStream_t stream);
}
static_assert(sizeof(mpcqp_band::Settings) == sizeof(Settings), "the band path reads mpcqp::Settings by layout");
static_assert(sizeof(mpcqp_dense::Settings) == sizeof(Settings), "the dense path reads mpcqp::Settings by layout");

namespace {
const double kBoundInf = 1e20;     // OSQP_INFTY is 1e30 (constants.h:78); IEEE inf is what the reference passes

// l / u of the stage structure -> x0, stage-uniform box, obstacle lower bounds.  Returns 0 or an error code.
int parse_bounds(const Shape& sh, const double* l, const double* u, double* x0, double* blo, double* bhi, double* low, double* obs_hi, std::string* err) {
  const int NS = sh.NS, N = NS - 1, R = sh.R;
  for (int i = 0; i < sh.m; ++i) if (l[i] > u[i]) { *err = "lower bound > upper bound at row " + std::to_string(i); return MPCQP_ERR_DATA; }
  for (int i = 0; i < 8 * NS; ++i) {
    if (l[i] != u[i]) { *err = "dynamics row " + std::to_string(i) + " is not an equality"; return MPCQP_ERR_STRUCTURE; }
    if (i < 8) x0[i] = -l[i]; else if (l[i] != 0.0) { *err = "dynamics row " + std::to_string(i) + " has a non-zero right-hand side"; return MPCQP_ERR_STRUCTURE; }
  }
  for (int k = 0; k < NS; ++k) for (int j = 0; j < 8; ++j) {
    const int i = 8 * NS + 8 * k + j;
    if (k == 0) { blo[j] = l[i]; bhi[j] = u[i]; }
    else if (l[i] != blo[j] || u[i] != bhi[j]) { *err = "state box is not stage-uniform (row " + std::to_string(i) + ")"; return MPCQP_ERR_STRUCTURE; }
  }
  for (int k = 0; k < N; ++k) for (int j = 0; j < 5; ++j) {
    const int i = 16 * NS + 5 * k + j;
    if (k == 0) { blo[8 + j] = l[i]; bhi[8 + j] = u[i]; }
    else if (l[i] != blo[8 + j] || u[i] != bhi[8 + j]) { *err = "input box is not stage-uniform (row " + std::to_string(i) + ")"; return MPCQP_ERR_STRUCTURE; }
  }
  for (int k = 0; k < N; ++k) for (int o = 0; o < R; ++o) {
    const int i = 16 * NS + 5 * N + k * R + o;
    if (!(u[i] >= kBoundInf)) { *err = "obstacle row " + std::to_string(i) + " has a finite upper bound"; return MPCQP_ERR_STRUCTURE; }
    // one "infinity" for all of them: IEEE inf (the reference) or a finite stand-in such as OSQP_INFTY (the certificates differ)
    if (k == 0 && o == 0) *obs_hi = u[i];
    else if (u[i] != *obs_hi) { *err = "obstacle rows mix different infinite upper bounds"; return MPCQP_ERR_STRUCTURE; }
    low[k * R + o] = l[i];
  }
  if (N * R == 0) *obs_hi = INFINITY;
  return MPCQP_OK;
}

bool set_coef(double* slot, bool* seen, double v) { if (!*seen) { *slot = v; *seen = true; return true; } return *slot == v; }

int parse_structure(int64_t n, int64_t m, const int64_t* Pp, const int64_t* Pi, const double* Px, const int64_t* Ap, const int64_t* Ai,
                    const double* Ax, Shape* sh, std::vector<double>* pd, std::vector<unsigned char>* slack, std::vector<double>* g,
                    std::string* err) {
  if (n < 34 || (n + 5) % 13 != 0) { *err = "n is not 8*horizon + 5*(horizon-1)"; return MPCQP_ERR_STRUCTURE; }
  const int NS = (int)((n + 5) / 13), N = NS - 1;
  const int64_t mr = m - 16 * NS - 5 * N;
  if (mr < 0 || mr % N != 0) { *err = "m is not 16*horizon + 5*(horizon-1) + num_obs*(horizon-1) (field-of-view half-space rows are not supported)"; return MPCQP_ERR_STRUCTURE; }
  const int R = (int)(mr / N);
  sh->NS = NS; sh->R = R; sh->n = (int)n; sh->m = (int)m;
  pd->assign((size_t)NS * 13, 0.0);
  auto stage_slot = [&](int64_t v) { return v < 8 * NS ? (int)(v / 8) * 13 + (int)(v % 8) : (int)((v - 8 * NS) / 5) * 13 + 8 + (int)((v - 8 * NS) % 5); };
  for (int64_t j = 0; j < n; ++j) for (int64_t t = Pp[j]; t < Pp[j + 1]; ++t) {
    if (Pi[t] != j) { if (Px[t] == 0.0) continue; *err = "P is not diagonal"; return MPCQP_ERR_STRUCTURE; }
    if (Px[t] < 0.0) { *err = "P has a negative diagonal entry"; return MPCQP_ERR_DATA; }
    (*pd)[stage_slot(j)] = Px[t];
  }
  slack->assign((size_t)N * (R > 0 ? R : 1), 255);
  g->assign((size_t)N * (R > 0 ? R : 1) * 3, 0.0);
  bool s_apv = false, s_bpa = false, s_bva = false;
  const int64_t base = 16 * NS + 5 * N;
  std::vector<char> have_neg((size_t)8 * NS, 0), have_id((size_t)n, 0);
  auto bad = [&](int64_t j, int64_t r) { *err = "A(" + std::to_string(r) + "," + std::to_string(j) + ") does not belong to the mpcPlanner constraint structure"; return MPCQP_ERR_STRUCTURE; };
  for (int64_t j = 0; j < n; ++j) {
    const bool is_state = j < 8 * NS;
    const int k = is_state ? (int)(j / 8) : (int)((j - 8 * NS) / 5), c = is_state ? (int)(j % 8) : (int)((j - 8 * NS) % 5);
    for (int64_t t = Ap[j]; t < Ap[j + 1]; ++t) {
      const int64_t r = Ai[t]; const double v = Ax[t];
      if (r < 0 || r >= m) { *err = "row index out of range"; return MPCQP_ERR_DATA; }
      if (v == 0.0) continue;
      if (r == 8 * NS + j) { if (v != 1.0) return bad(j, r); have_id[j] = 1; continue; }
      if (r >= base) {                                   // obstacle row of stage k only
        const int64_t kk = (r - base) / (R > 0 ? R : 1), o = (r - base) % (R > 0 ? R : 1);
        if (R == 0 || kk != k || k >= N) return bad(j, r);
        if (is_state) { if (c > 2) return bad(j, r); (*g)[(kk * R + o) * 3 + c] = v; }
        else { if (c < 3 || v != -1.0 || (*slack)[kk * R + o] != 255) return bad(j, r); (*slack)[kk * R + o] = (unsigned char)(c - 3); }
        continue;
      }
      if (r >= 8 * NS) return bad(j, r);
      const int rk = (int)(r / 8), ri = (int)(r % 8);
      if (is_state) {
        if (rk == k) { if (ri != c || v != -1.0) return bad(j, r); have_neg[r] = 1; continue; }
        if (rk != k + 1) return bad(j, r);
        if (c < 6 && ri == c) { if (v != 1.0) return bad(j, r); continue; }                 // Ad diagonal
        if (c >= 3 && c < 6 && ri == c - 3) { if (!set_coef(&sh->a_pv, &s_apv, v)) return bad(j, r); continue; }
        return bad(j, r);
      }
      if (rk != k + 1) return bad(j, r);
      if (c < 3 && ri == c) { if (!set_coef(&sh->b_pa, &s_bpa, v)) return bad(j, r); continue; }
      if (c < 3 && ri == 3 + c) { if (!set_coef(&sh->b_va, &s_bva, v)) return bad(j, r); continue; }
      if (c >= 3 && ri == 3 + c) { if (v != 1.0) return bad(j, r); continue; }
      return bad(j, r);
    }
  }
  for (size_t i = 0; i < have_neg.size(); ++i) if (!have_neg[i]) { *err = "dynamics row " + std::to_string(i) + " lacks its -1 entry"; return MPCQP_ERR_STRUCTURE; }
  for (int64_t j = 0; j < n; ++j) if (!have_id[j]) { *err = "box row of variable " + std::to_string(j) + " is missing"; return MPCQP_ERR_STRUCTURE; }
  for (int i = 0; i < N * R; ++i) if ((*slack)[i] == 255) { *err = "obstacle row " + std::to_string(i) + " has no slack entry"; return MPCQP_ERR_STRUCTURE; }
  if (!s_apv) sh->a_pv = 0.0;
  if (!s_bpa) sh->b_pa = 0.0;
  if (!s_bva) sh->b_va = 0.0;
  return MPCQP_OK;
}
}  // namespace

namespace {
// Generic path.  validate_data of osqp_setup: l <= u, P upper triangular, indices in range, monotone column pointers.
int validate_csc_pattern(mpcqp_engine* e, int64_t n, int64_t m, const int64_t* Pc, const int64_t* Pi, const int64_t* Ac, const int64_t* Ai) {
  if (Pc[0] != 0 || Ac[0] != 0) { e->err = "column pointers do not start at 0"; return MPCQP_ERR_DATA; }
  for (int64_t j = 0; j < n; ++j) if (Pc[j + 1] < Pc[j] || Ac[j + 1] < Ac[j]) { e->err = "column pointers are not monotone"; return MPCQP_ERR_DATA; }
  if ((Pc[n] > 0 && !Pi) || (Ac[n] > 0 && !Ai)) { e->err = "null row-index array"; return MPCQP_ERR_DATA; }
  for (int64_t j = 0; j < n; ++j) {
    if (Pc[j + 1] < Pc[j] || Ac[j + 1] < Ac[j]) { e->err = "column pointers are not monotone"; return MPCQP_ERR_DATA; }
    for (int64_t t = Pc[j]; t < Pc[j + 1]; ++t) if (Pi[t] < 0 || Pi[t] > j) { e->err = "P is not upper triangular (entry " + std::to_string(Pi[t]) + "," + std::to_string(j) + ")"; return MPCQP_ERR_DATA; }
    for (int64_t t = Ac[j]; t < Ac[j + 1]; ++t) if (Ai[t] < 0 || Ai[t] >= m) { e->err = "A row index out of range in column " + std::to_string(j); return MPCQP_ERR_DATA; }
  }
  return MPCQP_OK;
}
int validate_csc(mpcqp_engine* e, int64_t n, int64_t m, const int64_t* Pc, const int64_t* Pi, const int64_t* Ac, const int64_t* Ai) {
  if (n + m > 4096) { e->err = "unstructured problem with n + m > 4096: the dense generic kernel does not take it"; return MPCQP_ERR_STRUCTURE; }
  return validate_csc_pattern(e, n, m, Pc, Pi, Ac, Ai);
}

// B QPs sharing one CSC pattern, host arrays in, host arrays out: upload (the CSC data and the vectors are all HBM ever
// sees of the inputs), ONE launch of mpcqp_dense_solve_kernel (persistent, one CTA per QP at a time), download.
int run_dense(mpcqp_engine* e, DenseDev& d, const Settings& st, int B, int n, int m, const int64_t* Pc, const int64_t* Pi, const double* Px,
              const double* q, const int64_t* Ac, const int64_t* Ai, const double* Ax, const double* l, const double* u,
              const double* wx, const double* wy, double* x, double* y, int32_t* info_i, double* info_d) {
  const size_t one = sizeof(double), nnzP = (size_t)Pc[n], nnzA = (size_t)Ac[n], mm = (size_t)(m > 0 ? m : 1);
  CK(cudaSetDevice(e->device));
  // Sparse path first: a KKT matrix that is narrow-banded under reverse Cuthill-McKee (polyTrajSolver's is: half-bandwidth
  // 8 .. 12) is factored and solved in band form by one warp per QP (csrc/mpcqp_band.cuh); anything else runs dense.
  if (e->force_generic != 4) {
    std::vector<int> flat; int off[14], w = 0;
    if (mpcqp_band::pattern_build(n, m, Pc, Pi, Ac, Ai, &flat, off, &w)) {
      const int bgrid = mpcqp_band::grid_size(B, n + m, w, e->device);
      if (bgrid > 0) {
        auto upb = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
          cudaError_t r = b.need(bytes ? bytes : 8); if (r != cudaSuccess || !bytes) return r;
          return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream);
        };
        CK(upb(d.pat, flat.data(), flat.size() * sizeof(int)));
        CK(upb(d.Px, Px, (size_t)B * nnzP * one)); CK(upb(d.Ax, Ax, (size_t)B * nnzA * one));
        CK(upb(d.q, q, (size_t)B * n * one)); CK(upb(d.l, l, (size_t)B * m * one)); CK(upb(d.u, u, (size_t)B * m * one));
        if (wx) CK(upb(d.wx, wx, (size_t)B * n * one));
        if (wy && m > 0) CK(upb(d.wy, wy, (size_t)B * m * one));
        CK(d.x.need((size_t)B * n * one)); CK(d.y.need((size_t)B * mm * one)); CK(d.ii.need((size_t)B * 3 * sizeof(int32_t))); CK(d.dd.need((size_t)B * 3 * one));
        const size_t bws = mpcqp_band::ws_doubles(n, m, (int)nnzP, (int)nnzA);
        CK(d.ws.need((size_t)bgrid * bws * one));
        mpcqp_band::Batch bb; memset(&bb, 0, sizeof bb);
        bb.B = B;
        mpcqp_band::pattern_bind(&bb.pt, n, m, w, (int)nnzP, (int)nnzA, d.pat.as<int>(), off);
        bb.Px = d.Px.as<double>(); bb.Ax = d.Ax.as<double>(); bb.q = d.q.as<double>(); bb.l = d.l.as<double>(); bb.u = d.u.as<double>();
        bb.warm_x = wx ? d.wx.as<double>() : nullptr; bb.warm_y = (wy && m > 0) ? d.wy.as<double>() : nullptr;
        bb.ws = d.ws.as<double>(); bb.ws_stride = (long long)bws;
        bb.x = d.x.as<double>(); bb.y = (y && m > 0) ? d.y.as<double>() : nullptr; bb.info_i = d.ii.as<int32_t>(); bb.info_d = d.dd.as<double>();
        mpcqp_band::Settings bs; memcpy(&bs, &st, sizeof bs);
        e->last_launches = 0;
        CK(cudaEventRecord(e->ev0, e->stream));
        CK(cudaEventRecord(e->evs, e->stream));
        const int rc = mpcqp_band::launch(bb, bgrid, bs, e->stream);
        if (rc) { e->err = std::string("mpcqp_band_solve_kernel: ") + cudaGetErrorString((cudaError_t)rc); return MPCQP_ERR_CUDA; }
        CK(cudaEventRecord(e->ev1, e->stream));
        e->last_launches = 1; e->last_fast = 5;
        CK(cudaMemcpyAsync(x, d.x.p, (size_t)B * n * one, cudaMemcpyDeviceToHost, e->stream));
        if (y && m > 0) CK(cudaMemcpyAsync(y, d.y.p, (size_t)B * m * one, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(info_i, d.ii.p, (size_t)B * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(info_d, d.dd.p, (size_t)B * 3 * one, cudaMemcpyDeviceToHost, e->stream));
        return mpcqp_engine_sync(e);
      }
    }
  }
  const int grid = mpcqp_dense::grid_size(B, n, m, e->device);
  const size_t wsd = mpcqp_dense::ws_doubles(n, m);
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t r = b.need(bytes ? bytes : 8); if (r != cudaSuccess || !bytes) return r;
    return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream);
  };
  CK(up(d.Pc, Pc, (size_t)(n + 1) * 8)); CK(up(d.Pi, Pi, nnzP * 8)); CK(up(d.Px, Px, (size_t)B * nnzP * one));
  CK(up(d.Ac, Ac, (size_t)(n + 1) * 8)); CK(up(d.Ai, Ai, nnzA * 8)); CK(up(d.Ax, Ax, (size_t)B * nnzA * one));
  CK(up(d.q, q, (size_t)B * n * one)); CK(up(d.l, l, (size_t)B * m * one)); CK(up(d.u, u, (size_t)B * m * one));
  if (wx) CK(up(d.wx, wx, (size_t)B * n * one));
  if (wy && m > 0) CK(up(d.wy, wy, (size_t)B * m * one));
  CK(d.x.need((size_t)B * n * one)); CK(d.y.need((size_t)B * mm * one)); CK(d.ii.need((size_t)B * 3 * sizeof(int32_t))); CK(d.dd.need((size_t)B * 3 * one));
  CK(d.ws.need((size_t)grid * wsd * one));
  mpcqp_dense::Batch bt; memset(&bt, 0, sizeof bt);
  bt.B = B; bt.n = n; bt.m = m; bt.nnzP = (long long)nnzP; bt.nnzA = (long long)nnzA;
  bt.Pc = d.Pc.as<int64_t>(); bt.Pi = d.Pi.as<int64_t>(); bt.Ac = d.Ac.as<int64_t>(); bt.Ai = d.Ai.as<int64_t>();
  bt.Px = d.Px.as<double>(); bt.Ax = d.Ax.as<double>(); bt.q = d.q.as<double>(); bt.l = d.l.as<double>(); bt.u = d.u.as<double>();
  bt.warm_x = wx ? d.wx.as<double>() : nullptr; bt.warm_y = (wy && m > 0) ? d.wy.as<double>() : nullptr;
  bt.ws = d.ws.as<double>(); bt.ws_stride = (long long)wsd;
  bt.x = d.x.as<double>(); bt.y = (y && m > 0) ? d.y.as<double>() : nullptr; bt.info_i = d.ii.as<int32_t>(); bt.info_d = d.dd.as<double>();
  mpcqp_dense::Settings ds; memcpy(&ds, &st, sizeof ds);
  e->last_launches = 0;
  CK(cudaEventRecord(e->ev0, e->stream));
  CK(cudaEventRecord(e->evs, e->stream));
  const int rc = mpcqp_dense::launch(bt, grid, ds, e->stream);
  if (rc) { e->err = std::string("mpcqp_dense_solve_kernel: ") + cudaGetErrorString((cudaError_t)rc); return MPCQP_ERR_CUDA; }
  CK(cudaEventRecord(e->ev1, e->stream));
  e->last_launches = 1; e->last_fast = 4;
  CK(cudaMemcpyAsync(x, d.x.p, (size_t)B * n * one, cudaMemcpyDeviceToHost, e->stream));
  if (y && m > 0) CK(cudaMemcpyAsync(y, d.y.p, (size_t)B * m * one, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(info_i, d.ii.p, (size_t)B * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(info_d, d.dd.p, (size_t)B * 3 * one, cudaMemcpyDeviceToHost, e->stream));
  return mpcqp_engine_sync(e);
}

int setup_dense(mpcqp_engine* e, mpcqp_problem* pr, int64_t n, int64_t m, const int64_t* Pc, const int64_t* Pi, const double* Px,
                const double* q, const int64_t* Ac, const int64_t* Ai, const double* Ax, const double* l, const double* u) {
  int rc = validate_csc(e, n, m, Pc, Pi, Ac, Ai); if (rc) return rc;
  for (int64_t i = 0; i < m; ++i) if (l[i] > u[i]) { e->err = "lower bound > upper bound at row " + std::to_string(i); return MPCQP_ERR_DATA; }
  pr->dense = true;
  pr->sh = Shape(); pr->sh.n = (int)n; pr->sh.m = (int)m;
  pr->dPc.assign(Pc, Pc + n + 1); pr->dPi.assign(Pi, Pi + Pc[n]); pr->dPx.assign(Px, Px + Pc[n]);
  pr->dAc.assign(Ac, Ac + n + 1); pr->dAi.assign(Ai, Ai + Ac[n]); pr->dAx.assign(Ax, Ax + Ac[n]);
  pr->q.assign(q, q + n); pr->dl.assign(l, l + m); pr->du.assign(u, u + m);
  pr->sol_x.assign((size_t)n, 0.0); pr->sol_y.assign((size_t)(m > 0 ? m : 1), 0.0);
  memset(&pr->info, 0, sizeof pr->info);
  pr->info.status_val = MPCQP_UNSOLVED;
  return MPCQP_OK;
}
}  // namespace

extern "C" int mpcqp_setup(mpcqp_engine* e, mpcqp_problem** out, int64_t n, int64_t m, const int64_t* P_colptr,
                           const int64_t* P_rowidx, const double* P_val, const double* q, const int64_t* A_colptr,
                           const int64_t* A_rowidx, const double* A_val, const double* l, const double* u,
                           const mpcqp_settings* s) {
  if (!e) return MPCQP_ERR_ARG;
  if (!out) { e->err = "null out pointer"; return MPCQP_ERR_ARG; }
  *out = nullptr;
  if (n <= 0 || m < 0 || !P_colptr || !q || !A_colptr || (m > 0 && (!l || !u))) { e->err = "null array or non-positive size"; return MPCQP_ERR_DATA; }
  const auto t0 = std::chrono::steady_clock::now();
  mpcqp_problem* pr = new mpcqp_problem();
  pr->e = e; ++e->live_problems;
  int rc = check_settings(e, s, &pr->st);
  if (rc) { mpcqp_cleanup(pr); return rc; }
  pr->user = *s;
  rc = validate_csc_pattern(e, n, m, P_colptr, P_rowidx, A_colptr, A_rowidx);       // before any column is walked
  if (rc) { mpcqp_cleanup(pr); return rc; }
  rc = parse_structure(n, m, P_colptr, P_rowidx, P_val, A_colptr, A_rowidx, A_val, &pr->sh, &pr->pd, &pr->slack, &pr->g, &e->err);
  if (rc == MPCQP_OK) {
    const int NS = pr->sh.NS, N = NS - 1, R = pr->sh.R;
    pr->x0.assign(8, 0.0); pr->low.assign((size_t)N * (R > 0 ? R : 1), 0.0);
    rc = parse_bounds(pr->sh, l, u, pr->x0.data(), pr->sh.blo, pr->sh.bhi, pr->low.data(), &pr->sh.obs_hi, &e->err);
  }
  if (rc == MPCQP_ERR_STRUCTURE) {            // not an mpcPlanner QP (e.g. polyTrajSolver.cpp:162-222), or the planner's matrices with
                                              // bounds outside its pattern (non-uniform box, finite obstacle upper bound): generic kernel
    rc = setup_dense(e, pr, n, m, P_colptr, P_rowidx, P_val, q, A_colptr, A_rowidx, A_val, l, u);
    if (rc) { mpcqp_cleanup(pr); return rc; }
    pr->info.setup_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *out = pr;
    return MPCQP_OK;
  }
  if (rc) { mpcqp_cleanup(pr); return rc; }
  if (n + m <= 4096) {                        // kept for a later mpcqp_update_bounds that leaves the pattern
    pr->cPc.assign(P_colptr, P_colptr + n + 1); pr->cPi.assign(P_rowidx, P_rowidx + P_colptr[n]); pr->cPx.assign(P_val, P_val + P_colptr[n]);
    pr->cAc.assign(A_colptr, A_colptr + n + 1); pr->cAi.assign(A_rowidx, A_rowidx + A_colptr[n]); pr->cAx.assign(A_val, A_val + A_colptr[n]);
  }
  pr->q.assign(q, q + n);
  pr->sol_x.assign((size_t)n, 0.0); pr->sol_y.assign((size_t)m, 0.0);
  memset(&pr->info, 0, sizeof pr->info);
  pr->info.status_val = MPCQP_UNSOLVED;
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t r = b.need(bytes ? bytes : 8); if (r != cudaSuccess) return r;
    return bytes ? cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream) : cudaSuccess;
  };
  cudaError_t ce = cudaSetDevice(e->device);
  if (ce == cudaSuccess) ce = up(pr->d_pd, pr->pd.data(), pr->pd.size() * sizeof(double));
  if (ce == cudaSuccess) ce = up(pr->d_slack, pr->slack.data(), pr->slack.size());
  if (ce == cudaSuccess) ce = up(pr->d_q, pr->q.data(), pr->q.size() * sizeof(double));
  if (ce == cudaSuccess) ce = up(pr->d_x0, pr->x0.data(), 8 * sizeof(double));
  if (ce == cudaSuccess) ce = up(pr->d_g, pr->g.data(), pr->g.size() * sizeof(double));
  if (ce == cudaSuccess) ce = up(pr->d_low, pr->low.data(), pr->low.size() * sizeof(double));
  if (ce == cudaSuccess) ce = pr->d_x.need((size_t)n * sizeof(double));
  if (ce == cudaSuccess) ce = pr->d_y.need((size_t)(m > 0 ? m : 1) * sizeof(double));
  if (ce == cudaSuccess) ce = pr->d_i.need(3 * sizeof(int32_t));
  if (ce == cudaSuccess) ce = pr->d_d.need(3 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->err = std::string("mpcqp_setup: ") + cudaGetErrorString(ce); mpcqp_cleanup(pr); return MPCQP_ERR_CUDA; }
  pr->info.setup_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *out = pr;
  return MPCQP_OK;
}

extern "C" int mpcqp_warm_start(mpcqp_problem* pr, const double* x, const double* y) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  if (!x || !y) { pr->e->err = "null warm start"; return MPCQP_ERR_ARG; }
  pr->warm_x.assign(x, x + pr->sh.n); pr->warm_y.assign(y, y + pr->sh.m);
  pr->has_wx = pr->has_wy = true;
  return MPCQP_OK;
}

extern "C" int mpcqp_warm_start_x(mpcqp_problem* pr, const double* x) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  if (!x) { pr->e->err = "null warm start"; return MPCQP_ERR_ARG; }
  pr->warm_x.assign(x, x + pr->sh.n);
  pr->has_wx = true;
  return MPCQP_OK;
}

extern "C" int mpcqp_update_lin_cost(mpcqp_problem* pr, const double* q_new) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  mpcqp_engine* e = pr->e;
  if (!q_new) { e->err = "null q"; return MPCQP_ERR_ARG; }
  pr->q.assign(q_new, q_new + pr->sh.n);
  if (pr->dense) return MPCQP_OK;              // uploaded by the next solve
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(pr->d_q.p, pr->q.data(), pr->q.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MPCQP_OK;
}

extern "C" int mpcqp_update_bounds(mpcqp_problem* pr, const double* l_new, const double* u_new) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  mpcqp_engine* e = pr->e;
  if (!l_new || !u_new) { e->err = "null bounds"; return MPCQP_ERR_ARG; }
  if (pr->dense) {
    for (int i = 0; i < pr->sh.m; ++i) if (l_new[i] > u_new[i]) { e->err = "lower bound > upper bound at row " + std::to_string(i); return MPCQP_ERR_DATA; }
    pr->dl.assign(l_new, l_new + pr->sh.m); pr->du.assign(u_new, u_new + pr->sh.m);
    return MPCQP_OK;
  }
  Shape sh = pr->sh; std::vector<double> x0(8), low(pr->low.size());
  int rc = parse_bounds(sh, l_new, u_new, x0.data(), sh.blo, sh.bhi, low.data(), &sh.obs_hi, &e->err);
  if (rc == MPCQP_ERR_STRUCTURE && !pr->cPc.empty()) {
    // bounds outside the planner's pattern: osqp_update_bounds would take them, so the problem moves to the generic path
    const std::vector<double> q = pr->q, sx = pr->sol_x, sy = pr->sol_y;
    const mpcqp_info info = pr->info; const Shape keep = pr->sh;
    rc = setup_dense(e, pr, keep.n, keep.m, pr->cPc.data(), pr->cPi.data(), pr->cPx.data(), q.data(), pr->cAc.data(), pr->cAi.data(),
                     pr->cAx.data(), l_new, u_new);
    if (rc) { pr->dense = false; pr->sh = keep; return rc; }
    pr->sol_x = sx; pr->sol_y = sy; pr->info = info;        // the previous solution / a pending warm start survive (same n, m)
    return MPCQP_OK;
  }
  if (rc) return rc;
  pr->sh = sh; pr->x0 = x0; pr->low = low;
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(pr->d_x0.p, pr->x0.data(), 8 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  if (!pr->low.empty() && pr->sh.R > 0) CK(cudaMemcpyAsync(pr->d_low.p, pr->low.data(), pr->low.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MPCQP_OK;
}

extern "C" int mpcqp_solve(mpcqp_problem* pr) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  mpcqp_engine* e = pr->e;
  const int n = pr->sh.n, m = pr->sh.m;
  const auto t0 = std::chrono::steady_clock::now();
  CK(cudaSetDevice(e->device));
  // osqp_solve starts from the workspace iterates: an explicit warm start if one was given, else the previous
  // solution (OSQP keeps its iterates between solves when settings->warm_start is on), else zeros.
  const bool ws = pr->st.warm_start != 0;
  const double* wx = nullptr; const double* wy = nullptr;
  // an infeasible / non-convex outcome leaves no iterate behind: store_solution fills the solution with OSQP_NAN and
  // cold-starts the workspace (x = z = y = 0), so the next solve of the same object starts from zeros
  const int64_t ps = pr->info.status_val;
  const bool prev = pr->solved && !(ps == MPCQP_PRIMAL_INFEASIBLE || ps == MPCQP_PRIMAL_INFEASIBLE_INACCURATE || ps == MPCQP_DUAL_INFEASIBLE ||
                                    ps == MPCQP_DUAL_INFEASIBLE_INACCURATE || ps == MPCQP_NON_CVX);
  if (ws && pr->has_wx) wx = pr->warm_x.data(); else if (ws && prev) wx = pr->sol_x.data();
  // osqp_warm_start_x replaces x only: the dual iterate of the previous solve stays (osqp.h:165)
  if (ws && pr->has_wy) wy = pr->warm_y.data(); else if (ws && prev) wy = pr->sol_y.data();
  if (pr->dense) {
    int32_t hi[3]; double hd[3];
    int rc = run_dense(e, pr->dd, pr->st, 1, n, m, pr->dPc.data(), pr->dPi.data(), pr->dPx.data(), pr->q.data(), pr->dAc.data(), pr->dAi.data(),
                       pr->dAx.data(), pr->dl.data(), pr->du.data(), wx, wy, pr->sol_x.data(), pr->sol_y.data(), hi, hd);
    if (rc) return rc;
    pr->info.status_val = hi[0]; pr->info.iter = hi[1]; pr->info.rho_updates = hi[2];
    pr->info.obj_val = hd[0]; pr->info.pri_res = hd[1]; pr->info.dua_res = hd[2];
    pr->info.solve_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pr->solved = true; pr->has_wx = pr->has_wy = false;
    return MPCQP_OK;
  }
  if (wx) { CK(pr->d_wx.need((size_t)n * sizeof(double))); CK(cudaMemcpyAsync(pr->d_wx.p, wx, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream)); }
  if (wy && m > 0) { CK(pr->d_wy.need((size_t)m * sizeof(double))); CK(cudaMemcpyAsync(pr->d_wy.p, wy, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, e->stream)); }
  Batch bt; memset(&bt, 0, sizeof bt);
  bt.pd = pr->d_pd.as<double>(); bt.slack = pr->d_slack.as<unsigned char>(); bt.q = pr->d_q.as<double>(); bt.x0 = pr->d_x0.as<double>();
  bt.g = pr->d_g.as<double>(); bt.low = pr->d_low.as<double>();
  bt.warm_x = wx ? pr->d_wx.as<double>() : nullptr; bt.warm_y = (wy && m > 0) ? pr->d_wy.as<double>() : nullptr;
  bt.x = pr->d_x.as<double>(); bt.y = pr->d_y.as<double>();
  int32_t* di = pr->d_i.as<int32_t>(); double* dd = pr->d_d.as<double>();
  bt.status = di; bt.iter = di + 1; bt.rho_updates = di + 2; bt.obj = dd; bt.pri_res = dd + 1; bt.dua_res = dd + 2; bt.B = 1;
  e->last_launches = 0;
  CK(cudaEventRecord(e->ev0, e->stream));
  int rc = launch_solve(e, pr->sh, pr->st, bt); if (rc) return rc;
  CK(cudaEventRecord(e->ev1, e->stream));
  int32_t hi[3]; double hd[3];
  CK(cudaMemcpyAsync(pr->sol_x.data(), pr->d_x.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (m > 0) CK(cudaMemcpyAsync(pr->sol_y.data(), pr->d_y.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(hi, di, sizeof hi, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(hd, dd, sizeof hd, cudaMemcpyDeviceToHost, e->stream));
  rc = mpcqp_engine_sync(e); if (rc) return rc;
  pr->info.status_val = hi[0]; pr->info.iter = hi[1]; pr->info.rho_updates = hi[2];
  pr->info.obj_val = hd[0]; pr->info.pri_res = hd[1]; pr->info.dua_res = hd[2];
  pr->info.solve_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  pr->solved = true; pr->has_wx = pr->has_wy = false;
  return MPCQP_OK;
}

extern "C" int mpcqp_get_info(const mpcqp_problem* pr, mpcqp_info* info) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  if (!info) return MPCQP_ERR_ARG;
  *info = pr->info;
  return MPCQP_OK;
}

extern "C" int mpcqp_get_solution(const mpcqp_problem* pr, double* x, double* y) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  if (!pr->solved) { pr->e->err = "no solution yet: call mpcqp_solve first"; return MPCQP_ERR_NOT_INIT; }
  if (x) memcpy(x, pr->sol_x.data(), pr->sol_x.size() * sizeof(double));
  if (y) memcpy(y, pr->sol_y.data(), pr->sol_y.size() * sizeof(double));
  return MPCQP_OK;
}

// B unstructured QPs sharing one CSC pattern in one launch (include/mpcqp_b200.h section 3b).
extern "C" int mpcqp_solve_qp_batch_host(mpcqp_engine* e, const mpcqp_settings* s, int32_t B, int64_t n, int64_t m,
                                         const int64_t* P_colptr, const int64_t* P_rowidx, const double* P_val, const double* q,
                                         const int64_t* A_colptr, const int64_t* A_rowidx, const double* A_val, const double* l,
                                         const double* u, const double* warm_x, const double* warm_y, double* x, double* y,
                                         int32_t* status, int32_t* iter, int32_t* rho_updates, double* obj, double* pri_res, double* dua_res) {
  if (!e) return MPCQP_ERR_ARG;
  if (B <= 0 || n <= 0 || m < 0 || !P_colptr || !q || !A_colptr || !x || !status || (m > 0 && (!l || !u))) { e->err = "null array or non-positive size"; return MPCQP_ERR_DATA; }
  Settings st;
  int rc = check_settings(e, s, &st); if (rc) return rc;
  rc = validate_csc(e, n, m, P_colptr, P_rowidx, A_colptr, A_rowidx); if (rc) return rc;
  for (int64_t i = 0; i < (int64_t)B * m; ++i) if (l[i] > u[i]) { e->err = "lower bound > upper bound at row " + std::to_string(i % m) + " of problem " + std::to_string(i / m); return MPCQP_ERR_DATA; }
  std::vector<int32_t> ii((size_t)B * 3); std::vector<double> dd((size_t)B * 3);
  rc = run_dense(e, e->dense, st, B, (int)n, (int)m, P_colptr, P_rowidx, P_val, q, A_colptr, A_rowidx, A_val, l, u, warm_x, warm_y, x, y, ii.data(), dd.data());
  if (rc) return rc;
  for (int b = 0; b < B; ++b) {
    status[b] = ii[3 * b]; if (iter) iter[b] = ii[3 * b + 1]; if (rho_updates) rho_updates[b] = ii[3 * b + 2];
    if (obj) obj[b] = dd[3 * b]; if (pri_res) pri_res[b] = dd[3 * b + 1]; if (dua_res) dua_res[b] = dd[3 * b + 2];
  }
  return MPCQP_OK;
}

extern "C" int mpcqp_cleanup(mpcqp_problem* pr) {
  if (!pr) return MPCQP_ERR_NOT_INIT;
  cudaSetDevice(pr->e->device);
  for (DevBuf* b : pr->bufs.all) b->release();
  mpcqp_engine* e = pr->e;
  delete pr;
  if (e && --e->live_problems <= 0 && e->destroy_pending) { e->live_problems = 0; mpcqp_engine_destroy(e); }
  return MPCQP_OK;
}
