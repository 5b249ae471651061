// mpcqp_engine.cu — sm_100a kernels and the C ABI (include/mpcqp_b200.h) of the batched MPC QP engine.
//
// Kernels:
//   mpc_assemble_kernel  device-side builder: planner inputs -> structured QP data (q, x0, obstacle-row
//                        gradients and lower bounds), restating mpcPlanner.cpp:952-966, 1040-1071, 1114-1139.
//   mpcqp_solve_kernel   persistent, one warp per QP; body in mpcqp_core.cuh.
// There is no host solve path in this library.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/mpcqp_b200.h"
#include "mpcqp_core.cuh"

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
    } else {                                        // obstacle rows                   (MP.cpp:1040-1071, 1114-1139)
      const long long u = t - nq - n0;
      const long long bk = u / a.R;                 // b*N + k
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

// Compacts the indices of the instances flagged `hard` into order[0 .. cnt[0]).
__global__ void mpc_order_kernel(int B, const int* hard, int* order, int* cnt) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    if (hard[b]) order[atomicAdd(&cnt[0], 1)] = b;
  }
}

// ------------------------------------------------------------------------------------------------
// solve kernel: persistent, one warp (= one CTA) per QP at a time, work fetched from a global counter
// ------------------------------------------------------------------------------------------------
template <int NST, int RT>
__global__ void __launch_bounds__(32) mpcqp_solve_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Settings st,
                                                         const __grid_constant__ Batch bt, int ws_stride, int* counter) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  Qp<NST, RT> qp(smem, sh, st, bt, bt.ws + (size_t)blockIdx.x * ws_stride, lane);
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= bt.B) break;
    qp.run(bt, b);
  }
}

// CTA kernel: persistent, one 4-warp CTA per QP at a time (mode 2 of mpcqp_core.cuh), two CTAs per SM.
template <int RT>
__global__ void __launch_bounds__(128, 2) mpcqp_solve_cta_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Settings st,
                                                                 const __grid_constant__ Batch bt, int ws_stride, int* counter) {
  extern __shared__ double smem[];
  __shared__ int s_next, s_flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Qp<30, RT, kModeCta> qp(smem, sh, st, bt, bt.ws + (size_t)blockIdx.x * ws_stride, lane);
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(counter + (bt.queue == 2 ? 3 : 0), 1);
    __syncthreads();
    const int idx = s_next;
    __syncthreads();
    int b = idx;
    if (bt.queue == 1) { if (idx >= *bt.nhard) break; b = bt.order[idx]; }
    else {
      if (idx >= bt.B) break;
      if (bt.queue == 2 && bt.hard[idx]) continue;       // solved by the other launch
    }
    qp.run_cta(bt, b, warp, &s_flag);
  }
}

typedef void (*SolveKernel)(const Shape, const Settings, const Batch, int, int*);
// Fast-path instantiations (compile-time dims, register-resident iterates); anything else runs the generic kernel.
#ifndef MPCQP_FAST_R_LIST
#define MPCQP_FAST_R_LIST X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#endif
// want: 2 = CTA kernel if available, 1 = warp-fast kernel if available, 0 = generic.  *mode returns what was picked.
static SolveKernel pick_kernel(int NS, int R, int want, int* mode) {
  if (NS == 30 && want == kModeCta) {
    *mode = kModeCta;
    switch (R) {
#define X(r) case r: return mpcqp_solve_cta_kernel<r>;
      MPCQP_FAST_R_LIST
#undef X
      default: break;
    }
  }
  if (NS == 30 && want >= kModeWarp) {
    *mode = kModeWarp;
    switch (R) {
#define X(r) case r: return mpcqp_solve_kernel<30, r>;
      MPCQP_FAST_R_LIST
#undef X
      default: break;
    }
  }
  *mode = kModeGeneric;
  return mpcqp_solve_kernel<0, 0>;
}

}  // namespace mpcqp

// ------------------------------------------------------------------------------------------------
// host side: engine object + C ABI
// ------------------------------------------------------------------------------------------------
using namespace mpcqp;

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
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

struct mpcqp_engine {
  int device = 0, num_sms = 0, max_smem_optin = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;          // main stream; side stream for the second solve launch
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs = nullptr;   // batch start / end, solve-kernel start
  cudaEvent_t evf = nullptr, evj = nullptr;                  // fork / join of the side stream
  std::string err;
  double last_ms = 0.0, last_solve_ms = 0.0; long long last_launches = 0; int last_fast = 0; int force_generic = 0;
  // structured-problem buffers (device)
  DevBuf pd, slack, q, x0s, g, low, ws, counter, hard, order;
  // staging for the *_host entry point
  DevBuf in_x0, in_xref, in_c, in_semi, in_yaw, in_lin, in_warm, out_x, out_y, out_i, out_d;
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
  *out = e;
  return MPCQP_OK;
}

extern "C" int mpcqp_engine_destroy(mpcqp_engine* e) {
  if (!e) return MPCQP_ERR_ARG;
  cudaSetDevice(e->device);
  DevBuf* bufs[] = { &e->pd, &e->slack, &e->q, &e->x0s, &e->g, &e->low, &e->ws, &e->counter, &e->hard, &e->order, &e->in_x0, &e->in_xref, &e->in_c,
                     &e->in_semi, &e->in_yaw, &e->in_lin, &e->in_warm, &e->out_x, &e->out_y, &e->out_i, &e->out_d };
  for (DevBuf* b : bufs) b->release();
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
  SolveKernel kern = pick_kernel(sh.NS, sh.R, want, &mode);
  const int threads = mode == kModeCta ? 128 : 32;
  const size_t smem = (size_t)smem_doubles(sh.NS, sh.R, mode) * sizeof(double);
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
  const int wsd = ws_doubles(sh.NS, sh.R, mode);
  CK(e->counter.need(4 * sizeof(int)));
  CK(cudaMemsetAsync(e->counter.p, 0, sizeof(int), e->stream));
  e->last_fast = mode;
  bt.queue = 0; bt.nhard = nullptr;
  if (mode != kModeCta) { bt.order = nullptr; bt.hard = nullptr; }
  if (mode == kModeCta) {
    // A CTA that has an SM to itself iterates ~1.5x faster than two sharing one.  Small batches therefore run one CTA
    // per SM (the launch asks for the whole shared memory of the SM).  Larger batches run as two concurrent launches:
    // the instances flagged hard (they run to max_iter and would otherwise form the tail of the batch) one per SM on
    // the main stream, everything else two per SM on the side stream, on whatever SMs the first launch leaves free.
    const size_t smem_solo = (size_t)e->max_smem_optin - 2048;   // more than half an SM: nothing else fits beside it
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solo));
    if (bt.B <= e->num_sms || !bt.order) {
      const bool solo = bt.B <= e->num_sms;
      CK(e->ws.need((size_t)grid * wsd * sizeof(double)));
      bt.ws = e->ws.as<double>();
      CK(cudaEventRecord(e->evs, e->stream));
      kern<<<(unsigned)grid, threads, solo ? smem_solo : smem, e->stream>>>(sh, st, bt, wsd, e->counter.as<int>());
      CK(cudaGetLastError());
      e->last_launches += 1;
      return MPCQP_OK;
    }
    const long long gh = e->num_sms < bt.B ? e->num_sms : bt.B;
    CK(e->ws.need((size_t)(grid + gh) * wsd * sizeof(double)));
    bt.nhard = e->counter.as<int>() + 1;
    CK(cudaEventRecord(e->evs, e->stream));
    CK(cudaEventRecord(e->evf, e->stream));
    CK(cudaStreamWaitEvent(e->stream2, e->evf, 0));
    Batch bh = bt; bh.queue = 1; bh.ws = e->ws.as<double>();
    kern<<<(unsigned)gh, threads, smem_solo, e->stream>>>(sh, st, bh, wsd, e->counter.as<int>());
    CK(cudaGetLastError());
    Batch bn = bt; bn.queue = 2; bn.ws = e->ws.as<double>() + (size_t)gh * wsd;
    kern<<<(unsigned)grid, threads, smem, e->stream2>>>(sh, st, bn, wsd, e->counter.as<int>());
    CK(cudaGetLastError());
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
  std::vector<unsigned char> slk((size_t)N * (R > 0 ? R : 1), 0);
  for (int i = 0; i < N * R; ++i) slk[i] = obs_dyn_host[i] ? 0 : 1;   // MP.cpp:1064-1069
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
  bt.pd = e->pd.as<double>(); bt.slack = e->slack.as<unsigned char>(); bt.q = a.q; bt.x0 = a.x0s; bt.g = a.g; bt.low = a.low;
  bt.warm_x = warm_x; bt.x = x; bt.y = y; bt.status = status; bt.iter = iter; bt.rho_updates = rho_updates;
  bt.obj = obj; bt.pri_res = pri_res; bt.dua_res = dua_res; bt.B = B; bt.order = e->order.as<int>(); bt.hard = e->hard.as<int>();
  rc = launch_solve(e, sh, st, bt); if (rc) return rc;
  CK(cudaEventRecord(e->ev1, e->stream));
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
