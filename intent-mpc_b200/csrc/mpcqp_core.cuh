// mpcqp_core.cuh — ADMM solvers for the stage-structured MPC QP of trajPlanner::mpcPlanner::solveTraj (reference:
// trajectory_planner/include/trajectory_planner/mpcPlanner.cpp:375-541; QP layout mpcPlanner.cpp:932-1146), reproducing the
// iterate sequence of the OSQP 0.6.2 solver the reference calls through OsqpEigen::Solver (constants:
// third_party/osqp/constants.h:59-118, step structure: third_party/osqp/auxil.h:21-154).
//
// This is NOT a port of OSQP/QDLDL.  What all kernel generations in this file share:
//   * lane = horizon stage.  Per-stage data (x_k, u_k, the 21+R constraint rows of stage k, the factor blocks of stage k) sit
//     in columns [slot][stage], so every per-row / per-variable step of ADMM is a conflict-free lane-parallel loop and the
//     neighbour-stage coupling (dynamics rows) is a read of column k±1 (or a warp shuffle).
//   * The iteration runs in UN-scaled coordinates: xh = D x, zh = E^-1 z, uh = E^-1 (y/rho).  Ruiz equilibration (scaling.h:
//     scale_data) then only enters through Rh_i = rho_i E_i^2 and sigma/D_j^2, the constraint matrix keeps its exact constants
//     (+-1, ts, ts^2/2, obstacle gradients) and is never stored.  Algebraically OSQP's iteration; only rounding
//     differs.
//   * The KKT solve of update_xz_tilde (auxil.h:67) is done on the reduced SPD system (c P + sigma D^-2 + A' Rh A) xt = rhs.
//     Slack states, accelerations and slack inputs are "leaf" variables eliminated in closed form, leaving a 6x6
//     block-tridiagonal system in (p_k, v_k).
// Three ways to run it (Qp<NST, RT, QMODE, ASSIST>, see `Mem` below):
//   * mode 2, the CTA kernels (horizon 30: the hot path; also built for horizons 20 and 25): one 4-warp CTA per QP, warps 0-2 own one axis each, warp 3 the slack
//     variables; iterates in registers from the first to the last iteration; the block-tridiagonal system is solved by block
//     parallel cyclic reduction (pcr_factor_cta, solve_role) whose per-level 6x6 matrices fill shared memory; optional three
//     assistant warps keep the matrices of the upper levels in registers (launches with one CTA per SM); a "wide" variant
//     takes a run-time obstacle count per instance.
//   * mode 1, the one-warp register kernel (horizon <= 32; first generation, kept for A/B tests): twisted two-ended block
//     LDL' chain solved by 12 lanes with warp shuffles.
//   * mode 0, the generic one-warp kernel (any horizon / obstacle count): the same chain with the factor, the right-hand side
//     and x in shared memory and the streaming per-row data in L2-resident global scratch.
//
// The same source compiles for the host when MPCQP_HOST_EMUL is defined: lane loops become plain loops over all stages
// (mode 0, and the PCR linear algebra of mode 2 around the generic iteration).  That build exists only for tests/ (logic
// checks without a GPU); the shipped library contains no host solve path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MPCQP_HOST_EMUL)
#define MQ_DEV 1
#define MQ_HD __device__ __forceinline__
#define MQ_HHD __host__ __device__ inline
#define MQ_NOINL __device__ __forceinline__
#else
#define MQ_DEV 0
#define MQ_HD inline
#define MQ_HHD inline
#define MQ_NOINL inline
#endif

namespace mpcqp {

// third_party/osqp/constants.h:59-118
constexpr double kRhoMin = 1e-06, kRhoMax = 1e06, kRhoEqOverIneq = 1e03, kRhoTol = 1e-04;
constexpr double kMinScaling = 1e-04, kMaxScaling = 1e+04, kInfty = 1e30;
constexpr double kOsqpNan = 2143289344.0;  // constants.h:95-97: (c_float)0x7fc00000UL is this NUMBER
enum Status : int {                         // constants.h:18-30
  kDualInfInacc = 4, kPrimInfInacc = 3, kSolvedInacc = 2, kSolved = 1, kMaxIter = -2,
  kPrimInf = -3, kDualInf = -4, kNonCvx = -7, kUnsolved = -10,
  kSuspended = -100                         // internal: parked for another launch to resume (never reported)
};

constexpr int NX = 8, NU = 5, NV = 13;      // mpcPlanner.h:42-43
constexpr int NBR = 21;                     // dynamics (8) + box (13) rows per stage; obstacle rows follow

struct Shape {              // batch-uniform problem shape, passed by value
  int NS;                   // stages = horizon (N+1)
  int R;                    // obstacle rows per stage (numObs)
  int n, m;
  double a_pv, b_pa, b_va;  // dynamics coefficients as the reference inserts them (float-rounded, MP.cpp:1003,1014)
  double blo[NV], bhi[NV];  // box bounds on (x_k, u_k), stage-uniform (MP.cpp:904-921)
  double obs_hi;            // upper bound of every obstacle row: IEEE +inf as the reference writes it (MP.cpp:1139), or the finite
                            // "infinity" of a caller of the CSC entry points (OSQP_INFTY = 1e30): OSQP's primal-infeasibility
                            // certificate multiplies this bound by a zero multiplier — NaN with inf, 0 with 1e30
};

struct Settings {           // third_party/osqp/types.h:139-176 (subset that affects the iterates)
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination, warm_start;
};

struct Batch {              // device pointers
  const double* pd;             // [NS*13] diagonal of P per stage variable (batch-uniform)
  const unsigned char* slack;   // [(NS-1)*R] 0: row uses slack input 3 (dynamic), 1: slack input 4 (static); per instance if slack_stride
  const double* q;              // [B][n]   linear cost, reference variable order
  const double* x0;             // [B][8]   stage-0 equality right-hand side is -x0
  const double* g;              // [B][NS-1][R][3] obstacle-row gradients
  const double* low;            // [B][NS-1][R]    obstacle-row lower bounds (upper = +inf)
  const double* warm_x;         // [B][n] or nullptr
  const double* warm_y;         // [B][m] or nullptr (dual warm start; the reference always passes zeros, MP.cpp:487)
  double* x;                    // [B][n]
  double* y;                    // [B][m] or nullptr
  int* status; int* iter; int* rho_updates;
  double* obj; double* pri_res; double* dua_res;
  double* ws;                   // per-warp scratch, ws_doubles(shape) each
  const int* order;             // [nhard] indices of the instances flagged hard, or nullptr
  const int* hard;              // [B] hard flags, or nullptr
  const int* nhard;             // number of hard instances (device scalar), or nullptr
  int queue;                    // 0: one queue over all B; 1: hard instances only; 2: the others only
  int slack_stride;             // 0: one slack pattern for the batch; (NS-1)*R: one per instance
  // migration of long-running instances (CTA kernels): an instance that is still running after `suspend_at` iterations
  // parks its state in slot s = (*susp_count)++ (if s < susp_cap) of susp_cold / susp_scal / susp_list and a later launch
  // (queue 4) resumes it, bit-identically, where it stopped
  int suspend_at, susp_cap, susp_stride;
  double* susp_cold; double* susp_scal; int* susp_list; int* susp_count;
  // order in which the parked instances are resumed: susp_key[slot] = primal residual when the instance was parked (a large
  // one after a few dozen iterations marks the instances that run long), susp_order = slots by decreasing key (or nullptr:
  // slot order)
  double* susp_key; const int* susp_order;
  const double* limits;         // [B][2] per-instance (max_vel, max_acc) replacing the batch-uniform box on v and a (CTA kernels), or nullptr
  const int* nobs;              // [B] obstacle rows per stage of each instance (wide CTA kernel only; R is then the stride of
                                // g / low / slack and m the stride of y), or nullptr
  long long* dbg;               // [B][8] per-phase clock64 totals (only with MPCQP_PHASE_TIMING), or nullptr
  int B;
};

// Where each per-stage array lives.  Element (slot j, stage k) of array A is A[j*NS + k].
//   mode 0, generic (runtime dims, any horizon): everything the iteration touches is in shared memory; one warp
//           per QP; twisted block LDL' chain.
//   mode 1, warp-fast (compile-time dims, horizon <= 32): one warp per QP; the iterates x, z, u and the right-hand
//           side live in REGISTERS during a burst of iterations and are parked in the per-warp global scratch
//           between bursts; shared memory holds the read-only scaled data, the twisted factor and the 6-vector
//           exchange buffer.
//   mode 2, CTA (horizons 17..30, lane = stage): one 4-warp CTA per QP; the block-tridiagonal solve is a parallel cyclic reduction
//           whose per-level 6x6 matrices fill shared memory (86 KB); iterates live in registers split by axis
//           across the warps; everything cold (factor workspace, parked iterates) is in the per-CTA global scratch.
constexpr int kModeGeneric = 0, kModeWarp = 1, kModeCta = 2;
constexpr int kPcrLevels = 5, kPcrLevelDoubles = 3 * 12 * 30 * 2, kPcrDoubles = kPcrLevels * kPcrLevelDoubles;
struct Mem {
  double *X, *Z, *U, *B, *TD, *MA;                       // iterates, rhs, exchange (cold in modes 1, 2)
  double *RH, *SD, *CQ, *G3, *LO, *W;                    // read-only per iteration + work vector
  double *SI, *GG, *DSI, *ESD, *DGI, *FS, *DAI, *CV, *PK; // factor (+ 72-double parking area)
  double *E, *D, *DY, *DX;                               // always in global scratch
  double *PCR, *WK, *RA, *RS, *YB;                          // mode 2, shared: PCR matrices, r ping-pong, slack part of r, y
  double *PD, *PL, *PI, *OG, *OX, *OU;                           // mode 2, global: PCR factor workspace (D, L x2, D^-1); parked obstacle part of the rhs
  double *ROW;                                                   // mode 2 wide, shared: obstacle rows' z, u, Rh, low, gradient ([7 R] slots)
};
MQ_HHD int hot_slots(int R, int mode) {
  return (NBR + R) + NV + NV + 3 * R + R + (mode == kModeWarp ? 6 : NV) + 36 + 36 + 2 + 2 + 2 + 6 + 3 + 12;
}
MQ_HHD int iter_slots(int R) { return NV + 2 * (NBR + R) + NV + 8 + 3; }
// mode 2: the "cold block" (scalings, rho vector, scaled cost, obstacle rows, parked iterates) is one contiguous run of
// slots so that the setup phase can build it in shared memory (in the not-yet-used PCR region) and copy it out once.
MQ_HHD int cold_slots(int R) { return 2 * (NBR + R) + 2 * NV + (NBR + R) + NV + NV + 3 * R + R + iter_slots(R); }
constexpr int kWorkSlots = 108;             // PCR factor workspace: D / D^-1, L, next L (36 slots each); r / y exchange buffers alias it
// generic mode: slots that live in global scratch instead of shared memory (see map_memory)
MQ_HHD int generic_global_slots(int R) { return (NBR + R) + NV + NV + 3 * R + R + 2 * (NBR + R); }
MQ_HHD int smem_doubles(int NS, int R, int mode, bool wide = false) {
  if (mode == kModeCta) return kPcrDoubles + (kWorkSlots + (wide ? 7 * R : 0)) * NS;
  if (mode == kModeWarp) return hot_slots(R, mode) * NS + 72;
  return (hot_slots(R, mode) + iter_slots(R) - generic_global_slots(R)) * NS + 72;
}
// threads of a CTA-mode block: four solver warps; one-per-SM ("assist") blocks add three PCR assistants and, from four obstacle
// rows per stage on (compile-time R, not the wide kernel), the row helper (Qp::kCtaThreads is the same number)
MQ_HHD int cta_threads(int R, bool assist, bool wide) { return !assist ? 128 : ((!wide && R >= 4) ? 256 : 224); }
MQ_HHD int ws_doubles(int NS, int R, int mode) {
  const int base = 2 * (NBR + R) + 2 * NV;
  if (mode == kModeCta) return (cold_slots(R) + NV + 36 + 36 + 27 + 36 + 72 + 36 + 3 + NV + (NBR + R)) * NS;
  return (base + (mode == kModeWarp ? iter_slots(R) : generic_global_slots(R))) * NS;
}
// cold block at `g` (same order wherever it lives)
MQ_HHD double* map_cold(Mem& m, double* g, int NS, int R) {
  const int MK = NBR + R;
  m.E = g; g += MK * NS; m.D = g; g += NV * NS; m.DY = g; g += MK * NS; m.DX = g; g += NV * NS;
  m.RH = g; g += MK * NS; m.SD = g; g += NV * NS; m.CQ = g; g += NV * NS; m.G3 = g; g += 3 * R * NS; m.LO = g; g += R * NS;
  m.X = g; g += NV * NS; m.Z = g; g += MK * NS; m.U = g; g += MK * NS; m.B = g; g += NV * NS; m.TD = g; g += 8 * NS; m.MA = g; g += 3 * NS;
  return g;
}
MQ_HHD void map_memory(Mem& m, double* sm, double* ws, int NS, int R, int mode) {
  const int MK = NBR + R;
  double* p = sm;
  double* g = ws;
  if (mode == kModeCta) {
    m.PCR = p; p += kPcrDoubles;
    m.WK = p; m.RA = p; m.RS = p + 2 * 6 * NS; m.YB = p + 2 * 6 * NS + 4 * NS; m.ROW = p + kWorkSlots * NS;
    g = map_cold(m, g, NS, R);
    // T blocks (SI) and couplings (GG, 12 slots used) are staged in the factor workspace: LN slots / tail of the LC slots
    m.W = g; g += NV * NS; m.SI = m.WK + 72 * NS; m.GG = m.WK + 60 * NS; m.DSI = g; g += 2 * NS; m.ESD = g; g += 2 * NS;
    m.DGI = g; g += 2 * NS; m.FS = g; g += 6 * NS; m.DAI = g; g += 3 * NS; m.CV = g; g += 12 * NS; m.PK = nullptr;
    m.PD = g; g += 36 * NS; m.PL = g; g += 72 * NS; m.PI = g; g += 36 * NS; m.OG = g; g += 3 * NS; m.OX = g; g += NV * NS; m.OU = g; g += MK * NS;
    return;
  }
  m.E = g; g += MK * NS; m.D = g; g += NV * NS; m.DY = g; g += MK * NS; m.DX = g; g += NV * NS;
  const bool fast = mode == kModeWarp;
  m.PCR = m.WK = m.RA = m.RS = m.YB = m.PD = m.PL = m.PI = m.OG = m.OX = m.OU = m.ROW = nullptr;
  // generic mode: the per-row data that an iteration only streams through once (rho vector, sigma / D^2, scaled cost, obstacle
  // gradients and bounds, the constraint iterates z and u) live in the warp's L2-resident global scratch; shared memory keeps
  // what the serial chain and the neighbour-stage reads touch (factor, right-hand side, x): 2.4x less shared memory per QP, so
  // that three warps instead of one share an SM at horizon 60
  double*& r_ = fast ? p : g;
  m.RH = r_; r_ += MK * NS; m.SD = r_; r_ += NV * NS; m.CQ = r_; r_ += NV * NS; m.G3 = r_; r_ += 3 * R * NS; m.LO = r_; r_ += R * NS;
  m.W = p; p += (fast ? 6 : NV) * NS; m.SI = p; p += 36 * NS; m.GG = p; p += 36 * NS; m.DSI = p; p += 2 * NS; m.ESD = p; p += 2 * NS;
  m.DGI = p; p += 2 * NS; m.FS = p; p += 6 * NS; m.DAI = p; p += 3 * NS; m.CV = p; p += 12 * NS; m.PK = p; p += 72;
  if (fast) {
    m.X = g; g += NV * NS; m.Z = g; g += MK * NS; m.U = g; g += MK * NS; m.B = g; g += NV * NS; m.TD = g; g += 8 * NS; m.MA = g; g += 3 * NS;
  } else {
    m.X = p; p += NV * NS; m.Z = g; g += MK * NS; m.U = g; g += MK * NS; m.B = p; p += NV * NS; m.TD = p; p += 8 * NS; m.MA = p; p += 3 * NS;
  }
}

// Dimensions: compile-time in fast mode, runtime otherwise.
template <int NST, int RT> struct DimsT {
  static constexpr int NS = NST, N = NST - 1, R = RT, MK = NBR + RT;
  MQ_HHD DimsT(int, int) {}
};
// CTA mode with a run-time obstacle count (9 .. kWideMax rows per stage): the rows' state lives in shared memory
constexpr int kWideR = -1, kWideMax = 32;
template <int NST> struct DimsT<NST, kWideR> {
  static constexpr int NS = NST, N = NST - 1;
  int R, MK;
  MQ_HHD DimsT(int, int r) : R(r), MK(NBR + r) {}
};
template <> struct DimsT<0, 0> {
  int NS, N, R, MK;
  MQ_HHD DimsT(int ns, int r) : NS(ns), N(ns - 1), R(r), MK(NBR + r) {}
};

#if MQ_DEV
#define MQ_FOR_STAGES(k) for (int k = lane; k < NS; k += 32)
#define MQ_SYNC() __syncwarp()
// Barrier of the four solver warps of a CTA-mode block (threads 0..127).  The block may carry three more warps (PCR
// assistants, threads 128..223) and a row helper (threads 224..255) that never take part in it.
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0, 128;" ::: "memory"); }
#else
#define MQ_FOR_STAGES(k) for (int k = 0; k < NS; ++k)
#define MQ_SYNC() ((void)0)
#endif

#if MQ_DEV
// One bulk asynchronous copy (cp.async.bulk: the TMA engine moves the bytes, SASS UBLKCP) of `bytes` (a multiple of 16, both
// addresses 16-byte aligned) from shared to global memory, issued and awaited by the calling thread.  The block's writes to `src`
// must be ordered before the call by a barrier; the copy has landed when the call returns.
__device__ __forceinline__ void bulk_store_shared_to_global(double* dst, const double* src, unsigned bytes) {
  if (bytes & 8u) { bytes -= 8u; dst[bytes / 8] = src[bytes / 8]; }       // an odd number of doubles (odd horizon x odd slot count): the last one by hand
  const unsigned s = (unsigned)__cvta_generic_to_shared(src);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes to src -> visible to the async proxy
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async.global;" ::: "memory");               // the block reads dst with ordinary loads afterwards
}
#endif

MQ_HD double limit_scaling(double v) { v = v < kMinScaling ? 1.0 : v; return v > kMaxScaling ? kMaxScaling : v; }
// 1/sqrt(v) of scaling.h's scale_data (sqrt then reciprocal, as OSQP computes it)
#if MQ_DEV
MQ_HD double rsqrt_scaling(double v) { return rsqrt(v); }      // <= 1 ulp from 1/sqrt: below every tolerance that is compared
#else
MQ_HD double rsqrt_scaling(double v) { return 1.0 / sqrt(v); }
#endif

// In-place inverse of a symmetric positive definite 6x6 (Gauss-Jordan, no pivoting).
MQ_HD void inv6(double* a) {
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    double piv = 1.0 / a[p * 6 + p];
#pragma unroll
    for (int j = 0; j < 6; ++j) if (j != p) a[p * 6 + j] *= piv;
#pragma unroll
    for (int i = 0; i < 6; ++i) if (i != p) {
      double f = a[i * 6 + p];
#pragma unroll
      for (int j = 0; j < 6; ++j) if (j != p) a[i * 6 + j] -= f * a[p * 6 + j];
      a[i * 6 + p] = -f * piv;
    }
    a[p * 6 + p] = piv;
  }
}

template <int NST, int RT, int QMODE = (NST > 0 ? kModeWarp : kModeGeneric), bool ASSIST = false> struct Qp : DimsT<NST, RT> {
  using Dm = DimsT<NST, RT>;
  using Dm::NS; using Dm::N; using Dm::R; using Dm::MK;
  static constexpr bool kFast = QMODE == kModeWarp;
  static constexpr bool kCta = QMODE == kModeCta;
  static constexpr bool kWide = RT == kWideR;            // run-time R, obstacle rows in shared memory (CTA mode only)
  static constexpr int RC = kWide ? 0 : RT;              // compile-time R of the register-row code (none in wide mode)
  // One-per-SM blocks with at least four obstacle rows carry an eighth warp, the row helper: the rows that would belong to the
  // slack warp (o = 3, 7) live in shared memory (kHelpSlots values per row and stage) and the helper runs them in every iteration,
  // so that the slack warp — four variables, the longest instruction stream between y and the next right-hand side — is not
  // what the axis warps wait for.  Same arithmetic as the register rows (obstacle_rows): results do not depend on the variant.
  static constexpr bool kHelp = ASSIST && !kWide && RC >= 4;
  // up to kHelpAllMax rows per stage the helper runs ALL obstacle rows (their four chains interleave in one warp and are done
  // before the axis warps need them), which also frees the axis warps of a row's registers and instructions
  static constexpr int kHelpAllMax = 0;     // (measured at four rows: 1.67 us per iteration against 1.61 with the slack warp's row only)
  static constexpr bool kHelpAll = kHelp && RC <= kHelpAllMax;
  static constexpr int kHelpSlots = 12;
  static constexpr int kCtaThreads = kHelp ? 256 : (ASSIST ? 224 : 128);
  static_assert(!kWide || (QMODE == kModeCta && ASSIST), "wide mode is a one-per-SM CTA kernel");
  static_assert(QMODE == kModeGeneric || NST > 0, "modes 1 and 2 need compile-time dims");
  static_assert(QMODE != kModeCta || (NST >= 17 && NST <= 30), "the CTA path is built for horizons 17..30 (lane = stage, 5 PCR levels, matrices sized for 30 stages)");
  static constexpr int kWS = 6;    // fast mode: W holds one 6-vector per stage, stage-major (16-byte aligned rows)
  // Fast mode stores the per-stage chain data (G, W) in CHAIN order so that both chains of the twisted
  // recursion walk upward in memory: stages 0..mid-1 -> slots 0..mid-1, mid -> slot mid, stages N..mid+1 ->
  // slots mid+1..N.
  MQ_HD int cslot(int k) const { return k <= NS / 2 ? k : (NS / 2 + 1) + (N - k); }
  Mem m; const Shape& sh; const Settings& st; int lane;
  double lim_v = 0.0, lim_a = 0.0; bool has_lim = false;   // per-instance limits (CTA kernels, Batch::limits)
  // suspend / resume (Batch::suspend_at)
  bool allow_suspend_ = false, resume_ = false, resume_soft_ = false, susp_refactor_ = false;
  int susp_K_ = 0, susp_cap_ = 0, susp_slot_ = -1, resume_iter_ = 0, resume_rho_updates_ = 0;
  int* susp_count_ = nullptr;
  int Rs;                                        // stride of the obstacle-row inputs (= R unless instances carry their own count)
  double *smem0, *ws0;
  const double* pd; const unsigned char* slack; const double* x0p;
  double c, cinv, rho, nq, nq_s;                 // cost scaling, current rho, |q|_inf norms (unscaled / scaled)
  double pri_res, dua_res, obj, nAx, nZ, nPx, nAty, pri_s, dua_s, nAx_s, nZ_s, nPx_s, nAty_s;
  int status, info_iter, rho_updates;
#ifdef MPCQP_PHASE_TIMING
  long long dbg_ruiz = 0;
#endif

#define X_(j, k) m.X[(j) * NS + (k)]
#define Z_(i, k) m.Z[(i) * NS + (k)]
#define U_(i, k) m.U[(i) * NS + (k)]
#define RH_(i, k) m.RH[(i) * NS + (k)]
#define SD_(j, k) m.SD[(j) * NS + (k)]
#define CQ_(j, k) m.CQ[(j) * NS + (k)]
#define G3_(i, k) m.G3[(i) * NS + (k)]
#define LO_(o, k) m.LO[(o) * NS + (k)]
#define B_(j, k) m.B[(j) * NS + (k)]
#define W_(j, k) m.W[kFast ? cslot(k) * kWS + (j) : (j) * NS + (k)]
#define TD_(r, k) m.TD[(r) * NS + (k)]
#define MA_(c, k) m.MA[(c) * NS + (k)]
#define SI_(e, k) m.SI[(e) * NS + (k)]
#define GG_(e, k) m.GG[kFast ? cslot(k) * 36 + (e) : (e) * NS + (k)]
#define DSI_(t, k) m.DSI[(t) * NS + (k)]
#define ESD_(t, k) m.ESD[(t) * NS + (k)]
#define DGI_(t, k) m.DGI[(t) * NS + (k)]
#define FS_(e, k) m.FS[(e) * NS + (k)]
#define DAI_(c, k) m.DAI[(c) * NS + (k)]
#define CV_(e, k) m.CV[(e) * NS + (k)]
#define PK_(chain, e) m.PK[(chain) * 36 + (e)]
#define WSE_(i, k) m.E[(i) * NS + (k)]
#define WSD_(j, k) m.D[(j) * NS + (k)]
#define WSDY_(i, k) m.DY[(i) * NS + (k)]
#define WSDX_(j, k) m.DX[(j) * NS + (k)]
#define SLK_(o, k) ((int)slack[(k) * Rs + (o)])
#define PD_(e, k) m.PD[(e) * NS + (k)]
#define PL_(buf, e, k) m.PL[((buf) * 36 + (e)) * NS + (k)]
#define PI_(e, k) m.PI[(e) * NS + (k)]
#define OG_(c, k) m.OG[(c) * NS + (k)]
#define OX_(j, k) m.OX[(j) * NS + (k)]
#define OU_(i, k) m.OU[(i) * NS + (k)]

  MQ_HD int nrows(int k) const { return k < N ? MK : 16; }
  // position of row i of stage k in the reference's constraint ordering (MP.cpp:989-1071)
  MQ_HD int row_index(int k, int i) const {
    return i < 8 ? 8 * k + i : (i < 16 ? 8 * NS + 8 * k + (i - 8) : (i < NBR ? 16 * NS + 5 * k + (i - 16) : 16 * NS + 5 * N + k * R + (i - NBR)));
  }
  MQ_HD int nvars(int k) const { return k < N ? NV : NX; }

  // ---- warp reductions -------------------------------------------------------------------
  MQ_HD double wmax(double v) const {
#if MQ_DEV
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
    return v;
  }
  MQ_HD double wsum(double v) const {
#if MQ_DEV
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
  MQ_HD bool wany(bool p) const {
#if MQ_DEV
    return __any_sync(0xffffffffu, p);
#else
    return p;
#endif
  }

  // box on variable j of a stage (MP.cpp:904-921); CTA kernels may carry per-instance velocity / acceleration limits
  MQ_HD double box_lo(int j) const {
    if constexpr (kCta) { if (has_lim) { if (j >= 3 && j < 6) return -lim_v; if (j >= 8 && j < 11) return -lim_a; } }
    return sh.blo[j];
  }
  MQ_HD double box_hi(int j) const {
    if constexpr (kCta) { if (has_lim) { if (j >= 3 && j < 6) return lim_v; if (j >= 8 && j < 11) return lim_a; } }
    return sh.bhi[j];
  }
  // ---- un-scaled bounds of row i of stage k (MP.cpp:1074-1146) -----------------------------
  MQ_HD void row_bounds(int k, int i, double& lo, double& hi) const {
    if (i < 8) { double v = (k == 0) ? -x0p[i] : 0.0; lo = v; hi = v; }
    else if (i < NBR) { lo = box_lo(i - 8); hi = box_hi(i - 8); }
    else { lo = LO_(i - NBR, k); hi = sh.obs_hi; }
  }
  // rho class on SCALED bounds (auxil.h: set_rho_vec): -1 loose, 1 equality, 0 inequality
  MQ_HD int row_type(double e, double lo, double hi) const {
    double ls = e * lo, us = e * hi;
    if (ls < -kInfty * kMinScaling && us > kInfty * kMinScaling) return -1;
    if (us - ls < kRhoTol) return 1;
    return 0;
  }
  MQ_HD double rho_of_type(int t) const { return t < 0 ? kRhoMin : (t > 0 ? kRhoEqOverIneq * rho : rho); }

  // (A v)_i for row i of stage k; v(j,k) accessor.  Row content: MP.cpp:989-1071.
  template <class F> MQ_HD double row_ax(int k, int i, F v) const {
    if (i < 8) {
      double s = -v(i, k);
      if (k > 0) {
        if (i < 3) s += v(i, k - 1) + sh.a_pv * v(3 + i, k - 1) + sh.b_pa * v(8 + i, k - 1);
        else if (i < 6) s += v(i, k - 1) + sh.b_va * v(5 + i, k - 1);
        else s += v(5 + i, k - 1);
      }
      return s;
    } else if (i < NBR) return v(i - 8, k);
    int o = i - NBR;
    return G3_(3 * o, k) * v(0, k) + G3_(3 * o + 1, k) * v(1, k) + G3_(3 * o + 2, k) * v(2, k) - v(11 + SLK_(o, k), k);
  }
  // (A' t)_j for variable j of stage k; t(i,k) accessor over rows (must be readable for stage k+1).
  template <class F> MQ_HD double col_aty(int k, int j, F t) const {
    double s = t(8 + j, k);
    if (j < 8) s -= t(j, k);
    if (k < N) {
      if (j < 3) { s += t(j, k + 1); for (int o = 0; o < R; ++o) s += G3_(3 * o + j, k) * t(NBR + o, k); }
      else if (j < 6) s += sh.a_pv * t(j - 3, k + 1) + t(j, k + 1);
      else if (j < 8) {}
      else if (j < 11) s += sh.b_pa * t(j - 8, k + 1) + sh.b_va * t(j - 5, k + 1);
      else { s += t(j - 5, k + 1); for (int o = 0; o < R; ++o) if (SLK_(o, k) == j - 11) s -= t(NBR + o, k); }
    }
    return s;
  }

  // ---- setup: load, Ruiz equilibration (scaling.h: scale_data), rho vector, warm start --------
  // `wi` of `nw` cooperating warps (mode 2: the four warps of the CTA; otherwise one): the per-variable / per-row loops
  // of every stage are dealt round-robin to the warps, lane = stage in each of them.
  MQ_HD void setup_sync(int nw) const {
#if MQ_DEV
    if (nw > 1) cta_sync(); else __syncwarp();
#else
    (void)nw;
#endif
  }
  // sum and max over all cooperating warps; every thread gets the same values
  MQ_HD void setup_reduce(double& sm, double& mx, int wi, int nw) const {
    sm = wsum(sm); mx = wmax(mx);
#if MQ_DEV
    if (nw > 1) {
      // exchange buffer outside the (aliased) PCR region; the wide cold block spills into the workspace: use its tail
      double* red = kWide ? m.ROW + 7 * R * NS - 8 : m.YB;
      if (lane == 0) { red[2 * wi] = sm; red[2 * wi + 1] = mx; }
      cta_sync();
      double s2 = 0.0, m2 = 0.0;
      for (int w = 0; w < nw; ++w) { s2 += red[2 * w]; m2 = fmax(m2, red[2 * w + 1]); }
      cta_sync();
      sm = s2; mx = m2;
    }
#else
    (void)wi; (void)nw;
#endif
  }
  // NW = number of cooperating warps (compile time, so that the dealt-out loops unroll and their sqrt / divide chains
  // interleave); wi = this warp's index.
  template <int NW = 1> MQ_NOINL void load_and_scale(const Batch& bt, int b, const int wi = 0) {
    constexpr int nw = NW;
    const double* q = bt.q + (size_t)b * sh.n;
    const double* gp = bt.g + (size_t)b * N * Rs * 3;
    const double* lp = bt.low + (size_t)b * N * Rs;
    MQ_FOR_STAGES(k) {
      for (int j = wi; j < NV; j += nw) {
        bool ex = j < nvars(k);
        CQ_(j, k) = ex ? (j < 8 ? q[8 * k + j] : q[8 * NS + 5 * k + (j - 8)]) : 0.0;
        SD_(j, k) = 1.0;  // D during Ruiz
        B_(j, k) = 0.0; X_(j, k) = 0.0;
      }
      for (int i = wi; i < MK; i += nw) { RH_(i, k) = 1.0; Z_(i, k) = 0.0; U_(i, k) = 0.0; }  // RH holds E during Ruiz
      for (int o = wi; o < R; o += nw) {
        bool ex = k < N;
        G3_(3 * o, k) = ex ? gp[(k * Rs + o) * 3] : 0.0;
        G3_(3 * o + 1, k) = ex ? gp[(k * Rs + o) * 3 + 1] : 0.0;
        G3_(3 * o + 2, k) = ex ? gp[(k * Rs + o) * 3 + 2] : 0.0;
        LO_(o, k) = ex ? lp[k * Rs + o] : 0.0;
      }
      if (wi == 0) {
        for (int r = 0; r < 8; ++r) TD_(r, k) = 0.0;
        for (int cc = 0; cc < 3; ++cc) MA_(cc, k) = 0.0;
      }
    }
    c = 1.0;
    setup_sync(nw);
#ifdef MPCQP_PHASE_TIMING
    long long ts0 = clock64();
#endif
    const double apv = fabs(sh.a_pv), bpa = fabs(sh.b_pa), bva = fabs(sh.b_va);
    for (int pass = 0; pass < st.scaling; ++pass) {
      // column norms of [P A'; A 0] -> B (Dt), row norms of A -> Z (Et); D lives in SD, E in RH
      MQ_FOR_STAGES(k) {
        const int nv = nvars(k), nr = nrows(k);
#pragma unroll
        for (int jj = 0; jj < (NW > 1 ? (NV + NW - 1) / NW : NV); ++jj) {
          const int j = wi + nw * jj;
          if (j >= nv) continue;
          double dj = SD_(j, k);
          double an = RH_(8 + j, k);
          if (j < 8) an = fmax(an, RH_(j, k));
          if (k < N) {
            if (j < 3) { an = fmax(an, RH_(j, k + 1)); for (int o = 0; o < R; ++o) an = fmax(an, RH_(NBR + o, k) * fabs(G3_(3 * o + j, k))); }
            else if (j < 6) an = fmax(an, fmax(RH_(j - 3, k + 1) * apv, RH_(j, k + 1)));
            else if (j < 8) {}
            else if (j < 11) an = fmax(an, fmax(RH_(j - 8, k + 1) * bpa, RH_(j - 5, k + 1) * bva));
            else { an = fmax(an, RH_(j - 5, k + 1)); for (int o = 0; o < R; ++o) if (SLK_(o, k) == j - 11) an = fmax(an, RH_(NBR + o, k)); }
          }
          double pn = fabs(c * pd[k * NV + j]) * dj * dj;
          B_(j, k) = rsqrt_scaling(limit_scaling(fmax(pn, an * dj)));
        }
#pragma unroll 4
        for (int i = wi; i < nr; i += nw) {
          double rn;
          if (i < 8) {
            rn = SD_(i, k);
            if (k > 0) {
              if (i < 3) rn = fmax(rn, fmax(SD_(i, k - 1), fmax(apv * SD_(3 + i, k - 1), bpa * SD_(8 + i, k - 1))));
              else if (i < 6) rn = fmax(rn, fmax(SD_(i, k - 1), bva * SD_(5 + i, k - 1)));
              else rn = fmax(rn, SD_(5 + i, k - 1));
            }
          } else if (i < NBR) rn = SD_(i - 8, k);
          else {
            int o = i - NBR;
            rn = SD_(11 + SLK_(o, k), k);
            for (int cc = 0; cc < 3; ++cc) rn = fmax(rn, fabs(G3_(3 * o + cc, k)) * SD_(cc, k));
          }
          Z_(i, k) = rsqrt_scaling(limit_scaling(rn * RH_(i, k)));
        }
      }
      setup_sync(nw);
      double psum = 0.0, qmax = 0.0;
      MQ_FOR_STAGES(k) {
        const int nv = nvars(k), nr = nrows(k);
        for (int j = wi; j < nv; j += nw) {
          double dj = SD_(j, k) * B_(j, k);
          SD_(j, k) = dj;
          psum += fabs(c * pd[k * NV + j]) * dj * dj;
          qmax = fmax(qmax, fabs(c * CQ_(j, k) * dj));
        }
        for (int i = wi; i < nr; i += nw) RH_(i, k) *= Z_(i, k);
      }
      setup_reduce(psum, qmax, wi, nw);
      double ct = psum / (double)sh.n;
      double nqv = limit_scaling(qmax);
      if (nqv > ct) ct = nqv;
      ct = limit_scaling(ct);
      c *= 1.0 / ct;
      setup_sync(nw);
    }
#ifdef MPCQP_PHASE_TIMING
    dbg_ruiz = clock64() - ts0;
#endif
    cinv = 1.0 / c;
    // finalise: stash E, D (needed for rho estimates / rho updates), form Rh, sigma/D^2, c q
    rho = fmin(fmax(st.rho, kRhoMin), kRhoMax);
    double nq0 = 0.0, nq1 = 0.0;
    MQ_FOR_STAGES(k) {
      for (int j = wi; j < NV; j += nw) {
        double dj = SD_(j, k);
        WSD_(j, k) = dj;
        double cq = c * CQ_(j, k);
        nq0 = fmax(nq0, fabs(CQ_(j, k)));
        nq1 = fmax(nq1, fabs(dj * cq));
        CQ_(j, k) = cq;
        SD_(j, k) = st.sigma / (dj * dj);
        B_(j, k) = 0.0; WSDX_(j, k) = 0.0;
      }
      const int nr = nrows(k);
      for (int i = wi; i < MK; i += nw) {
        double e = RH_(i, k);
        WSE_(i, k) = e; WSDY_(i, k) = 0.0; Z_(i, k) = 0.0;
        if (i < nr) { double lo, hi; row_bounds(k, i, lo, hi); RH_(i, k) = rho_of_type(row_type(e, lo, hi)) * e * e; }
        else RH_(i, k) = 0.0;
      }
    }
    {
      double dummy = 0.0;
      setup_reduce(dummy, nq0, wi, nw); nq = nq0;
      dummy = 0.0;
      setup_reduce(dummy, nq1, wi, nw); nq_s = nq1;
    }
    // warm start (osqp.h:157): x given, y = 0 (MP.cpp:487), z = A x
    if (bt.warm_x && st.warm_start) {
      const double* wx = bt.warm_x + (size_t)b * sh.n;
      MQ_FOR_STAGES(k) { const int nv = nvars(k); for (int j = wi; j < nv; j += nw) X_(j, k) = j < 8 ? wx[8 * k + j] : wx[8 * NS + 5 * k + (j - 8)]; }
    }
    setup_sync(nw);
    MQ_FOR_STAGES(k) { const int nr = nrows(k); for (int i = wi; i < nr; i += nw) Z_(i, k) = row_ax(k, i, [&](int j, int kk) { return X_(j, kk); }); }
    if (bt.warm_y && st.warm_start) {   // y_s = c E^-1 y  =>  u = E^-1 y_s / rho = c y / Rh
      const double* wy = bt.warm_y + (size_t)b * sh.m;
      MQ_FOR_STAGES(k) {
        const int nr = nrows(k);
        for (int i = wi; i < nr; i += nw) U_(i, k) = c * wy[row_index(k, i)] / RH_(i, k);
      }
    }
    setup_sync(nw);
  }

  // ---- factorisation of the reduced KKT matrix ---------------------------------------------
  // Build leaf factors + T blocks per stage (parallel), then the twisted block recursion.
  MQ_NOINL void factor() {
    const double apv = sh.a_pv, bpa = sh.b_pa, bva = sh.b_va;
    // pass 1: leaves, own contributions to Tkk (-> SI) and Tnk (-> GG[0..11])
    MQ_FOR_STAGES(k) {
      double hd[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        double h = c * pd[k * NV + j] + SD_(j, k) + RH_(8 + j, k);
        if (j < 8) h += RH_(j, k);
        hd[j] = (j < 8 || k < N) ? h : 1.0;
      }
      double T[36];
#pragma unroll
      for (int e = 0; e < 36; ++e) T[e] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) T[i * 6 + i] = hd[i];
      // own slack states
      DSI_(0, k) = 1.0 / hd[6]; DSI_(1, k) = 1.0 / hd[7];
      double es0 = k > 0 ? -RH_(6, k) : 0.0, es1 = k > 0 ? -RH_(7, k) : 0.0;
      ESD_(0, k) = es0 / hd[6]; ESD_(1, k) = es1 / hd[7];
      if (k < N) {
        double rn[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) rn[r] = RH_(r, k + 1);
        // slack inputs: eliminate s_{k+1,t} first, then sigma_{k,t}
        double dsg[2], fs[6];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          double dsn = c * pd[(k + 1) * NV + 6 + t] + SD_(6 + t, k + 1) + RH_(14 + t, k + 1) + rn[6 + t];
          dsg[t] = hd[11 + t] + rn[6 + t] - rn[6 + t] * rn[6 + t] / dsn;
          fs[3 * t] = fs[3 * t + 1] = fs[3 * t + 2] = 0.0;
        }
        for (int o = 0; o < R; ++o) {
          double ro = RH_(NBR + o, k), g0 = G3_(3 * o, k), g1 = G3_(3 * o + 1, k), g2 = G3_(3 * o + 2, k);
          int t = SLK_(o, k);
          if (t == 0) { dsg[0] += ro; fs[0] -= ro * g0; fs[1] -= ro * g1; fs[2] -= ro * g2; }
          else { dsg[1] += ro; fs[3] -= ro * g0; fs[4] -= ro * g1; fs[5] -= ro * g2; }
          T[0] += ro * g0 * g0; T[1] += ro * g0 * g1; T[2] += ro * g0 * g2;
          T[7] += ro * g1 * g1; T[8] += ro * g1 * g2; T[14] += ro * g2 * g2;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          double di = 1.0 / dsg[t];
          DGI_(t, k) = di;
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) FS_(3 * t + cc, k) = fs[3 * t + cc];
          T[0] -= fs[3 * t] * fs[3 * t] * di; T[1] -= fs[3 * t] * fs[3 * t + 1] * di; T[2] -= fs[3 * t] * fs[3 * t + 2] * di;
          T[7] -= fs[3 * t + 1] * fs[3 * t + 1] * di; T[8] -= fs[3 * t + 1] * fs[3 * t + 2] * di; T[14] -= fs[3 * t + 2] * fs[3 * t + 2] * di;
        }
        T[6] = T[1]; T[12] = T[2]; T[13] = T[8];
        // accelerations + dynamics rows of stage k+1
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double r1 = rn[cc], r2 = rn[3 + cc];
          double da = hd[8 + cc] + r1 * bpa * bpa + r2 * bva * bva;
          double dai = 1.0 / da;
          double cv0 = r1 * bpa, cv1 = r1 * bpa * apv + r2 * bva, cv2 = -r1 * bpa, cv3 = -r2 * bva;
          DAI_(cc, k) = dai;
          CV_(4 * cc, k) = cv0; CV_(4 * cc + 1, k) = cv1; CV_(4 * cc + 2, k) = cv2; CV_(4 * cc + 3, k) = cv3;
          int ip = cc, iv = 3 + cc;
          T[ip * 6 + ip] += r1 - cv0 * cv0 * dai;
          double off = r1 * apv - cv0 * cv1 * dai;
          T[ip * 6 + iv] += off; T[iv * 6 + ip] += off;
          T[iv * 6 + iv] += r1 * apv * apv + r2 - cv1 * cv1 * dai;
          // coupling block (rows stage k+1, cols stage k), axis cc: [pp, pv, vp, vv]
          GG_(4 * cc, k) = -r1 - cv2 * cv0 * dai;
          GG_(4 * cc + 1, k) = -r1 * apv - cv2 * cv1 * dai;
          GG_(4 * cc + 2, k) = -cv3 * cv0 * dai;
          GG_(4 * cc + 3, k) = -r2 - cv3 * cv1 * dai;
        }
      } else {
        DGI_(0, k) = DGI_(1, k) = 1.0;
#pragma unroll
        for (int e = 0; e < 6; ++e) FS_(e, k) = 0.0;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { DAI_(cc, k) = 1.0; CV_(4 * cc, k) = CV_(4 * cc + 1, k) = CV_(4 * cc + 2, k) = CV_(4 * cc + 3, k) = 0.0; }
      }
#pragma unroll
      for (int e = 0; e < 36; ++e) SI_(e, k) = T[e];
    }
    MQ_SYNC();
    // pass 2: pull the acceleration-leaf Schur terms of stage k-1 into Tkk[k]
    MQ_FOR_STAGES(k) {
      if (k > 0) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double dai = DAI_(cc, k - 1), cv2 = CV_(4 * cc + 2, k - 1), cv3 = CV_(4 * cc + 3, k - 1);
          int ip = cc, iv = 3 + cc;
          SI_(ip * 6 + ip, k) -= cv2 * cv2 * dai;
          SI_(ip * 6 + iv, k) -= cv2 * cv3 * dai;
          SI_(iv * 6 + ip, k) -= cv2 * cv3 * dai;
          SI_(iv * 6 + iv, k) -= cv3 * cv3 * dai;
        }
      }
    }
    MQ_SYNC();
#if MQ_DEV
    if constexpr (kCta) return;                 // the whole CTA continues with pcr_factor_cta()
#else
    if constexpr (kCta) { pcr_factor(); return; }
#endif
    // pass 3: twisted recursion.  chain 0 walks k = 0..mid-1 upward, chain 1 walks k = N..mid+1 downward;
    // both run the same instruction stream on two lanes.  Their Schur corrections onto the middle block
    // are parked in the (idle during factorisation) B/W/TD/MA scratch columns mid and mid+1.
    const int mid = NS / 2;
#if MQ_DEV
    const int chain = lane == 0 ? 0 : (lane == 6 ? 1 : -1);
    if (chain >= 0) factor_chain(chain, mid);
#else
    factor_chain(0, mid); factor_chain(1, mid);
#endif
    MQ_SYNC();
#if MQ_DEV
    if (lane == 0)
#endif
    {
      double S[36];
#pragma unroll
      for (int e = 0; e < 36; ++e) {
        double v = SI_(e, mid);
        if (mid > 0) v -= PK_(0, e);
        if (N - mid > 0) v -= PK_(1, e);
        S[e] = v;
      }
      inv6(S);
#pragma unroll
      for (int e = 0; e < 36; ++e) SI_(e, mid) = S[e];
    }
    MQ_SYNC();
  }
  MQ_HD void factor_chain(int chain, int mid) {
    const int steps = chain == 0 ? mid : N - mid;
    if (steps <= 0) return;
    int k = chain == 0 ? 0 : N;
    const int dk = chain == 0 ? 1 : -1;
    double S[36], G[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) S[e] = SI_(e, k);
    for (int s = 0; s < steps; ++s) {
      inv6(S);
#pragma unroll
      for (int e = 0; e < 36; ++e) SI_(e, k) = S[e];
      // coupling C (rows: next stage in walking direction, cols: stage k); per axis [pp pv; vp vv]
      const int kc = chain == 0 ? k : k - 1;
      double cpp[3], cpv[3], cvp[3], cvv[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        double a = GG_(4 * cc, kc), bq = GG_(4 * cc + 1, kc), d = GG_(4 * cc + 2, kc), e2 = GG_(4 * cc + 3, kc);
        cpp[cc] = a; cvv[cc] = e2; cpv[cc] = chain == 0 ? bq : d; cvp[cc] = chain == 0 ? d : bq;
      }
      // G = C * Sinv
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          G[cc * 6 + j] = cpp[cc] * S[cc * 6 + j] + cpv[cc] * S[(3 + cc) * 6 + j];
          G[(3 + cc) * 6 + j] = cvp[cc] * S[cc * 6 + j] + cvv[cc] * S[(3 + cc) * 6 + j];
        }
      const int kn = k + dk;
      // S_next = Tkk[kn] - G * C'
      double Snext[36];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          Snext[i * 6 + cc] = G[i * 6 + cc] * cpp[cc] + G[i * 6 + 3 + cc] * cpv[cc];
          Snext[i * 6 + 3 + cc] = G[i * 6 + cc] * cvp[cc] + G[i * 6 + 3 + cc] * cvv[cc];
        }
      // store G of stage k.  Chain 0 read its coupling from this very column above; chain 1 read
      // it from column k-1, which it overwrites only at its next step.
#pragma unroll
      for (int e = 0; e < 36; ++e) GG_(e, k) = G[e];
      if (kn == mid) {
#pragma unroll
        for (int e = 0; e < 36; ++e) PK_(chain, e) = Snext[e];
      } else {
#pragma unroll
        for (int e = 0; e < 36; ++e) S[e] = SI_(e, kn) - Snext[e];
      }
      k = kn;
    }
  }


  // ---- mode 2: block parallel cyclic reduction (PCR) of the reduced KKT matrix --------------------------
  // The (p_k, v_k) system  L_k y_{k-s} + D_k y_k + L_{k+s}' y_{k+s} = r_k  (s = 1 initially, D_k = T_k from the
  // leaf elimination above, L_k = coupling of stages k-1 and k) is reduced in log2 steps: eliminating the two
  // neighbours at distance s gives the same form at distance 2s with
  //     alpha_k = L_k D_{k-s}^-1,  gamma_k = L_{k+s}' D_{k+s}^-1,
  //     r_k <- r_k - alpha_k r_{k-s} - gamma_k r_{k+s},   D_k <- D_k - alpha_k L_k' - gamma_k L_{k+s},
  //     L_k <- -alpha_k L_{k-s}.
  // With 30 stages, after strides 1, 2, 4, 8 every equation couples to ONE partner at distance 16 (or none), and
  // the last step is folded into the final solve:  y_k = D'^-1 r_k - (D'^-1 M_k) r_partner.  All stages advance
  // in parallel, so a solve is 5 dependent 6x6 mat-vec rounds instead of the 30 dependent steps of a block LDL'.
  // Storage (shared memory, 5 x [3][12][NS][2] doubles): matrices are kept in AXIS-MAJOR order
  // (p_x v_x p_y v_y p_z v_z) and split by row pair, so that warp w of the CTA reads rows 2w, 2w+1 of its stage
  // as 12 conflict-free 16-byte loads per level.
  MQ_HHD static int pcr_old(int a) { return (a >> 1) + 3 * (a & 1); }       // axis-major index -> (p, v)-major index
  // element [a][b] (axis-major) of matrix `half` of level l: half 0 = alpha (level 4: D'^-1), 1 = gamma (level 4: D'^-1 M)
  MQ_HD int pcr_idx(int l, int half, int a, int b, int k) const {
    return l * kPcrLevelDoubles + (((a >> 1) * 12 + half * 6 + (a & 1) * 3 + (b >> 1)) * NS + k) * 2 + (b & 1);
  }
  MQ_NOINL void pcr_factor() {
    MQ_FOR_STAGES(k) {
      for (int e = 0; e < 36; ++e) { PD_(e, k) = SI_(e, k); PL_(0, e, k) = 0.0; }
      if (k > 0) {
        for (int cc = 0; cc < 3; ++cc) {
          PL_(0, cc * 6 + cc, k) = GG_(4 * cc, k - 1); PL_(0, cc * 6 + 3 + cc, k) = GG_(4 * cc + 1, k - 1);
          PL_(0, (3 + cc) * 6 + cc, k) = GG_(4 * cc + 2, k - 1); PL_(0, (3 + cc) * 6 + 3 + cc, k) = GG_(4 * cc + 3, k - 1);
        }
      }
    }
    MQ_SYNC();
    int cur = 0;
    for (int l = 0, s = 1; l < kPcrLevels; ++l, s <<= 1) {
      const bool last = l == kPcrLevels - 1;
      MQ_FOR_STAGES(k) {
        double S[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) S[e] = PD_(e, k);
        inv6(S);
#pragma unroll
        for (int e = 0; e < 36; ++e) PI_(e, k) = S[e];
      }
      MQ_SYNC();
      MQ_FOR_STAGES(k) {
        const bool hm = k - s >= 0, hp = k + s <= N, hmm = k - 2 * s >= 0;
        const int km = hm ? k - s : k, kp = hp ? k + s : k;
        if (!last) {
          // row by row, so that only a few 6-vectors are live
#pragma unroll 1
          for (int i = 0; i < 6; ++i) {
            double a[6], g[6], d[6], ln[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sa = 0.0, sg = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) { sa += PL_(cur, i * 6 + t, k) * PI_(t * 6 + j, km); sg += PL_(cur, t * 6 + i, kp) * PI_(t * 6 + j, kp); }
              a[j] = hm ? sa : 0.0; g[j] = hp ? sg : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sd = PD_(i * 6 + j, k), sl = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) { sd -= a[t] * PL_(cur, j * 6 + t, k); sl -= a[t] * PL_(cur, t * 6 + j, km); }
#pragma unroll
              for (int t = 0; t < 6; ++t) sd -= g[t] * PL_(cur, t * 6 + j, kp);
              d[j] = sd; ln[j] = hmm ? sl : 0.0;
            }
            // D_k is read by its own lane only and row i is complete: update in place.  L goes to the other buffer.
            const int an = 2 * (i % 3) + i / 3;                      // axis-major position of (p,v)-major row i
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              PD_(i * 6 + j, k) = d[j]; PL_(1 - cur, i * 6 + j, k) = ln[j];
              const int bn = 2 * (j % 3) + j / 3;
              m.PCR[pcr_idx(l, 0, an, bn, k)] = a[j]; m.PCR[pcr_idx(l, 1, an, bn, k)] = g[j];
            }
          }
        } else {
          // single partner: M = alpha (partner k-s) or gamma (partner k+s); D' = D - M (.)', y = D'^-1 r - D'^-1 M r_partner
          double M[36], S[36];
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sa = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sa += hm ? PL_(cur, i * 6 + t, k) * PI_(t * 6 + j, km) : PL_(cur, t * 6 + i, kp) * PI_(t * 6 + j, kp);
              M[i * 6 + j] = (hm || hp) ? sa : 0.0;
            }
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sd = PD_(i * 6 + j, k);
#pragma unroll
              for (int t = 0; t < 6; ++t) sd -= M[i * 6 + t] * (hm ? PL_(cur, j * 6 + t, k) : PL_(cur, t * 6 + j, kp));
              S[i * 6 + j] = sd;
            }
          inv6(S);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int an = 2 * (i % 3) + i / 3;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int bn = 2 * (j % 3) + j / 3;
              double sm = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sm += S[i * 6 + t] * M[t * 6 + j];
              m.PCR[pcr_idx(l, 0, an, bn, k)] = S[i * 6 + j]; m.PCR[pcr_idx(l, 1, an, bn, k)] = sm;
            }
          }
        }
      }
      MQ_SYNC();
      cur ^= 1;
    }
  }
#if !MQ_DEV
  // Plain-loop PCR solve on W(0..5, .) — the executable specification of what the CTA kernel's three axis warps
  // do (host emulation only).
  void pcr_solve_ref() {
    double r[2][32][6];
    for (int k = 0; k < NS; ++k) for (int a = 0; a < 6; ++a) r[0][k][a] = W_(pcr_old(a), k);
    int cur = 0;
    for (int l = 0, s = 1; l < kPcrLevels - 1; ++l, s <<= 1) {
      for (int k = 0; k < NS; ++k) {
        const int km = k - s >= 0 ? k - s : k, kp = k + s <= N ? k + s : k;
        for (int a = 0; a < 6; ++a) {
          double v = r[cur][k][a];
          for (int b = 0; b < 6; ++b) v -= m.PCR[pcr_idx(l, 0, a, b, k)] * r[cur][km][b];
          for (int b = 0; b < 6; ++b) v -= m.PCR[pcr_idx(l, 1, a, b, k)] * r[cur][kp][b];
          r[1 - cur][k][a] = v;
        }
      }
      cur ^= 1;
    }
    const int l = kPcrLevels - 1, s = 1 << l;
    for (int k = 0; k < NS; ++k) {
      const int kq = k - s >= 0 ? k - s : (k + s <= N ? k + s : k);
      for (int a = 0; a < 6; ++a) {
        double v = 0.0;
        for (int b = 0; b < 6; ++b) v += m.PCR[pcr_idx(l, 0, a, b, k)] * r[cur][k][b];
        for (int b = 0; b < 6; ++b) v -= m.PCR[pcr_idx(l, 1, a, b, k)] * r[cur][kq][b];
        W_(pcr_old(a), k) = v;
      }
    }
  }
#endif

  // ---- one ADMM iteration (auxil.h:67-112: update_xz_tilde, update_x, update_z, update_y) --------
  // rows_phase<MODE>: MODE 0 = normal iteration tail, 1 = (re)build the right-hand side only (no iterate
  // update), 2 = normal + record delta_x / delta_y for the infeasibility tests.
  // Reads x~ from W (own stage and stage k-1), updates X, Z, U in place and leaves in B the part of the next
  // right-hand side  sigma D^-2 x - c q + A' Rh (z - u)  that comes from stage k's own rows; TD gets the
  // dynamics-row terms that stage k-1 must add (rhs_finish).
  template <int MODE> MQ_HD void rows_phase() {
    const double al = st.alpha, om = 1.0 - st.alpha;
    MQ_FOR_STAGES(k) {
      double xo[NV], xm[NV], racc[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) { racc[j] = 0.0; xo[j] = 0.0; xm[j] = 0.0; }
      if (MODE != 1) {
#pragma unroll
        for (int j = 0; j < NV; ++j) { xo[j] = W_(j, k); xm[j] = k > 0 ? W_(j, k - 1) : 0.0; }
      }
      // Generic mode keeps z, u and the rho vector in L2-resident global scratch: all 63 values of the stage's dynamics and box
      // rows are requested at once (one L2 round trip per stage instead of one per row: they were 25 % of the kernel's time)
      double zr[NBR], ur[NBR], rr[NBR];
#pragma unroll
      for (int i = 0; i < NBR; ++i) { zr[i] = Z_(i, k); ur[i] = U_(i, k); rr[i] = RH_(i, k); }
      // dynamics rows
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double z = zr[r], u = ur[r];
        if (MODE != 1) {
          double zt = -xo[r];
          if (r < 3) zt += xm[r] + sh.a_pv * xm[3 + r] + sh.b_pa * xm[8 + r];
          else if (r < 6) zt += xm[r] + sh.b_va * xm[5 + r];
          else zt += xm[5 + r];
          double bnd = (k == 0) ? -x0p[r] : 0.0;
          double v = al * zt + om * z + u;
          z = fmin(fmax(v, bnd), bnd);
          double un = v - z;
          if (MODE == 2) WSDY_(r, k) = rr[r] * (un - u);
          u = un;
          Z_(r, k) = z; U_(r, k) = u;
        }
        double t = rr[r] * (z - u);
        TD_(r, k) = t;
        racc[r] -= t;
      }
      // box rows
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (j < 8 || k < N) {
          double z = zr[8 + j], u = ur[8 + j];
          if (MODE != 1) {
            double v = al * xo[j] + om * z + u;
            z = fmin(fmax(v, box_lo(j)), box_hi(j));
            double un = v - z;
            if (MODE == 2) WSDY_(8 + j, k) = rr[8 + j] * (un - u);
            u = un;
            Z_(8 + j, k) = z; U_(8 + j, k) = u;
          }
          racc[j] += rr[8 + j] * (z - u);
        }
      }
      // obstacle rows (MP.cpp:1040-1071): grad . p_k - slack  >=  low; four rows' operands are requested together (see above)
      if (k < N) {
        for (int o0 = 0; o0 < R; o0 += 4) {
          double g_[4][3], lo_[4], z_[4], u_[4], rh_[4]; int sl_[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int o = o0 + q < R ? o0 + q : o0;
            g_[q][0] = G3_(3 * o, k); g_[q][1] = G3_(3 * o + 1, k); g_[q][2] = G3_(3 * o + 2, k);
            lo_[q] = LO_(o, k); z_[q] = Z_(NBR + o, k); u_[q] = U_(NBR + o, k); rh_[q] = RH_(NBR + o, k); sl_[q] = SLK_(o, k);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int o = o0 + q;
            if (o < R) {
              const int i = NBR + o;
              const double g0 = g_[q][0], g1 = g_[q][1], g2 = g_[q][2];
              const int sl = sl_[q];
              double z = z_[q], u = u_[q];
              if (MODE != 1) {
                double zt = g0 * xo[0] + g1 * xo[1] + g2 * xo[2] - (sl ? xo[12] : xo[11]);
                double v = al * zt + om * z + u;
                z = fmax(v, lo_[q]);
                double un = v - z;
                if (MODE == 2) WSDY_(i, k) = rh_[q] * (un - u);
                u = un;
                Z_(i, k) = z; U_(i, k) = u;
              }
              double t = rh_[q] * (z - u);
              racc[0] += g0 * t; racc[1] += g1 * t; racc[2] += g2 * t;
              if (sl) racc[12] -= t; else racc[11] -= t;
            }
          }
        }
      }
      // x update (auxil.h:83 update_x) and the variable part of the next rhs
      double sdv[NV], cqv[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) { sdv[j] = SD_(j, k); cqv[j] = CQ_(j, k); }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        double x = X_(j, k);
        if (MODE != 1) {
          double xn = al * xo[j] + om * x;
          if (MODE == 2) WSDX_(j, k) = xn - x;
          x = xn;
          X_(j, k) = x;
        }
        B_(j, k) = (j < 8 || k < N) ? racc[j] + sdv[j] * x - cqv[j] : 0.0;
      }
    }
    MQ_SYNC();
  }
  // Add the dynamics-row terms of stage k+1 to stage k's rhs; publish the acceleration-leaf message.
  MQ_HD void rhs_finish() {
    MQ_FOR_STAGES(k) {
      if (k < N) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double tp = TD_(cc, k + 1), tv = TD_(3 + cc, k + 1);
          B_(cc, k) += tp;
          B_(3 + cc, k) += sh.a_pv * tp + tv;
          double ba = B_(8 + cc, k) + sh.b_pa * tp + sh.b_va * tv;
          B_(8 + cc, k) = ba;
          MA_(cc, k) = DAI_(cc, k) * ba;
        }
        B_(11, k) += TD_(6, k + 1);
        B_(12, k) += TD_(7, k + 1);
      }
    }
    MQ_SYNC();
  }
  // Forward elimination of the leaf variables -> reduced 6-vector per stage in B(0..5).
  MQ_HD void leaf_forward() {
    MQ_FOR_STAGES(k) {
      double r[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) r[i] = B_(i, k);
      if (k < N) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          double r11 = B_(11 + t, k) - ESD_(t, k + 1) * B_(6 + t, k + 1);
          B_(11 + t, k) = r11;
          double f = DGI_(t, k) * r11;
          r[0] -= FS_(3 * t, k) * f; r[1] -= FS_(3 * t + 1, k) * f; r[2] -= FS_(3 * t + 2, k) * f;
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { double mm = MA_(cc, k); r[cc] -= CV_(4 * cc, k) * mm; r[3 + cc] -= CV_(4 * cc + 1, k) * mm; }
      }
      if (k > 0) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { double mm = MA_(cc, k - 1); r[cc] -= CV_(4 * cc + 2, k - 1) * mm; r[3 + cc] -= CV_(4 * cc + 3, k - 1) * mm; }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) W_(i, k) = r[i];
    }
    MQ_SYNC();
  }
  // Twisted block LDL' solve on W(0..5, .):  L w = r (two chains meeting at mid), v = S^-1 w, L' y = v.
  MQ_HD void chain_solve() {
    const int mid = NS / 2;
    const int s_top = mid, s_bot = N - mid;
    const int nst = s_top > s_bot ? s_top : s_bot;
#if MQ_DEV
    const int half = lane / 6, i = lane - 6 * half;
    const bool mine = half < 2;
    const int base = mine ? 6 * half : 0;
    const int dk = half == 0 ? 1 : -1;
    const int k0 = half == 0 ? 0 : N;
    const int mysteps = half == 0 ? s_top : (half == 1 ? s_bot : 0);
    // forward
    double wi = mine ? W_(i, k0) : 0.0, cm = 0.0;
    for (int s = 0; s < nst; ++s) {
      const bool act = s < mysteps;
      const int k = k0 + dk * s, kn = k + dk;
      double acc = 0.0;
      double gr[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) gr[j] = act ? GG_(i * 6 + j, k) : 0.0;
      double rn = (act && kn != mid) ? W_(i, kn) : 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) acc += gr[j] * __shfl_sync(0xffffffffu, wi, base + j);
      if (act) {
        if (kn != mid) { wi = rn - acc; W_(i, kn) = wi; }
        else cm = acc;
      }
    }
    {
      double cb = __shfl_sync(0xffffffffu, cm, 6 + (lane % 6));
      if (lane < 6) W_(lane, mid) = W_(lane, mid) - cm - cb;
    }
    MQ_SYNC();
    // v = Sinv w  (lane = stage)
    MQ_FOR_STAGES(k) {
      double w[6], v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) w[j] = W_(j, k);
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) s += SI_(a * 6 + j, k) * w[j];
        v[a] = s;
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) W_(j, k) = v[j];
    }
    MQ_SYNC();
    // backward
    double yi = W_(lane % 6, mid);
    for (int s = 0; s < nst; ++s) {
      const bool act = s < mysteps;
      const int k = half == 0 ? mid - 1 - s : mid + 1 + s;
      double gc[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) gc[j] = act ? GG_(j * 6 + i, k) : 0.0;
      double vk = act ? W_(i, k) : 0.0;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) acc += gc[j] * __shfl_sync(0xffffffffu, yi, base + j);
      if (act) { yi = vk - acc; W_(i, k) = yi; }
    }
    MQ_SYNC();
#else
    for (int k = 0; k < s_top; ++k) {
      double acc[6];
      for (int a = 0; a < 6; ++a) { acc[a] = 0; for (int j = 0; j < 6; ++j) acc[a] += GG_(a * 6 + j, k) * W_(j, k); }
      for (int a = 0; a < 6; ++a) W_(a, k + 1) -= acc[a];
    }
    double cb[6] = {0, 0, 0, 0, 0, 0};
    for (int k = N; k > mid; --k) {
      double acc[6];
      for (int a = 0; a < 6; ++a) { acc[a] = 0; for (int j = 0; j < 6; ++j) acc[a] += GG_(a * 6 + j, k) * W_(j, k); }
      for (int a = 0; a < 6; ++a) W_(a, k - 1) -= acc[a];
    }
    (void)cb;
    for (int k = 0; k < NS; ++k) {
      double v[6];
      for (int a = 0; a < 6; ++a) { v[a] = 0; for (int j = 0; j < 6; ++j) v[a] += SI_(a * 6 + j, k) * W_(j, k); }
      for (int a = 0; a < 6; ++a) W_(a, k) = v[a];
    }
    for (int k = mid - 1; k >= 0; --k)
      for (int a = 0; a < 6; ++a) { double s = 0; for (int j = 0; j < 6; ++j) s += GG_(j * 6 + a, k) * W_(j, k + 1); W_(a, k) -= s; }
    for (int k = mid + 1; k <= N; ++k)
      for (int a = 0; a < 6; ++a) { double s = 0; for (int j = 0; j < 6; ++j) s += GG_(j * 6 + a, k) * W_(j, k - 1); W_(a, k) -= s; }
#endif
  }
  // Back-substitution of the leaf variables: accelerations and slack inputs, then slack states.
  MQ_HD void leaf_backward() {
    MQ_FOR_STAGES(k) {
      if (k < N) {
        double yk[6], yn[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { yk[i] = W_(i, k); yn[i] = W_(i, k + 1); }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          W_(8 + cc, k) = DAI_(cc, k) * (B_(8 + cc, k) - CV_(4 * cc, k) * yk[cc] - CV_(4 * cc + 1, k) * yk[3 + cc] -
                                         CV_(4 * cc + 2, k) * yn[cc] - CV_(4 * cc + 3, k) * yn[3 + cc]);
#pragma unroll
        for (int t = 0; t < 2; ++t)
          W_(11 + t, k) = DGI_(t, k) * (B_(11 + t, k) - FS_(3 * t, k) * yk[0] - FS_(3 * t + 1, k) * yk[1] - FS_(3 * t + 2, k) * yk[2]);
      } else {
#pragma unroll
        for (int j = 8; j < NV; ++j) W_(j, k) = 0.0;
      }
    }
    MQ_SYNC();
    MQ_FOR_STAGES(k) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        double v = DSI_(t, k) * B_(6 + t, k);
        if (k > 0) v -= ESD_(t, k) * W_(11 + t, k - 1);
        W_(6 + t, k) = v;
      }
    }
    MQ_SYNC();
  }
  template <int MODE> MQ_HD void iterate() {
    leaf_forward();
#if !MQ_DEV
    if constexpr (kCta) pcr_solve_ref(); else
#endif
    chain_solve();
    leaf_backward();
    rows_phase<MODE>();
    rhs_finish();
  }

  // ---- update_info (auxil.h:154): un-scaled residuals, objective, and the scaled norms adapt_rho needs ----
  MQ_NOINL void update_info(int iter) {
    double red[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) red[e] = 0.0;
    double ob = 0.0;
    auto xv = [&](int j, int kk) { return X_(j, kk); };
    auto yv = [&](int i, int kk) { return RH_(i, kk) * U_(i, kk); };
    MQ_FOR_STAGES(k) {
      const int nr = nrows(k), nv = nvars(k);
      for (int i = 0; i < nr; ++i) {
        double ax = row_ax(k, i, xv), z = Z_(i, k), e = WSE_(i, k);
        double d = fabs(ax - z);
        red[0] = fmax(red[0], d); red[1] = fmax(red[1], fabs(ax)); red[2] = fmax(red[2], fabs(z));
        red[3] = fmax(red[3], e * d); red[4] = fmax(red[4], e * fabs(ax)); red[5] = fmax(red[5], e * fabs(z));
      }
      for (int j = 0; j < nv; ++j) {
        double x = X_(j, k), p = pd[k * NV + j];
        double px = c * p * x, aty = col_aty(k, j, yv), dj = WSD_(j, k);
        double d = fabs(px + CQ_(j, k) + aty);
        red[6] = fmax(red[6], d); red[7] = fmax(red[7], fabs(px)); red[8] = fmax(red[8], fabs(aty));
        red[9] = fmax(red[9], dj * d); red[10] = fmax(red[10], dj * fabs(px)); red[11] = fmax(red[11], dj * fabs(aty));
        ob += (0.5 * p * x + CQ_(j, k) * cinv) * x;
      }
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) red[e] = wmax(red[e]);
    obj = wsum(ob);
    pri_res = red[0]; nAx = red[1]; nZ = red[2]; pri_s = red[3]; nAx_s = red[4]; nZ_s = red[5];
    dua_res = red[6] * cinv; nPx = red[7] * cinv; nAty = red[8] * cinv; dua_s = red[9]; nPx_s = red[10]; nAty_s = red[11];
    info_iter = iter;
  }

  // auxil.h:137 is_primal_infeasible — delta_y certificate test in un-scaled coordinates (E cancels).
  MQ_NOINL bool is_primal_infeasible(double eps) {
    double nd = 0.0;
    MQ_FOR_STAGES(k) {
      const int nr = nrows(k);
      for (int i = 0; i < nr; ++i) {
        double lo, hi; row_bounds(k, i, lo, hi);
        double e = WSE_(i, k), ls = e * lo, us = e * hi, dy = WSDY_(i, k);
        if (us > kInfty * kMinScaling) { if (ls < -kInfty * kMinScaling) dy = 0.0; else dy = fmin(dy, 0.0); }
        else if (ls < -kInfty * kMinScaling) dy = fmax(dy, 0.0);
        WSDY_(i, k) = dy;
        nd = fmax(nd, fabs(dy));
      }
    }
    nd = wmax(nd);
    MQ_SYNC();
    if (!(nd > eps)) return false;
    // IEEE semantics kept on purpose: an infinite bound times a zero multiplier is NaN, which makes the
    // comparison false — with the reference's +-inf bounds (MP.cpp:913-914) OSQP never declares primal
    // infeasibility, and neither do we.
    double lhs = 0.0;
    MQ_FOR_STAGES(k) {
      const int nr = nrows(k);
      for (int i = 0; i < nr; ++i) {
        double lo, hi; row_bounds(k, i, lo, hi);
        double dy = WSDY_(i, k);
        lhs += hi * fmax(dy, 0.0) + lo * fmin(dy, 0.0);
      }
    }
    lhs = wsum(lhs);
    if (!(lhs < -eps * nd)) return false;
    double na = 0.0;
    auto dv = [&](int i, int kk) { return WSDY_(i, kk); };
    MQ_FOR_STAGES(k) { const int nv = nvars(k); for (int j = 0; j < nv; ++j) na = fmax(na, fabs(col_aty(k, j, dv))); }
    na = wmax(na);
    return na < eps * nd;
  }
  // auxil.h:148 is_dual_infeasible — delta_x certificate test.
  MQ_NOINL bool is_dual_infeasible(double eps) {
    double nd = 0.0, qd = 0.0, pm = 0.0;
    MQ_FOR_STAGES(k) {
      const int nv = nvars(k);
      for (int j = 0; j < nv; ++j) {
        double dx = WSDX_(j, k);
        nd = fmax(nd, fabs(dx)); qd += CQ_(j, k) * dx; pm = fmax(pm, fabs(c * pd[k * NV + j] * dx));
      }
    }
    nd = wmax(nd); qd = wsum(qd); pm = wmax(pm);
    if (!(nd > eps)) return false;
    if (!(qd < -c * eps * nd)) return false;
    if (!(pm < c * eps * nd)) return false;
    bool bad = false;
    auto dv = [&](int j, int kk) { return WSDX_(j, kk); };
    MQ_FOR_STAGES(k) {
      const int nr = nrows(k);
      for (int i = 0; i < nr; ++i) {
        double lo, hi; row_bounds(k, i, lo, hi);
        double e = WSE_(i, k), adx = row_ax(k, i, dv);
        if ((e * hi < kInfty * kMinScaling && adx > eps * nd) || (e * lo > -kInfty * kMinScaling && adx < -eps * nd)) bad = true;
      }
    }
    return !wany(bad);
  }
  // auxil.h:133 check_termination
  MQ_NOINL bool check_termination(bool approx) {
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (pri_res > kInfty || dua_res > kInfty) { status = kNonCvx; obj = kOsqpNan; return true; }
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool pr = false, dr = false, pinf = false, dinf = false;
    if (sh.m == 0) pr = true;
    else { double eps_p = ea + er * fmax(nZ, nAx); if (pri_res < eps_p) pr = true; else pinf = is_primal_infeasible(epi); }
    double eps_d = ea + er * fmax(nq, fmax(nAty, nPx));
    if (dua_res < eps_d) dr = true; else dinf = is_dual_infeasible(edi);
    if (pr && dr) { status = approx ? kSolvedInacc : kSolved; return true; }
    if (pinf) { status = approx ? kPrimInfInacc : kPrimInf; obj = kInfty; return true; }
    if (dinf) { status = approx ? kDualInfInacc : kDualInf; obj = -kInfty; return true; }
    return false;
  }
  // auxil.h:21-38 compute_rho_estimate / adapt_rho, osqp.h osqp_update_rho
  MQ_NOINL bool adapt_rho() {
    double pn = pri_s / (fmax(nZ_s, nAx_s) + 1e-10);
    double dn = dua_s / (fmax(nq_s, fmax(nAty_s, nPx_s)) + 1e-10);
    double rn = rho * sqrt(pn / (dn + 1e-10));
    rn = fmin(fmax(rn, kRhoMin), kRhoMax);
    if (!(rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance)) return false;
    rho = rn;
    MQ_FOR_STAGES(k) {
      const int nr = nrows(k);
      for (int i = 0; i < nr; ++i) {
        double lo, hi; row_bounds(k, i, lo, hi);
        double e = WSE_(i, k);
        int t = row_type(e, lo, hi);
        if (t >= 0) {
          double rold = RH_(i, k), rnew = rho_of_type(t) * e * e;
          U_(i, k) = U_(i, k) * rold / rnew;   // y is kept across a rho update; u = y / Rh
          RH_(i, k) = rnew;
        }
      }
    }
    MQ_SYNC();
    rho_updates += 1;
    return true;                                 // caller refactorises and rebuilds the rhs
  }

  // ---- bursts of iterations ------------------------------------------------------------------------
  // A burst is a run of ADMM iterations with no termination check in between; its last iteration records
  // delta_x / delta_y for the infeasibility tests.
  MQ_HD void burst(int niter) {
#if MQ_DEV
    if constexpr (kFast) { burst_fast(niter); return; }
#endif
    for (int it = 0; it < niter; ++it) { if (it == niter - 1) iterate<2>(); else iterate<0>(); }
  }

#if MQ_DEV
  // Fast path: stage k = lane; x, z, u and the rhs stay in registers for the whole burst, neighbour-stage
  // values travel by warp shuffle, shared memory is only read (scaled data, factor) except for the
  // 6-vector exchange with the 12 chain lanes.
  MQ_HD void burst_fast(int niter) {
    static_assert(!kFast || NST <= 32, "fast path needs horizon <= 32");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int mid = NS / 2, s_top = mid, s_bot = N - mid;
    const int k = lane < NS ? lane : NS - 1;          // ghost lanes shadow the last stage, never write
    const bool live = lane < NS, hasu = lane < N, notfirst = lane > 0 && live;
    const int kp = k < N ? k + 1 : k, km = k > 0 ? k - 1 : 0;
    const double al = st.alpha, om = 1.0 - st.alpha, apv = sh.a_pv, bpa = sh.b_pa, bva = sh.b_va;
    double x[NV], b[NV], z[MK], u[MK], bnd[8];
    int sl[R > 0 ? R : 1];
#pragma unroll
    for (int j = 0; j < NV; ++j) { x[j] = X_(j, k); b[j] = B_(j, k); }
#pragma unroll
    for (int i = 0; i < MK; ++i) { z[i] = Z_(i, k); u[i] = U_(i, k); }
#pragma unroll
    for (int r = 0; r < 8; ++r) bnd[r] = (k == 0) ? -x0p[r] : 0.0;
#pragma unroll
    for (int o = 0; o < R; ++o) sl[o] = hasu ? SLK_(o, k) : 0;
    // chain-lane geometry.  Lanes 0-5: upper chain (stages 0..mid-1), lanes 6-11: lower chain (stages N..mid+1);
    // lanes 12-31 shadow lanes 0-11 with their writes disabled so the whole warp runs one instruction stream.
    static_assert(NS % 2 == 0, "fast path assumes an even number of stages (16-byte aligned rows)");
    constexpr int nst = s_top > s_bot ? s_top : s_bot;
    const int cl = lane % 12, chalf = cl / 6, ci = cl - 6 * chalf;
    const bool cwr = lane < 12;
    const int mysteps = chalf == 0 ? s_top : s_bot;
    const int slot0 = chalf == 0 ? 0 : mid + 1;             // first source slot of the forward walk
    const int slotT = chalf == 0 ? mid - 1 : N;             // first target slot of the backward walk
    const double* pGr = m.GG + slot0 * 36 + ci * 6;         // row ci of G at the forward source slot
    double* pWc = m.W + slot0 * kWS;
    const double* pGt = m.GG + slotT * 36 + ci;             // column ci of G at the backward target slot
    double* pWt = m.W + slotT * kWS;
    double* dummy = m.PK + 8 * (lane & 7);                  // scratch slot for lanes whose result is not needed
    double* pWk = live ? m.W + cslot(k) * kWS : dummy;      // this stage's 6-vector (ghost lanes: dummy)

    for (int it = 0; it < niter; ++it) {
      if (it == niter - 1 && live) {                 // park the old x, u: deltas are formed after the update
#pragma unroll
        for (int j = 0; j < NV; ++j) WSDX_(j, k) = x[j];
#pragma unroll
        for (int i = 0; i < MK; ++i) WSDY_(i, k) = u[i];
      }
      // ---- forward elimination of the leaves -> reduced rhs r[0..5]
      double r[6], r11[2], ma[3];
#pragma unroll
      for (int i = 0; i < 6; ++i) r[i] = b[i];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        double bn = __shfl_down_sync(FULL, b[6 + t], 1);
        double v = b[11 + t] - ESD_(t, kp) * bn;
        r11[t] = hasu ? v : 0.0;
        double f = DGI_(t, k) * r11[t];
        r[0] -= FS_(3 * t, k) * f; r[1] -= FS_(3 * t + 1, k) * f; r[2] -= FS_(3 * t + 2, k) * f;
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        ma[cc] = DAI_(cc, k) * b[8 + cc];
        r[cc] -= CV_(4 * cc, k) * ma[cc]; r[3 + cc] -= CV_(4 * cc + 1, k) * ma[cc];
        double mm = __shfl_up_sync(FULL, ma[cc], 1);
        mm = notfirst ? mm : 0.0;
        r[cc] -= CV_(4 * cc + 2, km) * mm; r[3 + cc] -= CV_(4 * cc + 3, km) * mm;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) pWk[i] = r[i];
      __syncwarp();
      // ---- twisted chain, forward: L w = r.  6 lanes per chain (lane = row of the 6x6 block); lanes 0-5 walk up
      // from stage 0, lanes 6-11 walk down from stage N, both upward in chain-slot order.  The running vector is
      // exchanged through the W row that has to be written anyway (1 STS + 3 broadcast LDS.128 per step).
      {
        double cm = 0.0;
        // software pipeline: the G row and the rhs entry of step s+1 are loaded before the store of step s
        double2 ga = *reinterpret_cast<const double2*>(pGr), gb = *reinterpret_cast<const double2*>(pGr + 2),
                gc = *reinterpret_cast<const double2*>(pGr + 4);
        double rn = pWc[kWS + ci];
#pragma unroll
        for (int s = 0; s < nst; ++s) {
          const bool act = s < mysteps, lastst = s == mysteps - 1;
          if (lastst && chalf == 1) rn = 0.0;
          const double2 wa = *reinterpret_cast<const double2*>(pWc + s * kWS), wb = *reinterpret_cast<const double2*>(pWc + s * kWS + 2),
                        wc = *reinterpret_cast<const double2*>(pWc + s * kWS + 4);
          double a0 = -ga.x * wa.x, a1 = fma(-ga.y, wa.y, rn);
          a0 = fma(-gb.x, wb.x, a0); a1 = fma(-gb.y, wb.y, a1);
          a0 = fma(-gc.x, wc.x, a0); a1 = fma(-gc.y, wc.y, a1);
          if (s + 1 < nst) {
            ga = *reinterpret_cast<const double2*>(pGr + (s + 1) * 36); gb = *reinterpret_cast<const double2*>(pGr + (s + 1) * 36 + 2);
            gc = *reinterpret_cast<const double2*>(pGr + (s + 1) * 36 + 4);
            rn = pWc[(s + 2) * kWS + ci];
          }
          const double wn = a0 + a1;
          double* dst = (act && !lastst && cwr) ? pWc + (s + 1) * kWS + ci : dummy;   // branch-free: idle lanes hit a dummy slot
          *dst = wn;
          cm = (act && lastst) ? wn : cm;
          __syncwarp();
        }
        const double cb = __shfl_sync(FULL, cm, 6 + ci);
        double* dst = lane < 6 ? m.W + mid * kWS + lane : dummy;
        *dst = cm + (s_bot > 0 ? cb : 0.0);
      }
      __syncwarp();
      // ---- v = S^-1 w per stage
      {
        double w[6], v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = W_(j, k);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double s0 = SI_(a * 6, k) * w[0], s1 = SI_(a * 6 + 1, k) * w[1];
          s0 = fma(SI_(a * 6 + 2, k), w[2], s0); s1 = fma(SI_(a * 6 + 3, k), w[3], s1);
          s0 = fma(SI_(a * 6 + 4, k), w[4], s0); s1 = fma(SI_(a * 6 + 5, k), w[5], s1);
          v[a] = s0 + s1;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) pWk[j] = v[j];
      }
      __syncwarp();
      // ---- twisted chain, backward: L' y = v (lane = column of G)
      {
        double g0 = pGt[0], g1 = pGt[6], g2 = pGt[12], g3 = pGt[18], g4 = pGt[24], g5 = pGt[30];
        double vk = pWt[ci];
#pragma unroll
        for (int s = 0; s < nst; ++s) {
          const bool act = s < mysteps;
          const double* ps = s == 0 ? m.W + mid * kWS : pWt - (s - 1) * kWS;      // y of the inner neighbour
          const double2 ya = *reinterpret_cast<const double2*>(ps), yb = *reinterpret_cast<const double2*>(ps + 2),
                        yc = *reinterpret_cast<const double2*>(ps + 4);
          double a0 = -g0 * ya.x, a1 = fma(-g1, ya.y, vk);
          a0 = fma(-g2, yb.x, a0); a1 = fma(-g3, yb.y, a1);
          a0 = fma(-g4, yc.x, a0); a1 = fma(-g5, yc.y, a1);
          if (s + 1 < nst) {
            const double* pgc = pGt - (s + 1) * 36;
            g0 = pgc[0]; g1 = pgc[6]; g2 = pgc[12]; g3 = pgc[18]; g4 = pgc[24]; g5 = pgc[30];
            vk = pWt[-(s + 1) * kWS + ci];
          }
          double* dst = (act && cwr) ? pWt - s * kWS + ci : dummy;
          *dst = a0 + a1;
          __syncwarp();
        }
      }
      __syncwarp();
      // ---- back-substitution of the leaves -> x~ (xt)
      double xt[NV];
      {
        double yn[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { xt[i] = W_(i, k); yn[i] = __shfl_down_sync(FULL, xt[i], 1); }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          xt[8 + cc] = DAI_(cc, k) * (b[8 + cc] - CV_(4 * cc, k) * xt[cc] - CV_(4 * cc + 1, k) * xt[3 + cc] -
                                      CV_(4 * cc + 2, k) * yn[cc] - CV_(4 * cc + 3, k) * yn[3 + cc]);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          xt[11 + t] = DGI_(t, k) * (r11[t] - FS_(3 * t, k) * xt[0] - FS_(3 * t + 1, k) * xt[1] - FS_(3 * t + 2, k) * xt[2]);
          double xp = __shfl_up_sync(FULL, xt[11 + t], 1);
          xt[6 + t] = DSI_(t, k) * b[6 + t] - (notfirst ? ESD_(t, k) * xp : 0.0);
        }
        if (!hasu) {
#pragma unroll
          for (int j = 8; j < NV; ++j) xt[j] = 0.0;
        }
      }
      // ---- z, u, x updates and the next right-hand side.  Clamps are spelled as OSQP's c_min/c_max macros
      // (a > b ? a : b), which is both the reference semantics for NaN and cheaper than fmin/fmax.
      double racc[NV], td[8];
#pragma unroll
      for (int j = 0; j < NV; ++j) racc[j] = 0.0;
      {
        // what stage k-1 contributes to stage k's dynamics rows: Ad x~_{k-1} + Bd u~_{k-1}  (8 values, one shuffle each)
        double pr[8];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) { pr[cc] = xt[cc] + apv * xt[3 + cc] + bpa * xt[8 + cc]; pr[3 + cc] = xt[3 + cc] + bva * xt[8 + cc]; }
        pr[6] = xt[11]; pr[7] = xt[12];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          double pv = __shfl_up_sync(FULL, pr[rr], 1);
          double zt = (notfirst ? pv : 0.0) - xt[rr];
          double v = al * zt + om * z[rr] + u[rr];
          z[rr] = bnd[rr];                         // equality row: the projection onto [b, b] is b
          u[rr] = v - bnd[rr];
          double t = RH_(rr, k) * (bnd[rr] - u[rr]);
          td[rr] = t; racc[rr] -= t;
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        double v = al * xt[j] + om * z[8 + j] + u[8 + j];
        const double lo = sh.blo[j], hi = sh.bhi[j];
        double zn = v > lo ? v : lo;
        zn = zn < hi ? zn : hi;
        z[8 + j] = zn; u[8 + j] = v - zn;
        racc[j] += RH_(8 + j, k) * (zn - u[8 + j]);
      }
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const int i = NBR + o;
        double g0 = G3_(3 * o, k), g1 = G3_(3 * o + 1, k), g2 = G3_(3 * o + 2, k);
        double zt = g0 * xt[0] + g1 * xt[1] + g2 * xt[2] - (sl[o] ? xt[12] : xt[11]);
        double v = al * zt + om * z[i] + u[i];
        const double lo = LO_(o, k);
        double zn = v > lo ? v : lo;
        z[i] = zn; u[i] = v - zn;
        double t = RH_(i, k) * (zn - u[i]);
        racc[0] += g0 * t; racc[1] += g1 * t; racc[2] += g2 * t;
        if (sl[o]) racc[12] -= t; else racc[11] -= t;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        x[j] = al * xt[j] + om * x[j];
        b[j] = (j < 8 || hasu) ? racc[j] + SD_(j, k) * x[j] - CQ_(j, k) : 0.0;
      }
      // dynamics rows of stage k+1 feed stage k's rhs
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        double tp = __shfl_down_sync(FULL, td[cc], 1), tv = __shfl_down_sync(FULL, td[3 + cc], 1);
        if (hasu) { b[cc] += tp; b[3 + cc] += apv * tp + tv; b[8 + cc] += bpa * tp + bva * tv; }
      }
      {
        double t6 = __shfl_down_sync(FULL, td[6], 1), t7 = __shfl_down_sync(FULL, td[7], 1);
        if (hasu) { b[11] += t6; b[12] += t7; }
      }
    }
    // park the state; form the deltas of the last iteration
    if (live) {
#pragma unroll
      for (int j = 0; j < NV; ++j) { X_(j, k) = x[j]; B_(j, k) = b[j]; WSDX_(j, k) = x[j] - WSDX_(j, k); }
#pragma unroll
      for (int i = 0; i < MK; ++i) { Z_(i, k) = z[i]; U_(i, k) = u[i]; WSDY_(i, k) = RH_(i, k) * (u[i] - WSDY_(i, k)); }
    }
    __syncwarp();
  }
#endif


#if MQ_DEV
  // ================================================================================================
  // mode 2: one 4-warp CTA per QP.  lane = stage in every warp; warps 0..2 own one axis each (variables p, v, a of
  // that axis, their two dynamics rows and three box rows, and rows 2w, 2w+1 of the PCR solve); warp 3 owns the
  // slack states / slack inputs and their rows; obstacle row o belongs to warp o % 4.  All iterates live in registers during
  // a burst; the warps exchange the reduced right-hand side, the PCR intermediate vectors and the solution through small
  // shared-memory buffers with named barriers (kBarRS, kBarAll, [kBarAxis | kBarH1, kBarA], kBarY, kBarT per iteration).
  // One-per-SM blocks add three PCR assistants (assist_role) and a row helper (helper_role).
  // ================================================================================================
  static MQ_HD void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
  static MQ_HD void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
  static MQ_HD double up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
  static MQ_HD double dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
  static MQ_HD double clampd(double v, double lo, double hi) { double z = v > lo ? v : lo; return z < hi ? z : hi; }
  static constexpr int kBarAll = 1, kBarAxis = 2, kBarY = 3, kBarT = 4, kBarCmd = 5, kBarH1 = 6, kBarRS = 7, kBarA = 8;

#ifdef MPCQP_PHASE_TIMING
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // setup, leaf, pcr factor, load, iterate, info+check, park/adapt, store
#define MQ_T0() long long t0_ = clock64()
#define MQ_T(i) do { long long t1_ = clock64(); tacc[i] += t1_ - t0_; t0_ = t1_; } while (0)
#else
#define MQ_T0() ((void)0)
#define MQ_T(i) ((void)0)
#endif
  // PCR factorisation by the whole CTA (same recurrences as pcr_factor() above, which stays as the host-emulated
  // specification).  Work split: axis warp w owns rows 2w, 2w+1 (axis-major) of every 6x6 product of its stage and
  // keeps its two rows of D in registers across levels; warp 3 inverts the D blocks (lane = stage) while the axis
  // warps wait.  Workspace in shared memory, [entry][stage]: DI (D, then D^-1 in place), LC (current L), LN (next L).
  MQ_HD void pcr_factor_cta(const int warp) {
    const int k = lane < NS ? lane : NS - 1;
    const bool live = lane < NS;
    double* DI = m.WK; double* LC = m.WK + 36 * NS; double* LN = m.WK + 72 * NS;
    cta_sync();                                      // warp 0 has written T (SI_) and the couplings (GG_)
    MQ_T0();
    // initial D = T (leaf elimination, (p,v)-major) and L = coupling of stages k-1, k, permuted to axis-major.  T and
    // the couplings are staged in the workspace itself (SI_ in the LN slots, GG_ in the tail of LC): read, then write.
    double dr[12];
    {
      double dv[9], lv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const int e = warp * 9 + q, a = e / 6, b2 = e - 6 * a, oa = pcr_old(a), ob = pcr_old(b2);
        dv[q] = SI_(oa * 6 + ob, k);
        lv[q] = (k > 0 && (oa % 3) == (ob % 3)) ? GG_(4 * (oa % 3) + (oa / 3) * 2 + (ob / 3), k - 1) : 0.0;
      }
      if (warp < 3) {
#pragma unroll
        for (int ar = 0; ar < 2; ++ar)
#pragma unroll
          for (int j = 0; j < 6; ++j) dr[ar * 6 + j] = SI_(pcr_old(2 * warp + ar) * 6 + pcr_old(j), k);
      }
      cta_sync();
      if (live) {
#pragma unroll
        for (int q = 0; q < 9; ++q) { const int e = warp * 9 + q; DI[e * NS + k] = dv[q]; LC[e * NS + k] = lv[q]; }
      }
    }
    for (int l = 0, s = 1; l < kPcrLevels; ++l, s <<= 1) {
      const bool last = l == kPcrLevels - 1;
      cta_sync();                                    // DI = D of every stage, LC = L
      MQ_T(l == 0 ? 8 : 10);
      if (warp == 3) {
        double S[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) S[e] = DI[e * NS + k];
        inv6(S);
        if (live) {
#pragma unroll
          for (int e = 0; e < 36; ++e) DI[e * NS + k] = S[e];
        }
      }
      cta_sync();                                    // DI = D^-1
      MQ_T(9);
      const bool hm = k - s >= 0, hp = k + s <= N, hmm = k - 2 * s >= 0;
      const int km = hm ? k - s : k, kp = hp ? k + s : k;
      double2* const P2 = reinterpret_cast<double2*>(m.PCR) + l * (kPcrLevelDoubles / 2) + (warp * 12) * NS + k;
      double out_l[12];                                   // what goes to LN: next L rows (or M rows on the last level)
      if (warp < 3) {
        const int a0 = 2 * warp;
        double al_[12], ga_[12], M[36];
        // every product reads its operands into registers first: nothing below is stored until all loads are done
        if (!last) {
          // alpha rows = L_k rows . D_{k-s}^-1
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = DI[e * NS + km];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar) {
            double lr[6];
#pragma unroll
            for (int t = 0; t < 6; ++t) lr[t] = LC[((a0 + ar) * 6 + t) * NS + k];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sa = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sa = fma(lr[t], M[t * 6 + j], sa);
              al_[ar * 6 + j] = hm ? sa : 0.0;
            }
          }
          // gamma rows = (L_{k+s}')rows . D_{k+s}^-1
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = DI[e * NS + kp];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar) {
            double lc[6];
#pragma unroll
            for (int t = 0; t < 6; ++t) lc[t] = LC[(t * 6 + a0 + ar) * NS + kp];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sg = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sg = fma(lc[t], M[t * 6 + j], sg);
              ga_[ar * 6 + j] = hp ? sg : 0.0;
            }
          }
          // D' rows -= alpha L_k'
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = LC[e * NS + k];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sd = dr[ar * 6 + j];
#pragma unroll
              for (int t = 0; t < 6; ++t) sd = fma(-al_[ar * 6 + t], M[j * 6 + t], sd);
              dr[ar * 6 + j] = sd;
            }
          // D' rows -= gamma L_{k+s}
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = LC[e * NS + kp];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sd = dr[ar * 6 + j];
#pragma unroll
              for (int t = 0; t < 6; ++t) sd = fma(-ga_[ar * 6 + t], M[t * 6 + j], sd);
              dr[ar * 6 + j] = sd;
            }
          // next L rows = -alpha L_{k-s}
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = LC[e * NS + km];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sl = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sl = fma(-al_[ar * 6 + t], M[t * 6 + j], sl);
              out_l[ar * 6 + j] = hmm ? sl : 0.0;
            }
          if (live) {
#pragma unroll
            for (int ar = 0; ar < 2; ++ar)
#pragma unroll
              for (int bp = 0; bp < 3; ++bp) {
                P2[(ar * 3 + bp) * NS] = make_double2(al_[ar * 6 + 2 * bp], al_[ar * 6 + 2 * bp + 1]);
                P2[(6 + ar * 3 + bp) * NS] = make_double2(ga_[ar * 6 + 2 * bp], ga_[ar * 6 + 2 * bp + 1]);
              }
          }
        } else {
          // single partner: M = alpha (partner k-s) or gamma (partner k+s);  D' = D - M (.)'
          const int kq2 = hm ? km : kp;
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = DI[e * NS + kq2];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar) {
            double lr[6];
#pragma unroll
            for (int t = 0; t < 6; ++t) lr[t] = hm ? LC[((a0 + ar) * 6 + t) * NS + k] : LC[(t * 6 + a0 + ar) * NS + kp];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sa = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) sa = fma(lr[t], M[t * 6 + j], sa);
              out_l[ar * 6 + j] = (hm || hp) ? sa : 0.0;      // M rows
            }
          }
#pragma unroll
          for (int e = 0; e < 36; ++e) M[e] = hm ? LC[e * NS + k] : LC[e * NS + kp];
#pragma unroll
          for (int ar = 0; ar < 2; ++ar)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double sd = dr[ar * 6 + j];
#pragma unroll
              for (int t = 0; t < 6; ++t) sd = fma(-out_l[ar * 6 + t], hm ? M[j * 6 + t] : M[t * 6 + j], sd);
              dr[ar * 6 + j] = sd;
            }
        }
        if (live) {
#pragma unroll
          for (int e = 0; e < 12; ++e) LN[(a0 * 6 + e) * NS + k] = out_l[e];
        }
      }
      cta_sync();                                    // everyone is done with D^-1 and L of this level
      if (warp < 3 && live) {
#pragma unroll
        for (int e = 0; e < 12; ++e) DI[((2 * warp) * 6 + e) * NS + k] = dr[e];
      }
      double* t_ = LC; LC = LN; LN = t_;
    }
    // D' of the last level is in DI, M in LC: invert, then rows of D'^-1 and D'^-1 M go to level 4 of the PCR store
    cta_sync();
    MQ_T(10);
    if (warp == 3) {
      double S[36];
#pragma unroll
      for (int e = 0; e < 36; ++e) S[e] = DI[e * NS + k];
      inv6(S);
      if (live) {
#pragma unroll
        for (int e = 0; e < 36; ++e) DI[e * NS + k] = S[e];
      }
    }
    cta_sync();
    if (warp < 3 && live) {
      double2* const P2 = reinterpret_cast<double2*>(m.PCR) + (kPcrLevels - 1) * (kPcrLevelDoubles / 2) + (warp * 12) * NS + k;
      double M[36];
#pragma unroll
      for (int e = 0; e < 36; ++e) M[e] = LC[e * NS + k];
#pragma unroll
      for (int ar = 0; ar < 2; ++ar) {
        const int a = 2 * warp + ar;
        double di[6], dm[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) di[j] = DI[(a * 6 + j) * NS + k];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double sm = 0.0;
#pragma unroll
          for (int t = 0; t < 6; ++t) sm = fma(di[t], M[t * 6 + j], sm);
          dm[j] = sm;
        }
#pragma unroll
        for (int bp = 0; bp < 3; ++bp) {
          P2[(ar * 3 + bp) * NS] = make_double2(di[2 * bp], di[2 * bp + 1]);
          P2[(6 + ar * 3 + bp) * NS] = make_double2(dm[2 * bp], dm[2 * bp + 1]);
        }
      }
    }
    cta_sync();
    // Levels 3 (stride 8) and 4 (stride 16, fused with D'^-1) are applied as ONE operator in the solve: with r the level-2
    // result, kq the level-4 partner, A/G the level-3 alpha/gamma,
    //   y_k = Di_k r_k - Di_k A_k r_{k-8} - Di_k G_k r_{k+8} - DM_k (r_kq - A_kq r_{kq-8} - G_kq r_{kq+8}),
    // and all these stages lie on the chain c0 = k mod 8, c0 + 8, c0 + 16, c0 + 24.  W[p] (this thread's two rows) is the
    // coefficient of the chain's p-th stage; the 24 pairs replace the level-3 / level-4 entries of the PCR store.
    {
      const double2* const L3 = reinterpret_cast<const double2*>(m.PCR) + (kPcrLevels - 2) * (kPcrLevelDoubles / 2);
      double2* const L34 = reinterpret_cast<double2*>(m.PCR) + (kPcrLevels - 2) * (kPcrLevelDoubles / 2);
      double W[4][12];
      const int c0 = k & 7, pk = k >> 3;
      const int kq = k >= 16 ? k - 16 : (k + 16 <= N ? k + 16 : -1);
      if (warp < 3) {
#pragma unroll
        for (int p_ = 0; p_ < 4; ++p_)
#pragma unroll
          for (int e = 0; e < 12; ++e) W[p_][e] = 0.0;
        const double2* const P4 = L3 + (kPcrLevelDoubles / 2) + (warp * 12) * NS + k;
        double di[12], dm[12];
#pragma unroll
        for (int ar = 0; ar < 2; ++ar)
#pragma unroll
          for (int bp = 0; bp < 3; ++bp) {
            const double2 a_ = P4[(ar * 3 + bp) * NS], b_ = P4[(6 + ar * 3 + bp) * NS];
            di[ar * 6 + 2 * bp] = a_.x; di[ar * 6 + 2 * bp + 1] = a_.y; dm[ar * 6 + 2 * bp] = b_.x; dm[ar * 6 + 2 * bp + 1] = b_.y;
          }
        // full 6x6 level-3 matrix (alpha: g = 0, gamma: g = 6) of stage t from the store (rows spread over the three warps)
        auto mat3 = [&](int t, int g, double (&X)[36]) {
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int bp = 0; bp < 3; ++bp) {
              const double2 v = L3[((r >> 1) * 12 + g + (r & 1) * 3 + bp) * NS + t];
              X[r * 6 + 2 * bp] = v.x; X[r * 6 + 2 * bp + 1] = v.y;
            }
        };
        // W[p] += sgn * rows . X
        auto acc = [&](int p_, const double (&rows)[12], const double (&X)[36], double sgn) {
#pragma unroll
          for (int q_ = 0; q_ < 4; ++q_) if (q_ == p_) {
#pragma unroll
            for (int ar = 0; ar < 2; ++ar)
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                double sm_ = 0.0;
#pragma unroll
                for (int t = 0; t < 6; ++t) sm_ = fma(rows[ar * 6 + t], X[t * 6 + j], sm_);
                W[q_][ar * 6 + j] += sgn * sm_;
              }
          }
        };
        auto accv = [&](int p_, const double (&rows)[12], double sgn) {
#pragma unroll
          for (int q_ = 0; q_ < 4; ++q_) if (q_ == p_) {
#pragma unroll
            for (int e = 0; e < 12; ++e) W[q_][e] += sgn * rows[e];
          }
        };
        double X[36];
        accv(pk, di, 1.0);
        if (k - 8 >= 0) { mat3(k, 0, X); acc(pk - 1, di, X, -1.0); }
        if (k + 8 <= N) { mat3(k, 6, X); acc(pk + 1, di, X, -1.0); }
        if (kq >= 0) {
          const int pq = kq >> 3;
          accv(pq, dm, -1.0);
          if (kq - 8 >= 0) { mat3(kq, 0, X); acc(pq - 1, dm, X, 1.0); }
          if (kq + 8 <= N) { mat3(kq, 6, X); acc(pq + 1, dm, X, 1.0); }
        }
      }
      cta_sync();                                  // every thread has read what it needs of levels 3, 4
      if (warp < 3 && live) {
#pragma unroll
        for (int p_ = 0; p_ < 4; ++p_)
#pragma unroll
          for (int ar = 0; ar < 2; ++ar)
#pragma unroll
            for (int bp = 0; bp < 3; ++bp) {
              const int e = p_ * 6 + ar * 3 + bp;
              L34[(e / 12) * (kPcrLevelDoubles / 2) + (warp * 12 + e % 12) * NS + k] = make_double2(W[p_][ar * 6 + 2 * bp], W[p_][ar * 6 + 2 * bp + 1]);
            }
      }
      (void)c0;
    }
    cta_sync();
    MQ_T(11);
  }

  // CTA-wide reduction of NM maxima and NS_ sums; every thread ends with the same values (fixed combination order).
  template <int NM, int NS_> MQ_HD void cta_reduce(double (&mx)[NM], double (&sm)[NS_], int warp) {
    static_assert(4 * (NM + NS_) <= 6 * NST, "reduction scratch is the y buffer");
    // maxima: the values are non-negative and never NaN (see the callers' `up`), so IEEE order is the order of the bit
    // patterns: two integer warp reductions (high word, then low word among the lanes that hold the maximal high word)
#pragma unroll
    for (int e = 0; e < NM; ++e) {
      const unsigned hi = (unsigned)__double2hiint(mx[e]);
      const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned)__double2loint(mx[e]) : 0u);
      mx[e] = __hiloint2double((int)mh, (int)ml);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int e = 0; e < NS_; ++e) sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], o);
    }
    double* red = m.YB;
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < NM; ++e) red[warp * (NM + NS_) + e] = mx[e];
#pragma unroll
      for (int e = 0; e < NS_; ++e) red[warp * (NM + NS_) + NM + e] = sm[e];
    }
    cta_sync();
#pragma unroll
    for (int e = 0; e < NM; ++e) {
      const double a0 = red[e], a1 = red[(NM + NS_) + e], a2 = red[2 * (NM + NS_) + e], a3 = red[3 * (NM + NS_) + e];
      const double b0 = a0 > a1 ? a0 : a1, b1 = a2 > a3 ? a2 : a3;
      mx[e] = b0 > b1 ? b0 : b1;
    }
#pragma unroll
    for (int e = 0; e < NS_; ++e) sm[e] = (red[NM + e] + red[(NM + NS_) + NM + e]) + (red[2 * (NM + NS_) + NM + e] + red[3 * (NM + NS_) + NM + e]);
    cta_sync();
  }

  // The whole osqp_solve of one QP for one role (AX: axis warp `warp` in 0..2; !AX: slack/obstacle warp).  The
  // iterates stay in registers from the first iteration to the last; residuals, termination and the rho estimate are
  // evaluated from registers by all four warps (update_info / check_termination / compute_rho_estimate of auxil.h);
  // only a rho change (re-factorisation) or a suspected infeasibility certificate parks the state and runs the
  // single-warp array code on warp 0.
  template <bool AX> MQ_HD void solve_role(const int warp, volatile int* flag, volatile int* cmd) {
    constexpr int NVR = AX ? 3 : 4;
    // obstacle row o belongs to warp o % 4: in registers (NOW rows, run by this warp), or — the rows the helper runs: the slack
    // warp's, or everybody's (kHelpAll) — in shared memory (NHR rows; this warp only loads / parks / checks / rescales them)
    constexpr bool kSmemRows = kHelpAll || (kHelp && !AX);
    constexpr int NOW = kSmemRows ? 0 : (RC + 3) / 4;
    constexpr int NHR = kSmemRows ? (RC + 3) / 4 : 0;
    constexpr bool kRows = kWide || RC > 0;             // the QP has obstacle rows
    constexpr int RW = NOW > 0 ? NOW : 1;
    const int cc = warp;
    const int k = lane < NS ? lane : NS - 1;            // ghost lanes shadow the last stage, never write
    const bool live = lane < NS, hasu = lane < N, notfirst = lane > 0 && live;
    const int kp = k < N ? k + 1 : k, km = k > 0 ? k - 1 : 0;
    const double al = st.alpha, om = 1.0 - st.alpha, apv = sh.a_pv, bpa = sh.b_pa, bva = sh.b_va;
    double2* const RA2 = reinterpret_cast<double2*>(m.RA);      // [2][NS][3] pairs
    double2* const RS2 = reinterpret_cast<double2*>(m.RS);      // [NS][2] pairs: (x, y), (z, -)
    double2* const YB2 = reinterpret_cast<double2*>(m.YB);      // [NS][3] pairs
    double* const XR = m.YB + 6 * NS;                           // [2][NS]   r11 of the slack-input elimination
    double* const T4 = XR + 2 * NS;                             // [R][4][NS] obstacle rows: (g0 t, g1 t, g2 t, t)
    // wide mode: the rows of a warp (o = warp, warp + 4, ...) are kept in shared memory and each warp hands over only
    // its partial sums TP[warp][5] = (sum w0 t, sum w1 t, sum w2 t, sum t | sigma_d, sum t | sigma_s)
    double* const TP = T4;
    double* const ZO = m.ROW; double* const UO = ZO + (kWide ? R * NS : 0); double* const ORH = UO + (kWide ? R * NS : 0);
    double* const OLO = ORH + (kWide ? R * NS : 0); double* const OG3 = OLO + (kWide ? R * NS : 0);
    const double2* const M = reinterpret_cast<const double2*>(m.PCR) + ((AX ? cc : 0) * 12) * NS + k;
    const int kq = k >= 16 ? k - 16 : (k + 16 <= N ? k + 16 : k);
    auto vj = [&](int e) { return AX ? (e == 0 ? cc : (e == 1 ? 3 + cc : 8 + cc)) : (e < 2 ? 6 + e : 9 + e); };
    auto di = [&](int t) { return AX ? cc + 3 * t : 6 + t; };

    double x[NVR], zb[NVR], ub[NVR], b[NVR], rhb[NVR], sd[NVR], cq[NVR], lo[AX ? 3 : 1], hi[AX ? 3 : 1];
    double zd[2], ud[2], rhd[2], bnd[2];
    // box bounds: the axis role's variable indices depend on the warp (registers); the slack role's are compile-time
    auto blo = [&](int e) { if constexpr (AX) return lo[e]; else return sh.blo[e < 2 ? 6 + e : 9 + e]; };
    auto bhi = [&](int e) { if constexpr (AX) return hi[e]; else return sh.bhi[e < 2 ? 6 + e : 9 + e]; };
    double dai = 0, cv0 = 0, cv1 = 0, cv2 = 0, cv3 = 0, cv2m = 0, cv3m = 0, ogp = 0;        // AX (ogp: parked obstacle part of b_p)
    double dsi[2], esd[2], esdn[2], dgi[2], fs[6];                                              // !AX (dgi, fs: both roles in wide mode)
    // row helper's rows: z, u, Rh, low, gradient (3), dgi and fs (3) of the row's slack type, the type itself
    auto HR = [&](int o, int slot) -> double& { return m.ROW[(o * kHelpSlots + slot) * NS + k]; };
    unsigned slmask = 0;                                // !AX: bit o set when row o is softened by sigma_s (slack input 4)
    // owned obstacle rows (both roles): row o = 4 q + warp
    double zo[RW], uo[RW], orh[RW], og3[3 * RW], olo[RW], odg[RW], ofs[3 * RW];
    int osl[RW]; bool oex[RW];

    // b (slack inputs) stays WITHOUT the obstacle sums in registers; ps holds the sums of the last completed iteration
    // (the position rows carry the matching part in ogp), so a burst can resume where the previous one stopped.
    double r11[2] = {0.0, 0.0}, ps[2] = {0.0, 0.0};

    auto load = [&]() {
#pragma unroll
      for (int e = 0; e < NVR; ++e) {
        const int j = vj(e);
        x[e] = X_(j, k); zb[e] = Z_(8 + j, k); ub[e] = U_(8 + j, k); b[e] = B_(j, k); rhb[e] = RH_(8 + j, k);
        sd[e] = SD_(j, k); cq[e] = CQ_(j, k);
        if constexpr (AX) { lo[e] = box_lo(j); hi[e] = box_hi(j); }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = di(t);
        zd[t] = Z_(i, k); ud[t] = U_(i, k); rhd[t] = RH_(i, k); bnd[t] = (k == 0) ? -x0p[i] : 0.0;
      }
      if constexpr (AX) {
        dai = DAI_(cc, k); cv0 = CV_(4 * cc, k); cv1 = CV_(4 * cc + 1, k); cv2 = CV_(4 * cc + 2, k); cv3 = CV_(4 * cc + 3, k);
        cv2m = notfirst ? CV_(4 * cc + 2, km) : 0.0; cv3m = notfirst ? CV_(4 * cc + 3, km) : 0.0;
        ogp = OG_(cc, k);
        if constexpr (kWide) {
#pragma unroll
          for (int t = 0; t < 2; ++t) dgi[t] = DGI_(t, k);
#pragma unroll
          for (int e = 0; e < 6; ++e) fs[e] = FS_(e, k);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 2; ++t) { dsi[t] = DSI_(t, k); esd[t] = notfirst ? ESD_(t, k) : 0.0; esdn[t] = ESD_(t, kp); dgi[t] = DGI_(t, k); }
#pragma unroll
        for (int e = 0; e < 6; ++e) fs[e] = FS_(e, k);
        ps[0] = ps[1] = 0.0;                          // b comes back complete from the array code
        if constexpr (!kWide) {
          slmask = 0;
#pragma unroll
          for (int o = 0; o < R; ++o) slmask |= (hasu && SLK_(o, k)) ? (1u << o) : 0u;
        }
      }
      if constexpr (kWide) {
        slmask = 0;
        for (int o = 0; o < R; ++o) slmask |= (hasu && SLK_(o, k)) ? (1u << o) : 0u;
        if (live) for (int o = warp; o < R; o += 4) {
          ZO[o * NS + k] = Z_(NBR + o, k); UO[o * NS + k] = U_(NBR + o, k); ORH[o * NS + k] = RH_(NBR + o, k); OLO[o * NS + k] = LO_(o, k);
          OG3[(3 * o) * NS + k] = G3_(3 * o, k); OG3[(3 * o + 1) * NS + k] = G3_(3 * o + 1, k); OG3[(3 * o + 2) * NS + k] = G3_(3 * o + 2, k);
        }
      }
#pragma unroll
      for (int q = 0; q < NOW; ++q) {
        const int o = 4 * q + warp;
        oex[q] = o < R && hasu;
        const int oo = o < R ? o : 0;
        zo[q] = Z_(NBR + oo, k); uo[q] = U_(NBR + oo, k); orh[q] = RH_(NBR + oo, k); olo[q] = LO_(oo, k);
        og3[3 * q] = G3_(3 * oo, k); og3[3 * q + 1] = G3_(3 * oo + 1, k); og3[3 * q + 2] = G3_(3 * oo + 2, k);
        osl[q] = hasu ? SLK_(oo, k) : 0;
        odg[q] = DGI_(osl[q], k); ofs[3 * q] = FS_(3 * osl[q], k); ofs[3 * q + 1] = FS_(3 * osl[q] + 1, k); ofs[3 * q + 2] = FS_(3 * osl[q] + 2, k);
      }
#pragma unroll
      for (int h = 0; h < NHR; ++h) {                   // the row helper's rows: same values, in shared memory
        const int o = 4 * h + warp;
        if (o < R && live) {
          const int sl = hasu ? SLK_(o, k) : 0;
          HR(o, 0) = Z_(NBR + o, k); HR(o, 1) = U_(NBR + o, k); HR(o, 2) = RH_(NBR + o, k); HR(o, 3) = LO_(o, k);
          HR(o, 4) = G3_(3 * o, k); HR(o, 5) = G3_(3 * o + 1, k); HR(o, 6) = G3_(3 * o + 2, k);
          HR(o, 7) = DGI_(sl, k); HR(o, 8) = FS_(3 * sl, k); HR(o, 9) = FS_(3 * sl + 1, k); HR(o, 10) = FS_(3 * sl + 2, k);
          HR(o, 11) = (double)sl;
        }
      }
    };
    // iterates (+ the deltas of the last iteration) back to the arrays the single-warp code and store() read
    auto park = [&]() {
      if (live) {
#pragma unroll
        for (int e = 0; e < NVR; ++e) {
          const int j = vj(e);
          X_(j, k) = x[e]; Z_(8 + j, k) = zb[e]; U_(8 + j, k) = ub[e]; B_(j, k) = b[e];
          WSDX_(j, k) = x[e] - OX_(j, k); WSDY_(8 + j, k) = rhb[e] * (ub[e] - OU_(8 + j, k));
          if (e < 2 || hasu) RH_(8 + j, k) = rhb[e];
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) { const int i = di(t); Z_(i, k) = zd[t]; U_(i, k) = ud[t]; WSDY_(i, k) = rhd[t] * (ud[t] - OU_(i, k)); RH_(i, k) = rhd[t]; }
        if constexpr (AX) OG_(cc, k) = ogp;
#pragma unroll
        for (int q = 0; q < NOW; ++q) {
          const int o = 4 * q + warp;
          if (o < R) { Z_(NBR + o, k) = zo[q]; U_(NBR + o, k) = uo[q]; WSDY_(NBR + o, k) = orh[q] * (uo[q] - OU_(NBR + o, k)); if (hasu) RH_(NBR + o, k) = orh[q]; }
        }
        if constexpr (kWide) {
          for (int o = warp; o < R; o += 4) {
            const double zv = ZO[o * NS + k], uv = UO[o * NS + k], rv = ORH[o * NS + k];
            Z_(NBR + o, k) = zv; U_(NBR + o, k) = uv; WSDY_(NBR + o, k) = rv * (uv - OU_(NBR + o, k)); if (hasu) RH_(NBR + o, k) = rv;
          }
        }
#pragma unroll
        for (int h = 0; h < NHR; ++h) {
          const int o = 4 * h + warp;
          if (o < R) {
            const double zv = HR(o, 0), uv = HR(o, 1), rv = HR(o, 2);
            Z_(NBR + o, k) = zv; U_(NBR + o, k) = uv; WSDY_(NBR + o, k) = rv * (uv - OU_(NBR + o, k)); if (hasu) RH_(NBR + o, k) = rv;
          }
        }
      }
    };

    // ---- one burst of ADMM iterations (auxil.h:67-112) in registers; the last one records delta_x, delta_u
    // Slack role: r11 (rhs of the slack-input elimination) and the position-row terms rs that it produces are formed at
    // the END of an iteration from the part of the rhs that does not depend on this iteration's obstacle rows, so that
    // the axis warps never wait for them; the obstacle rows' part reaches the position rows through T4 (the row owner
    // folds the slack elimination into what it writes) and r11 itself once all rows are in (after the next barrier).
    auto slack_forward = [&]() {                      // !AX: r11, rs from b -> RS2
      if constexpr (!AX) {
        double f[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const double bn = dn1(b[t]);
          const double v = b[2 + t] - esdn[t] * bn;
          r11[t] = hasu ? v : 0.0;
          f[t] = dgi[t] * r11[t];
        }
        const double rs0 = -fs[0] * f[0] - fs[3] * f[1], rs1 = -fs[1] * f[0] - fs[4] * f[1], rs2 = -fs[2] * f[0] - fs[5] * f[1];
        if (live) { RS2[k * 2] = make_double2(rs0, rs1); RS2[k * 2 + 1] = make_double2(rs2, 0.0); }
      }
    };
    auto slack_obstacle_sums = [&]() {                // !AX: the obstacle rows' part of b (slack inputs) and of r11
      if constexpr (!AX && kRows) {
        double s0 = 0.0, s1 = 0.0;                    // sum of t over the rows softened by sigma_d / sigma_s
        if constexpr (kWide) {
#pragma unroll
          for (int w = 0; w < 4; ++w) { s0 += TP[(5 * w + 3) * NS + k]; s1 += TP[(5 * w + 4) * NS + k]; }
        } else {
#pragma unroll
          for (int o = 0; o < R; ++o) { const double t = T4[(4 * o + 3) * NS + k]; if ((slmask >> o) & 1u) s1 += t; else s0 += t; }
        }
        ps[0] = hasu ? s0 : 0.0; ps[1] = hasu ? s1 : 0.0;
      }
    };
    // ---- owned obstacle rows (both roles) of one iteration: from the positions yp of this stage to the terms T4 / TP that the
    // position rows and the slack warp pick up.  Needs only yp and r11 (XR).
    auto obstacle_rows = [&](const double yp0, const double yp1, const double yp2) {
      // (MP.cpp:1040-1071): grad . p_k - slack  >=  low.  The slack input of the row's type is re-derived from r11 (4 flops)
      // instead of being fetched from warp 3.
#pragma unroll
      for (int q = 0; q < NOW; ++q) {
        const int o = 4 * q + warp;
        if (o < R) {
          const double r11s = XR[osl[q] * NS + k];
          const double sg = oex[q] ? odg[q] * (r11s - ofs[3 * q] * yp0 - ofs[3 * q + 1] * yp1 - ofs[3 * q + 2] * yp2) : 0.0;
          const double zt = og3[3 * q] * yp0 + og3[3 * q + 1] * yp1 + og3[3 * q + 2] * yp2 - sg;
          const double v = al * zt + om * zo[q] + uo[q];
          const double zn = v > olo[q] ? v : olo[q];
          zo[q] = zn; uo[q] = v - zn;
          const double t = orh[q] * (zn - uo[q]);
          // position rows get grad' t directly and, through the eliminated slack input of the row, fs dg t
          const double ts_ = oex[q] ? odg[q] * t : 0.0;
          if (live) {
            T4[(4 * o) * NS + k] = fma(ofs[3 * q], ts_, og3[3 * q] * t); T4[(4 * o + 1) * NS + k] = fma(ofs[3 * q + 1], ts_, og3[3 * q + 1] * t);
            T4[(4 * o + 2) * NS + k] = fma(ofs[3 * q + 2], ts_, og3[3 * q + 2] * t); T4[(4 * o + 3) * NS + k] = t;
          }
        }
      }
      if constexpr (kWide) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, as0 = 0.0, as1 = 0.0;
        const double r11d = XR[k], r11st = XR[NS + k];
        const double sgd = hasu ? dgi[0] * (r11d - fs[0] * yp0 - fs[1] * yp1 - fs[2] * yp2) : 0.0;     // slack inputs of this stage
        const double sgs = hasu ? dgi[1] * (r11st - fs[3] * yp0 - fs[4] * yp1 - fs[5] * yp2) : 0.0;
#pragma unroll 4
        for (int o = warp; o < R; o += 4) {
          const bool st_ = (slmask >> o) & 1u;
          const double g0 = OG3[(3 * o) * NS + k], g1 = OG3[(3 * o + 1) * NS + k], g2 = OG3[(3 * o + 2) * NS + k];
          const double zv = ZO[o * NS + k], uv = UO[o * NS + k], rv = ORH[o * NS + k], lv = OLO[o * NS + k];
          const double zt = g0 * yp0 + g1 * yp1 + g2 * yp2 - (st_ ? sgs : sgd);
          const double v = al * zt + om * zv + uv;
          const double zn = v > lv ? v : lv;
          const double un = v - zn;
          if (live) { ZO[o * NS + k] = zn; UO[o * NS + k] = un; }
          const double t = rv * (zn - un);
          const double ts_ = hasu ? (st_ ? dgi[1] : dgi[0]) * t : 0.0;
          a0 += fma(st_ ? fs[3] : fs[0], ts_, g0 * t); a1 += fma(st_ ? fs[4] : fs[1], ts_, g1 * t); a2 += fma(st_ ? fs[5] : fs[2], ts_, g2 * t);
          if (st_) as1 += t; else as0 += t;
        }
        if (live) {
          TP[(5 * warp) * NS + k] = a0; TP[(5 * warp + 1) * NS + k] = a1; TP[(5 * warp + 2) * NS + k] = a2;
          TP[(5 * warp + 3) * NS + k] = as0; TP[(5 * warp + 4) * NS + k] = as1;
        }
      }
    };
    // u of the helper's rows before the last iteration of a burst (delta_y of the certificates); called behind the first barrier
    // of the iteration: the helper's update of the previous iteration is ordered before it, the next one comes after kBarY
    auto snapshot_helper_rows = [&](const bool lastit) {
      if constexpr (NHR > 0) {
        if (lastit && live) {
#pragma unroll
          for (int h = 0; h < NHR; ++h) { const int o = 4 * h + warp; if (o < R) OU_(NBR + o, k) = HR(o, 1); }
        }
      }
    };
    auto iterate = [&](const int niter) {
      if constexpr (!AX) { slack_forward(); bar_arrive(kBarRS, 128); r11[0] -= ps[0]; r11[1] -= ps[1]; if (live) { XR[k] = r11[0]; XR[NS + k] = r11[1]; } }
      for (int left = niter; left > 0; --left) {
        const bool lastit = left == 1;
        const int it = niter - left;
        if (lastit && live) {                       // iterate k-1 for the infeasibility certificates (delta_x, delta_y)
#pragma unroll
          for (int e = 0; e < NVR; ++e) { const int j = vj(e); OX_(j, k) = x[e]; OU_(8 + j, k) = ub[e]; }
          OU_(di(0), k) = ud[0]; OU_(di(1), k) = ud[1];
#pragma unroll
          for (int q = 0; q < NOW; ++q) { const int o = 4 * q + warp; if (o < R) OU_(NBR + o, k) = uo[q]; }
          if constexpr (kWide) { for (int o = warp; o < R; o += 4) OU_(NBR + o, k) = UO[o * NS + k]; }
        }
        double xt[NVR], td[2], racc[NVR], yp0, yp1, yp2;
        if constexpr (AX) {
          // level-0 matrices are fetched before anything else so that they are in flight across the first barrier
          double2 mt[12];
#pragma unroll
          for (int h = 0; h < 12; ++h) mt[h] = M[h * NS];
          // ---- leaf forward (this axis' acceleration): reduced rhs rows (p, v); the slack-input part comes from warp 3
          double r0, r1;
          {
            const double ma = dai * b[2];
            const double mm = up1(ma);
            r0 = (b[0] + ogp) - cv0 * ma; r1 = b[1] - cv1 * ma;
            r0 -= cv2m * mm; r1 -= cv3m * mm;             // cv2m = cv3m = 0 on stage 0
          }
          // the slack warp's part of this stage's position row (written at the end of its previous iteration, long ago): folded
          // in BEFORE the exchange, so that level 0 reads complete neighbour vectors
          bar_sync(kBarRS, 128);
          r0 += m.RS[k * 4 + cc];
          if (live) RA2[k * 3 + cc] = make_double2(r0, r1);
          bar_sync(kBarAll, 128);
          snapshot_helper_rows(lastit);
          // ---- PCR levels 0..2 (with assistants: level 0 only; they run levels 1, 2 and the final operator from registers)
          constexpr int kRowLevels = ASSIST ? 1 : kPcrLevels - 2;
#pragma unroll
          for (int l = 0; l < kRowLevels; ++l) {
            const int s = 1 << l, cur = l & 1;
            const int kmm = k - s >= 0 ? k - s : k, kpp = k + s <= N ? k + s : k;
            const double2* ra = RA2 + cur * 3 * NS;
            double2 a0 = ra[kmm * 3], a1 = ra[kmm * 3 + 1], a2 = ra[kmm * 3 + 2];
            double2 c0 = ra[kpp * 3], c1 = ra[kpp * 3 + 1], c2 = ra[kpp * 3 + 2];
            double sa0 = mt[0].x * a0.x, sa1 = mt[3].x * a0.x, sg0 = mt[6].x * c0.x, sg1 = mt[9].x * c0.x;
            sa0 = fma(mt[0].y, a0.y, sa0); sa1 = fma(mt[3].y, a0.y, sa1); sg0 = fma(mt[6].y, c0.y, sg0); sg1 = fma(mt[9].y, c0.y, sg1);
            sa0 = fma(mt[1].x, a1.x, sa0); sa1 = fma(mt[4].x, a1.x, sa1); sg0 = fma(mt[7].x, c1.x, sg0); sg1 = fma(mt[10].x, c1.x, sg1);
            sa0 = fma(mt[1].y, a1.y, sa0); sa1 = fma(mt[4].y, a1.y, sa1); sg0 = fma(mt[7].y, c1.y, sg0); sg1 = fma(mt[10].y, c1.y, sg1);
            sa0 = fma(mt[2].x, a2.x, sa0); sa1 = fma(mt[5].x, a2.x, sa1); sg0 = fma(mt[8].x, c2.x, sg0); sg1 = fma(mt[11].x, c2.x, sg1);
            sa0 = fma(mt[2].y, a2.y, sa0); sa1 = fma(mt[5].y, a2.y, sa1); sg0 = fma(mt[8].y, c2.y, sg0); sg1 = fma(mt[11].y, c2.y, sg1);
            r0 = (r0 - sa0) - sg0; r1 = (r1 - sa1) - sg1;
            if constexpr (ASSIST) {
              if (live) RA2[(1 - cur) * 3 * NS + k * 3 + cc] = make_double2(r0, r1);
              bar_arrive(kBarH1, 192);                     // hand over to the assistants: they deliver y
            } else {
              // next level's matrices: issued before the barrier, consumed after it
              const double2* mn = M + (l + 1) * (kPcrLevelDoubles / 2);
#pragma unroll
              for (int h = 0; h < 12; ++h) mt[h] = mn[h * NS];
              if (live) RA2[(1 - cur) * 3 * NS + k * 3 + cc] = make_double2(r0, r1);
              bar_sync(kBarAxis, 96);
            }
          }
          // ---- levels 3 and 4 as one operator over the stride-8 chain of this stage (pcr_factor_cta): y = sum_p W_p r_{c0+8p},
          // r = the level-2 result (buffer 1); the second half of W is fetched while the first is applied
          double y0, y1;
          if constexpr (!ASSIST) {
            const double2* const rb = RA2 + 3 * NS;
            const int c0 = k & 7;
            const int t0 = c0, t1 = c0 + 8, t2 = c0 + 16 <= N ? c0 + 16 : c0, t3 = c0 + 24 <= N ? c0 + 24 : c0;   // (W_p = 0 where the chain is shorter)
            double2 a0 = rb[t0 * 3], a1 = rb[t0 * 3 + 1], a2 = rb[t0 * 3 + 2];
            double2 c0_ = rb[t1 * 3], c1 = rb[t1 * 3 + 1], c2 = rb[t1 * 3 + 2];
            const double2* mn = M + (kPcrLevels - 1) * (kPcrLevelDoubles / 2);
            double2 m2[12];
#pragma unroll
            for (int h = 0; h < 12; ++h) m2[h] = mn[h * NS];
            double sa0 = mt[0].x * a0.x, sa1 = mt[3].x * a0.x, sg0 = mt[6].x * c0_.x, sg1 = mt[9].x * c0_.x;
            sa0 = fma(mt[0].y, a0.y, sa0); sa1 = fma(mt[3].y, a0.y, sa1); sg0 = fma(mt[6].y, c0_.y, sg0); sg1 = fma(mt[9].y, c0_.y, sg1);
            sa0 = fma(mt[1].x, a1.x, sa0); sa1 = fma(mt[4].x, a1.x, sa1); sg0 = fma(mt[7].x, c1.x, sg0); sg1 = fma(mt[10].x, c1.x, sg1);
            sa0 = fma(mt[1].y, a1.y, sa0); sa1 = fma(mt[4].y, a1.y, sa1); sg0 = fma(mt[7].y, c1.y, sg0); sg1 = fma(mt[10].y, c1.y, sg1);
            sa0 = fma(mt[2].x, a2.x, sa0); sa1 = fma(mt[5].x, a2.x, sa1); sg0 = fma(mt[8].x, c2.x, sg0); sg1 = fma(mt[11].x, c2.x, sg1);
            sa0 = fma(mt[2].y, a2.y, sa0); sa1 = fma(mt[5].y, a2.y, sa1); sg0 = fma(mt[8].y, c2.y, sg0); sg1 = fma(mt[11].y, c2.y, sg1);
            a0 = rb[t2 * 3]; a1 = rb[t2 * 3 + 1]; a2 = rb[t2 * 3 + 2];
            c0_ = rb[t3 * 3]; c1 = rb[t3 * 3 + 1]; c2 = rb[t3 * 3 + 2];
            sa0 = fma(m2[0].x, a0.x, sa0); sa1 = fma(m2[3].x, a0.x, sa1); sg0 = fma(m2[6].x, c0_.x, sg0); sg1 = fma(m2[9].x, c0_.x, sg1);
            sa0 = fma(m2[0].y, a0.y, sa0); sa1 = fma(m2[3].y, a0.y, sa1); sg0 = fma(m2[6].y, c0_.y, sg0); sg1 = fma(m2[9].y, c0_.y, sg1);
            sa0 = fma(m2[1].x, a1.x, sa0); sa1 = fma(m2[4].x, a1.x, sa1); sg0 = fma(m2[7].x, c1.x, sg0); sg1 = fma(m2[10].x, c1.x, sg1);
            sa0 = fma(m2[1].y, a1.y, sa0); sa1 = fma(m2[4].y, a1.y, sa1); sg0 = fma(m2[7].y, c1.y, sg0); sg1 = fma(m2[10].y, c1.y, sg1);
            sa0 = fma(m2[2].x, a2.x, sa0); sa1 = fma(m2[5].x, a2.x, sa1); sg0 = fma(m2[8].x, c2.x, sg0); sg1 = fma(m2[11].x, c2.x, sg1);
            sa0 = fma(m2[2].y, a2.y, sa0); sa1 = fma(m2[5].y, a2.y, sa1); sg0 = fma(m2[8].y, c2.y, sg0); sg1 = fma(m2[11].y, c2.y, sg1);
            y0 = sa0 + sg0; y1 = sa1 + sg1;
            if (live) YB2[k * 3 + cc] = make_double2(y0, y1);
          }
          bar_sync(kBarY, kCtaThreads);
          if constexpr (ASSIST) { const double2 yv = YB2[k * 3 + cc]; y0 = yv.x; y1 = yv.y; }
          if constexpr (kRows) { yp0 = m.YB[k * 6]; yp1 = m.YB[k * 6 + 2]; yp2 = m.YB[k * 6 + 4]; }
          // ---- leaf backward: acceleration of this axis
          xt[0] = y0; xt[1] = y1;
          {
            double yn0, yn1;                              // y of stage k+1
            if constexpr (ASSIST) { const double2 yn = YB2[kp * 3 + cc]; yn0 = yn.x; yn1 = yn.y; }   // (same values as a shuffle, one round trip fewer)
            else { yn0 = dn1(y0); yn1 = dn1(y1); }
            const double a = dai * (b[2] - cv0 * y0 - cv1 * y1 - cv2 * yn0 - cv3 * yn1);
            xt[2] = hasu ? a : 0.0;
          }
          // ---- dynamics rows (p, v) of this stage: Ad x~_{k-1} + Bd u~_{k-1} - x~_k, projected onto the equality
          {
            const double pp = up1(xt[0] + apv * xt[1] + bpa * xt[2]), pv = up1(xt[1] + bva * xt[2]);
            const double zt0 = (notfirst ? pp : 0.0) - xt[0], zt1 = (notfirst ? pv : 0.0) - xt[1];
            const double v0 = al * zt0 + om * zd[0] + ud[0], v1 = al * zt1 + om * zd[1] + ud[1];
            zd[0] = bnd[0]; zd[1] = bnd[1];
            ud[0] = v0 - bnd[0]; ud[1] = v1 - bnd[1];
            td[0] = rhd[0] * (bnd[0] - ud[0]); td[1] = rhd[1] * (bnd[1] - ud[1]);
            racc[0] = -td[0]; racc[1] = -td[1]; racc[2] = 0.0;
          }
        } else {
          // ---- leaf forward, slack part (eliminate s_{k+1,t} then sigma_{k,t}): RS2 was written at the end of the last
          // iteration (or by the prologue); once every warp is past its obstacle rows, their part is added to r11
          bar_sync(kBarAll, 128);
          snapshot_helper_rows(lastit);
          if (it > 0) { slack_obstacle_sums(); r11[0] -= ps[0]; r11[1] -= ps[1]; if (live) { XR[k] = r11[0]; XR[NS + k] = r11[1]; } }
          bar_sync(kBarY, kCtaThreads);
          yp0 = m.YB[k * 6]; yp1 = m.YB[k * 6 + 2]; yp2 = m.YB[k * 6 + 4];
          // ---- leaf backward: slack inputs, then slack states
          double xp[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double v = dgi[t] * (r11[t] - fs[3 * t] * yp0 - fs[3 * t + 1] * yp1 - fs[3 * t + 2] * yp2);
            xt[2 + t] = hasu ? v : 0.0;
            const double q = up1(xt[2 + t]);
            xp[t] = notfirst ? q : 0.0;
            xt[t] = dsi[t] * b[t] - esd[t] * xp[t];
          }
          // ---- dynamics rows 6, 7:  sigma~_{k-1,t} - s~_{k,t}
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double zt = xp[t] - xt[t];
            const double v = al * zt + om * zd[t] + ud[t];
            zd[t] = bnd[t]; ud[t] = v - bnd[t];
            td[t] = rhd[t] * (bnd[t] - ud[t]);
            racc[t] = -td[t]; racc[2 + t] = 0.0;
          }
        }
        if constexpr (!kSmemRows) obstacle_rows(yp0, yp1, yp2);
        // ---- box rows
#pragma unroll
        for (int e = 0; e < NVR; ++e) {
          const double v = al * xt[e] + om * zb[e] + ub[e];
          const double zn = clampd(v, blo(e), bhi(e));
          zb[e] = zn; ub[e] = v - zn;
          racc[e] += rhb[e] * (zn - ub[e]);
        }
        // ---- x update (auxil.h:83), next right-hand side
#pragma unroll
        for (int e = 0; e < NVR; ++e) {
          x[e] = al * xt[e] + om * x[e];
          b[e] = (e < 2 || hasu) ? racc[e] + sd[e] * x[e] - cq[e] : 0.0;
        }
        {
          const double t0 = dn1(td[0]), t1 = dn1(td[1]);
          if (hasu) {
            if constexpr (AX) { b[0] += t0; b[1] += apv * t0 + t1; b[2] += bpa * t0 + bva * t1; }
            else { b[2] += t0; b[3] += t1; }
          }
        }
        // ---- obstacle rows' contributions to the next right-hand side: positions (axis warps) now; the slack warp only
        // announces its own row and goes on to the next iteration's r11 / rs (its sums follow after the next barrier)
        if constexpr (AX) {
          if constexpr (kRows) {
            bar_sync(kBarT, 128);
            double sgo = 0.0;
            if constexpr (kWide) {
#pragma unroll
              for (int w = 0; w < 4; ++w) sgo += TP[(5 * w + cc) * NS + k];
            } else {
#pragma unroll
              for (int o = 0; o < R; ++o) sgo += T4[(4 * o + cc) * NS + k];
            }
            ogp = hasu ? sgo : 0.0;
          }
        } else {
          // (with a row helper the slack warp owns no rows and the helper announces them)
          if constexpr (kRows && !kHelp) bar_arrive(kBarT, 128);
          slack_forward();
          if (!lastit) bar_arrive(kBarRS, 128);         // (the next burst's prologue announces its own)
        }
      }
      // burst end: all obstacle rows of the last iteration are in; the slack warp completes its rhs
      if constexpr (kRows) { cta_sync(); slack_obstacle_sums(); }
    };

    // screening values for the infeasibility certificates (filled by info())
    double s_ndy = 0.0, s_lhs = 0.0, s_ndx = 0.0, s_qd = 0.0, s_pm = 0.0;
    // ---- update_info (auxil.h:154) from registers, all warps; every thread ends with identical scalars
    auto info = [&](const int iter) {
      // cold per-row / per-variable data (E, D, P, previous iterate) first, so that the loads overlap the exchange
      double er_[NVR], ed_[2], eo_[RW], dv_[NVR], pv_[NVR], oxv[NVR], oub_[NVR], oud_[2], ouo_[RW];
#pragma unroll
      for (int e = 0; e < NVR; ++e) {
        const int j = vj(e);
        er_[e] = WSE_(8 + j, k); dv_[e] = WSD_(j, k); pv_[e] = pd[k * NV + j]; oxv[e] = OX_(j, k); oub_[e] = OU_(8 + j, k);
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) { ed_[t] = WSE_(di(t), k); oud_[t] = OU_(di(t), k); }
#pragma unroll
      for (int q = 0; q < NOW; ++q) { const int o = 4 * q + warp, oo = o < R ? o : 0; eo_[q] = WSE_(NBR + oo, k); ouo_[q] = OU_(NBR + oo, k); }
      double eh_[NHR > 0 ? NHR : 1], ouh_[NHR > 0 ? NHR : 1];
#pragma unroll
      for (int h = 0; h < NHR; ++h) { const int o = 4 * h + warp, oo = o < R ? o : 0; eh_[h] = WSE_(NBR + oo, k); ouh_[h] = OU_(NBR + oo, k); }
      // exchange: positions for the obstacle rows' A x; obstacle multipliers for the position / slack-input columns' A' y.
      // The buffers are the iteration's: wait until every warp has consumed the last iteration's obstacle terms.
      cta_sync();
      if constexpr (AX) { if (live) RA2[k * 3 + cc] = make_double2(x[0], x[1]); }
      else { if (live) { XR[k] = x[2]; XR[NS + k] = x[3]; } }
#pragma unroll
      for (int q = 0; q < NOW; ++q) {
        const int o = 4 * q + warp;
        if (o < R && live) {
          const double yo = orh[q] * uo[q];
          T4[(4 * o) * NS + k] = og3[3 * q] * yo; T4[(4 * o + 1) * NS + k] = og3[3 * q + 1] * yo; T4[(4 * o + 2) * NS + k] = og3[3 * q + 2] * yo;
          T4[(4 * o + 3) * NS + k] = yo;
        }
      }
#pragma unroll
      for (int h = 0; h < NHR; ++h) {
        const int o = 4 * h + warp;
        if (o < R && live) {
          const double yo = HR(o, 2) * HR(o, 1);
          T4[(4 * o) * NS + k] = HR(o, 4) * yo; T4[(4 * o + 1) * NS + k] = HR(o, 5) * yo; T4[(4 * o + 2) * NS + k] = HR(o, 6) * yo;
          T4[(4 * o + 3) * NS + k] = yo;
        }
      }
      if constexpr (kWide) {                     // partial sums of A' y over this warp's rows
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, as0 = 0.0, as1 = 0.0;
        for (int o = warp; o < R; o += 4) {
          const double yo = ORH[o * NS + k] * UO[o * NS + k];
          a0 += OG3[(3 * o) * NS + k] * yo; a1 += OG3[(3 * o + 1) * NS + k] * yo; a2 += OG3[(3 * o + 2) * NS + k] * yo;
          if ((slmask >> o) & 1u) as1 += yo; else as0 += yo;
        }
        if (live) {
          TP[(5 * warp) * NS + k] = a0; TP[(5 * warp + 1) * NS + k] = a1; TP[(5 * warp + 2) * NS + k] = a2;
          TP[(5 * warp + 3) * NS + k] = as0; TP[(5 * warp + 4) * NS + k] = as1;
        }
      }
      cta_sync();
      double mx[15], sm[3];
#pragma unroll
      for (int e = 0; e < 15; ++e) mx[e] = 0.0;
      sm[0] = sm[1] = sm[2] = 0.0;
      auto up = [&](double& mref, double v) { mref = v > mref ? v : mref; };      // max of non-negative, non-NaN values
      auto row = [&](double ax, double z, double e) {
        const double d = fabs(ax - z);
        up(mx[0], d); up(mx[1], fabs(ax)); up(mx[2], fabs(z));
        up(mx[3], e * d); up(mx[4], e * fabs(ax)); up(mx[5], e * fabs(z));
      };
      auto var = [&](double xv, double p, double cqv, double aty, double dj, double dxv) {
        const double px = c * p * xv;
        const double d = fabs(px + cqv + aty);
        up(mx[6], d); up(mx[7], fabs(px)); up(mx[8], fabs(aty));
        up(mx[9], dj * d); up(mx[10], dj * fabs(px)); up(mx[11], dj * fabs(aty));
        sm[0] += (0.5 * p * xv + cqv * cinv) * xv;
        // is_dual_infeasible screening (auxil.h:148): |dx|_inf, q'dx, |P dx|_inf
        up(mx[13], fabs(dxv)); up(mx[14], fabs(c * p * dxv));
        sm[2] += cqv * dxv;
      };
      // is_primal_infeasible screening (auxil.h:137): projected delta_y, its norm and the support-function value
      auto cert = [&](double dy, double e, double lo_, double hi_) {
        const double ls = e * lo_, us = e * hi_;
        if (us > kInfty * kMinScaling) { if (ls < -kInfty * kMinScaling) dy = 0.0; else dy = dy < 0.0 ? dy : 0.0; }
        else if (ls < -kInfty * kMinScaling) dy = dy > 0.0 ? dy : 0.0;
        up(mx[12], fabs(dy));
        sm[1] += hi_ * (dy > 0.0 ? dy : 0.0) + lo_ * (dy < 0.0 ? dy : 0.0);
      };
      const double yd0 = rhd[0] * ud[0], yd1 = rhd[1] * ud[1];
      const double yn0 = dn1(yd0), yn1 = dn1(yd1);
      double aty[NVR];
      if constexpr (AX) {
        const double pp = up1(x[0] + apv * x[1] + bpa * x[2]), pv = up1(x[1] + bva * x[2]);
        row((notfirst ? pp : 0.0) - x[0], zd[0], ed_[0]); row((notfirst ? pv : 0.0) - x[1], zd[1], ed_[1]);
        aty[0] = rhb[0] * ub[0] - yd0; aty[1] = rhb[1] * ub[1] - yd1; aty[2] = rhb[2] * ub[2];
        if (hasu) {
          double gy = 0.0;
          if constexpr (kWide) {
#pragma unroll
            for (int w = 0; w < 4; ++w) gy += TP[(5 * w + cc) * NS + k];
          } else {
#pragma unroll
            for (int o = 0; o < R; ++o) gy += T4[(4 * o + cc) * NS + k];
          }
          aty[0] += yn0 + gy; aty[1] += apv * yn0 + yn1; aty[2] += bpa * yn0 + bva * yn1;
        }
      } else {
        const double q0 = up1(x[2]), q1 = up1(x[3]);
        row((notfirst ? q0 : 0.0) - x[0], zd[0], ed_[0]); row((notfirst ? q1 : 0.0) - x[1], zd[1], ed_[1]);
        aty[0] = rhb[0] * ub[0] - yd0; aty[1] = rhb[1] * ub[1] - yd1; aty[2] = rhb[2] * ub[2]; aty[3] = rhb[3] * ub[3];
        if (hasu) {
          aty[2] += yn0; aty[3] += yn1;
          if constexpr (kWide) {
#pragma unroll
            for (int w = 0; w < 4; ++w) { aty[2] -= TP[(5 * w + 3) * NS + k]; aty[3] -= TP[(5 * w + 4) * NS + k]; }
          } else {
#pragma unroll
            for (int o = 0; o < R; ++o) { const double yo = T4[(4 * o + 3) * NS + k]; if ((slmask >> o) & 1u) aty[3] -= yo; else aty[2] -= yo; }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NOW; ++q) {
        const int o = 4 * q + warp;
        if (o < R && hasu) {
          const double x0v = m.RA[k * 6], x1v = m.RA[k * 6 + 2], x2v = m.RA[k * 6 + 4], xs = XR[osl[q] * NS + k];
          row(og3[3 * q] * x0v + og3[3 * q + 1] * x1v + og3[3 * q + 2] * x2v - xs, zo[q], eo_[q]);
          cert(orh[q] * (uo[q] - ouo_[q]), eo_[q], olo[q], sh.obs_hi);
        }
      }
#pragma unroll
      for (int h = 0; h < NHR; ++h) {
        const int o = 4 * h + warp;
        if (o < R && hasu) {
          const double x0v = m.RA[k * 6], x1v = m.RA[k * 6 + 2], x2v = m.RA[k * 6 + 4], xs = XR[(int)HR(o, 11) * NS + k];
          row(HR(o, 4) * x0v + HR(o, 5) * x1v + HR(o, 6) * x2v - xs, HR(o, 0), eh_[h]);
          cert(HR(o, 2) * (HR(o, 1) - ouh_[h]), eh_[h], HR(o, 3), sh.obs_hi);
        }
      }
      if constexpr (kWide) {
        if (hasu) {
          const double x0v = m.RA[k * 6], x1v = m.RA[k * 6 + 2], x2v = m.RA[k * 6 + 4], xd = XR[k], xs_ = XR[NS + k];
          for (int o = warp; o < R; o += 4) {
            const double e_ = WSE_(NBR + o, k), uv = UO[o * NS + k], rv = ORH[o * NS + k];
            row(OG3[(3 * o) * NS + k] * x0v + OG3[(3 * o + 1) * NS + k] * x1v + OG3[(3 * o + 2) * NS + k] * x2v - (((slmask >> o) & 1u) ? xs_ : xd),
                ZO[o * NS + k], e_);
            cert(rv * (uv - OU_(NBR + o, k)), e_, OLO[o * NS + k], sh.obs_hi);
          }
        }
      }
      cert(rhd[0] * (ud[0] - oud_[0]), ed_[0], bnd[0], bnd[0]); cert(rhd[1] * (ud[1] - oud_[1]), ed_[1], bnd[1], bnd[1]);
#pragma unroll
      for (int e = 0; e < NVR; ++e) {
        if (e < 2 || hasu) {
          row(x[e], zb[e], er_[e]); cert(rhb[e] * (ub[e] - oub_[e]), er_[e], blo(e), bhi(e));
          var(x[e], pv_[e], cq[e], aty[e], dv_[e], x[e] - oxv[e]);
        }
      }
      if (!live) {                                // ghost lanes carry a meaningless copy of the last stage
#pragma unroll
        for (int e = 0; e < 15; ++e) mx[e] = 0.0;
        sm[0] = sm[1] = sm[2] = 0.0;
      }
      cta_reduce(mx, sm, warp);
      obj = sm[0];
      pri_res = mx[0]; nAx = mx[1]; nZ = mx[2]; pri_s = mx[3]; nAx_s = mx[4]; nZ_s = mx[5];
      dua_res = mx[6] * cinv; nPx = mx[7] * cinv; nAty = mx[8] * cinv; dua_s = mx[9]; nPx_s = mx[10]; nAty_s = mx[11];
      s_ndy = mx[12]; s_lhs = sm[1]; s_ndx = mx[13]; s_pm = mx[14]; s_qd = sm[2];
      info_iter = iter;
    };
    // check_termination (auxil.h:133).  The residual tests are exact; the infeasibility certificates are screened
    // with their first (necessary) conditions and only evaluated in full, by warp 0 from the parked arrays, when
    // the screening passes.
    auto check = [&](const bool approx) -> bool {
      double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
      if (pri_res > kInfty || dua_res > kInfty) { status = kNonCvx; obj = kOsqpNan; return true; }
      if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
      const bool pr = sh.m == 0 || pri_res < ea + er * fmax(nZ, nAx);
      const bool dr = dua_res < ea + er * fmax(nq, fmax(nAty, nPx));
      if (pr && dr) { status = approx ? kSolvedInacc : kSolved; return true; }
      bool maybe = false;
      if (!pr) maybe = maybe || ((s_ndy > epi) && (s_lhs < -epi * s_ndy));
      if (!dr) maybe = maybe || ((s_ndx > edi) && (s_qd < -c * edi * s_ndx) && (s_pm < c * edi * s_ndx));
      if (!maybe) return false;
      park();
      cta_sync();
      if (warp == 0) { const bool d = check_termination(approx); if (lane == 0) *flag = d ? 1 : 0; }
      cta_sync();
      const int f = *flag;
      cta_sync();
      return f != 0;
    };

    // ---- osqp_solve (osqp.h:78): same control flow as solve() below
    status = kUnsolved; rho_updates = resume_ ? resume_rho_updates_ : 0; info_iter = 0; obj = 0.0; pri_res = 0.0; dua_res = 0.0;
    int iter = resume_ ? resume_iter_ : 0;
    // a resumed instance re-creates the factorisation from its parked rho vector; unless a rho update was pending when it
    // was parked, it then continues from the parked right-hand side (soft) exactly as if it had never stopped
    bool soft = resume_ && resume_soft_;
    bool refactor = true, last_checked = false, approx = false, reload = false;
    for (;;) {
      bool do_info, do_check, do_adapt = false;
      const bool final_pass = iter >= st.max_iter;
      if (!final_pass) {
        if (refactor) {
          cta_sync();                                // parked iterates / new Rh of all warps are visible to warp 0
          MQ_T0();
          if (warp == 0) {
            factor();                                     // leaf elimination + T blocks
            if (!soft) {
              rows_phase<1>(); rhs_finish();              // whole right-hand side into B_
              MQ_FOR_STAGES(kk) { OG_(0, kk) = 0.0; OG_(1, kk) = 0.0; OG_(2, kk) = 0.0; }
            }
          }
          MQ_T(1);
          pcr_factor_cta(warp);                           // starts and ends with a CTA barrier
          MQ_T(2);
          load();
          if (soft) {                                     // parked obstacle parts of the right-hand side (see the suspend exit)
            if constexpr (AX) ogp = TD_(cc, k); else { ps[0] = TD_(3, k); ps[1] = TD_(4, k); }
            soft = false;
          }
          MQ_T(3);
          refactor = false; reload = true;
        }
        int nb = st.max_iter;
        if (st.check_termination) { int c2 = (iter / st.check_termination + 1) * st.check_termination; if (c2 < nb) nb = c2; }
        if (st.adaptive_rho && st.adaptive_rho_interval) { int c2 = (iter / st.adaptive_rho_interval + 1) * st.adaptive_rho_interval; if (c2 < nb) nb = c2; }
        if constexpr (ASSIST) {                          // tell the assistants how long the burst is / to re-read the matrices
          if (warp == 0 && lane == 0) { cmd[0] = nb - iter; cmd[1] = reload ? 1 : 0; }
          bar_sync(kBarCmd, kCtaThreads);
          reload = false;
        }
        { MQ_T0(); iterate(nb - iter); MQ_T(4); }
        iter = nb;
        do_check = st.check_termination && (iter % st.check_termination == 0);
        do_adapt = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
        do_info = do_check || do_adapt;
        last_checked = do_check;
      } else if (!approx) {
        do_info = !last_checked; do_check = !last_checked;
      } else {
        do_info = false; do_check = true;
      }
      MQ_T0();
      if (do_info) info(iter);
      const bool done_ = do_check && check(approx);
      MQ_T(5);
      if (done_) break;
      if (final_pass) {
        if (approx) { status = kMaxIter; break; }
        approx = true;
        continue;
      }
      if (do_adapt) {
        // compute_rho_estimate / adapt_rho (auxil.h:21-38): the decision is taken by every thread on identical scalars
        const double pn = pri_s / (fmax(nZ_s, nAx_s) + 1e-10);
        const double dn = dua_s / (fmax(nq_s, fmax(nAty_s, nPx_s)) + 1e-10);
        double rn = rho * sqrt(pn / (dn + 1e-10));
        rn = fmin(fmax(rn, kRhoMin), kRhoMax);
        if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
          MQ_T0();
          // osqp_update_rho: new rho vector (auxil.h set_rho_vec classes on the SCALED bounds), y kept => u = y / Rh rescaled
          rho = rn; rho_updates += 1;
          auto rescale = [&](double& rh, double& u, double e, double lo_, double hi_) {
            const int t = row_type(e, lo_, hi_);
            if (t >= 0) { const double rnew = rho_of_type(t) * e * e; u = u * rh / rnew; rh = rnew; }
          };
#pragma unroll
          for (int e = 0; e < NVR; ++e) if (e < 2 || hasu) rescale(rhb[e], ub[e], WSE_(8 + vj(e), k), blo(e), bhi(e));
#pragma unroll
          for (int t = 0; t < 2; ++t) rescale(rhd[t], ud[t], WSE_(di(t), k), bnd[t], bnd[t]);
#pragma unroll
          for (int q = 0; q < NOW; ++q) { const int o = 4 * q + warp; if (o < R && hasu) rescale(orh[q], uo[q], WSE_(NBR + o, k), olo[q], sh.obs_hi); }
#pragma unroll
          for (int h = 0; h < NHR; ++h) {
            const int o = 4 * h + warp;
            if (o < R && hasu) {
              double rv = HR(o, 2), uv = HR(o, 1);
              rescale(rv, uv, WSE_(NBR + o, k), HR(o, 3), sh.obs_hi);
              HR(o, 2) = rv; HR(o, 1) = uv;
            }
          }
          if constexpr (kWide) {
            if (hasu) for (int o = warp; o < R; o += 4) {
              double rv = ORH[o * NS + k], uv = UO[o * NS + k];
              rescale(rv, uv, WSE_(NBR + o, k), OLO[o * NS + k], sh.obs_hi);
              ORH[o * NS + k] = rv; UO[o * NS + k] = uv;
            }
          }
          park();                                          // also writes the new Rh for the factorisation
          refactor = true;
          MQ_T(6);
        }
      }
      if (allow_suspend_ && iter == susp_K_ && iter < st.max_iter) {
        // still running after K iterations: hand the instance to a later launch if a slot is free
        cta_sync();
        if (warp == 0 && lane == 0) *flag = atomicAdd(susp_count_, 1);
        cta_sync();
        const int slot = *flag;
        cta_sync();
        if (slot < susp_cap_) { susp_slot_ = slot; susp_refactor_ = refactor; status = kSuspended; break; }
        allow_suspend_ = false;
      }
    }
    park();
    if (status == kSuspended && live) {
      if constexpr (AX) TD_(cc, k) = ogp; else { TD_(3, k) = ps[0]; TD_(4, k) = ps[1]; }
    }
  }

  // Setup only (scaling.h: scale_data, auxil.h: set_rho_vec, warm start) for instance b, by a CTA of a dedicated high-occupancy
  // kernel: the cold block is built in shared memory (`smem0`: cold_slots(R) * NS doubles + 8 of reduction scratch) and left in
  // slot `slot` exactly as a parked instance at iteration 0 with a factorisation pending, for run_cta_resume to pick up.
  MQ_HD void run_setup_only(const Batch& bt, int b, int slot, int warp) {
    x0p = bt.x0 + (size_t)b * 8;
    if (bt.limits) { lim_v = bt.limits[2 * (size_t)b]; lim_a = bt.limits[2 * (size_t)b + 1]; has_lim = true; }
    slack = bt.slack + (size_t)b * bt.slack_stride;
    const Mem keep = m;
    map_cold(m, smem0, NS, R);
    m.YB = smem0 + cold_slots(R) * NS;
    load_and_scale<4>(bt, b, warp);
    double* dst = bt.susp_cold + (size_t)slot * bt.susp_stride;
    if (threadIdx.x == 0) {
      // the cold block (55 KB at four obstacle rows) leaves shared memory as ONE bulk copy
      bulk_store_shared_to_global(dst, smem0, (unsigned)(cold_slots(R) * NS * sizeof(double)));
      double* sc = bt.susp_scal + (size_t)slot * 8;
      sc[0] = c; sc[1] = cinv; sc[2] = rho; sc[3] = nq; sc[4] = nq_s; sc[5] = 0.0; sc[6] = 1.0; sc[7] = 0.0;
      bt.susp_list[slot] = b;
    }
    m = keep;
    cta_sync();
  }

  // Resume the instance parked in `slot` by another launch (run_cta's suspend exit): same state, same arithmetic from here on.
  MQ_HD void run_cta_resume(const Batch& bt, int slot, int warp, volatile int* flag, volatile int* cmd = nullptr) {
    const int b = bt.susp_list[slot];
    x0p = bt.x0 + (size_t)b * 8;
    if (bt.limits) { lim_v = bt.limits[2 * (size_t)b]; lim_a = bt.limits[2 * (size_t)b + 1]; has_lim = true; }
    slack = bt.slack + (size_t)b * bt.slack_stride;
    const double* src = bt.susp_cold + (size_t)slot * bt.susp_stride;
    for (int i = threadIdx.x; i < cold_slots(R) * NS; i += 128) m.E[i] = src[i];
    const double* sc = bt.susp_scal + (size_t)slot * 8;
    c = sc[0]; cinv = sc[1]; rho = sc[2]; nq = sc[3]; nq_s = sc[4];
    resume_ = true; resume_rho_updates_ = (int)sc[5]; resume_soft_ = sc[6] == 0.0; resume_iter_ = (int)sc[7];
    allow_suspend_ = false;
    cta_sync();
    if (warp < 3) solve_role<true>(warp, flag, cmd); else solve_role<false>(warp, flag, cmd);
    resume_ = false;
    cta_sync();
    if (warp == 0) store(bt, b);
    cta_sync();
  }

  // PCR assistant (threads 128..223 of an ASSIST block): axis warp `a` keeps its rows of the level 1, 2 matrices and of the
  // final operator (levels 3 + 4 composed) in registers (192 of them) and runs them in every solve; the row warps hand
  // the level-0 result over in buffer 1 (kBarH1) and find y in the y buffer at kBarY.  Commands arrive through cmd[]:
  // cmd[0] = iterations of the next burst (< 0: the block is done), cmd[1] = the factorisation changed.
  MQ_HD void assist_role(const int a, volatile int* cmd) {
    const int k = lane < NS ? lane : NS - 1;
    const bool live = lane < NS;
    double2* const RA2 = reinterpret_cast<double2*>(m.RA);
    double2* const YB2 = reinterpret_cast<double2*>(m.YB);
    const double2* const M = reinterpret_cast<const double2*>(m.PCR) + (a * 12) * NS + k;
    const int c0 = k & 7;
    const int t0 = c0, t1 = c0 + 8, t2 = c0 + 16 <= N ? c0 + 16 : c0, t3 = c0 + 24 <= N ? c0 + 24 : c0;     // the stage's stride-8 chain (W_p = 0 where it is shorter)
    double2 mt[2][12], wt[24];
#pragma unroll
    for (int l = 0; l < 2; ++l)
#pragma unroll
      for (int h = 0; h < 12; ++h) mt[l][h] = make_double2(0.0, 0.0);
#pragma unroll
    for (int h = 0; h < 24; ++h) wt[h] = make_double2(0.0, 0.0);
    for (;;) {
      bar_sync(kBarCmd, kCtaThreads);
      const int n = cmd[0], rl = cmd[1];
      if (n < 0) break;
      if (rl) {
#pragma unroll
        for (int l = 0; l < 2; ++l)
#pragma unroll
          for (int h = 0; h < 12; ++h) mt[l][h] = M[(l + 1) * (kPcrLevelDoubles / 2) + h * NS];
#pragma unroll
        for (int h = 0; h < 24; ++h) wt[h] = M[(3 + h / 12) * (kPcrLevelDoubles / 2) + (h % 12) * NS];
      }
      for (int it = 0; it < n; ++it) {
        bar_sync(kBarH1, 192);
        double r0, r1;
        { const double2 own = RA2[3 * NS + k * 3 + a]; r0 = own.x; r1 = own.y; }
#pragma unroll
        for (int l = 1; l < kPcrLevels - 2; ++l) {
          const int s = 1 << l, cur = l & 1;
          const int kmm = k - s >= 0 ? k - s : k, kpp = k + s <= N ? k + s : k;
          const double2* ra = RA2 + cur * 3 * NS;
          const double2 a0 = ra[kmm * 3], a1 = ra[kmm * 3 + 1], a2 = ra[kmm * 3 + 2];
          const double2 c0_ = ra[kpp * 3], c1 = ra[kpp * 3 + 1], c2 = ra[kpp * 3 + 2];
          const double2* w = mt[l - 1];
          double sa0 = w[0].x * a0.x, sa1 = w[3].x * a0.x, sg0 = w[6].x * c0_.x, sg1 = w[9].x * c0_.x;
          sa0 = fma(w[0].y, a0.y, sa0); sa1 = fma(w[3].y, a0.y, sa1); sg0 = fma(w[6].y, c0_.y, sg0); sg1 = fma(w[9].y, c0_.y, sg1);
          sa0 = fma(w[1].x, a1.x, sa0); sa1 = fma(w[4].x, a1.x, sa1); sg0 = fma(w[7].x, c1.x, sg0); sg1 = fma(w[10].x, c1.x, sg1);
          sa0 = fma(w[1].y, a1.y, sa0); sa1 = fma(w[4].y, a1.y, sa1); sg0 = fma(w[7].y, c1.y, sg0); sg1 = fma(w[10].y, c1.y, sg1);
          sa0 = fma(w[2].x, a2.x, sa0); sa1 = fma(w[5].x, a2.x, sa1); sg0 = fma(w[8].x, c2.x, sg0); sg1 = fma(w[11].x, c2.x, sg1);
          sa0 = fma(w[2].y, a2.y, sa0); sa1 = fma(w[5].y, a2.y, sa1); sg0 = fma(w[8].y, c2.y, sg0); sg1 = fma(w[11].y, c2.y, sg1);
          r0 = (r0 - sa0) - sg0; r1 = (r1 - sa1) - sg1;
          if (live) RA2[(1 - cur) * 3 * NS + k * 3 + a] = make_double2(r0, r1);
          bar_sync(kBarA, 96);
        }
        // levels 3 and 4 as one operator over the chain (level-2 result in buffer 1): y = sum_p W_p r_{c0+8p}
        {
          const double2* const rb = RA2 + 3 * NS;
          double sa0 = 0.0, sa1 = 0.0, sg0 = 0.0, sg1 = 0.0;
          const int ts_[4] = {t0, t1, t2, t3};
#pragma unroll
          for (int p_ = 0; p_ < 4; p_ += 2) {
            const double2 a0 = rb[ts_[p_] * 3], a1 = rb[ts_[p_] * 3 + 1], a2 = rb[ts_[p_] * 3 + 2];
            const double2 c0_ = rb[ts_[p_ + 1] * 3], c1 = rb[ts_[p_ + 1] * 3 + 1], c2 = rb[ts_[p_ + 1] * 3 + 2];
            const double2* w = wt + 6 * p_;
            sa0 = fma(w[0].x, a0.x, sa0); sa1 = fma(w[3].x, a0.x, sa1); sg0 = fma(w[6].x, c0_.x, sg0); sg1 = fma(w[9].x, c0_.x, sg1);
            sa0 = fma(w[0].y, a0.y, sa0); sa1 = fma(w[3].y, a0.y, sa1); sg0 = fma(w[6].y, c0_.y, sg0); sg1 = fma(w[9].y, c0_.y, sg1);
            sa0 = fma(w[1].x, a1.x, sa0); sa1 = fma(w[4].x, a1.x, sa1); sg0 = fma(w[7].x, c1.x, sg0); sg1 = fma(w[10].x, c1.x, sg1);
            sa0 = fma(w[1].y, a1.y, sa0); sa1 = fma(w[4].y, a1.y, sa1); sg0 = fma(w[7].y, c1.y, sg0); sg1 = fma(w[10].y, c1.y, sg1);
            sa0 = fma(w[2].x, a2.x, sa0); sa1 = fma(w[5].x, a2.x, sa1); sg0 = fma(w[8].x, c2.x, sg0); sg1 = fma(w[11].x, c2.x, sg1);
            sa0 = fma(w[2].y, a2.y, sa0); sa1 = fma(w[5].y, a2.y, sa1); sg0 = fma(w[8].y, c2.y, sg0); sg1 = fma(w[11].y, c2.y, sg1);
          }
          if (live) YB2[k * 3 + a] = make_double2(sa0 + sg0, sa1 + sg1);
          bar_arrive(kBarY, kCtaThreads);
        }
      }
    }
  }

  // Row helper (threads 224..255 of a kHelp block): the obstacle rows o = 3, 7 — or all of them (kHelpAll) — of every iteration,
  // on the values their owners left in shared memory (solve_role: load), exactly as obstacle_rows() runs a register row.  It
  // waits for y like everyone else, announces the rows' terms at kBarT and follows the same burst commands as the assistants.
  MQ_HD void helper_role(volatile int* cmd) {
    constexpr int NROW = kHelpAll ? RC : (kHelp ? RC / 4 : 0), NR_ = NROW > 0 ? NROW : 1;
    const int k = lane < NS ? lane : NS - 1;
    const bool live = lane < NS, hasu = lane < N;
    const double al = st.alpha, om = 1.0 - st.alpha;
    double* const XR = m.YB + 6 * NS;
    double* const T4 = XR + 2 * NS;
    auto HR = [&](int o, int slot) -> double& { return m.ROW[(o * kHelpSlots + slot) * NS + k]; };
    for (;;) {
      bar_sync(kBarCmd, kCtaThreads);
      const int n = cmd[0];
      if (n < 0) break;
      for (int it = 0; it < n; ++it) {
        // everything but y and r11 is fetched before the barrier
        double zo[NR_], uo[NR_], orh[NR_], olo[NR_], og3[3 * NR_], odg[NR_], ofs[3 * NR_];
        int osl[NR_];
#pragma unroll
        for (int h = 0; h < NROW; ++h) {
          const int o = kHelpAll ? h : 4 * h + 3;
          zo[h] = HR(o, 0); uo[h] = HR(o, 1); orh[h] = HR(o, 2); olo[h] = HR(o, 3);
          og3[3 * h] = HR(o, 4); og3[3 * h + 1] = HR(o, 5); og3[3 * h + 2] = HR(o, 6);
          odg[h] = HR(o, 7); ofs[3 * h] = HR(o, 8); ofs[3 * h + 1] = HR(o, 9); ofs[3 * h + 2] = HR(o, 10);
          osl[h] = (int)HR(o, 11);
        }
        bar_sync(kBarY, kCtaThreads);
        const double yp0 = m.YB[k * 6], yp1 = m.YB[k * 6 + 2], yp2 = m.YB[k * 6 + 4];
        // loads, arithmetic and stores in separate passes: the rows' chains then interleave (a store between two rows would
        // order the next row's loads behind it)
        double r11s[NR_], w0[NR_], w1[NR_], w2[NR_], w3[NR_];
#pragma unroll
        for (int h = 0; h < NROW; ++h) r11s[h] = XR[osl[h] * NS + k];
#pragma unroll
        for (int h = 0; h < NROW; ++h) {
          const double sg = hasu ? odg[h] * (r11s[h] - ofs[3 * h] * yp0 - ofs[3 * h + 1] * yp1 - ofs[3 * h + 2] * yp2) : 0.0;
          const double zt = og3[3 * h] * yp0 + og3[3 * h + 1] * yp1 + og3[3 * h + 2] * yp2 - sg;
          const double v = al * zt + om * zo[h] + uo[h];
          const double zn = v > olo[h] ? v : olo[h];
          zo[h] = zn; uo[h] = v - zn;
          const double t = orh[h] * (zn - uo[h]);
          const double ts_ = hasu ? odg[h] * t : 0.0;
          w0[h] = fma(ofs[3 * h], ts_, og3[3 * h] * t); w1[h] = fma(ofs[3 * h + 1], ts_, og3[3 * h + 1] * t);
          w2[h] = fma(ofs[3 * h + 2], ts_, og3[3 * h + 2] * t); w3[h] = t;
        }
        if (live) {
#pragma unroll
          for (int h = 0; h < NROW; ++h) {
            const int o = kHelpAll ? h : 4 * h + 3;
            HR(o, 0) = zo[h]; HR(o, 1) = uo[h];
            T4[(4 * o) * NS + k] = w0[h]; T4[(4 * o + 1) * NS + k] = w1[h]; T4[(4 * o + 2) * NS + k] = w2[h]; T4[(4 * o + 3) * NS + k] = w3[h];
          }
        }
        bar_arrive(kBarT, 128);
      }
    }
  }

  MQ_HD void run_cta(const Batch& bt, int b, int warp, volatile int* flag, volatile int* cmd = nullptr) {
    x0p = bt.x0 + (size_t)b * 8;
    if (bt.limits) { lim_v = bt.limits[2 * (size_t)b]; lim_a = bt.limits[2 * (size_t)b + 1]; has_lim = true; }
    slack = bt.slack + (size_t)b * bt.slack_stride;
    if constexpr (kWide) {
      if (bt.nobs) { Dm::R = bt.nobs[b]; Dm::MK = NBR + Dm::R; map_memory(m, smem0, ws0, NS, Dm::R, QMODE); }
    }
    // setup (scaling.h: scale_data, auxil.h: set_rho_vec, warm start) builds the cold block in the still unused PCR
    // region of shared memory; the CTA then copies it to its global home in one pass
    {
      MQ_T0();
      const Mem keep = m;
      map_cold(m, keep.PCR, NS, R);
      load_and_scale<4>(bt, b, warp);           // all four warps; every thread ends with the same c, rho, |q| norms
      m = keep;
      if (threadIdx.x == 0) bulk_store_shared_to_global(m.E, m.PCR, (unsigned)(cold_slots(R) * NS * sizeof(double)));   // one bulk copy
      cta_sync();
      MQ_T(0);
    }
    allow_suspend_ = bt.suspend_at > 0 && bt.suspend_at < st.max_iter && !kWide;
    susp_K_ = bt.suspend_at; susp_cap_ = bt.susp_cap; susp_count_ = bt.susp_count; resume_ = false;
    if (warp < 3) solve_role<true>(warp, flag, cmd); else solve_role<false>(warp, flag, cmd);
    cta_sync();
    if (status == kSuspended) {                            // park the cold block and the scalars in the slot; no result yet
      double* dst = bt.susp_cold + (size_t)susp_slot_ * bt.susp_stride;
      for (int i = threadIdx.x; i < cold_slots(R) * NS; i += 128) dst[i] = m.E[i];
      if (threadIdx.x == 0) {
        double* sc = bt.susp_scal + (size_t)susp_slot_ * 8;
        sc[0] = c; sc[1] = cinv; sc[2] = rho; sc[3] = nq; sc[4] = nq_s; sc[5] = (double)rho_updates; sc[6] = susp_refactor_ ? 1.0 : 0.0;
        sc[7] = (double)susp_K_;
        bt.susp_list[susp_slot_] = b;
        if (bt.susp_key) bt.susp_key[susp_slot_] = pri_res == pri_res ? pri_res : 1e300;
      }
      cta_sync();
    } else {
      MQ_T0();
      if (warp == 0) store(bt, b);
      cta_sync();
      MQ_T(7);
    }
#ifdef MPCQP_PHASE_TIMING
    tacc[12] = dbg_ruiz;
    if (threadIdx.x == 0 && bt.dbg) { for (int i = 0; i < 16; ++i) bt.dbg[(size_t)b * 16 + i] = tacc[i]; }
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#endif
  }
#endif

  // ---- osqp_solve (osqp.h:78) ---------------------------------------------------------------------
  // Written as one loop with a single call site per phase (factor, burst, update_info, check_termination,
  // adapt_rho) so that everything inlines into the kernel once: shared-memory addresses then fold to
  // immediates and the settings / shape fields to constant-bank operands.
  MQ_HD void solve() {
    status = kUnsolved; rho_updates = 0; info_iter = 0; obj = 0.0; pri_res = 0.0; dua_res = 0.0;
    int iter = 0;
    bool refactor = true, last_checked = false, approx = false;
    for (;;) {
      bool do_info, do_check, do_adapt = false;
      const bool final_pass = iter >= st.max_iter;
      if (!final_pass) {
        if (refactor) { factor(); rows_phase<1>(); rhs_finish(); refactor = false; }
        int nb = st.max_iter;
        if (st.check_termination) { int c2 = (iter / st.check_termination + 1) * st.check_termination; if (c2 < nb) nb = c2; }
        if (st.adaptive_rho && st.adaptive_rho_interval) { int c2 = (iter / st.adaptive_rho_interval + 1) * st.adaptive_rho_interval; if (c2 < nb) nb = c2; }
        burst(nb - iter);
        iter = nb;
        do_check = st.check_termination && (iter % st.check_termination == 0);
        do_adapt = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
        do_info = do_check || do_adapt;
        last_checked = do_check;
      } else if (!approx) {          // loop ran out: one exact check if the last iteration was not checked
        do_info = !last_checked; do_check = !last_checked;
      } else {                       // approximate (10x) check decides between *_inaccurate and max_iter
        do_info = false; do_check = true;
      }
      if (do_info) update_info(iter);
      if (do_check && check_termination(approx)) break;
      if (final_pass) {
        if (approx) { status = kMaxIter; break; }
        approx = true;
        continue;
      }
      if (do_adapt && adapt_rho()) refactor = true;
    }
  }

  // auxil.h:118 store_solution (+ scaling.h unscale_solution): x = D x_s = xh, y = E y_s / c = Rh u / c
  MQ_NOINL void store(const Batch& bt, int b) {
    const bool has = status != kPrimInf && status != kPrimInfInacc && status != kDualInf && status != kDualInfInacc && status != kNonCvx;
    double* xo = bt.x + (size_t)b * sh.n;
    double* yo = bt.y ? bt.y + (size_t)b * sh.m : nullptr;
    MQ_FOR_STAGES(k) {
      const int nv = nvars(k), nr = nrows(k);
      for (int j = 0; j < nv; ++j) xo[j < 8 ? 8 * k + j : 8 * NS + 5 * k + (j - 8)] = has ? X_(j, k) : kOsqpNan;
      if (yo) for (int i = 0; i < nr; ++i) {
        yo[row_index(k, i)] = has ? RH_(i, k) * U_(i, k) * cinv : kOsqpNan;
      }
    }
#if MQ_DEV
    if (lane == 0)
#endif
    {
      bt.status[b] = status; bt.iter[b] = info_iter; bt.rho_updates[b] = rho_updates;
      bt.obj[b] = obj; bt.pri_res[b] = pri_res; bt.dua_res[b] = dua_res;
    }
  }

  MQ_HD Qp(double* smem, const Shape& shape, const Settings& set, const Batch& bt, double* ws, int lane_)
      : Dm(shape.NS, shape.R), sh(shape), st(set) {
    lane = lane_; Rs = shape.R; smem0 = smem; ws0 = ws;
    map_memory(m, smem, ws, NS, R, QMODE);
    pd = bt.pd; slack = bt.slack;
  }
  MQ_HD void run(const Batch& bt, int b) {
    x0p = bt.x0 + (size_t)b * 8;
    slack = bt.slack + (size_t)b * bt.slack_stride;
    load_and_scale(bt, b);
    solve();
    store(bt, b);
    MQ_SYNC();
  }
};

}  // namespace mpcqp
