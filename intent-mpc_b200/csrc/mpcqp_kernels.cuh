// mpcqp_kernels.cuh — the solve kernels (bodies in mpcqp_core.cuh) and the table through which the host code reaches
// their instantiations.  Every instantiation is a separate translation unit of mpcqp_kernels.cu (one nvcc process per
// obstacle count, see __graft_entry__.build), so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>
#include "mpcqp_core.cuh"

namespace mpcqp {

typedef void (*SolveKernel)(const Shape, const Settings, const Batch, int, int*);
// Fast-path instantiations (compile-time dims, register-resident iterates); anything else runs the generic kernel.
#ifndef MPCQP_FAST_R_LIST
#define MPCQP_FAST_R_LIST X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#endif
// getters defined by the translation units of mpcqp_kernels.cu
#define X(r) SolveKernel mpcqp_kernel_cta_##r(bool assist); SolveKernel mpcqp_kernel_warp_##r(); SolveKernel mpcqp_kernel_setup_##r();
MPCQP_FAST_R_LIST
#undef X
SolveKernel mpcqp_kernel_cta_wide();
SolveKernel mpcqp_kernel_generic();
// CTA kernel (two-per-SM variant, also used alone on an SM) for horizons other than 30 — 20 is the code default of
// mpcPlanner::initParam (mpcPlanner.cpp:19-173), 25 the case where 8 * horizon % 5 != 0 (castMPCToQPHessian's index quirk).  Every
// (horizon, obstacle count) pair is a translation unit of its own (mpcqp_kernels.cu, -DMPCQP_GROUP_ALT_NS / _R) that registers its
// kernel here when the library is loaded; the lookup returns nullptr for pairs that were not built.
constexpr int kAltNsMax = 30, kAltRMax = 8;
struct AltKernelRegistration { AltKernelRegistration(int ns, int r, SolveKernel k); };
SolveKernel mpcqp_kernel_cta_alt(int ns, int r);

#ifdef MPCQP_KERNEL_BODIES
// ------------------------------------------------------------------------------------------------
// solve kernel: persistent, one warp (= one CTA) per QP at a time, work fetched from a global counter
// ------------------------------------------------------------------------------------------------
template <int NST, int RT>
__global__ void __launch_bounds__(32) mpcqp_solve_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Settings st,
                                                         const __grid_constant__ Batch bt, int ws_stride, int* counter) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x;
  Qp<NST, RT> qp(smem, sh, st, bt, bt.ws + (size_t)blockIdx.x * ws_stride, lane);
  // queue 3 (bt.order given): the instances flagged hard first, then the others in natural order; otherwise natural order
  bool hardq = bt.order != nullptr;
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(counter + (hardq || !bt.order ? 0 : 3), 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    int b = idx;
    if (hardq) {
      if (idx >= *bt.nhard) { hardq = false; continue; }
      b = bt.order[idx];
    } else {
      if (idx >= bt.B) break;
      if (bt.order && bt.hard[idx]) continue;
    }
    qp.run(bt, b);
  }
}

// CTA kernel: persistent, one 4-warp CTA per QP at a time (mode 2 of mpcqp_core.cuh), two CTAs per SM.  ASSIST: the
// block has the SM to itself and carries three more warps that keep the PCR matrices of levels 1..3 in registers, and — with
// four or more obstacle rows per stage — an eighth warp that runs the slack warp's obstacle rows (Qp::helper_role).
template <int RT, bool ASSIST, int NST = 30>
__global__ void __launch_bounds__(Qp<NST, RT, kModeCta, ASSIST>::kCtaThreads, ASSIST ? 1 : 2) mpcqp_solve_cta_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Settings st,
                                                                                        const __grid_constant__ Batch bt, int ws_stride, int* counter) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next, s_flag, s_cmd[2];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // one-per-SM blocks: opaque to the compiler, which would otherwise re-derive the lane (S2R + mask + compare, ~40 dependent
  // cycles) at every use of a lane predicate in the iteration loop instead of keeping it in a register (two-per-SM blocks are
  // better off with the register: 16k batch 445k -> 408k QPs/s with it)
  if constexpr (ASSIST) asm volatile("" : "+r"(lane), "+r"(warp));
  using Q = Qp<NST, RT, kModeCta, ASSIST>;
  Q qp(smem, sh, st, bt, bt.ws + (size_t)blockIdx.x * ws_stride, lane);
  if constexpr (ASSIST) {
    if (warp >= 4) {
      if (warp < 7) qp.assist_role(warp - 4, s_cmd);
      else if constexpr (Q::kHelp) qp.helper_role(s_cmd);
      return;
    }
  }
  // queue 0: all instances in natural order.  queue 1: the hard list only.  queue 2: everything not flagged hard, then
  // whatever is left of the hard list (so an over-long hard list does not serialise on the one-per-SM launch).
  // queue 3: the hard list first, then everything else (one launch; used when every CTA has an SM to itself anyway).
  if (bt.queue == 4) {                                     // resume the instances parked by an earlier launch
    int n = *bt.susp_count; if (n > bt.susp_cap) n = bt.susp_cap;
    for (;;) {
      if (threadIdx.x == 0) s_next = atomicAdd(counter + 2, 1);
      cta_sync();
      const int slot = s_next;
      cta_sync();
      if (slot >= n) break;
      qp.run_cta_resume(bt, bt.susp_order ? bt.susp_order[slot] : slot, warp, &s_flag, s_cmd);
    }
  } else {
  bool natural = bt.queue != 1 && bt.queue != 3;
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(counter + ((natural && bt.queue >= 2) ? 3 : 0), 1);
    cta_sync();
    const int idx = s_next;
    cta_sync();
    int b = idx;
    if (natural) {
      if (idx >= bt.B) { if (bt.queue == 2) { natural = false; continue; } break; }
      if (bt.queue >= 2 && bt.hard[idx]) continue;       // on the hard list
    } else {
      if (idx >= *bt.nhard) { if (bt.queue == 3) { natural = true; continue; } break; }
      b = bt.order[idx];
    }
    qp.run_cta(bt, b, warp, &s_flag, s_cmd);
  }
  }
  if constexpr (ASSIST) {                                  // release the assistants
    if (threadIdx.x == 0) s_cmd[0] = -1;
    Q::bar_sync(5, Q::kCtaThreads);
  }
}

// Setup kernel: persistent, one 4-warp CTA per instance at a time, three CTAs per SM (the cold block in shared memory is all it
// needs); instances are taken hard list first, then the rest, and slot numbers are handed out in that order so that the solve
// launch that follows (queue 4) starts the long ones first.  counter[0]: fetch, counter[3]: natural fetch, susp_count: slots.
template <int RT>
__global__ void __launch_bounds__(128, 3) mpcqp_setup_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Settings st,
                                                             const __grid_constant__ Batch bt, int ws_stride, int* counter) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next, s_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Qp<30, RT, kModeCta, false> qp(smem, sh, st, bt, bt.ws + (size_t)blockIdx.x * ws_stride, lane);
  bool natural = bt.nhard == nullptr;
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(counter + (natural ? 3 : 0), 1);
    cta_sync();
    const int idx = s_next;
    cta_sync();
    int b = idx;
    if (natural) {
      if (idx >= bt.B) break;
      if (bt.nhard && bt.hard[idx]) continue;
    } else {
      if (idx >= *bt.nhard) { natural = true; continue; }
      b = bt.order[idx];
    }
    if (threadIdx.x == 0) s_slot = atomicAdd(bt.susp_count, 1);
    cta_sync();
    const int slot = s_slot;
    cta_sync();
    qp.run_setup_only(bt, b, slot, warp);
  }
}
#endif  // MPCQP_KERNEL_BODIES

}  // namespace mpcqp
