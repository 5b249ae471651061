// mpcqp_kernels.cu — instantiations of the solve kernels.  Compiled once per group:
//   -DMPCQP_GROUP_R_LIST="X(4)"   the CTA kernel (with and without assistant warps) and the one-warp register kernel for
//                                 the listed obstacle counts;
//   -DMPCQP_GROUP_MISC            the wide CTA kernel (run-time obstacle count) and the generic one-warp kernel;
//   -DMPCQP_GROUP_ALT_NS=20 -DMPCQP_GROUP_ALT_R=4   the CTA kernel for that horizon (17..30) and obstacle count.
// A single-process build passes both (and several counts) at once.
#define MPCQP_KERNEL_BODIES
#include "mpcqp_kernels.cuh"

namespace mpcqp {
#ifdef MPCQP_GROUP_R_LIST
#define X(r) SolveKernel mpcqp_kernel_cta_##r(bool assist) { return assist ? mpcqp_solve_cta_kernel<r, true> : mpcqp_solve_cta_kernel<r, false>; } \
             SolveKernel mpcqp_kernel_warp_##r() { return mpcqp_solve_kernel<30, r>; } \
             SolveKernel mpcqp_kernel_setup_##r() { return mpcqp_setup_kernel<r>; }
MPCQP_GROUP_R_LIST
#undef X
#endif
#ifdef MPCQP_GROUP_ALT_NS          // -DMPCQP_GROUP_ALT_NS=20 -DMPCQP_GROUP_ALT_R=4: the CTA kernel at that horizon / obstacle count
static AltKernelRegistration alt_registration_(MPCQP_GROUP_ALT_NS, MPCQP_GROUP_ALT_R, mpcqp_solve_cta_kernel<MPCQP_GROUP_ALT_R, false, MPCQP_GROUP_ALT_NS>);
#endif
#ifdef MPCQP_GROUP_MISC
SolveKernel mpcqp_kernel_cta_wide() { return mpcqp_solve_cta_kernel<kWideR, true>; }
SolveKernel mpcqp_kernel_generic() { return mpcqp_solve_kernel<0, 0>; }
#endif
}  // namespace mpcqp
