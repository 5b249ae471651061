// mpcqp_dense.cu — kernel + launcher of the generic (unstructured) solve path (body: mpcqp_dense.cuh).  Its own translation
// unit: nothing here touches the register allocation of the stage-structured kernels.
#include <cuda_runtime.h>
#include "mpcqp_dense.cuh"

namespace mpcqp_dense {

// Persistent: one CTA per QP at a time, QPs taken round-robin (b = blockIdx.x, += gridDim.x).  The dense matrices live in
// the CTA's private global workspace (2 (n+m)^2 + n^2 + 2nm doubles: 2.8 MB at n = 200, m = 150, 0.3 MB at n = 64, m = 48; L2-resident while grid x that stays below the 126 MB);
// right-hand sides, the pivot column and the reduction scratch in shared memory.  HBM sees the CSC inputs and x, y once.
__global__ void __launch_bounds__(512) mpcqp_dense_solve_kernel(const __grid_constant__ Batch bt, const __grid_constant__ Settings st) {
  extern __shared__ double dq_smem[];
  Solver sv;
  double* ws = bt.ws + (size_t)blockIdx.x * bt.ws_stride;
  for (int b = blockIdx.x; b < bt.B; b += gridDim.x) {
    sv.run_batch_item(bt, b, ws, st, dq_smem, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
  }
}

int block_threads(int n, int m) {
  // row threads (n + m rounded up to a warp) x kMaxParts k-slices, at most 512 (126 registers per thread)
  const int nr = ((n + m + 31) / 32) * 32;
  int parts = 512 / nr; if (parts > kMaxParts) parts = kMaxParts; if (parts < 1) parts = 1;
  int threads = nr * parts;
  if (threads > 512) threads = 512;
  if (threads < 128) threads = 128;
  return threads;
}

// How many CTAs the launch will use for B problems (the caller sizes bt.ws = grid * ws_stride doubles).  0 on error.
int grid_size(int B, int n, int m, int device) {
  const size_t smem = smem_doubles(n, m) * sizeof(double);
  if (smem > 48 * 1024 && cudaFuncSetAttribute(mpcqp_dense_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpcqp_dense_solve_kernel, block_threads(n, m), smem) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (per_sm < 1) return 0;
  const long long cap = (long long)per_sm * sms;
  return (int)(B < cap ? B : cap);
}

int launch(const Batch& bt, int grid, const Settings& st, cudaStream_t stream) {
  const size_t smem = smem_doubles(bt.n, bt.m) * sizeof(double);
  mpcqp_dense_solve_kernel<<<grid, block_threads(bt.n, bt.m), smem, stream>>>(bt, st);
  return (int)cudaGetLastError();
}

}  // namespace mpcqp_dense
