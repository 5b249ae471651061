// mpcqp_band.cu — kernel, host-side pattern analysis (reverse Cuthill-McKee ordering, band slots, row lists) and launcher
// of the sparse generic solve path (body: mpcqp_band.cuh).  Its own translation unit.
#include <cuda_runtime.h>
#include <vector>

#include "mpcqp_band.cuh"
#include "mpcqp_band_host.hpp"

namespace mpcqp_band {

// Persistent: one warp (= one CTA of 32 threads) per QP at a time, QPs taken round-robin.  Shared memory: the band factor
// and two N-vectors; everything else in the warp's private global workspace (L2-resident).
__global__ void __launch_bounds__(32, 12) mpcqp_band_solve_kernel(const __grid_constant__ Batch bt, const __grid_constant__ Settings st) {
  extern __shared__ double bq_smem[];
  Solver sv;
  double* ws = bt.ws + (size_t)blockIdx.x * bt.ws_stride;
  for (int b = blockIdx.x; b < bt.B; b += gridDim.x) {
    sv.run(bt, b, st, ws, bq_smem, (int)threadIdx.x);
    __syncwarp();
  }
}

void pattern_bind(Pattern* pt, int n, int m, int w, int nnzP, int nnzA, const int* dev_flat, const int* off) {
  pt->n = n; pt->m = m; pt->N = n + m; pt->w = w; pt->nnzP = nnzP; pt->nnzA = nnzA;
  pt->Pc = dev_flat + off[0]; pt->Pi = dev_flat + off[1]; pt->Ac = dev_flat + off[2]; pt->Ai = dev_flat + off[3];
  pt->Pr_ptr = dev_flat + off[4]; pt->Pr_pos = dev_flat + off[5]; pt->Pr_col = dev_flat + off[6];
  pt->Ar_ptr = dev_flat + off[7]; pt->Ar_pos = dev_flat + off[8]; pt->Ar_col = dev_flat + off[9];
  pt->slotP = dev_flat + off[10]; pt->slotA = dev_flat + off[11]; pt->perm = dev_flat + off[12]; pt->iperm = dev_flat + off[13];
}

// How many warps (CTAs) the launch will use for B problems; 0 when the band does not fit shared memory.
int grid_size(int B, int N, int w, int device) {
  const size_t smem = smem_doubles(N, w) * sizeof(double);
  int optin = 0, sms = 0;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (smem > (size_t)optin) return 0;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(mpcqp_band_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpcqp_band_solve_kernel, 32, smem) != cudaSuccess || per_sm < 1) return 0;
  const long long cap = (long long)per_sm * sms;
  return (int)(B < cap ? B : cap);
}

int launch(const Batch& bt, int grid, const Settings& st, cudaStream_t stream) {
  const size_t smem = smem_doubles(bt.pt.N, bt.pt.w) * sizeof(double);
  mpcqp_band_solve_kernel<<<grid, 32, smem, stream>>>(bt, st);
  return (int)cudaGetLastError();
}

}  // namespace mpcqp_band
