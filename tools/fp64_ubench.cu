#include <cstdio>
#include <cuda_runtime.h>
// single-warp microbenchmarks: cycles per instruction for dependent / independent DFMA, SHFL, LDS
__global__ void k_dfma_dep(double* out, long long* cyc, int n) {
  double x = threadIdx.x * 1e-3, a = 0.999999, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> __global__ void k_dfma_ilp(double* out, long long* cyc, int n) {
  double x[ILP]; double a = 0.999999, b = 1e-9;
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < ILP; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma_ilp(float* out, long long* cyc, int n) {
  float x[8]; float a = 0.999999f, b = 1e-9f;
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], a, b);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_dep(double* out, long long* cyc, int n) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_ind(double* out, long long* cyc, int n) {
  double x[8]; for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_dep(double* out, long long* cyc, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i + 32) % 1024;
  __syncwarp();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = idx[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_ind(double* out, long long* cyc, int n) {
  __shared__ double buf[2048];
  for (int i = threadIdx.x; i < 2048; i += 32) buf[i] = i;
  __syncwarp();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; u += 4) {
      s0 += buf[threadIdx.x + 32 * u]; s1 += buf[threadIdx.x + 32 * (u + 1)];
      s2 += buf[threadIdx.x + 32 * (u + 2)]; s3 += buf[threadIdx.x + 32 * (u + 3)];
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = s0 + s1 + s2 + s3; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_mix(double* out, long long* cyc, int n) {   // LDS + DFMA stream like the solver rows phase
  __shared__ double buf[2048];
  for (int i = threadIdx.x; i < 2048; i += 32) buf[i] = 1e-3 * i;
  __syncwarp();
  double x[8]; for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fma(x[j], buf[threadIdx.x + 32 * (8 * u + j)], 1e-9);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
  float* outf; cudaMalloc(&outf, 4096);
  const int n = 4096;
#define RUN(name, launch, ops) launch; cudaDeviceSynchronize(); launch; cudaDeviceSynchronize(); printf("%-28s %8.2f cycles/op\n", name, (double)cyc[0] / ((double)n * (ops)));
  RUN("dfma dependent", (k_dfma_dep<<<1, 32>>>(out, cyc, n)), 16)
  RUN("dfma ilp2", (k_dfma_ilp<2><<<1, 32>>>(out, cyc, n)), 8)
  RUN("dfma ilp4", (k_dfma_ilp<4><<<1, 32>>>(out, cyc, n)), 16)
  RUN("dfma ilp8", (k_dfma_ilp<8><<<1, 32>>>(out, cyc, n)), 32)
  RUN("dfma ilp8 1 thread", (k_dfma_ilp<8><<<1, 1>>>(out, cyc, n)), 32)
  RUN("dfma ilp8 2 warps/SMSP(8w)", (k_dfma_ilp<8><<<1, 256>>>(out, cyc, n)), 32)
  RUN("ffma ilp8", (k_ffma_ilp<<<1, 32>>>(outf, cyc, n)), 32)
  RUN("shfl.f64+dadd dependent", (k_shfl_dep<<<1, 32>>>(out, cyc, n)), 16)
  RUN("shfl.f64 independent", (k_shfl_ind<<<1, 32>>>(out, cyc, n)), 16)
  RUN("lds dependent (pointer chase)", (k_lds_dep<<<1, 32>>>(out, cyc, n)), 16)
  RUN("lds.64 independent + dadd", (k_lds_ind<<<1, 32>>>(out, cyc, n)), 16)
  RUN("lds.64 + dfma ilp8", (k_mix<<<1, 32>>>(out, cyc, n)), 16)
  return 0;
}
