// Test driver for intent-mpc_b200/host/MultiGpuB200.hpp (one process, one host thread + one engine per worker):
//   multi_gpu_test <workers> <in.bin> <out.bin> [repeats]
// in.bin: int64 header {B, R, horizon} then x0, xref, obs_c, obs_semi, obs_yaw, lin_pt, warm_x (doubles) and obs_dyn (int32 [N][R]).
// out.bin: x [B][n] doubles, then status, iter, rho_updates as doubles [B] each, obj [B], then per-worker kernel ms.
// Worker g runs on CUDA device g % device_count, so on a one-GPU box several engines share the device and on an 8-GPU box every
// worker has its own.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../intent-mpc_b200/host/MultiGpuB200.hpp"

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: multi_gpu_test <workers> <in.bin> <out.bin> [repeats]\n"); return 1; }
  const int G = atoi(argv[1]); const int reps = argc > 4 ? atoi(argv[4]) : 1;
  FILE* f = fopen(argv[2], "rb"); if (!f) return 2;
  std::vector<long long> h = rd<long long>(f, 3);
  const int B = (int)h[0], R = (int)h[1], NS = (int)h[2], N = NS - 1, n = 8 * NS + 5 * N;
  auto x0 = rd<double>(f, (size_t)B * 6), xref = rd<double>(f, (size_t)B * NS * 3), oc = rd<double>(f, (size_t)B * N * R * 3), os = rd<double>(f, (size_t)B * N * R * 3),
       oy = rd<double>(f, (size_t)B * N * R), lp = rd<double>(f, (size_t)B * N * 3), wx = rd<double>(f, (size_t)B * n);
  auto od = rd<int32_t>(f, (size_t)N * R);
  fclose(f);
  const int ndev = mpcqp_device_count();
  if (ndev < 1) { fprintf(stderr, "no CUDA device (no CPU fallback)\n"); return 3; }
  std::vector<int> devs; for (int g = 0; g < G; ++g) devs.push_back(g % ndev);
  mpcqpB200::MultiGpuBatchSolver solver(devs);
  if (!solver.ok()) { fprintf(stderr, "%s\n", solver.lastError().c_str()); return 4; }
  mpcqp_mpc_params p; mpcqp_default_mpc_params(&p); p.horizon = NS;
  mpcqp_settings s; mpcqp_set_default_settings(&s);
  std::vector<double> x((size_t)B * n), obj((size_t)B), pr((size_t)B), du((size_t)B);
  std::vector<int32_t> st((size_t)B), it((size_t)B), ru((size_t)B);
  for (int r = 0; r < reps; ++r) {
    const int rc = solver.solveMpcBatch(&p, &s, B, R, x0.data(), xref.data(), R ? oc.data() : nullptr, R ? os.data() : nullptr, R ? oy.data() : nullptr, R ? od.data() : nullptr,
                                        lp.data(), wx.data(), x.data(), nullptr, st.data(), it.data(), ru.data(), obj.data(), pr.data(), du.data());
    if (rc != MPCQP_OK) { fprintf(stderr, "solve failed (%d): %s\n", rc, solver.lastError().c_str()); return 5; }
  }
  FILE* g = fopen(argv[3], "wb"); if (!g) return 6;
  fwrite(x.data(), 8, x.size(), g);
  std::vector<double> t((size_t)B);
  for (const std::vector<int32_t>* v : {&st, &it, &ru}) {
    for (int b = 0; b < B; ++b) t[(size_t)b] = (*v)[(size_t)b];
    fwrite(t.data(), 8, t.size(), g);
  }
  fwrite(obj.data(), 8, obj.size(), g);
  for (int w = 0; w < G; ++w) { double k = solver.kernelMs(w); fwrite(&k, 8, 1, g); }
  fclose(g);
  printf("workers %d devices %d kernel ms max %.3f\n", G, ndev, solver.lastKernelMsMax());
  return 0;
}
