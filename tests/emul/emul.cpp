// TEST INFRASTRUCTURE ONLY: compiles intent-mpc_b200/csrc/mpcqp_core.cuh for the host (lane loops become
// plain loops) so the kernel's logic can be checked against the oracle in the CPU-only test tier.
// Never linked into the shipped library; the product has no host solve path.
#define MPCQP_HOST_EMUL 1
#include "../../intent-mpc_b200/csrc/mpcqp_core.cuh"
#include <vector>
#include <cstring>

extern "C" int emul_solve_batch(int linsys, int NS, int R, int B, double a_pv, double b_pa, double b_va, const double* blo,
                                const double* bhi, double obs_hi, const double* settings_d, const int* settings_i,
                                const double* pd, const unsigned char* slack, const double* q, const double* x0,
                                const double* g, const double* low, const double* warm_x, double* x, double* y,
                                int* status, int* iter, int* rho_updates, double* obj, double* pri_res,
                                double* dua_res) {
  using namespace mpcqp;
  Shape sh; sh.NS = NS; sh.R = R; sh.n = 8 * NS + 5 * (NS - 1); sh.m = 16 * NS + 5 * (NS - 1) + R * (NS - 1);
  sh.a_pv = a_pv; sh.b_pa = b_pa; sh.b_va = b_va; sh.obs_hi = obs_hi;
  for (int j = 0; j < NV; ++j) { sh.blo[j] = blo[j]; sh.bhi[j] = bhi[j]; }
  Settings st;
  st.rho = settings_d[0]; st.sigma = settings_d[1]; st.alpha = settings_d[2]; st.eps_abs = settings_d[3];
  st.eps_rel = settings_d[4]; st.eps_prim_inf = settings_d[5]; st.eps_dual_inf = settings_d[6];
  st.adaptive_rho_tolerance = settings_d[7];
  st.max_iter = settings_i[0]; st.scaling = settings_i[1]; st.adaptive_rho = settings_i[2];
  st.adaptive_rho_interval = settings_i[3]; st.check_termination = settings_i[4]; st.warm_start = settings_i[5];
  Batch bt; memset(&bt, 0, sizeof bt);
  bt.pd = pd; bt.slack = slack; bt.q = q; bt.x0 = x0; bt.g = g; bt.low = low; bt.warm_x = warm_x;
  bt.x = x; bt.y = y; bt.status = status; bt.iter = iter; bt.rho_updates = rho_updates; bt.obj = obj;
  bt.pri_res = pri_res; bt.dua_res = dua_res; bt.B = B;
  if (linsys == 0) {
    std::vector<double> sm((size_t)smem_doubles(NS, R, kModeGeneric)), ws((size_t)ws_doubles(NS, R, kModeGeneric));
    Qp<0, 0> qp(sm.data(), sh, st, bt, ws.data(), 0);
    for (int b = 0; b < B; ++b) qp.run(bt, b);
    return 0;
  }
  // linsys 1: the CTA path's linear algebra (PCR factor + plain-loop PCR solve) with the generic iteration around it
  if (NS != 30) return -1;
  std::vector<double> sm((size_t)smem_doubles(NS, R, kModeCta)), ws((size_t)ws_doubles(NS, R, kModeCta));
  switch (R) {
#define X(r) case r: { Qp<30, r, kModeCta> qp(sm.data(), sh, st, bt, ws.data(), 0); for (int b = 0; b < B; ++b) qp.run(bt, b); return 0; }
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
    default: return -2;
  }
}

// ---- generic (unstructured) path: intent-mpc_b200/csrc/mpcqp_dense.cuh with one "thread" -------------------------------
#include "../../intent-mpc_b200/csrc/mpcqp_dense.cuh"

extern "C" int emul_dense_solve(int n, int m, const long long* Pc, const long long* Pi, const double* Px, const long long* Ac,
                                const long long* Ai, const double* Ax, const double* q, const double* l,
                                const double* u, const double* warm_x, const double* warm_y, const double* settings_d,
                                const int* settings_i, double* x, double* y, int* info_i, double* info_d) {
  namespace dq = mpcqp_dense;
  dq::Settings st;
  st.rho = settings_d[0]; st.sigma = settings_d[1]; st.alpha = settings_d[2]; st.eps_abs = settings_d[3];
  st.eps_rel = settings_d[4]; st.eps_prim_inf = settings_d[5]; st.eps_dual_inf = settings_d[6];
  st.adaptive_rho_tolerance = settings_d[7];
  st.max_iter = settings_i[0]; st.scaling = settings_i[1]; st.adaptive_rho = settings_i[2];
  st.adaptive_rho_interval = settings_i[3]; st.check_termination = settings_i[4]; st.warm_start = settings_i[5];
  std::vector<double> ws(dq::ws_doubles(n, m)), sm(dq::smem_doubles(n, m));
  int32_t ii[3];
  dq::Problem pb;
  pb.n = n; pb.m = m; pb.Pc = (const int64_t*)Pc; pb.Pi = (const int64_t*)Pi; pb.Px = Px; pb.Ac = (const int64_t*)Ac; pb.Ai = (const int64_t*)Ai; pb.Ax = Ax;
  pb.q0 = q; pb.l0 = l; pb.u0 = u;
  pb.warm_x = warm_x; pb.warm_y = warm_y; pb.ws = ws.data(); pb.x = x; pb.y = y; pb.info_i = ii; pb.info_d = info_d;
  dq::Solver sv;
  sv.run(pb, st, sm.data(), 0, 1);
  info_i[0] = ii[0]; info_i[1] = ii[1]; info_i[2] = ii[2];
  return 0;
}

// ---- sparse generic path: intent-mpc_b200/csrc/mpcqp_band.cuh with one "lane" (host pattern analysis as the library does) ----
#include "../../intent-mpc_b200/csrc/mpcqp_band_host.hpp"

// returns the half-bandwidth found (solve done), or -(bandwidth) when the pattern is not eligible for the band path
extern "C" int emul_band_solve(int n, int m, const long long* Pc, const long long* Pi, const double* Px, const long long* Ac,
                               const long long* Ai, const double* Ax, const double* q, const double* l,
                               const double* u, const double* warm_x, const double* warm_y, const double* settings_d,
                               const int* settings_i, double* x, double* y, int* info_i, double* info_d) {
  namespace bq = mpcqp_band;
  bq::Settings st;
  st.rho = settings_d[0]; st.sigma = settings_d[1]; st.alpha = settings_d[2]; st.eps_abs = settings_d[3];
  st.eps_rel = settings_d[4]; st.eps_prim_inf = settings_d[5]; st.eps_dual_inf = settings_d[6];
  st.adaptive_rho_tolerance = settings_d[7];
  st.max_iter = settings_i[0]; st.scaling = settings_i[1]; st.adaptive_rho = settings_i[2];
  st.adaptive_rho_interval = settings_i[3]; st.check_termination = settings_i[4]; st.warm_start = settings_i[5];
  std::vector<int> flat; int off[14], w = 0;
  if (!bq::pattern_build(n, m, (const int64_t*)Pc, (const int64_t*)Pi, (const int64_t*)Ac, (const int64_t*)Ai, &flat, off, &w)) return -w;
  bq::Batch bt; memset(&bt, 0, sizeof bt);
  bt.B = 1;
  bq::Pattern& pt = bt.pt;
  pt.n = n; pt.m = m; pt.N = n + m; pt.w = w; pt.nnzP = (int)Pc[n]; pt.nnzA = (int)Ac[n];
  const int* f = flat.data();
  pt.Pc = f + off[0]; pt.Pi = f + off[1]; pt.Ac = f + off[2]; pt.Ai = f + off[3]; pt.Pr_ptr = f + off[4]; pt.Pr_pos = f + off[5]; pt.Pr_col = f + off[6];
  pt.Ar_ptr = f + off[7]; pt.Ar_pos = f + off[8]; pt.Ar_col = f + off[9]; pt.slotP = f + off[10]; pt.slotA = f + off[11]; pt.perm = f + off[12]; pt.iperm = f + off[13];
  bt.Px = Px; bt.Ax = Ax; bt.q = q; bt.l = l; bt.u = u; bt.warm_x = warm_x; bt.warm_y = warm_y; bt.x = x; bt.y = y;
  int32_t ii[3]; bt.info_i = ii; bt.info_d = info_d;
  std::vector<double> ws(bq::ws_doubles(n, m, pt.nnzP, pt.nnzA)), sm(bq::smem_doubles(n + m, w));
  bq::Solver sv;
  sv.run(bt, 0, st, ws.data(), sm.data(), 0);
  info_i[0] = ii[0]; info_i[1] = ii[1]; info_i[2] = ii[2];
  return w;
}
