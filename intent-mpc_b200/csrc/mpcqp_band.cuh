// mpcqp_band.cuh — the SPARSE generic solve path behind mpcqp_setup / mpcqp_solve / mpcqp_solve_qp_batch_host: convex QPs in
// CSC form WITHOUT the mpcPlanner stage structure whose KKT matrix [P + sigma I, A'; A, -diag(1/rho)] (third_party/osqp/
// kkt.h:15-18) becomes narrow-banded under a reverse Cuthill-McKee ordering — which is what the reference's second consumer
// of the OsqpEigen::Solver boundary produces: polyTrajSolver's minimum-snap QPs (trajectory_planner/include/
// trajectory_planner/polyTrajSolver.cpp:14-37, 162-239, 848-900) chain 8-coefficient segment blocks through continuity rows;
// their KKT systems (n + m = 42 ... 350 for K = 3 ... 25 segments) have half-bandwidth 8 ... 12 after RCM.
//
// ONE WARP per QP, many warps per SM.  Same iterate sequence as OSQP 0.6.2 (auxil.h:21-154, constants.h:59-118) — Ruiz
// equilibration on the CSC data in OSQP's own per-entry operation order (scaling.h: scale_data), rho vector by constraint
// class, ADMM with relaxation, unscaled residual test, both infeasibility certificates, rho adaptation with
// re-factorisation, OSQP_NAN fill — around a banded L D L' of the permuted KKT matrix held in SHARED memory (N (w+1) doubles:
// 30 KB at K = 25) and two substitution sweeps per iteration in which lane (row mod 32) owns the running sum of its row: one
// broadcast shuffle and one FMA per row, no reduction.  Everything else (vectors, the scaled CSC values) sits in the warp's
// private L2-resident workspace.  Problems whose bandwidth exceeds 31 or whose band does not fit shared memory go to the
// dense kernel (mpcqp_dense.cuh).
//
// The file also compiles for the host (MPCQP_HOST_EMUL: one "lane", shuffles become plain reads) so that the CPU test tier can
// check the logic against the reference binary's golden vectors; the shipped library has no host solve path.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef MPCQP_HOST_EMUL
#define BQ_FN inline
#define BQ_SYNC() ((void)0)
#define BQ_LANES 1
#else
#define BQ_FN __device__ __forceinline__
#define BQ_SYNC() __syncwarp()
#define BQ_LANES 32
#endif
#define BQ_FOR(i, cnt) for (int i = lane; i < (cnt); i += BQ_LANES)

namespace mpcqp_band {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4, kMinScaling = 1e-4,
                 kMaxScaling = 1e4, kInfty = 1e30, kOsqpNan = 2143289344.0;   // OSQP_NAN is this NUMBER (constants.h:95-97)
enum { kSolved = 1, kSolvedInacc = 2, kPrimInfInacc = 3, kDualInfInacc = 4, kMaxIter = -2, kPrimInf = -3, kDualInf = -4,
       kNonCvx = -7, kUnsolved = -10 };
constexpr int kMaxBand = 31;      // half-bandwidth limit of the lane-per-row substitution

struct Settings {   // same layout as mpcqp::Settings (types.h:139-176 subset)
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination, warm_start;
};

// Built on the host from the CSC pattern (shared by the batch).  All index arrays are int32.
struct Pattern {
  int n, m, N, w, nnzP, nnzA;
  const int *Pc, *Pi;            // P upper triangle, CSC: column pointers [n+1], row indices [nnzP]
  const int *Ac, *Ai;            // A, CSC
  const int *Pr_ptr, *Pr_pos;    // rows of the STRICTLY upper part of P: row j -> CSC positions of entries (j, k > j), ascending k
  const int *Pr_col;             //   and their columns k
  const int *Ar_ptr, *Ar_pos;    // rows of A: row i -> CSC positions of its entries, ascending column
  const int *Ar_col;             //   and their columns
  const int *slotP, *slotA;      // band slot r (w+1) + (w - (r - c)) of the lower-triangle image (r >= c, permuted) of every entry
  const int *perm, *iperm;       // perm[new] = old KKT index (0..n-1 variables, n..N-1 constraints); iperm[old] = new
};

struct Batch {
  int B;
  Pattern pt;
  const double *Px, *Ax, *q, *l, *u;          // [B][nnzP], [B][nnzA], [B][n], [B][m], [B][m]
  const double *warm_x, *warm_y;              // [B][n], [B][m] or nullptr
  double* ws; long long ws_stride;            // per resident warp
  double *x, *y;                              // [B][n], [B][m] or nullptr
  int32_t* info_i; double* info_d;            // [B][3] each
};

inline size_t ws_doubles(int n, int m, int nnzP, int nnzA) { return (size_t)nnzP + nnzA + 14 * (size_t)n + 18 * (size_t)m + 64; }
inline size_t smem_doubles(int N, int w) { return (size_t)N * (w + 1) + 2 * (size_t)N; }

struct Solver {
  int lane;
  Pattern p;
  int n, m, N, w, W1;
  Settings s;
  double *Px, *Ax, *q, *l, *u;                                   // scaled copies (global)
  double *D, *Dinv, *E, *Einv, *rho, *rho_inv, *ctype;
  double *x, *z, *y, *xp, *zp, *Axv, *Pxv, *Aty, *dy, *dx, *Atdy, *Pdx, *Adx, *tmpn, *tmpm;
  double *Lb, *dinv, *sol;                                       // shared: band factor (unit L below the diagonal), 1 / D, rhs -> solution
  double c, cinv, pri_res, dua_res, obj;
  int status, rho_updates;

  // ---- warp-wide reductions; every lane gets the result ---------------------------------------------------------------
  BQ_FN double wmax(double v) const {
#ifndef MPCQP_HOST_EMUL
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
    return v;
  }
  BQ_FN double wsum(double v) const {
#ifndef MPCQP_HOST_EMUL
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
  BQ_FN double norm_inf(const double* v, int cnt) const { double a = 0; BQ_FOR(i, cnt) a = fmax(a, fabs(v[i])); return wmax(a); }
  BQ_FN double scaled_norm_inf(const double* S, const double* v, int cnt) const { double a = 0; BQ_FOR(i, cnt) a = fmax(a, fabs(S[i] * v[i])); return wmax(a); }
  static BQ_FN double limit_scaling(double v) { v = v < kMinScaling ? 1.0 : v; return v > kMaxScaling ? kMaxScaling : v; }

  // ---- sparse mat-vecs: one lane per output entry, gathers in ascending index order (deterministic) -------------------------
  BQ_FN void mv_A(const double* v, double* out) const {
    BQ_FOR(i, m) { double a = 0; for (int t = p.Ar_ptr[i]; t < p.Ar_ptr[i + 1]; ++t) a += Ax[p.Ar_pos[t]] * v[p.Ar_col[t]]; out[i] = a; }
    BQ_SYNC();
  }
  BQ_FN void mv_At(const double* v, double* out) const {
    BQ_FOR(j, n) { double a = 0; for (int t = p.Ac[j]; t < p.Ac[j + 1]; ++t) a += Ax[t] * v[p.Ai[t]]; out[j] = a; }
    BQ_SYNC();
  }
  BQ_FN void mv_P(const double* v, double* out) const {      // symmetric P from its upper triangle
    BQ_FOR(j, n) {
      double a = 0;
      for (int t = p.Pc[j]; t < p.Pc[j + 1]; ++t) a += Px[t] * v[p.Pi[t]];                       // entries (i <= j, j)
      for (int t = p.Pr_ptr[j]; t < p.Pr_ptr[j + 1]; ++t) a += Px[p.Pr_pos[t]] * v[p.Pr_col[t]];  // entries (j, k > j)
      out[j] = a;
    }
    BQ_SYNC();
  }
  // column inf-norms of the symmetric P (upper triangle stored)
  BQ_FN double p_col_norm(int j) const {
    double a = 0;
    for (int t = p.Pc[j]; t < p.Pc[j + 1]; ++t) a = fmax(a, fabs(Px[t]));
    for (int t = p.Pr_ptr[j]; t < p.Pr_ptr[j + 1]; ++t) a = fmax(a, fabs(Px[p.Pr_pos[t]]));
    return a;
  }

  BQ_FN void carve(double* ws, double* smem) {
    n = p.n; m = p.m; N = p.N; w = p.w; W1 = w + 1;
    double* g = ws;
    Px = g; g += p.nnzP; Ax = g; g += p.nnzA; q = g; g += n; l = g; g += m; u = g; g += m;
    D = g; g += n; Dinv = g; g += n; x = g; g += n; xp = g; g += n; Pxv = g; g += n; Aty = g; g += n; dx = g; g += n; Atdy = g; g += n;
    Pdx = g; g += n; tmpn = g; g += n;
    E = g; g += m; Einv = g; g += m; rho = g; g += m; rho_inv = g; g += m; ctype = g; g += m; z = g; g += m; zp = g; g += m; y = g; g += m;
    Axv = g; g += m; dy = g; g += m; Adx = g; g += m; tmpm = g; g += m;
    Lb = smem; dinv = smem + (size_t)N * W1; sol = dinv + N;
  }

  // ---- scaling.h: scale_data (Ruiz passes + cost normalisation) on the CSC values, OSQP's per-entry operation order -----------
  BQ_FN void scale_data() {
    BQ_FOR(j, n) D[j] = 1.0;
    BQ_FOR(i, m) E[i] = 1.0;
    c = 1.0;
    for (int pass = 0; pass < s.scaling; ++pass) {
      BQ_SYNC();
      BQ_FOR(j, n) {        // column inf-norms of [P A'; A 0], first n columns
        double a = p_col_norm(j);
        for (int t = p.Ac[j]; t < p.Ac[j + 1]; ++t) a = fmax(a, fabs(Ax[t]));
        tmpn[j] = 1.0 / sqrt(limit_scaling(a));
      }
      BQ_FOR(i, m) {        // last m columns = rows of A
        double a = 0;
        for (int t = p.Ar_ptr[i]; t < p.Ar_ptr[i + 1]; ++t) a = fmax(a, fabs(Ax[p.Ar_pos[t]]));
        tmpm[i] = 1.0 / sqrt(limit_scaling(a));
      }
      BQ_SYNC();
      BQ_FOR(j, n) {        // P <- D P D, A <- E A D: row factor first, then column factor (mat_premult_diag, mat_postmult_diag)
        for (int t = p.Pc[j]; t < p.Pc[j + 1]; ++t) Px[t] = (Px[t] * tmpn[p.Pi[t]]) * tmpn[j];
        for (int t = p.Ac[j]; t < p.Ac[j + 1]; ++t) Ax[t] = (Ax[t] * tmpm[p.Ai[t]]) * tmpn[j];
        q[j] *= tmpn[j]; D[j] *= tmpn[j];
      }
      BQ_FOR(i, m) E[i] *= tmpm[i];
      BQ_SYNC();
      double part = 0, nq = 0;
      BQ_FOR(j, n) { part += p_col_norm(j); nq = fmax(nq, fabs(q[j])); }
      double c_temp = wsum(part) / (double)n;
      nq = limit_scaling(wmax(nq));
      if (nq > c_temp) c_temp = nq;
      c_temp = 1.0 / limit_scaling(c_temp);
      BQ_SYNC();
      BQ_FOR(t, p.nnzP) Px[t] *= c_temp;
      BQ_FOR(j, n) q[j] *= c_temp;
      c *= c_temp;
    }
    BQ_SYNC();
    cinv = 1.0 / c;
    BQ_FOR(j, n) Dinv[j] = 1.0 / D[j];
    BQ_FOR(i, m) { Einv[i] = 1.0 / E[i]; l[i] *= E[i]; u[i] *= E[i]; }
    BQ_SYNC();
  }

  // ---- auxil.h: set_rho_vec (classes from the SCALED bounds) -------------------------------------------------------------
  BQ_FN void set_rho_vec() {
    s.rho = fmin(fmax(s.rho, kRhoMin), kRhoMax);
    BQ_FOR(i, m) {
      if (l[i] < -kInfty * kMinScaling && u[i] > kInfty * kMinScaling) { ctype[i] = -1; rho[i] = kRhoMin; }
      else if (u[i] - l[i] < kRhoTol) { ctype[i] = 1; rho[i] = kRhoEqOverIneq * s.rho; }
      else { ctype[i] = 0; rho[i] = s.rho; }
      rho_inv[i] = 1.0 / rho[i];
    }
    BQ_SYNC();
  }

  // ---- permuted KKT matrix into the band, banded L D L' in place ------------------------------------------------------------
  // Lb[r (w+1) + (w - d)] = K(r, r - d), d = 0 .. w.  After the factorisation the diagonal slot holds D_r and the others the unit
  // lower factor L(r, r - d); dinv[r] = 1 / D_r.
  BQ_FN void factor() {
    BQ_SYNC();
    BQ_FOR(t, N * W1) Lb[t] = 0.0;
    BQ_SYNC();
    BQ_FOR(r, N) { const int o = p.perm[r]; Lb[r * W1 + w] = o < n ? s.sigma : -rho_inv[o - n]; }
    BQ_SYNC();
    // one entry per slot in a duplicate-free CSC pattern (what OsqpEigen hands over); lanes take disjoint entries
#ifdef MPCQP_HOST_EMUL
    for (int t = 0; t < p.nnzP; ++t) Lb[p.slotP[t]] += Px[t];
    for (int t = 0; t < p.nnzA; ++t) Lb[p.slotA[t]] += Ax[t];
#else
    BQ_FOR(t, p.nnzP) atomicAdd(&Lb[p.slotP[t]], Px[t]);
    BQ_FOR(t, p.nnzA) atomicAdd(&Lb[p.slotA[t]], Ax[t]);
#endif
    BQ_SYNC();
    // right-looking: pivot k, column c_i = K(k+i, k), i = 1..w; K(k+i, k+j) -= c_i c_j / d for 1 <= j <= i; then L(k+i, k) = c_i / d.
    // Lane i-1 owns row k+i; the pivot column is read through `sol` (shared, broadcast reads).
    for (int k = 0; k < N; ++k) {
      const double d = Lb[k * W1 + w];
      const double di = 1.0 / d;
      const int cnt = (N - 1 - k) < w ? (N - 1 - k) : w;
      BQ_FOR(i0, cnt) sol[i0] = Lb[(k + 1 + i0) * W1 + (w - 1 - i0)];
      BQ_SYNC();
      BQ_FOR(i0, cnt) {
        const double ci = sol[i0] * di;
        double* row = Lb + (k + 1 + i0) * W1 + (w - i0);          // slot of column k+1 in row k+1+i0 ... up to the diagonal
        for (int j0 = 0; j0 <= i0; ++j0) row[j0] -= ci * sol[j0];
        Lb[(k + 1 + i0) * W1 + (w - 1 - i0)] = ci;
      }
      if (lane == 0) dinv[k] = di;
      BQ_SYNC();
    }
  }

  // sol <- K^-1 sol (permuted order), in place: L y = b, y /= D, L' x = y.  Lane (row mod 32) owns the running sum of its row.
  BQ_FN void band_solve() {
    BQ_SYNC();
#ifdef MPCQP_HOST_EMUL
    for (int k = 0; k < N; ++k) { double a = sol[k]; for (int d = 1; d <= w && d <= k; ++d) a -= Lb[k * W1 + (w - d)] * sol[k - d]; sol[k] = a; }
    for (int k = 0; k < N; ++k) sol[k] *= dinv[k];
    for (int k = N - 1; k >= 0; --k) { double a = sol[k]; for (int d = 1; d <= w && k + d < N; ++d) a -= Lb[(k + d) * W1 + (w - d)] * sol[k + d]; sol[k] = a; }
#else
    {   // forward: when y_k is final every row i in (k, k + w] adds L(i, k) y_k to its running sum
      double acc = 0.0;
      for (int k = 0; k < N; ++k) {
        const int owner = k & 31;
        const int dd = (lane - owner) & 31;                      // this lane's row is k + dd
        const int i = k + dd;
        const double lik = (dd >= 1 && dd <= w && i < N) ? Lb[i * W1 + (w - dd)] : 0.0;
        double yk = sol[k] - acc;                                // meaningful on the owner lane
        yk = __shfl_sync(0xffffffffu, yk, owner);
        if (lane == owner) { sol[k] = yk; acc = 0.0; } else acc = fma(lik, yk, acc);
      }
    }
    __syncwarp();
    for (int k = lane; k < N; k += 32) sol[k] *= dinv[k];
    __syncwarp();
    {   // backward: when x_i is final every row k in [i - w, i) adds L(i, k) x_i (row i of the band: contiguous)
      double acc = 0.0;
      for (int i = N - 1; i >= 0; --i) {
        const int owner = i & 31;
        const int dd = (owner - lane) & 31;                      // this lane's row is i - dd
        const double lik = (dd >= 1 && dd <= w && i - dd >= 0) ? Lb[i * W1 + (w - dd)] : 0.0;
        double xi = sol[i] - acc;
        xi = __shfl_sync(0xffffffffu, xi, owner);
        if (lane == owner) { sol[i] = xi; acc = 0.0; } else acc = fma(lik, xi, acc);
      }
    }
#endif
    BQ_SYNC();
  }

  // ---- auxil.h:67-112, one ADMM iteration ------------------------------------------------------------------------------
  BQ_FN void iterate() {
    double* t;
    t = x; x = xp; xp = t;
    t = z; z = zp; zp = t;
    BQ_FOR(r, N) { const int o = p.perm[r]; sol[r] = o < n ? s.sigma * xp[o] - q[o] : zp[o - n] - rho_inv[o - n] * y[o - n]; }
    band_solve();
    BQ_FOR(i, n) {
      const double xt = sol[p.iperm[i]];
      x[i] = s.alpha * xt + (1.0 - s.alpha) * xp[i];
      dx[i] = x[i] - xp[i];
    }
    BQ_FOR(i, m) {
      const double nu = sol[p.iperm[n + i]];
      const double zt = (zp[i] - rho_inv[i] * y[i]) + rho_inv[i] * nu;
      const double zr = s.alpha * zt + (1.0 - s.alpha) * zp[i];
      const double zn = fmin(fmax(zr + rho_inv[i] * y[i], l[i]), u[i]);      // proj.h: project
      z[i] = zn;
      dy[i] = rho[i] * (zr - zn);
      y[i] += dy[i];
    }
    BQ_SYNC();
  }

  // ---- update_info: unscaled residuals; the scaled ones stay in zp / xp for the rho estimate -------------------------------
  BQ_FN void update_info() {
    BQ_SYNC();
    if (m > 0) {
      mv_A(x, Axv);
      BQ_FOR(i, m) zp[i] = Axv[i] - z[i];
      BQ_SYNC();
      pri_res = s.scaling ? scaled_norm_inf(Einv, zp, m) : norm_inf(zp, m);
    } else pri_res = 0.0;
    mv_P(x, Pxv);
    mv_At(y, Aty);
    BQ_FOR(j, n) xp[j] = (q[j] + Pxv[j]) + Aty[j];
    BQ_SYNC();
    dua_res = s.scaling ? cinv * scaled_norm_inf(Dinv, xp, n) : norm_inf(xp, n);
  }

  BQ_FN bool is_primal_infeasible(double eps) {
    BQ_FOR(i, m) {
      if (u[i] > kInfty * kMinScaling) { if (l[i] < -kInfty * kMinScaling) dy[i] = 0.0; else dy[i] = fmin(dy[i], 0.0); }
      else if (l[i] < -kInfty * kMinScaling) dy[i] = fmax(dy[i], 0.0);
    }
    BQ_SYNC();
    const double norm_dy = s.scaling ? scaled_norm_inf(E, dy, m) : norm_inf(dy, m);
    if (norm_dy > eps) {
      double part = 0;       // IEEE: +inf * 0 = NaN keeps the comparison false, as in OSQP
      BQ_FOR(i, m) part += u[i] * fmax(dy[i], 0.0) + l[i] * fmin(dy[i], 0.0);
      const double lhs = wsum(part);
      if (lhs < -eps * norm_dy) {
        mv_At(dy, Atdy);
        const double nn = s.scaling ? scaled_norm_inf(Dinv, Atdy, n) : norm_inf(Atdy, n);
        return nn < eps * norm_dy;
      }
    }
    return false;
  }

  BQ_FN bool is_dual_infeasible(double eps) {
    const double norm_dx = s.scaling ? scaled_norm_inf(D, dx, n) : norm_inf(dx, n);
    const double cs = s.scaling ? c : 1.0;
    if (norm_dx > eps) {
      double part = 0;
      BQ_FOR(j, n) part += q[j] * dx[j];
      if (wsum(part) < -cs * eps * norm_dx) {
        mv_P(dx, Pdx);
        const double np = s.scaling ? scaled_norm_inf(Dinv, Pdx, n) : norm_inf(Pdx, n);
        if (np < cs * eps * norm_dx) {
          mv_A(dx, Adx);
          double bad = 0;
          BQ_FOR(i, m) {
            const double a = s.scaling ? Adx[i] * Einv[i] : Adx[i];
            if ((u[i] < kInfty * kMinScaling && a > eps * norm_dx) || (l[i] > -kInfty * kMinScaling && a < -eps * norm_dx)) bad = 1.0;
          }
          return wmax(bad) == 0.0;
        }
      }
    }
    return false;
  }

  BQ_FN bool check_termination(bool approximate) {
    double eps_abs = s.eps_abs, eps_rel = s.eps_rel, epi = s.eps_prim_inf, edi = s.eps_dual_inf;
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (pri_res > kInfty || dua_res > kInfty) { status = kNonCvx; obj = kOsqpNan; return true; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
    if (m == 0) prim_ok = true;
    else {
      const double a = s.scaling ? scaled_norm_inf(Einv, z, m) : norm_inf(z, m);
      const double b = s.scaling ? scaled_norm_inf(Einv, Axv, m) : norm_inf(Axv, m);
      if (pri_res < eps_abs + eps_rel * fmax(a, b)) prim_ok = true; else prim_inf = is_primal_infeasible(epi);
    }
    double t;
    if (s.scaling) {
      t = scaled_norm_inf(Dinv, q, n);
      t = fmax(t, scaled_norm_inf(Dinv, Aty, n));
      t = fmax(t, scaled_norm_inf(Dinv, Pxv, n));
      t *= cinv;
    } else { t = norm_inf(q, n); t = fmax(t, norm_inf(Aty, n)); t = fmax(t, norm_inf(Pxv, n)); }
    if (dua_res < eps_abs + eps_rel * t) dual_ok = true; else dual_inf = is_dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approximate ? kSolvedInacc : kSolved; return true; }
    if (prim_inf) { status = approximate ? kPrimInfInacc : kPrimInf; obj = kInfty; return true; }
    if (dual_inf) { status = approximate ? kDualInfInacc : kDualInf; obj = -kInfty; return true; }
    return false;
  }

  BQ_FN void adapt_rho() {
    double pr = norm_inf(zp, m), du = norm_inf(xp, n);
    const double pn = fmax(norm_inf(z, m), norm_inf(Axv, m));
    pr /= (pn + 1e-10);
    double dn = norm_inf(q, n); dn = fmax(dn, norm_inf(Aty, n)); dn = fmax(dn, norm_inf(Pxv, n));
    du /= (dn + 1e-10);
    const double rho_new = fmin(fmax(s.rho * sqrt(pr / (du + 1e-10)), kRhoMin), kRhoMax);
    if (rho_new > s.rho * s.adaptive_rho_tolerance || rho_new < s.rho / s.adaptive_rho_tolerance) {
      s.rho = rho_new;
      BQ_FOR(i, m) {
        if (ctype[i] == 0) { rho[i] = s.rho; rho_inv[i] = 1.0 / s.rho; }
        else if (ctype[i] == 1) { rho[i] = kRhoEqOverIneq * s.rho; rho_inv[i] = 1.0 / rho[i]; }
      }
      BQ_SYNC();
      factor();
      rho_updates += 1;
    }
  }

  BQ_FN bool has_solution() const {
    return status != kPrimInf && status != kPrimInfInacc && status != kDualInf && status != kDualInfInacc && status != kNonCvx;
  }

  // ---- osqp_setup + osqp_warm_start + osqp_solve + store_solution for problem b of the batch ---------------------------------
  BQ_FN void run(const Batch& bt, int b, const Settings& st, double* ws, double* smem, int lane_) {
    lane = lane_; s = st; p = bt.pt;
    carve(ws, smem);
    const double* Px0 = bt.Px + (size_t)b * p.nnzP; const double* Ax0 = bt.Ax + (size_t)b * p.nnzA;
    const double* q0 = bt.q + (size_t)b * n; const double* l0 = bt.l + (size_t)b * m; const double* u0 = bt.u + (size_t)b * m;
    const double* wx = bt.warm_x ? bt.warm_x + (size_t)b * n : nullptr; const double* wy = bt.warm_y ? bt.warm_y + (size_t)b * m : nullptr;
    BQ_FOR(t, p.nnzP) Px[t] = Px0[t];
    BQ_FOR(t, p.nnzA) Ax[t] = Ax0[t];
    BQ_FOR(j, n) q[j] = q0[j];
    BQ_FOR(i, m) { l[i] = l0[i]; u[i] = u0[i]; }
    BQ_SYNC();
    if (s.scaling) scale_data();
    else { c = cinv = 1.0; BQ_FOR(j, n) D[j] = Dinv[j] = 1.0; BQ_FOR(i, m) E[i] = Einv[i] = 1.0; BQ_SYNC(); }
    set_rho_vec();
    factor();
    // iterates: cold start or osqp_warm_start (x <- D^-1 x, y <- c E^-1 y, z <- A x)
    BQ_FOR(j, n) { x[j] = (wx && s.warm_start) ? wx[j] * Dinv[j] : 0.0; xp[j] = 0.0; }
    BQ_FOR(i, m) { y[i] = (wy && s.warm_start) ? (wy[i] * Einv[i]) * c : 0.0; zp[i] = 0.0; z[i] = 0.0; }
    BQ_SYNC();
    if (wx && s.warm_start) mv_A(x, z);
    status = kUnsolved; rho_updates = 0; obj = 0.0; pri_res = dua_res = 0.0;
    int iter, last = 0; bool can_check = false, done = false;
    for (iter = 1; iter <= s.max_iter; ++iter) {
      iterate();
      last = iter;
      can_check = s.check_termination && (iter % s.check_termination == 0);
      if (can_check) { update_info(); if (check_termination(false)) { done = true; break; } }
      if (s.adaptive_rho && s.adaptive_rho_interval && (iter % s.adaptive_rho_interval == 0)) {
        if (!can_check) update_info();
        adapt_rho();
      }
    }
    if (!done && !can_check) { update_info(); check_termination(false); }
    if (has_solution()) {
      mv_P(x, Pxv);
      double part = 0;
      BQ_FOR(j, n) part += (0.5 * Pxv[j] + q[j]) * x[j];
      obj = wsum(part) * cinv;
    }
    if (status == kUnsolved) { if (!check_termination(true)) status = kMaxIter; }
    // store_solution (auxil.h:118)
    const bool ok = has_solution();
    double* xo = bt.x + (size_t)b * n; double* yo = bt.y ? bt.y + (size_t)b * m : nullptr;
    BQ_FOR(j, n) xo[j] = ok ? (s.scaling ? D[j] * x[j] : x[j]) : kOsqpNan;
    if (yo) BQ_FOR(i, m) yo[i] = ok ? (s.scaling ? E[i] * y[i] * cinv : y[i]) : kOsqpNan;
    if (lane == 0) {
      bt.info_i[3 * (size_t)b] = status; bt.info_i[3 * (size_t)b + 1] = last; bt.info_i[3 * (size_t)b + 2] = rho_updates;
      bt.info_d[3 * (size_t)b] = obj; bt.info_d[3 * (size_t)b + 1] = pri_res; bt.info_d[3 * (size_t)b + 2] = dua_res;
    }
    BQ_SYNC();
  }
};

}  // namespace mpcqp_band
