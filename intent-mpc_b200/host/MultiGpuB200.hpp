// MultiGpuB200.hpp — one process, one host thread per GPU (BASELINE.json north_star: "independent QPs
// shard by batch index across the 8 B200s of one box, one host thread per GPU, with no inter-GPU traffic on the solve path").
// Uses include/mpcqp_b200.h only.  Every worker thread owns one engine (its own CUDA stream, events and device staging
// buffers) and solves the contiguous index range [g*B/G, (g+1)*B/G) of each batch through mpcqp_solve_mpc_batch_host; the
// caller's arrays are read and written in place (disjoint ranges, no copies between threads, no collective).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mpcqp_b200.h"

namespace mpcqpB200 {

class MultiGpuBatchSolver {
 public:
  // devices[g] = CUDA device of worker g (a device may appear more than once: several engines on one GPU)
  explicit MultiGpuBatchSolver(const std::vector<int>& devices) : w_(devices.size()) {
    for (size_t g = 0; g < w_.size(); ++g) { w_[g].device = devices[g]; w_[g].th = std::thread([this, g] { loop(g); }); }
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { for (auto& w : w_) if (!w.ready) return false; return true; });
  }
  ~MultiGpuBatchSolver() {
    { std::lock_guard<std::mutex> lk(mu_); quit_ = true; ++epoch_; }
    go_.notify_all();
    for (auto& w : w_) if (w.th.joinable()) w.th.join();
  }
  MultiGpuBatchSolver(const MultiGpuBatchSolver&) = delete;
  MultiGpuBatchSolver& operator=(const MultiGpuBatchSolver&) = delete;

  bool ok() const { for (auto& w : w_) if (!w.eng) return false; return !w_.empty(); }
  int numWorkers() const { return (int)w_.size(); }
  std::string lastError() const { for (auto& w : w_) if (!w.err.empty()) return w.err; return ""; }
  // shard g of B instances over G workers: [g*B/G, (g+1)*B/G)
  static void shardBounds(int64_t B, int G, int g, int64_t* lo, int64_t* hi) { *lo = B * g / G; *hi = B * (g + 1) / G; }
  double lastKernelMsMax() const { double m = 0; for (auto& w : w_) if (w.kernel_ms > m) m = w.kernel_ms; return m; }
  double kernelMs(int g) const { return w_[(size_t)g].kernel_ms; }

  // Same arguments as mpcqp_solve_mpc_batch_host (obs_dyn: the batch-uniform [N][R] pattern).  Returns MPCQP_OK or the first
  // worker's error code.
  int solveMpcBatch(const mpcqp_mpc_params* p, const mpcqp_settings* s, int32_t B, int32_t num_obs, const double* x0, const double* xref,
                    const double* obs_c, const double* obs_semi, const double* obs_yaw, const int32_t* obs_dyn, const double* lin_pt,
                    const double* warm_x, double* x, double* y, int32_t* status, int32_t* iter, int32_t* rho_updates, double* obj,
                    double* pri_res, double* dua_res) {
    if (!ok()) return MPCQP_ERR_CUDA;
    job_ = Job{p, s, B, num_obs, x0, xref, obs_c, obs_semi, obs_yaw, obs_dyn, lin_pt, warm_x, x, y, status, iter, rho_updates, obj, pri_res, dua_res};
    { std::lock_guard<std::mutex> lk(mu_); pending_ = (int)w_.size(); ++epoch_; }
    go_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    for (auto& w : w_) if (w.rc != MPCQP_OK) return w.rc;
    return MPCQP_OK;
  }

 private:
  struct Job {
    const mpcqp_mpc_params* p; const mpcqp_settings* s; int32_t B, R;
    const double *x0, *xref, *obs_c, *obs_semi, *obs_yaw; const int32_t* obs_dyn; const double *lin_pt, *warm_x;
    double *x, *y; int32_t *status, *iter, *rho_updates; double *obj, *pri_res, *dua_res;
  };
  struct Worker { int device = 0; mpcqp_engine* eng = nullptr; std::thread th; bool ready = false; int rc = MPCQP_OK; double kernel_ms = 0; std::string err; };

  void loop(size_t g) {
    Worker& w = w_[g];
    if (mpcqp_engine_create(w.device, &w.eng) != MPCQP_OK) { w.eng = nullptr; w.err = "mpcqp_engine_create failed on device " + std::to_string(w.device) + " (no CPU fallback)"; }
    uint64_t seen = 0;
    { std::lock_guard<std::mutex> lk(mu_); w.ready = true; seen = epoch_; }
    done_.notify_all();
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        go_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (quit_) break;
      }
      const Job j = job_;
      int64_t lo, hi; shardBounds(j.B, (int)w_.size(), (int)g, &lo, &hi);
      w.rc = MPCQP_OK; w.kernel_ms = 0;
      if (hi > lo && w.eng) {
        const int NS = j.p->horizon, N = NS - 1, R = j.R;
        const int64_t n = 8 * NS + 5 * N, m = 16 * NS + 5 * N + (int64_t)R * N;
        auto off = [&](const double* a, int64_t stride) { return a ? a + lo * stride : nullptr; };
        w.rc = mpcqp_solve_mpc_batch_host(w.eng, j.p, j.s, (int32_t)(hi - lo), R, off(j.x0, 6), off(j.xref, 3 * NS), off(j.obs_c, (int64_t)N * R * 3),
                                          off(j.obs_semi, (int64_t)N * R * 3), off(j.obs_yaw, (int64_t)N * R), j.obs_dyn, off(j.lin_pt, 3 * N), off(j.warm_x, n),
                                          j.x + lo * n, j.y ? j.y + lo * m : nullptr, j.status + lo, j.iter + lo, j.rho_updates + lo, j.obj + lo, j.pri_res + lo,
                                          j.dua_res + lo);
        if (w.rc != MPCQP_OK) w.err = mpcqp_engine_last_error(w.eng); else w.kernel_ms = mpcqp_engine_last_kernel_ms(w.eng);
      } else if (!w.eng) w.rc = MPCQP_ERR_CUDA;
      { std::lock_guard<std::mutex> lk(mu_); --pending_; }
      done_.notify_all();
    }
    if (w.eng) { mpcqp_engine_destroy(w.eng); w.eng = nullptr; }
  }

  std::vector<Worker> w_;
  std::mutex mu_; std::condition_variable go_, done_;
  uint64_t epoch_ = 0; int pending_ = 0; bool quit_ = false;
  Job job_{};
};

}  // namespace mpcqpB200
