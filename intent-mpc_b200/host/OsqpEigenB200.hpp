// OsqpEigenB200.hpp — an OsqpEigen::Solver-shaped C++ facade over the C ABI of the B200 engine
// (include/mpcqp_b200.h).  Drop-in for the way trajPlanner::mpcPlanner::solveTraj uses OsqpEigen
// (reference: trajectory_planner/include/trajectory_planner/mpcPlanner.cpp:436-527):
//
//     OsqpEigen::Solver solver;
//     solver.settings()->setVerbosity(false);  solver.settings()->setWarmStart(true);
//     solver.data()->setNumberOfVariables(n);  solver.data()->setNumberOfConstraints(m);
//     solver.data()->setHessianMatrix(P);      solver.data()->setGradient(q);
//     solver.data()->setLinearConstraintsMatrix(A);
//     solver.data()->setLowerBound(l);         solver.data()->setUpperBound(u);
//     solver.initSolver();  solver.setWarmStart(x, y);  solver.solveProblem();  solver.getSolution();
//
// Class and method names, argument meaning, return conventions (bool / ErrorExitFlag / Status) follow
// third_party/OsqpEigen/{Solver,Data,Settings,Constants}.hpp.  Differences, all forced by the engine:
//   * every solve runs on the GPU; there is no CPU path.  A problem with the mpcPlanner stage structure runs on the
//     stage-structured kernels, any other (polyTrajSolver.cpp:162-239: setUpProblem / updateProblem with updateBounds)
//     on the dense generic kernel; only an unstructured problem with n + m > 4096 makes initSolver() return false
//     (mpcqp_setup -> MPCQP_ERR_STRUCTURE);
//   * adaptive_rho_interval defaults to 25 and time_limit to 0 (the two determinism pins); setting a time
//     limit is accepted and ignored with a message on debugStream(), as mpcPlanner sets one (mpcPlanner.cpp:442-444);
//   * q, l, u are copied when set (the reference keeps pointers until initSolver, Data.hpp:92-122).
// Eigen is optional: matrices / vectors are taken as any type with Eigen's accessor names (rows, cols, nonZeros,
// outerIndexPtr, innerIndexPtr, valuePtr, coeff-less access / data, size), so Eigen::SparseMatrix<double> and
// Eigen::VectorXd work unchanged where Eigen exists; OsqpEigen::SparseMatrix / OsqpEigen::Vector below are minimal
// stand-ins for builds without it.  Define MPCQP_WITH_EIGEN before including to get Eigen return types.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mpcqp_b200.h"

#ifdef MPCQP_WITH_EIGEN
#include <Eigen/Dense>
#include <Eigen/Sparse>
#endif

namespace OsqpEigen {

using c_float = double;          // osqp glob_opts.h:87
using c_int = long long;         // osqp glob_opts.h:80
constexpr c_float INFTY = 1e30;  // OSQP_INFTY, osqp constants.h:78

enum class Status : int {        // OsqpEigen/Constants.hpp:25-41
  DualInfeasibleInaccurate = 4, PrimalInfeasibleInaccurate = 3, SolvedInaccurate = 2, Solved = 1, MaxIterReached = -2,
  PrimalInfeasible = -3, DualInfeasible = -4, Sigint = -5, TimeLimitReached = -6, NonCvx = -7, Unsolved = -10
};
enum class ErrorExitFlag : int { // OsqpEigen/Constants.hpp:46-56
  NoError = 0, DataValidationError = 1, SettingsValidationError = 2, LinsysSolverLoadError = 3, LinsysSolverInitError = 4,
  NonCvxError = 5, MemAllocError = 6, WorkspaceNotInitError = 7
};

inline std::ostream& debugStream() { return std::cerr; }

#ifdef MPCQP_WITH_EIGEN
using Vector = Eigen::Matrix<c_float, Eigen::Dynamic, 1>;
#else
// Minimal dense vector with the accessors the facade and mpcPlanner-style callers use.
class Vector {
 public:
  Vector() {}
  explicit Vector(size_t n, double v = 0.0) : v_(n, v) {}
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  long size() const { return (long)v_.size(); }
  long rows() const { return (long)v_.size(); }
  void resize(size_t n) { v_.resize(n); }
  void setZero() { for (double& x : v_) x = 0.0; }
  double& operator()(long i) { return v_[(size_t)i]; }
  double operator()(long i) const { return v_[(size_t)i]; }
  double& operator[](long i) { return v_[(size_t)i]; }
  double operator[](long i) const { return v_[(size_t)i]; }
 private:
  std::vector<double> v_;
};
#endif

// Minimal compressed column matrix with Eigen::SparseMatrix's accessor names.
class SparseMatrix {
 public:
  SparseMatrix() : r_(0), c_(0), outer_(1, 0) {}
  SparseMatrix(long rows, long cols) : r_(rows), c_(cols), outer_((size_t)cols + 1, 0) {}
  // from triplets (row, col, value); rows are sorted within a column (no duplicate entries expected)
  void setFromTriplets(const std::vector<long>& rows, const std::vector<long>& cols, const std::vector<double>& vals) {
    std::vector<std::vector<std::pair<long, double>>> colv((size_t)c_);
    for (size_t t = 0; t < rows.size(); ++t) colv[(size_t)cols[t]].push_back({rows[t], vals[t]});
    inner_.clear(); val_.clear(); outer_.assign((size_t)c_ + 1, 0);
    for (long j = 0; j < c_; ++j) {
      auto& cv = colv[(size_t)j];
      for (size_t a = 1; a < cv.size(); ++a) { auto key = cv[a]; size_t b = a; while (b > 0 && cv[b - 1].first > key.first) { cv[b] = cv[b - 1]; --b; } cv[b] = key; }
      for (auto& e : cv) { inner_.push_back((int)e.first); val_.push_back(e.second); }
      outer_[(size_t)j + 1] = (int)inner_.size();
    }
  }
  long rows() const { return r_; }
  long cols() const { return c_; }
  long nonZeros() const { return (long)val_.size(); }
  bool isCompressed() const { return true; }
  const int* outerIndexPtr() const { return outer_.data(); }
  const int* innerIndexPtr() const { return inner_.data(); }
  const double* valuePtr() const { return val_.data(); }
 private:
  long r_, c_;
  std::vector<int> outer_, inner_;
  std::vector<double> val_;
};

// ---- Settings (OsqpEigen/Settings.hpp:43-196) --------------------------------------------------------
class Settings {
 public:
  Settings() { resetDefaultSettings(); }
  void resetDefaultSettings() { mpcqp_set_default_settings(&s_); }
  void setRho(const double rho) { s_.rho = rho; }
  void setSigma(const double sigma) { s_.sigma = sigma; }
  void setScaling(const int scaling) { s_.scaling = scaling; }
  void setAdaptiveRho(const bool on) { s_.adaptive_rho = on ? 1 : 0; }
  void setAdaptiveRhoInterval(const int rhoInterval) { s_.adaptive_rho_interval = rhoInterval; }
  void setAdaptiveRhoTolerance(const double v) { s_.adaptive_rho_tolerance = v; }
  void setAdaptiveRhoFraction(const double v) { s_.adaptive_rho_fraction = v; }
  void setMaxIteration(const int maxIteration) { s_.max_iter = maxIteration; }
  void setAbsoluteTolerance(const double v) { s_.eps_abs = v; }
  void setRelativeTolerance(const double v) { s_.eps_rel = v; }
  void setPrimalInfeasibilityTollerance(const double v) { s_.eps_prim_inf = v; }
  void setPrimalInfeasibilityTolerance(const double v) { s_.eps_prim_inf = v; }
  void setDualInfeasibilityTollerance(const double v) { s_.eps_dual_inf = v; }
  void setDualInfeasibilityTolerance(const double v) { s_.eps_dual_inf = v; }
  void setAlpha(const double alpha) { s_.alpha = alpha; }
  void setLinearSystemSolver(const int) {}       // one linear solver: block-tridiagonal PCR / LDL' on the GPU
  void setDelta(const double delta) { s_.delta = delta; }
  void setPolish(const bool polish) { s_.polish = polish ? 1 : 0; }
  void setPolishRefineIter(const int v) { s_.polish_refine_iter = v; }
  void setVerbosity(const bool isVerbose) { s_.verbose = isVerbose ? 1 : 0; }
  void setScaledTerimination(const bool v) { s_.scaled_termination = v ? 1 : 0; }
  void setCheckTermination(const int v) { s_.check_termination = v; }
  void setWarmStart(const bool warmStart) { s_.warm_start = warmStart ? 1 : 0; }
  void setTimeLimit(const double timeLimit) {
    // wall-clock termination is not reproducible and not implemented; mpcPlanner sets it after its first solve
    if (timeLimit != 0.0 && !warned_) { debugStream() << "[OsqpEigenB200::Settings::setTimeLimit] time limit ignored (engine runs to OSQP's termination criteria).\n"; warned_ = true; }
  }
  mpcqp_settings* getSettings() { return &s_; }
  const mpcqp_settings* getSettings() const { return &s_; }
 private:
  mpcqp_settings s_;
  bool warned_ = false;
};

// ---- Data (OsqpEigen/Data.hpp:44-151) ----------------------------------------------------------------
class Data {
 public:
  Data() {}
  Data(int n, int m) : n_(n), m_(m) {}
  void clearHessianMatrix() { hessSet_ = false; Pp_.clear(); Pi_.clear(); Px_.clear(); }
  void clearLinearConstraintsMatrix() { linSet_ = false; Ap_.clear(); Ai_.clear(); Ax_.clear(); }
  void setNumberOfVariables(int n) { n_ = n; }
  void setNumberOfConstraints(int m) { m_ = m; }
  int getNumberOfVariables() const { return n_; }
  int getNumberOfConstraints() const { return m_; }

  // The upper triangle is kept, as Data.tpp:38-39 does before handing P to OSQP.
  template <class Mat> bool setHessianMatrix(const Mat& H) {
    if (hessSet_) { debugStream() << "[OsqpEigen::Data::setHessianMatrix] The hessian matrix was already set. Please use clearHessianMatrix() method to deallocate memory.\n"; return false; }
    if (!H.isCompressed()) { debugStream() << "[OsqpEigen::Data::setHessianMatrix] Please set the hessian matrix in a compressed form.\n"; return false; }
    if (H.rows() != n_ || H.cols() != n_) { debugStream() << "[OsqpEigen::Data::setHessianMatrix] The Hessian matrix has to be a n x n size matrix.\n"; return false; }
    copyCsc(H, true, Pp_, Pi_, Px_);
    hessSet_ = true;
    return true;
  }
  template <class Mat> bool setLinearConstraintsMatrix(const Mat& A) {
    if (linSet_) { debugStream() << "[OsqpEigen::Data::setLinearConstraintsMatrix] The linear constraint matrix was already set. Please use clearLinearConstraintsMatrix() method to deallocate memory.\n"; return false; }
    if (!A.isCompressed()) { debugStream() << "[OsqpEigen::Data::setLinearConstraintsMatrix] Please set the matrix in a compressed form.\n"; return false; }
    if (A.rows() != m_ || A.cols() != n_) { debugStream() << "[OsqpEigen::Data::setLinearConstraintsMatrix] The Linear constraints matrix has to be a m x n size matrix.\n"; return false; }
    copyCsc(A, false, Ap_, Ai_, Ax_);
    linSet_ = true;
    return true;
  }
  template <class Vec> bool setGradient(const Vec& g) {
    if ((long)g.size() != n_) { debugStream() << "[OsqpEigen::Data::setGradient] The size of the gradient must be equal to the number of the variables.\n"; return false; }
    q_.assign(g.data(), g.data() + n_); gradSet_ = true; return true;
  }
  template <class Vec> bool setLowerBound(const Vec& l) {
    if ((long)l.size() != m_) { debugStream() << "[OsqpEigen::Data::setLowerBound] The size of the lower bound must be equal to the number of the constraints.\n"; return false; }
    l_.assign(l.data(), l.data() + m_); lowSet_ = true; return true;
  }
  template <class Vec> bool setUpperBound(const Vec& u) {
    if ((long)u.size() != m_) { debugStream() << "[OsqpEigen::Data::setUpperBound] The size of the upper bound must be equal to the number of the constraints.\n"; return false; }
    u_.assign(u.data(), u.data() + m_); upSet_ = true; return true;
  }
  template <class Vec> bool setBounds(const Vec& l, const Vec& u) { return setLowerBound(l) && setUpperBound(u); }
  bool isSet() const { return n_ > 0 && m_ >= 0 && hessSet_ && gradSet_ && linSet_ && lowSet_ && upSet_; }

  // raw views for the solver
  const std::vector<int64_t>& Pp() const { return Pp_; } const std::vector<int64_t>& Pi() const { return Pi_; } const std::vector<double>& Px() const { return Px_; }
  const std::vector<int64_t>& Ap() const { return Ap_; } const std::vector<int64_t>& Ai() const { return Ai_; } const std::vector<double>& Ax() const { return Ax_; }
  std::vector<double>& q() { return q_; } std::vector<double>& l() { return l_; } std::vector<double>& u() { return u_; }

 private:
  template <class Mat> static void copyCsc(const Mat& M, bool upper, std::vector<int64_t>& p, std::vector<int64_t>& i, std::vector<double>& x) {
    const long cols = (long)M.cols();
    p.assign((size_t)cols + 1, 0); i.clear(); x.clear();
    const auto* op = M.outerIndexPtr(); const auto* ip = M.innerIndexPtr(); const auto* vp = M.valuePtr();
    for (long j = 0; j < cols; ++j) {
      for (long t = (long)op[j]; t < (long)op[j + 1]; ++t) {
        if (upper && (long)ip[t] > j) continue;
        i.push_back((int64_t)ip[t]); x.push_back((double)vp[t]);
      }
      p[(size_t)j + 1] = (int64_t)i.size();
    }
  }
  int n_ = 0, m_ = 0;
  bool hessSet_ = false, gradSet_ = false, linSet_ = false, lowSet_ = false, upSet_ = false;
  std::vector<int64_t> Pp_, Pi_, Ap_, Ai_;
  std::vector<double> Px_, Ax_, q_, l_, u_;
};

// One engine per (host thread, device): mpcPlanner creates a fresh Solver every control step
// (mpcPlanner.cpp:436), which must not pay for stream / event creation each time.  Engines are kept per device for the
// life of the thread; a Solver that outlives its thread's engines (a static Solver, a polyTrajSolver member destroyed late)
// stays valid because mpcqp_engine_destroy defers while problems of that engine are alive.
inline mpcqp_engine* threadEngine(int device = 0) {
  struct Holder {
    std::vector<std::pair<int, mpcqp_engine*>> engines;
    ~Holder() { for (auto& kv : engines) if (kv.second) mpcqp_engine_destroy(kv.second); }
  };
  static thread_local Holder h;
  for (auto& kv : h.engines) if (kv.first == device) return kv.second;
  mpcqp_engine* e = nullptr;
  if (mpcqp_engine_create(device, &e) != MPCQP_OK) return nullptr;
  h.engines.emplace_back(device, e);
  return e;
}

// ---- Solver (OsqpEigen/Solver.hpp:87-249) ------------------------------------------------------------
class Solver {
 public:
  Solver() : settings_(new Settings()), data_(new Data()), work_(nullptr, [](mpcqp_problem* p) { if (p) mpcqp_cleanup(p); }) {}
  void setDevice(int device) { device_ = device; }

  bool initSolver() {
    if (isInitialized()) { debugStream() << "[OsqpEigen::Solver::initSolver] The solver has been already initialized. Please use clearSolver() method to deallocate memory.\n"; return false; }
    if (!data_->isSet()) { debugStream() << "[OsqpEigen::Solver::initSolver] Some data are not set.\n"; return false; }
    mpcqp_engine* e = threadEngine(device_);
    if (!e) { debugStream() << "[OsqpEigen::Solver::initSolver] No usable CUDA device: this engine has no CPU fallback.\n"; return false; }
    mpcqp_problem* p = nullptr;
    const int rc = mpcqp_setup(e, &p, data_->getNumberOfVariables(), data_->getNumberOfConstraints(), data_->Pp().data(), data_->Pi().data(),
                               data_->Px().data(), data_->q().data(), data_->Ap().data(), data_->Ai().data(), data_->Ax().data(),
                               data_->l().data(), data_->u().data(), settings_->getSettings());
    if (rc != MPCQP_OK) { debugStream() << "[OsqpEigen::Solver::initSolver] Unable to setup the workspace: " << mpcqp_engine_last_error(e) << "\n"; lastError_ = rc; return false; }
    work_.reset(p);
    const size_t n = (size_t)data_->getNumberOfVariables(), m = (size_t)data_->getNumberOfConstraints();
    primal_.resize(n); dual_.resize(m); sol_.resize(n); dsol_.resize(m);
    for (size_t i = 0; i < n; ++i) primal_[i] = 0.0;
    for (size_t i = 0; i < m; ++i) dual_[i] = 0.0;
    return true;
  }
  bool isInitialized() { return (bool)work_; }
  void clearSolver() { work_.reset(); }
  bool clearSolverVariables() {
    if (!isInitialized()) { debugStream() << "[OsqpEigen::Solver::clearSolverVariables] Unable to clear the solver variables.\n"; return false; }
    for (double& v : primal_) v = 0.0;
    for (double& v : dual_) v = 0.0;
    return mpcqp_warm_start(work_.get(), primal_.data(), dual_.data()) == MPCQP_OK;
  }
  bool solve() { return solveProblem() == ErrorExitFlag::NoError && (getStatus() == Status::Solved); }
  ErrorExitFlag solveProblem() {
    if (!isInitialized()) { debugStream() << "[OsqpEigen::Solver::solveProblem] The solve has not been initialized yet. Please call initSolver() method.\n"; return ErrorExitFlag::WorkspaceNotInitError; }
    const int rc = mpcqp_solve(work_.get());
    if (rc != MPCQP_OK) { debugStream() << "[OsqpEigen::Solver::solveProblem] " << mpcqp_engine_last_error(threadEngine(device_)) << "\n"; return rc == MPCQP_ERR_NOT_INIT ? ErrorExitFlag::WorkspaceNotInitError : ErrorExitFlag::LinsysSolverInitError; }
    mpcqp_get_solution(work_.get(), sol_.data(), dsol_.data());
    return ErrorExitFlag::NoError;
  }
  Status getStatus() const { mpcqp_info i; if (!work_ || mpcqp_get_info(work_.get(), &i) != MPCQP_OK) return Status::Unsolved; return (Status)(int)i.status_val; }
  c_float getObjValue() const { mpcqp_info i; if (!work_ || mpcqp_get_info(work_.get(), &i) != MPCQP_OK) return 0.0; return i.obj_val; }
  bool getInfo(mpcqp_info* out) const { return work_ && mpcqp_get_info(work_.get(), out) == MPCQP_OK; }
  const Vector& getSolution() { copyOut(sol_, solV_); return solV_; }
  const Vector& getDualSolution() { copyOut(dsol_, dsolV_); return dsolV_; }

  template <class Vec> bool updateGradient(const Vec& g) {
    if (!isInitialized() || (long)g.size() != data_->getNumberOfVariables()) { debugStream() << "[OsqpEigen::Solver::updateGradient] The size of the gradient must be equal to the number of the variables.\n"; return false; }
    data_->q().assign(g.data(), g.data() + g.size());
    return mpcqp_update_lin_cost(work_.get(), data_->q().data()) == MPCQP_OK;
  }
  template <class Vec> bool updateLowerBound(const Vec& l) {
    if (!isInitialized() || (long)l.size() != data_->getNumberOfConstraints()) { debugStream() << "[OsqpEigen::Solver::updateLowerBound] The size of the lower bound must be equal to the number of the variables.\n"; return false; }
    data_->l().assign(l.data(), l.data() + l.size());
    return pushBounds();
  }
  template <class Vec> bool updateUpperBound(const Vec& u) {
    if (!isInitialized() || (long)u.size() != data_->getNumberOfConstraints()) { debugStream() << "[OsqpEigen::Solver::updateUpperBound] The size of the upper bound must be equal to the number of the variables.\n"; return false; }
    data_->u().assign(u.data(), u.data() + u.size());
    return pushBounds();
  }
  template <class Vec> bool updateBounds(const Vec& l, const Vec& u) {
    if (!isInitialized() || (long)l.size() != data_->getNumberOfConstraints() || (long)u.size() != data_->getNumberOfConstraints()) { debugStream() << "[OsqpEigen::Solver::updateBounds] The size of the bounds must be equal to the number of the constraints.\n"; return false; }
    data_->l().assign(l.data(), l.data() + l.size()); data_->u().assign(u.data(), u.data() + u.size());
    return pushBounds();
  }
  // A new Hessian / constraint matrix re-runs setup (OSQP re-factorises as well); the solver keeps its warm start.
  template <class Mat> bool updateHessianMatrix(const Mat& H) { data_->clearHessianMatrix(); if (!data_->setHessianMatrix(H)) return false; return reinit(); }
  template <class Mat> bool updateLinearConstraintsMatrix(const Mat& A) { data_->clearLinearConstraintsMatrix(); if (!data_->setLinearConstraintsMatrix(A)) return false; return reinit(); }

  template <class V1, class V2> bool setWarmStart(const V1& primal, const V2& dual) {
    if (!isInitialized() || (long)primal.size() != data_->getNumberOfVariables() || (long)dual.size() != data_->getNumberOfConstraints()) { debugStream() << "[OsqpEigen::Solver::setWarmStart] The size of the vectors has to be equal to the number of variables / constraints.\n"; return false; }
    primal_.assign(primal.data(), primal.data() + primal.size()); dual_.assign(dual.data(), dual.data() + dual.size());
    return mpcqp_warm_start(work_.get(), primal_.data(), dual_.data()) == MPCQP_OK;
  }
  template <class V1> bool setPrimalVariable(const V1& primal) {
    if (!isInitialized() || (long)primal.size() != data_->getNumberOfVariables()) { debugStream() << "[OsqpEigen::Solver::setPrimalVariable] The size of the vector has to be equal to the number of variables.\n"; return false; }
    primal_.assign(primal.data(), primal.data() + primal.size());
    return mpcqp_warm_start_x(work_.get(), primal_.data()) == MPCQP_OK;
  }
  template <class V2> bool setDualVariable(const V2& dual) {
    if (!isInitialized() || (long)dual.size() != data_->getNumberOfConstraints()) { debugStream() << "[OsqpEigen::Solver::setDualVariable] The size of the vector has to be equal to the number of constraints.\n"; return false; }
    dual_.assign(dual.data(), dual.data() + dual.size());
    return mpcqp_warm_start(work_.get(), primal_.data(), dual_.data()) == MPCQP_OK;
  }
  template <class V1> bool getPrimalVariable(V1& primal) { if ((long)primal.size() != (long)primal_.size()) return false; for (size_t i = 0; i < primal_.size(); ++i) primal.data()[i] = primal_[i]; return true; }
  template <class V2> bool getDualVariable(V2& dual) { if ((long)dual.size() != (long)dual_.size()) return false; for (size_t i = 0; i < dual_.size(); ++i) dual.data()[i] = dual_[i]; return true; }

  const std::unique_ptr<Settings>& settings() const { return settings_; }
  const std::unique_ptr<Data>& data() const { return data_; }
  const std::unique_ptr<mpcqp_problem, std::function<void(mpcqp_problem*)>>& workspace() const { return work_; }
  int lastEngineError() const { return lastError_; }

 private:
  bool pushBounds() { const int rc = mpcqp_update_bounds(work_.get(), data_->l().data(), data_->u().data()); if (rc != MPCQP_OK) debugStream() << "[OsqpEigen::Solver::updateBounds] " << mpcqp_engine_last_error(threadEngine(device_)) << "\n"; return rc == MPCQP_OK; }
  bool reinit() { if (!isInitialized()) return true; std::vector<double> px = primal_, dy = dual_; clearSolver(); if (!initSolver()) return false; primal_ = px; dual_ = dy; return mpcqp_warm_start(work_.get(), primal_.data(), dual_.data()) == MPCQP_OK; }
  static void copyOut(const std::vector<double>& src, Vector& dst) { if ((size_t)dst.size() != src.size()) dst.resize(src.size()); for (size_t i = 0; i < src.size(); ++i) dst.data()[i] = src[i]; }
  std::unique_ptr<Settings> settings_;
  std::unique_ptr<Data> data_;
  std::unique_ptr<mpcqp_problem, std::function<void(mpcqp_problem*)>> work_;
  std::vector<double> primal_, dual_, sol_, dsol_;
  Vector solV_, dsolV_;
  int device_ = 0, lastError_ = 0;
};

}  // namespace OsqpEigen
