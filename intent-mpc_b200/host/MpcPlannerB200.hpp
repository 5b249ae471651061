// MpcPlannerB200.hpp — ROS-free, Eigen-free mirror of trajPlanner::mpcPlanner's plan / trajectory interface
// (reference: trajectory_planner/include/trajectory_planner/mpcPlanner.h:108-175, mpcPlanner.cpp) on top of the
// batched entry point of the B200 engine.  Same method names, argument meaning and return conventions; Eigen::Vector3d
// becomes std::array<double,3>, nav_msgs::Path becomes std::vector<Vec3>.
//
// What runs where:
//   * reference window, intent-combination enumeration, obstacle parameters (including the isDynamic quirk of
//     mpcPlanner.cpp:1194), candidate scoring and selection: host code below, restating mpcPlanner.cpp:571-887,
//     1148-1231;
//   * QP assembly (castMPCToQP*, mpcPlanner.cpp:932-1146) and the OSQP solve (mpcPlanner.cpp:436-527): on the GPU through
//     mpcqp_solve_mpc_batch_host — the up-to-six candidate QPs of one control step go down as ONE batch per obstacle
//     count (candidates 4 and 5 carry the closest obstacle twice, mpcPlanner.cpp:737-741, so they have one more row per
//     stage than candidates 0-3).  This is legitimate because the reference's sequential solves all share the warm start
//     and the linearisation point of the previous step (currentStatesSol_ changes only after selection, :629-639).
// Not mirrored: RViz publishers, point-cloud clustering (disabled in the reference, mpcPlanner.cpp:189-194; static
// obstacles are passed in with updateStaticObstacles), the wall-clock cut-off between candidates (:613) and the field-of-view
// half-space rows (3-argument updateCurrStates; off in production).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mpcqp_b200.h"

namespace trajPlannerB200 {

using Vec3 = std::array<double, 3>;
constexpr int numStates = 8, numControls = 5;          // mpcPlanner.h:42-43
enum Intent { FORWARD = 0, LEFT = 1, RIGHT = 2, STOP = 3 };   // dynamic_predictor/utils.h:15-20

struct staticObstacle { Vec3 centroid; Vec3 size; double yaw; };   // clustering/obstacleClustering.h

inline double norm3(const Vec3& a, const Vec3& b) { return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2])); }

class mpcPlanner {
 public:
  using ObTraj = std::vector<Vec3>;                    // one obstacle, one intent: position (or size) per prediction step

  explicit mpcPlanner(int device = 0) { mpcqp_default_mpc_params(&p_); mpcqp_set_default_settings(&s_); ok_ = mpcqp_engine_create(device, &eng_) == MPCQP_OK; }
  ~mpcPlanner() { if (eng_) mpcqp_engine_destroy(eng_); }
  mpcPlanner(const mpcPlanner&) = delete;
  mpcPlanner& operator=(const mpcPlanner&) = delete;
  bool engineReady() const { return ok_; }
  const char* lastError() const { return eng_ ? mpcqp_engine_last_error(eng_) : "no CUDA device (no CPU fallback)"; }

  // ---- parameters (initParam keys, mpcPlanner.cpp:19-173) ----------------------------------------
  mpcqp_mpc_params& params() { return p_; }
  mpcqp_settings& settings() { return s_; }
  void updateMaxVel(double maxVel) { p_.max_vel = maxVel; }                              // mpcPlanner.cpp:249
  void updateMaxAcc(double maxAcc) { p_.max_acc = maxAcc; }                              // :253
  void updateCurrStates(const Vec3& pos, const Vec3& vel) { currPos_ = pos; currVel_ = vel; trajHist_.push_back(pos); stateReceived_ = true; }   // :257-263
  void updatePath(const std::vector<Vec3>& path, double ts) { p_.ts = ts; inputTraj_ = path; firstTime_ = true; stateReceived_ = false; trajHist_.clear(); lastRefStartIdx_ = 0; }   // :307-314
  void updateStaticObstacles(const std::vector<staticObstacle>& obs) { staticObstacles_ = obs; }
  void updateDynamicObstacles(const std::vector<Vec3>& pos, const std::vector<Vec3>& vel, const std::vector<Vec3>& size) {   // :316-341
    (void)vel;
    dynamicObstaclesPos_.assign(pos.size(), ObTraj()); dynamicObstaclesSize_.assign(pos.size(), ObTraj());
    for (size_t i = 0; i < pos.size(); ++i) { dynamicObstaclesPos_[i].assign((size_t)p_.horizon, pos[i]); dynamicObstaclesSize_[i].assign((size_t)p_.horizon, size[i]); }
  }
  // predPos[ob][intent][step], predSize likewise, intentProb[ob][4]   (:343-373)
  void updatePredObstacles(const std::vector<std::vector<ObTraj>>& predPos, const std::vector<std::vector<ObTraj>>& predSize, const std::vector<std::array<double, 4>>& intentProb) {
    dynamicObstaclesPos_.clear(); dynamicObstaclesSize_.clear();
    if (!predPos.empty()) {
      for (size_t i = 0; i < predPos.size(); ++i) { dynamicObstaclesPos_.push_back(ObTraj((size_t)p_.horizon, predPos[i][0][0])); dynamicObstaclesSize_.push_back(ObTraj((size_t)p_.horizon, predSize[i][0][0])); }
      obPredPos_ = predPos; obPredSize_ = predSize; obIntentProb_ = intentProb;
    } else { obPredPos_.clear(); obPredSize_.clear(); obIntentProb_.clear(); }
  }

  // ---- planning ------------------------------------------------------------------------------------
  bool makePlan() {                                                                     // :543-569
    if (firstTime_) { currentStatesSol_.clear(); currentControlsSol_.clear(); ref_.clear(); }
    std::vector<staticObstacle> so = staticObstacles_; std::vector<ObTraj> dp = dynamicObstaclesPos_, ds = dynamicObstaclesSize_;
    if (firstTime_) { so.clear(); dp.clear(); ds.clear(); }
    std::vector<Vec3> xRef; getReferenceTraj(xRef);
    std::vector<std::vector<double>> st, ct;
    const bool ok = solveTraj(so, dp, ds, st, ct, xRef);
    candidateStatus_ = lastStatus_; candidateIter_ = lastIter_;
    if (ok) { currentStatesSol_ = st; currentControlsSol_ = ct; firstTime_ = false; ref_ = xRef; }
    return ok;
  }

  bool makePlanWithPred() {                                                             // :571-661
    if (firstTime_) { candidateStates_.clear(); candidateControls_.clear(); trajWeightedScore_.clear(); trajScore_.clear(); currentStatesSol_.clear(); currentControlsSol_.clear(); ref_.clear(); }
    std::vector<staticObstacle> so; std::vector<ObTraj> dp, ds;
    if (!firstTime_) { so = staticObstacles_; dp = dynamicObstaclesPos_; ds = dynamicObstaclesSize_; }
    std::vector<Vec3> xRef; getReferenceTraj(xRef);
    bool valid;
    if (!obPredPos_.empty() && !firstTime_) {
      int obIdx; std::vector<std::vector<ObTraj>> combPos, combSize;
      getIntentComb(obIdx, combPos, combSize);
      // the candidates of one control step as batches of equal obstacle count
      std::vector<std::vector<std::vector<double>>> st(combPos.size()), ct(combPos.size());
      std::vector<char> solved(combPos.size(), 0);
      std::vector<int> cst(combPos.size(), MPCQP_UNSOLVED), cit(combPos.size(), 0);
      std::vector<size_t> counts;
      for (auto& c : combPos) if (std::find(counts.begin(), counts.end(), c.size()) == counts.end()) counts.push_back(c.size());
      for (size_t cnt : counts) {
        std::vector<int> idx;
        for (size_t i = 0; i < combPos.size(); ++i) if (combPos[i].size() == cnt) idx.push_back((int)i);
        std::vector<const std::vector<ObTraj>*> pp, ss;
        for (int i : idx) { pp.push_back(&combPos[(size_t)i]); ss.push_back(&combSize[(size_t)i]); }
        std::vector<std::vector<std::vector<double>>> bst, bct;
        if (solveBatch(so, pp, ss, xRef, bst, bct)) for (size_t a = 0; a < idx.size(); ++a) {
          st[(size_t)idx[a]] = bst[a]; ct[(size_t)idx[a]] = bct[a]; solved[(size_t)idx[a]] = 1;
          cst[(size_t)idx[a]] = lastStatus_[a]; cit[(size_t)idx[a]] = lastIter_[a];
        }
      }
      candidateStatus_.clear(); candidateIter_.clear();
      for (size_t i = 0; i < combPos.size(); ++i) if (solved[i]) { candidateStatus_.push_back(cst[i]); candidateIter_.push_back(cit[i]); }
      std::vector<std::vector<std::vector<double>>> cs, cc; std::vector<Vec3> score; std::vector<int> intentType;
      for (size_t i = 0; i < combPos.size(); ++i) if (solved[i]) {
        cs.push_back(st[i]); cc.push_back(ct[i]);
        score.push_back(getTrajectoryScore(st[i], so, combPos[i], combSize[i], xRef)); intentType.push_back((int)i);
      }
      candidateStates_ = cs; candidateControls_ = cc;
      if (!cs.empty()) {
        firstTime_ = false; valid = true;
        const int best = evaluateTraj(score, obIdx, intentType);
        currentStatesSol_ = candidateStates_[(size_t)best]; currentControlsSol_ = candidateControls_[(size_t)best];
        trajScore_ = score; ref_ = xRef; bestIdx_ = best;
      } else valid = false;
    } else {
      candidateStates_.clear(); candidateControls_.clear(); trajWeightedScore_.clear(); trajScore_.clear();
      std::vector<std::vector<double>> st, ct;
      valid = solveTraj(so, dp, ds, st, ct, xRef);
      candidateStatus_ = lastStatus_; candidateIter_ = lastIter_;
      if (valid) { currentStatesSol_ = st; currentControlsSol_ = ct; firstTime_ = false; ref_ = xRef; }
    }
    return valid;
  }

  // One QP (mpcPlanner.cpp:375-541).  states: horizon x 8, controls: (horizon-1) x 5.
  bool solveTraj(const std::vector<staticObstacle>& so, const std::vector<ObTraj>& dynPos, const std::vector<ObTraj>& dynSize,
                 std::vector<std::vector<double>>& statesSol, std::vector<std::vector<double>>& controlsSol, const std::vector<Vec3>& xRef) {
    std::vector<const std::vector<ObTraj>*> pp{&dynPos}, ss{&dynSize};
    std::vector<std::vector<std::vector<double>>> st, ct;
    if (!solveBatch(so, pp, ss, xRef, st, ct)) return false;
    statesSol = st[0]; controlsSol = ct[0];
    return true;
  }

  // ---- results -------------------------------------------------------------------------------------
  void getTrajectory(std::vector<Vec3>& traj) const { traj.clear(); for (auto& s : currentStatesSol_) traj.push_back({s[0], s[1], s[2]}); }   // :1234-1242
  Vec3 getPos(double t) const { if (currentStatesSol_.empty()) return currPos_; return interp(currentStatesSol_, 0, t); }                     // :1257-1274
  Vec3 getVel(double t) const { if (currentStatesSol_.empty()) return {0, 0, 0}; return interp(currentStatesSol_, 3, t); }                    // :1276-1292
  Vec3 getAcc(double t) const { if (currentControlsSol_.empty()) return {0, 0, 0}; return interp(currentControlsSol_, 0, t); }                // :1294-1310
  Vec3 getRef(double t) const {                                                                                                                // :1312-1327
    if (ref_.empty()) return currPos_;
    std::vector<std::vector<double>> r; for (auto& v : ref_) r.push_back({v[0], v[1], v[2]});
    return interp(r, 0, t);
  }
  double getTs() const { return p_.ts; }
  double getHorizon() const { return p_.horizon; }
  double getLastQpSolveTime() const { return lastQpSolveTime_; }       // device + transfer time of the last batched solve, seconds
  const std::vector<int>& lastStatus() const { return candidateStatus_; }   // OSQP status per QP of the last plan call, in candidate order (the reference never looks)
  const std::vector<int>& lastIterations() const { return candidateIter_; }
  const std::vector<Vec3>& trajScore() const { return trajScore_; }       // (consistency, detour, safety) per candidate
  const std::vector<std::vector<std::vector<double>>>& candidateStates() const { return candidateStates_; }
  const std::vector<std::vector<std::vector<double>>>& candidateControls() const { return candidateControls_; }
  int closestObstacle() const { return obIdx_; }
  const std::vector<double>& trajWeightedScore() const { return trajWeightedScore_; }
  int bestCandidate() const { return bestIdx_; }
  const std::vector<std::vector<double>>& currentStates() const { return currentStatesSol_; }
  const std::vector<std::vector<double>>& currentControls() const { return currentControlsSol_; }

  // ---- pieces exposed for the tests -----------------------------------------------------------------
  void getReferenceTraj(std::vector<Vec3>& ref) {                                       // :1199-1231
    ref.clear();
    if (inputTraj_.empty()) { for (int i = 0; i < p_.horizon; ++i) ref.push_back(currPos_); return; }
    double least = 1.7976931348623157e308; const int maxFwd = (int)(3.0 / p_.ts);
    int startIdx = lastRefStartIdx_; const int end = std::min(lastRefStartIdx_ + maxFwd, (int)inputTraj_.size());
    for (int i = lastRefStartIdx_; i < end; ++i) { const double d = norm3(currPos_, inputTraj_[(size_t)i]); if (d < least) { least = d; startIdx = i; } }
    lastRefStartIdx_ = startIdx;
    for (int i = startIdx; i < startIdx + p_.horizon; ++i) ref.push_back(i < (int)inputTraj_.size() ? inputTraj_[(size_t)i] : inputTraj_.back());
  }
  void getIntentComb(int& obIdx, std::vector<std::vector<ObTraj>>& combPos, std::vector<std::vector<ObTraj>>& combSize) {   // :710-769
    findClosestObstacle(obIdx); obIdx_ = obIdx;
    const auto& pr = obIntentProb_[(size_t)obIdx];
    std::vector<std::pair<double, int>> w = {{pr[STOP], 0}, {pr[LEFT], 1}, {pr[RIGHT], 2}, {pr[FORWARD], 3}, {std::max(pr[LEFT], pr[FORWARD]), 4}, {std::max(pr[RIGHT], pr[FORWARD]), 5}};
    std::sort(w.begin(), w.end());
    const int first[6] = {STOP, LEFT, RIGHT, FORWARD, LEFT, RIGHT};
    std::vector<std::vector<ObTraj>> tp(6), ts(6);
    for (int i = 0; i < 6; ++i) {
      tp[(size_t)i].push_back(obPredPos_[(size_t)obIdx][(size_t)first[i]]); ts[(size_t)i].push_back(obPredSize_[(size_t)obIdx][(size_t)first[i]]);
      if (i >= 4) { tp[(size_t)i].push_back(obPredPos_[(size_t)obIdx][FORWARD]); ts[(size_t)i].push_back(obPredSize_[(size_t)obIdx][FORWARD]); }
    }
    combPos.assign(6, {}); combSize.assign(6, {});
    for (int i = 0; i < 6; ++i) { combPos[(size_t)i] = tp[(size_t)w[(size_t)(5 - i)].second]; combSize[(size_t)i] = ts[(size_t)w[(size_t)(5 - i)].second]; }
    for (size_t i = 0; i < 6; ++i) for (size_t j = 0; j < obPredPos_.size(); ++j) if ((int)j != obIdx_) {
      const auto& q = obIntentProb_[j];
      const int mi = (int)(std::max_element(q.begin(), q.end()) - q.begin());
      combPos[i].push_back(obPredPos_[j][(size_t)mi]); combSize[i].push_back(obPredSize_[j][(size_t)mi]);
    }
  }

 private:
  void findClosestObstacle(int& obIdx) const {                                          // :663-708
    obIdx = -1; double minDist = INFINITY;
    if (firstTime_ || currentStatesSol_.size() < 2) {
      for (size_t i = 0; i < dynamicObstaclesPos_.size(); ++i) { const double d = norm3(currPos_, dynamicObstaclesPos_[i][0]); if (d < minDist) { minDist = d; obIdx = (int)i; } }
      return;
    }
    for (size_t i = 0; i < dynamicObstaclesPos_.size(); ++i) {
      double dist = 0;
      for (int j = 0; j < (int)(currentStatesSol_.size() / 3); ++j) {
        const Vec3 s{currentStatesSol_[0][0], currentStatesSol_[0][1], currentStatesSol_[0][2]}, nx{currentStatesSol_[1][0], currentStatesSol_[1][1], currentStatesSol_[1][2]};
        const double ta = std::atan2(nx[1] - s[1], nx[0] - s[0]), oa = std::atan2(dynamicObstaclesPos_[i][0][1] - s[1], dynamicObstaclesPos_[i][0][0] - s[0]);
        dist += std::exp(-(double)j) * norm3(s, dynamicObstaclesPos_[i][0]) * (3.0 - std::cos(ta - oa));
        if (dist > minDist) break;
      }
      if (dist < minDist) { minDist = dist; obIdx = (int)i; }
    }
  }
  Vec3 getTrajectoryScore(const std::vector<std::vector<double>>& st, const std::vector<staticObstacle>& so, const std::vector<ObTraj>& op, const std::vector<ObTraj>& os, const std::vector<Vec3>& xRef) const {   // :771-852
    double cons = 0;
    if (!(firstTime_ || currentStatesSol_.empty() || st.empty())) {
      const int ms = std::min(10, std::min((int)currentStatesSol_.size(), (int)st.size()));
      if (ms > 0) { for (int i = 0; i < ms; ++i) cons += norm3({currentStatesSol_[(size_t)i][0], currentStatesSol_[(size_t)i][1], currentStatesSol_[(size_t)i][2]}, {st[(size_t)i][0], st[(size_t)i][1], st[(size_t)i][2]}); cons = std::max(cons / ms, 0.1); }
    }
    double det = 0;
    for (size_t i = 0; i < st.size(); ++i) det += norm3(xRef[i], {st[i][0], st[i][1], st[i][2]});
    det = std::max(det / (double)st.size(), 0.1);
    double saf = 0;
    for (size_t i = 0; i < st.size(); ++i) {
      double dist = 0, tw = 0; const Vec3 pos{st[i][0], st[i][1], 0};
      for (size_t j = 0; j < op.size(); ++j) {
        Vec3 o = op[j][i]; o[2] = 0;
        const double ms = std::sqrt(os[j][i][0] * os[j][i][0] + os[j][i][1] * os[j][i][1]), d = norm3(pos, o), w = 1 - std::tanh(std::atanh(0.5) / (p_.dynamic_safety_dist + ms) * d);
        dist += d * w; tw += w;
      }
      for (size_t j = 0; j < so.size(); ++j) {
        const Vec3 o{so[j].centroid[0], so[j].centroid[1], 0};
        const double ms = std::sqrt(so[j].size[0] / 2 * so[j].size[0] / 2 + so[j].size[1] / 2 * so[j].size[1] / 2), d = norm3(pos, o), w = 1 - std::tanh(std::atanh(0.5) / (p_.static_safety_dist + ms) * d);
        dist += d * w; tw += w;
      }
      saf += dist / tw;
    }
    saf /= (double)st.size();
    return {cons, det, saf};
  }
  int evaluateTraj(const std::vector<Vec3>& score, int obIdx, const std::vector<int>& intentType) {   // :854-887
    trajWeightedScore_.clear();
    double ca = 0, da = 0, sa = 0;
    for (auto& s : score) { ca += s[0]; da += s[1]; sa += s[2]; }
    ca /= (double)score.size(); da /= (double)score.size(); sa /= (double)score.size();
    const auto& pr = obIntentProb_[(size_t)obIdx];
    const double w[6] = {pr[STOP], pr[LEFT], pr[RIGHT], pr[FORWARD], std::max(pr[LEFT], pr[FORWARD]), std::max(pr[RIGHT], pr[FORWARD])};
    // weightedScore.maxCoeff(&bestTrajIdx): Eigen 3.3's visitor starts from element 0 and replaces on a strict `>`
    int best = 0; double bs = 0.0;
    for (size_t i = 0; i < score.size(); ++i) {
      const double v = w[intentType[i]] * (1.0 * (ca / score[i][0]) + 1.0 * (da / score[i][1]) + 1.0 * (score[i][2] / sa));
      trajWeightedScore_.push_back(v);
      if (i == 0) bs = v; else if (v > bs) { bs = v; best = (int)i; }
    }
    return best;
  }
  Vec3 interp(const std::vector<std::vector<double>>& seq, int off, double t) const {
    int idx = (int)std::floor(t / p_.ts); const double dt = t - idx * p_.ts;
    idx = std::max(0, std::min(idx, (int)seq.size() - 1)); const int nx = std::min(idx + 1, (int)seq.size() - 1);
    Vec3 r;
    for (int c = 0; c < 3; ++c) r[(size_t)c] = seq[(size_t)idx][(size_t)(off + c)] + (seq[(size_t)nx][(size_t)(off + c)] - seq[(size_t)idx][(size_t)(off + c)]) / p_.ts * dt;
    return r;
  }

  // B candidates with the same obstacle count in one engine call: updateObstacleParam (:1148-1197) + linearisation point
  // (:1042-1051) + warm start (:485-509) on the host; assembly and solve on the device.
  bool solveBatch(const std::vector<staticObstacle>& so, const std::vector<const std::vector<ObTraj>*>& dynPos, const std::vector<const std::vector<ObTraj>*>& dynSize,
                  const std::vector<Vec3>& xRef, std::vector<std::vector<std::vector<double>>>& states, std::vector<std::vector<std::vector<double>>>& controls) {
    if (!ok_) return false;
    if (firstTime_) { currentStatesSol_.clear(); currentControlsSol_.clear(); }          // :378-381
    const int B = (int)dynPos.size(), NS = p_.horizon, N = NS - 1, D = (int)dynPos[0]->size(), S = (int)so.size(), R = D + S;
    const int n = numStates * NS + numControls * N;
    std::vector<double> x0((size_t)B * 6), xr((size_t)B * NS * 3), oc((size_t)B * N * std::max(R, 1) * 3), os(oc.size()), oy((size_t)B * N * std::max(R, 1)), lp((size_t)B * N * 3), wx((size_t)B * n, 0.0);
    std::vector<int32_t> od((size_t)N * std::max(R, 1), 0);
    for (int b = 0; b < B; ++b) {
      for (int c = 0; c < 3; ++c) { x0[(size_t)b * 6 + c] = currPos_[(size_t)c]; x0[(size_t)b * 6 + 3 + c] = currVel_[(size_t)c]; }
      for (int k = 0; k < NS; ++k) for (int c = 0; c < 3; ++c) xr[((size_t)b * NS + k) * 3 + c] = xRef[(size_t)k][(size_t)c];
      for (int k = 0; k < N; ++k) {
        for (int i = 0; i < D; ++i) {
          const ObTraj& tp = (*dynPos[(size_t)b])[(size_t)i]; const ObTraj& tz = (*dynSize[(size_t)b])[(size_t)i];
          const Vec3& pp = k < (int)tp.size() ? tp[(size_t)k] : tp.back(); const Vec3& zz = k < (int)tp.size() ? tz[(size_t)k] : tz.back();
          const size_t u = ((size_t)b * N + k) * R + i;
          for (int c = 0; c < 3; ++c) { oc[u * 3 + c] = pp[(size_t)c]; os[u * 3 + c] = zz[(size_t)c] / 2 + p_.dynamic_safety_dist; }
          oy[u] = 0.0;
        }
        for (int i = 0; i < S; ++i) {
          const size_t u = ((size_t)b * N + k) * R + D + i;
          for (int c = 0; c < 3; ++c) { oc[u * 3 + c] = so[(size_t)i].centroid[(size_t)c]; os[u * 3 + c] = so[(size_t)i].size[(size_t)c] / 2 + p_.static_safety_dist; }
          oy[u] = so[(size_t)i].yaw;
        }
        // linearisation point: previous plan at the same stage, unshifted, else the current position
        const bool have = !firstTime_ && k < (int)currentStatesSol_.size();
        for (int c = 0; c < 3; ++c) lp[((size_t)b * N + k) * 3 + c] = have ? currentStatesSol_[(size_t)k][(size_t)c] : currPos_[(size_t)c];
      }
      if (!firstTime_) {
        for (int k = 0; k < NS && k < (int)currentStatesSol_.size(); ++k) for (int j = 0; j < numStates; ++j) wx[(size_t)b * n + numStates * k + j] = currentStatesSol_[(size_t)k][(size_t)j];
        for (int k = 0; k < N && k < (int)currentControlsSol_.size(); ++k) for (int j = 0; j < numControls; ++j) wx[(size_t)b * n + numStates * NS + numControls * k + j] = currentControlsSol_[(size_t)k][(size_t)j];
      }
    }
    // isDynamic with the reference's index quirk (:1194): the static loop clears entries [0, S), not [D, D+S)
    for (int k = 0; k < N; ++k) for (int i = 0; i < R; ++i) od[(size_t)k * R + i] = (i < D && i >= S) ? 1 : 0;
    std::vector<double> x((size_t)B * n), obj((size_t)B), pr((size_t)B), du((size_t)B);
    std::vector<int32_t> status((size_t)B), iter((size_t)B), ru((size_t)B);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = mpcqp_solve_mpc_batch_host(eng_, &p_, &s_, B, R, x0.data(), xr.data(), R ? oc.data() : nullptr, R ? os.data() : nullptr, R ? oy.data() : nullptr,
                                              R ? od.data() : nullptr, R ? lp.data() : nullptr, wx.data(), x.data(), nullptr, status.data(), iter.data(), ru.data(),
                                              obj.data(), pr.data(), du.data());
    lastQpSolveTime_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != MPCQP_OK) return false;                                                    // mpcPlanner maps any solver failure to `return 0` (:514-518)
    lastStatus_.assign(status.begin(), status.end()); lastIter_.assign(iter.begin(), iter.end());
    states.assign((size_t)B, {}); controls.assign((size_t)B, {});
    for (int b = 0; b < B; ++b) {
      for (int k = 0; k < NS; ++k) states[(size_t)b].push_back(std::vector<double>(x.begin() + (size_t)b * n + numStates * k, x.begin() + (size_t)b * n + numStates * (k + 1)));
      for (int k = 0; k < N; ++k) controls[(size_t)b].push_back(std::vector<double>(x.begin() + (size_t)b * n + numStates * NS + numControls * k, x.begin() + (size_t)b * n + numStates * NS + numControls * (k + 1)));
    }
    return true;
  }

  mpcqp_engine* eng_ = nullptr; bool ok_ = false;
  mpcqp_mpc_params p_; mpcqp_settings s_;
  Vec3 currPos_{0, 0, 0}, currVel_{0, 0, 0};
  bool firstTime_ = true, stateReceived_ = false;
  int lastRefStartIdx_ = 0, obIdx_ = -1, bestIdx_ = -1;
  double lastQpSolveTime_ = 0.0;
  std::vector<Vec3> inputTraj_, trajHist_, ref_, trajScore_;
  std::vector<staticObstacle> staticObstacles_;
  std::vector<ObTraj> dynamicObstaclesPos_, dynamicObstaclesSize_;
  std::vector<std::vector<ObTraj>> obPredPos_, obPredSize_;
  std::vector<std::array<double, 4>> obIntentProb_;
  std::vector<std::vector<double>> currentStatesSol_, currentControlsSol_;
  std::vector<std::vector<std::vector<double>>> candidateStates_, candidateControls_;
  std::vector<double> trajWeightedScore_;
  std::vector<int> lastStatus_, lastIter_, candidateStatus_, candidateIter_;
};

}  // namespace trajPlannerB200
