// Test-only C shim around the mpcPlanner mirror (intent-mpc_b200/host/MpcPlannerB200.hpp) so that tests/test_planner_parity.py
// can drive it from Python in lockstep with the oracle restatement (oracle/mpc_planner.py) and the device kernels.
// Built by the test into tests/cpp/libplanner_capi.so; not part of the product library.
#include <cstring>
#include <vector>

#include "../../intent-mpc_b200/host/MpcPlannerB200.hpp"

using namespace trajPlannerB200;
using P = mpcPlanner;

static std::vector<std::vector<P::ObTraj>> unpack(const double* a, int D, int T) {   // [D][4][T][3]
  std::vector<std::vector<P::ObTraj>> out((size_t)D, std::vector<P::ObTraj>(4));
  for (int d = 0; d < D; ++d) for (int it = 0; it < 4; ++it) for (int k = 0; k < T; ++k) {
    const double* q = a + ((((size_t)d * 4 + it) * T) + k) * 3;
    out[(size_t)d][(size_t)it].push_back({q[0], q[1], q[2]});
  }
  return out;
}

extern "C" {
void* pl_create(int device) { P* p = new P(device); return p; }
void pl_destroy(void* h) { delete (P*)h; }
int pl_ready(void* h) { return ((P*)h)->engineReady() ? 1 : 0; }
const char* pl_last_error(void* h) { return ((P*)h)->lastError(); }
void pl_set_params(void* h, const mpcqp_mpc_params* p) { ((P*)h)->params() = *p; }
void pl_update_max_vel(void* h, double v) { ((P*)h)->updateMaxVel(v); }
void pl_update_max_acc(void* h, double a) { ((P*)h)->updateMaxAcc(a); }
void pl_update_path(void* h, const double* path, int n, double ts) {
  std::vector<Vec3> v; for (int i = 0; i < n; ++i) v.push_back({path[3 * i], path[3 * i + 1], path[3 * i + 2]});
  ((P*)h)->updatePath(v, ts);
}
void pl_update_curr_states(void* h, const double* pos, const double* vel) { ((P*)h)->updateCurrStates({pos[0], pos[1], pos[2]}, {vel[0], vel[1], vel[2]}); }
void pl_update_static(void* h, const double* csy, int S) {           // [S][7] centroid, size, yaw
  std::vector<staticObstacle> so;
  for (int i = 0; i < S; ++i) { const double* q = csy + 7 * i; so.push_back({{q[0], q[1], q[2]}, {q[3], q[4], q[5]}, q[6]}); }
  ((P*)h)->updateStaticObstacles(so);
}
void pl_update_dynamic(void* h, const double* pos, const double* vel, const double* size, int D) {
  std::vector<Vec3> p, v, s;
  for (int i = 0; i < D; ++i) { p.push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}); v.push_back({vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]}); s.push_back({size[3 * i], size[3 * i + 1], size[3 * i + 2]}); }
  ((P*)h)->updateDynamicObstacles(p, v, s);
}
void pl_update_pred(void* h, const double* pp, const double* ps, const double* prob, int D, int T) {
  std::vector<std::array<double, 4>> pr;
  for (int d = 0; d < D; ++d) pr.push_back({prob[4 * d], prob[4 * d + 1], prob[4 * d + 2], prob[4 * d + 3]});
  if (D == 0) { ((P*)h)->updatePredObstacles({}, {}, {}); return; }
  ((P*)h)->updatePredObstacles(unpack(pp, D, T), unpack(ps, D, T), pr);
}
int pl_make_plan(void* h) { return ((P*)h)->makePlan() ? 1 : 0; }
int pl_make_plan_with_pred(void* h) { return ((P*)h)->makePlanWithPred() ? 1 : 0; }

static void flat(const std::vector<std::vector<double>>& v, double* out) { for (auto& r : v) for (double x : r) *out++ = x; }
int pl_get_plan(void* h, double* states, double* controls) {
  P* p = (P*)h; flat(p->currentStates(), states); flat(p->currentControls(), controls); return (int)p->currentStates().size();
}
int pl_num_candidates(void* h) { return (int)((P*)h)->candidateStates().size(); }
void pl_get_candidate(void* h, int c, double* states, double* controls) { P* p = (P*)h; flat(p->candidateStates()[(size_t)c], states); flat(p->candidateControls()[(size_t)c], controls); }
int pl_get_scores(void* h, double* score, double* weighted) {
  P* p = (P*)h;
  for (size_t i = 0; i < p->trajScore().size(); ++i) for (int c = 0; c < 3; ++c) score[3 * i + c] = p->trajScore()[i][(size_t)c];
  for (size_t i = 0; i < p->trajWeightedScore().size(); ++i) weighted[i] = p->trajWeightedScore()[i];
  return (int)p->trajScore().size();
}
int pl_best(void* h) { return ((P*)h)->bestCandidate(); }
int pl_closest_obstacle(void* h) { return ((P*)h)->closestObstacle(); }
int pl_last_status(void* h, int* status, int* iter) {
  P* p = (P*)h;
  for (size_t i = 0; i < p->lastStatus().size(); ++i) { status[i] = p->lastStatus()[i]; iter[i] = p->lastIterations()[i]; }
  return (int)p->lastStatus().size();
}
void pl_get_pos(void* h, double t, double* o) { Vec3 v = ((P*)h)->getPos(t); memcpy(o, v.data(), 24); }
void pl_get_vel(void* h, double t, double* o) { Vec3 v = ((P*)h)->getVel(t); memcpy(o, v.data(), 24); }
void pl_get_acc(void* h, double t, double* o) { Vec3 v = ((P*)h)->getAcc(t); memcpy(o, v.data(), 24); }
void pl_get_ref(void* h, double t, double* o) { Vec3 v = ((P*)h)->getRef(t); memcpy(o, v.data(), 24); }
int pl_get_trajectory(void* h, double* out) { std::vector<Vec3> t; ((P*)h)->getTrajectory(t); for (size_t i = 0; i < t.size(); ++i) memcpy(out + 3 * i, t[i].data(), 24); return (int)t.size(); }
int pl_get_reference_traj(void* h, double* out) { std::vector<Vec3> t; ((P*)h)->getReferenceTraj(t); for (size_t i = 0; i < t.size(); ++i) memcpy(out + 3 * i, t[i].data(), 24); return (int)t.size(); }
double pl_get_ts(void* h) { return ((P*)h)->getTs(); }
double pl_get_horizon(void* h) { return ((P*)h)->getHorizon(); }
// the hypothesis rows of getIntentComb for the current state: obstacle position / size of candidate c, list entry r, step k
int pl_intent_comb(void* h, int* ob_idx, int* rows_per_cand, double* pos, double* size, int max_rows, int T) {   // pos/size [6][max_rows][T][3]
  P* p = (P*)h; std::vector<std::vector<P::ObTraj>> cp, cs; int ob = -1;
  p->getIntentComb(ob, cp, cs); *ob_idx = ob;
  for (size_t c = 0; c < cp.size(); ++c) {
    rows_per_cand[c] = (int)cp[c].size();
    for (size_t r = 0; r < cp[c].size() && (int)r < max_rows; ++r) for (int k = 0; k < T && k < (int)cp[c][r].size(); ++k) {
      const size_t o = (((c * (size_t)max_rows + r) * (size_t)T) + (size_t)k) * 3;
      memcpy(pos + o, cp[c][r][(size_t)k].data(), 24); memcpy(size + o, cs[c][r][(size_t)k].data(), 24);
    }
  }
  return (int)cp.size();
}
}
