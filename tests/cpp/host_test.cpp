// Test driver for the C++ host layer (intent-mpc_b200/host/*.hpp): run by tests/test_host_cpp.py.
//   host_test facade <problem.bin> <out.bin>   solve one CSC QP through the OsqpEigen::Solver-shaped facade exactly as
//                                              mpcPlanner::solveTraj drives OsqpEigen (mpcPlanner.cpp:436-527)
//   host_test planner <steps> <out.bin>        mpc_node-style receding-horizon loop (mpc_node.cpp:209-236) through the
//                                              mpcPlanner mirror; dumps the last control step's batch for an oracle check
//   host_test nogpu                            the facade and the planner must fail loudly without a CUDA device
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../intent-mpc_b200/host/OsqpEigenB200.hpp"
#include "../../intent-mpc_b200/host/MpcPlannerB200.hpp"

static std::vector<double> rd(FILE* f, size_t n) { std::vector<double> v(n); if (n && fread(v.data(), 8, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
static std::vector<long> rl(FILE* f, size_t n) { std::vector<long> v(n); if (n && fread(v.data(), 8, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
static void wd(FILE* f, const double* p, size_t n) { fwrite(p, 8, n, f); }

static int run_facade(const char* in, const char* out) {
  FILE* f = fopen(in, "rb"); if (!f) return 2;
  std::vector<long> h = rl(f, 4);                 // n, m, nnzP, nnzA
  const long n = h[0], m = h[1], nzp = h[2], nza = h[3];
  std::vector<long> Pr = rl(f, nzp), Pc = rl(f, nzp); std::vector<double> Pv = rd(f, nzp);
  std::vector<long> Ar = rl(f, nza), Ac = rl(f, nza); std::vector<double> Av = rd(f, nza);
  std::vector<double> q = rd(f, n), l = rd(f, m), u = rd(f, m), wx = rd(f, n);
  fclose(f);
  OsqpEigen::SparseMatrix P(n, n), A(m, n);
  P.setFromTriplets(Pr, Pc, Pv); A.setFromTriplets(Ar, Ac, Av);
  OsqpEigen::Vector gq(n), lo(m), up(m), primal(n), dual(m);
  for (long i = 0; i < n; ++i) { gq[i] = q[i]; primal[i] = wx[i]; }
  for (long i = 0; i < m; ++i) { lo[i] = l[i]; up[i] = u[i]; dual[i] = 0.0; }

  OsqpEigen::Solver solver;
  solver.settings()->setVerbosity(false);
  solver.settings()->setWarmStart(true);
  solver.data()->setNumberOfVariables((int)n);
  solver.data()->setNumberOfConstraints((int)m);
  if (!solver.data()->setHessianMatrix(P)) return 3;
  if (!solver.data()->setGradient(gq)) return 3;
  if (!solver.data()->setLinearConstraintsMatrix(A)) return 3;
  if (!solver.data()->setLowerBound(lo)) return 3;
  if (!solver.data()->setUpperBound(up)) return 3;
  if (solver.solveProblem() != OsqpEigen::ErrorExitFlag::WorkspaceNotInitError) return 4;     // not initialised yet
  if (!solver.initSolver()) return 5;
  if (solver.initSolver()) return 6;                                                            // already initialised
  if (!solver.setWarmStart(primal, dual)) return 7;
  if (solver.solveProblem() != OsqpEigen::ErrorExitFlag::NoError) return 8;
  // getSolution() returns a reference to a solver-owned buffer that the next solve overwrites (as in OsqpEigen): copy
  std::vector<double> x(n), y(m);
  { const OsqpEigen::Vector& xs = solver.getSolution(); const OsqpEigen::Vector& ys = solver.getDualSolution();
    for (long i = 0; i < n; ++i) x[i] = xs[i];
    for (long i = 0; i < m; ++i) y[i] = ys[i]; }
  mpcqp_info info; solver.getInfo(&info);
  // re-solve with a changed gradient and bounds (polyTrajSolver-style updates)
  OsqpEigen::Vector gq2(n); for (long i = 0; i < n; ++i) gq2[i] = 0.5 * q[i];
  double st2 = -99, it2 = -1; std::vector<double> x2(n, 0.0);
  if (solver.updateGradient(gq2) && solver.updateBounds(lo, up) && solver.setWarmStart(primal, dual) && solver.solveProblem() == OsqpEigen::ErrorExitFlag::NoError) {
    mpcqp_info i2; solver.getInfo(&i2); st2 = (double)i2.status_val; it2 = (double)i2.iter;
    const OsqpEigen::Vector& xs = solver.getSolution(); for (long i = 0; i < n; ++i) x2[i] = xs[i];
  }
  FILE* g = fopen(out, "wb"); if (!g) return 9;
  const double head[6] = {(double)info.status_val, (double)info.iter, (double)info.rho_updates, info.obj_val, st2, it2};
  wd(g, head, 6); wd(g, x.data(), n); wd(g, y.data(), m); wd(g, x2.data(), n);
  fclose(g);
  solver.clearSolver();
  if (solver.isInitialized()) return 10;
  // an unstructured problem (polyTrajSolver-style use of the facade) runs on the dense generic kernel
  OsqpEigen::Solver bad;
  bad.data()->setNumberOfVariables(2); bad.data()->setNumberOfConstraints(1);
  OsqpEigen::SparseMatrix P2(2, 2), A2(1, 2);
  P2.setFromTriplets({0, 1}, {0, 1}, {1.0, 1.0}); A2.setFromTriplets({0, 0}, {0, 1}, {1.0, 1.0});
  OsqpEigen::Vector q2(2), l2(1), u2(1); q2[0] = q2[1] = 1.0; l2[0] = 0.0; u2[0] = 1.0;
  bad.data()->setHessianMatrix(P2); bad.data()->setGradient(q2); bad.data()->setLinearConstraintsMatrix(A2); bad.data()->setLowerBound(l2); bad.data()->setUpperBound(u2);
  // min 1/2 (a^2 + b^2) + a + b  s.t. 0 <= a + b <= 1  ->  a = b = 0 (the unconstrained minimum (-1,-1) is cut off)
  if (!bad.initSolver()) return 11;
  if (bad.solveProblem() != OsqpEigen::ErrorExitFlag::NoError) return 12;
  if (bad.getStatus() != OsqpEigen::Status::Solved) return 13;
  const OsqpEigen::Vector& xb = bad.getSolution();
  if (fabs(xb[0]) > 5e-3 || fabs(xb[1]) > 5e-3) return 14;
  // polyTrajSolver::updateProblem (polyTrajSolver.cpp:225-239): new bounds on the same solver object, solved again;
  // 0.5 <= a + b <= 1  ->  a = b = 0.25
  OsqpEigen::Vector l3(1), u3(1); l3[0] = 0.5; u3[0] = 1.0;
  if (!bad.updateBounds(l3, u3)) return 15;
  if (!bad.solve()) return 16;
  const OsqpEigen::Vector& xc = bad.getSolution();
  if (fabs(xc[0] - 0.25) > 5e-3 || fabs(xc[1] - 0.25) > 5e-3) return 17;
  // polyTrajSolver::setUpProblem (:180-191) on a used solver: clearSolver + clear*Matrix + set* + initSolver again
  bad.clearSolver(); bad.data()->clearHessianMatrix(); bad.data()->clearLinearConstraintsMatrix();
  bad.data()->setNumberOfVariables(2); bad.data()->setNumberOfConstraints(1);
  if (!bad.data()->setHessianMatrix(P2) || !bad.data()->setGradient(q2) || !bad.data()->setLinearConstraintsMatrix(A2) ||
      !bad.data()->setLowerBound(l2) || !bad.data()->setUpperBound(u2) || !bad.initSolver() || !bad.solve()) return 18;
  if (fabs(bad.getSolution()[0]) > 5e-3) return 19;
  return 0;
}

static int run_planner(int steps, const char* out) {
  using namespace trajPlannerB200;
  mpcPlanner mpc(0);
  if (!mpc.engineReady()) { fprintf(stderr, "%s\n", mpc.lastError()); return 20; }
  mpc.updateMaxVel(5.0); mpc.updateMaxAcc(20.0);
  const double dt = 0.1;
  std::vector<Vec3> path;                              // straight reference at 2.5 m/s along x (ref_trajectory_dynus_benchmark spacing is coarser; speed shape only)
  for (int i = 0; i < 600; ++i) path.push_back({0.25 * i, 0.0, 2.0});
  mpc.updatePath(path, dt);
  Vec3 pos{0.0, 0.2, 2.0}, vel{2.0, 0.0, 0.0};
  std::vector<staticObstacle> so{{{9.0, 0.6, 2.0}, {0.4, 0.4, 4.0}, 0.3}};
  mpc.updateStaticObstacles(so);
  FILE* g = fopen(out, "wb"); if (!g) return 21;
  for (int step = 0; step < steps; ++step) {
    // two dynamic obstacles crossing the path, 4 intents each, 31 prediction steps (predictor_param.yaml:2-3)
    std::vector<std::vector<mpcPlanner::ObTraj>> pp(2), ps(2); std::vector<std::array<double, 4>> prob(2);
    for (int ob = 0; ob < 2; ++ob) {
      const Vec3 c{6.0 + 7.0 * ob + 0.02 * step, (ob ? -2.5 : 2.5) + (ob ? 0.08 : -0.08) * step, 2.0};
      const double vy = ob ? 0.8 : -0.8;
      pp[ob].resize(4); ps[ob].resize(4);
      for (int it = 0; it < 4; ++it) for (int k = 0; k <= 30; ++k) {
        const double t = 0.1 * k;
        Vec3 q = c;
        if (it == FORWARD) q[1] += vy * t; else if (it == LEFT) { q[0] -= 0.5 * t; q[1] += 0.7 * vy * t; } else if (it == RIGHT) { q[0] += 0.5 * t; q[1] += 0.7 * vy * t; }
        pp[ob][it].push_back(q); ps[ob][it].push_back({1.3 + 0.01 * k, 1.3 + 0.01 * k, 1.1});
      }
      prob[ob] = ob ? std::array<double, 4>{0.5, 0.2, 0.2, 0.1} : std::array<double, 4>{0.3, 0.4, 0.1, 0.2};
    }
    mpc.updateCurrStates(pos, vel);
    mpc.updatePredObstacles(pp, ps, prob);
    if (!mpc.makePlanWithPred()) { fprintf(stderr, "plan failed at step %d: %s\n", step, mpc.lastError()); fclose(g); return 22; }
    const double row[10] = {(double)step, pos[0], pos[1], pos[2], (double)mpc.lastStatus().size(), (double)mpc.bestCandidate(),
                            (double)(mpc.lastStatus().empty() ? 0 : mpc.lastStatus()[0]), (double)(mpc.lastIterations().empty() ? 0 : mpc.lastIterations()[0]), mpc.getLastQpSolveTime(), mpc.getRef(0.0)[0]};
    wd(g, row, 10);
    pos = mpc.getPos(dt); vel = mpc.getVel(dt);          // perfect tracking roll-forward (mpc_node.cpp:223-224)
  }
  fclose(g);
  return 0;
}

static int run_nogpu() {
  trajPlannerB200::mpcPlanner mpc(0);
  if (mpc.engineReady()) return 0;                       // a GPU is present: nothing to check here
  if (mpc.makePlan()) return 30;                         // must not produce a plan from a CPU path
  OsqpEigen::Solver s; s.data()->setNumberOfVariables(2); s.data()->setNumberOfConstraints(1);
  OsqpEigen::SparseMatrix P2(2, 2), A2(1, 2);
  P2.setFromTriplets({0, 1}, {0, 1}, {1.0, 1.0}); A2.setFromTriplets({0, 0}, {0, 1}, {1.0, 1.0});
  OsqpEigen::Vector q2(2), l2(1), u2(1);
  s.data()->setHessianMatrix(P2); s.data()->setGradient(q2); s.data()->setLinearConstraintsMatrix(A2); s.data()->setLowerBound(l2); s.data()->setUpperBound(u2);
  if (s.initSolver()) return 31;
  if (s.solveProblem() != OsqpEigen::ErrorExitFlag::WorkspaceNotInitError) return 32;
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && !strcmp(argv[1], "facade")) return run_facade(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "planner")) return run_planner(atoi(argv[2]), argv[3]);
  if (argc >= 2 && !strcmp(argv[1], "nogpu")) return run_nogpu();
  fprintf(stderr, "usage: host_test facade|planner|nogpu ...\n");
  return 1;
}
