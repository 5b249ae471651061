"""QPs on which the reference's OSQP binary reports the statuses no mpcPlanner instance with IEEE-inf bounds ever reaches:
primal infeasible (-3), dual infeasible (-4) and their `inaccurate` forms (3, 4; third_party/osqp/constants.h:18-30), with
the OSQP_NAN fill of x and y (constants.h:95-97: the NUMBER 2143289344.0) and obj = +-OSQP_INFTY.  Shared by the golden
generator (make_golden_infeasible.py) and tests/test_infeasible.py.  Each case: (QpBatch, settings overrides)."""
import dataclasses

import numpy as np

from intent_mpc_b200 import workloads as W
from intent_mpc_b200.polytraj_workload import QpBatch
from tests.helpers import to_qp_batch


def _csc(M, upper=False):
    M = np.asarray(M, dtype=np.float64)
    n = M.shape[1]; cp = [0]; ri = []; v = []
    for j in range(n):
        rows = np.nonzero(M[:, j])[0]
        if upper:
            rows = rows[rows <= j]
        ri += rows.tolist(); v += M[rows, j].tolist(); cp.append(len(ri))
    return np.array(cp, dtype=np.int64), np.array(ri, dtype=np.int64), np.array(v, dtype=np.float64)


def _qp(P, q, A, l, u):
    n, m = len(q), len(l)
    Pc, Pr, Pv = _csc(P, True); Ac, Ar, Av = _csc(A)
    return QpBatch(n=n, m=m, P_colptr=Pc, P_rowidx=Pr, P_val=Pv[None], q=np.asarray(q, float)[None], A_colptr=Ac, A_rowidx=Ar,
                   A_val=Av[None], l=np.asarray(l, float)[None], u=np.asarray(u, float)[None], warm_x=np.zeros((1, n)))


def _random_infeasible(seed, n=24, m=36):
    """A strictly convex QP whose constraint set is empty: row m-1 is the negative of row 0 with a gap of 1."""
    r = np.random.default_rng(seed)
    G = r.normal(size=(n, n)); P = G @ G.T / n + 0.1 * np.eye(n)
    A = r.normal(size=(m, n)) * (r.uniform(size=(m, n)) < 0.3)
    A[:, 0] += 1.0
    x = r.normal(size=n)
    l = A @ x - r.uniform(0.1, 1.0, m); u = A @ x + r.uniform(0.1, 1.0, m)
    A[m - 1] = A[0]; l[m - 1] = u[0] + 1.0; u[m - 1] = u[0] + 2.0
    return _qp(P, r.normal(size=n), A, l, u)


def _random_unbounded(seed, n=20, m=24):
    """Positive semidefinite P with a null direction d, q'd < 0 and no constraint bounding d: dual infeasible."""
    r = np.random.default_rng(seed)
    d = np.zeros(n); d[0] = 1.0
    G = r.normal(size=(n, n)); G[:, 0] = 0.0; G[0, :] = 0.0
    P = G @ G.T / n
    A = r.normal(size=(m, n)) * (r.uniform(size=(m, n)) < 0.4); A[:, 0] = 0.0
    A[0, 0] = 1.0; A[0, 1:] = 0.0
    l = -np.ones(m); u = np.ones(m)
    l[0] = 0.0; u[0] = np.inf                                      # x_0 >= 0 only: the cost -x_0 runs away
    q = r.normal(size=n) * 0.1; q[0] = -1.0
    return _qp(P, q, A, l, u)


def _mpc_finite(horizon, B):
    """The stress set's mpcPlanner QPs (tight limits, starts outside the box, an obstacle around the start) with their
    infinite bounds written as +-OSQP_INFTY = 1e30 instead of IEEE inf: OSQP's certificates then work (with inf they evaluate
    inf * 0 = NaN and never fire), and the reference declares the instances that cannot be feasible."""
    qb = to_qp_batch(W.stress_batch(B, horizon=horizon))
    return dataclasses.replace(qb, l=np.where(np.isinf(qb.l), -1e30, qb.l), u=np.where(np.isinf(qb.u), 1e30, qb.u))


def cases():
    pinf = _qp(np.eye(2), [1, 1], [[1, 0], [0, 1], [1, 1]], [0, 0, -10], [5, 5, -6])
    dinf = _qp(np.diag([0.0, 1.0]), [-1, 0.5], [[1, 0], [0, 1]], [0, -1], [1e30, 1])
    dinf_ieee = _qp(np.diag([0.0, 1.0]), [-1, 0.5], [[1, 0], [0, 1]], [0, -1], [np.inf, 1])
    return {
        "pinf2": (pinf, {}),                                        # -3 at the first check (iteration 25)
        "pinf2_cut20": (pinf, {"max_iter": 20}),                    # 3: only the 10x relaxed certificate holds when max_iter hits
        "pinf2_cut24": (pinf, {"max_iter": 24}),                    # -3 from the final exact check at max_iter
        "pinf2_cut10": (pinf, {"max_iter": 10}),                    # -2: too early for either
        "dinf2": (dinf, {}),                                        # -4
        "dinf2_ieee": (dinf_ieee, {}),                              # -4 with an IEEE-inf upper bound
        "dinf2_cut7": (dinf, {"max_iter": 7}),                      # 4
        "dinf2_tight": (dinf, {"max_iter": 15, "eps_dual_inf": 1e-7}),   # 4
        "dinf2_cut5": (dinf, {"max_iter": 5}),                      # -2
        "pinf_rand": (_random_infeasible(1), {}),
        "pinf_rand_b": (_random_infeasible(2, n=40, m=70), {}),
        "dinf_rand": (_random_unbounded(3), {}),
        "dinf_rand_b": (_random_unbounded(4, n=36, m=30), {}),
        "mpc_finite_h30": (_mpc_finite(30, 16), {}),                # structured mpcPlanner QPs (stage kernels): 1 / -3 / -2 mixed
        "mpc_finite_h60": (_mpc_finite(60, 12), {}),
    }
