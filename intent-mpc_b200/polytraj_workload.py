"""Workload generator (inputs only, no solving): the QP that `polyTrajSolver` hands to OsqpEigen — the second in-tree
consumer of the solver boundary — assembled with numpy.

Follows trajectory_planner/include/trajectory_planner/polyTrajSolver.cpp: `constructP` (:241-272, per-segment
minimum-`diffDegree` Hessian on normalised time, full block), `constructQ` (:309-312, zero), `constructA` (:314-575:
position end points / mid points / continuity, velocity, acceleration, jerk and snap continuity scaled by the segment
durations, optional corridor rows) and `constructBound` (:578-760: equalities from the path and the boundary
velocities / accelerations, `softConstraint_` boxes around the mid points, corridor boxes), `avgTimeAllocation`
(:125-138) and `getConstraintNum` (:156-160).  One QP per axis; x, y, z share P and A and differ in the bounds.

Returns a `QpBatch` (B = 3: the x, y and z problems) with the field names of the oracle's container, so the oracle
drivers and `Engine.solve_qp_batch` take it unchanged.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class QpBatch:
    """B problems sharing one CSC pattern (what OsqpEigen::Data would hold per problem)."""
    n: int
    m: int
    P_colptr: np.ndarray   # [n+1] int64, upper-triangular P
    P_rowidx: np.ndarray   # [nnzP]
    P_val: np.ndarray      # [B, nnzP]
    q: np.ndarray          # [B, n]
    A_colptr: np.ndarray   # [n+1]
    A_rowidx: np.ndarray   # [nnzA]
    A_val: np.ndarray      # [B, nnzA]
    l: np.ndarray          # [B, m]
    u: np.ndarray          # [B, m]
    warm_x: np.ndarray     # [B, n]


def _csc(dense, upper=False):
    n = dense.shape[1]
    colptr = [0]; rowidx = []; val = []
    for j in range(n):
        rows = np.nonzero(dense[:, j])[0]
        if upper:
            rows = rows[rows <= j]
        rowidx.extend(rows.tolist()); val.extend(dense[rows, j].tolist()); colptr.append(len(rowidx))
    return np.array(colptr, np.int64), np.array(rowidx, np.int64), np.array(val, np.float64)


def assemble(path, poly_degree=7, diff_degree=4, continuity_degree=4, desired_vel=1.0, init_vel=(0, 0, 0),
             end_vel=(0, 0, 0), init_acc=(0, 0, 0), end_acc=(0, 0, 0), soft_deviation=None, corridor=None):
    """path [K+1, 3] way points.  soft_deviation (3,) -> mid-point boxes (`softConstraint_`, :644-655);
    corridor = (times per segment [S], half-width) -> rows `sum_d t^d c_d` within +-half-width of the straight
    segment (`corridorConstraint_`, :552-572)."""
    path = np.asarray(path, float)
    K = len(path) - 1
    nc = poly_degree + 1
    n = nc * K
    # avgTimeAllocation (:125-138)
    T = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(path, axis=0), axis=1) / desired_vel)])
    P = np.zeros((n, n))
    for s in range(K):                                     # constructP (:257-270)
        for i in range(diff_degree, nc):
            for j in range(diff_degree, nc):
                f = 1.0
                for d in range(diff_degree):
                    f *= float((i - d) * (j - d))
                P[s * nc + i, s * nc + j] = f / float(i + j - 2 * diff_degree + 1)
    rows = []; lo = []; hi = []

    def deriv_row(seg, t, order, scale=1.0):
        r = np.zeros(n)
        for d in range(order, nc):
            f = 1.0
            for e in range(order):
                f *= (d - e)
            r[seg * nc + d] = f * (t ** (d - order)) * scale
        return r

    def eq(r, v):
        rows.append(r); lo.append(np.asarray(v, float) * np.ones(3)); hi.append(np.asarray(v, float) * np.ones(3))

    # position: 2 end points, K-1 mid points (right end of the first K-1 segments), K-1 continuity rows (:318-384)
    eq(deriv_row(0, 0.0, 0), path[0]); eq(deriv_row(K - 1, 1.0, 0), path[K])
    for i in range(K - 1):
        r = deriv_row(i, 1.0, 0)
        if soft_deviation is None:
            eq(r, path[i + 1])
        else:
            rows.append(r); lo.append(path[i + 1] - np.asarray(soft_deviation)); hi.append(path[i + 1] + np.asarray(soft_deviation))
    for i in range(K - 1):
        eq(deriv_row(i, 1.0, 0) - deriv_row(i + 1, 0.0, 0), 0.0)
    # velocity / acceleration: 2 end points + K-1 continuity rows scaled by the neighbours' durations (:388-498)
    for order, v0, v1 in ((1, init_vel, end_vel), (2, init_acc, end_acc)):
        eq(deriv_row(0, 0.0, order), v0); eq(deriv_row(K - 1, 1.0, order), v1)
        for i in range(K - 1):
            dl, dr = T[i + 1] - T[i], T[i + 2] - T[i + 1]
            eq(deriv_row(i, 1.0, order, dr ** order) - deriv_row(i + 1, 0.0, order, dl ** order), 0.0)
    # jerk / snap continuity (:501-548)
    for order in (3, 4):
        if continuity_degree >= order:
            for i in range(K - 1):
                dl, dr = T[i + 1] - T[i], T[i + 2] - T[i + 1]
                eq(deriv_row(i, 1.0, order, dr ** order) - deriv_row(i + 1, 0.0, order, dl ** order), 0.0)
    if corridor is not None:                               # (:552-572) + bounds (:735-760): a box around the chord
        times, half = corridor
        for i in range(K):
            for t in times:
                rows.append(deriv_row(i, t, 0)); c = path[i] + t * (path[i + 1] - path[i])
                lo.append(c - half); hi.append(c + half)
    A = np.array(rows); lo = np.array(lo); hi = np.array(hi)
    m = A.shape[0]
    Pc, Pr, Pv = _csc(P, upper=True)
    Ac, Ar, Av = _csc(A)
    return QpBatch(n=n, m=m, P_colptr=Pc, P_rowidx=Pr, P_val=np.tile(Pv, (3, 1)), q=np.zeros((3, n)),
                   A_colptr=Ac, A_rowidx=Ar, A_val=np.tile(Av, (3, 1)), l=np.ascontiguousarray(lo.T),
                   u=np.ascontiguousarray(hi.T), warm_x=np.zeros((3, n)))


def random_path(seed, K, step=2.5):
    """Way points of an A*/RRT-like path: K segments of roughly `step` metres with bounded turning."""
    rng = np.random.default_rng(seed)
    p = [np.array([rng.uniform(0, 5), rng.uniform(-3, 3), rng.uniform(1, 3)])]
    h = rng.uniform(-0.5, 0.5)
    for _ in range(K):
        h += rng.uniform(-0.6, 0.6)
        p.append(p[-1] + step * rng.uniform(0.6, 1.4) * np.array([np.cos(h), np.sin(h), rng.uniform(-0.15, 0.15)]))
    return np.array(p)


def cases():
    """Named polyTrajSolver-shaped problems used by the parity tests (each: 3 QPs, one per axis)."""
    return {
        "poly_k3": assemble(random_path(1, 3)),
        "poly_k6": assemble(random_path(2, 6), init_vel=(1.0, 0.2, 0.0)),
        "poly_k10_soft": assemble(random_path(3, 10), soft_deviation=(0.3, 0.3, 0.2)),
        "poly_k5_corridor": assemble(random_path(4, 5), corridor=((0.25, 0.5, 0.75), 0.4)),
        "poly_k12_d5": assemble(random_path(5, 12), poly_degree=5, diff_degree=3, continuity_degree=3),
        "poly_k25": assemble(random_path(6, 25), init_vel=(0.5, 0.0, 0.0)),
    }


def path_batch(num_paths, K=8, seed0=0, **kw):
    """`num_paths` candidate paths of K segments each -> one QpBatch of 3 * num_paths QPs.  The CSC pattern of A is shared
    (the duration-scaled continuity rows differ in value only), which is what the batched entry point needs."""
    parts = [assemble(random_path(seed0 + i, K), **kw) for i in range(num_paths)]
    first = parts[0]
    for p in parts:
        assert (p.A_colptr == first.A_colptr).all() and (p.A_rowidx == first.A_rowidx).all()
    cat = lambda k: np.ascontiguousarray(np.concatenate([getattr(p, k) for p in parts], axis=0))
    return QpBatch(n=first.n, m=first.m, P_colptr=first.P_colptr, P_rowidx=first.P_rowidx, P_val=cat("P_val"), q=cat("q"),
                   A_colptr=first.A_colptr, A_rowidx=first.A_rowidx, A_val=cat("A_val"), l=cat("l"), u=cat("u"),
                   warm_x=cat("warm_x"))
