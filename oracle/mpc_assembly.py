"""ORACLE (test infrastructure, not product): numpy restatement of mpcPlanner's QP assembly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates, function by function, how the reference turns one control step into the
OSQP problem (P, q, A, l, u):

  trajectory_planner/include/trajectory_planner/mpcPlanner.cpp   (abbreviated MP.cpp below)
    setDynamicsMatrices            MP.cpp:891-901
    setInequalityConstraints       MP.cpp:904-921
    setWeightMatrices              MP.cpp:925-931
    castMPCToQPHessian             MP.cpp:932-951   (float32 rounding, global-index R rotation quirk)
    castMPCToQPGradient            MP.cpp:952-966
    castMPCToQPConstraintMatrix    MP.cpp:984-1072  (float32 rounding of Ad/Bd entries)
    castMPCToQPConstraintVectors   MP.cpp:1074-1146
    updateObstacleParam            MP.cpp:1148-1197 (isDyamic[j][i]=0 index quirk at :1194)
    solveTraj                      MP.cpp:375-541   (problem sizes :450-452, warm start :485-509)

Everything is batched over a leading axis B so that thousands of instances assemble in
vectorised numpy; the sparsity pattern depends only on (horizon, numObs, numHalfSpace, slack
column choice) and the CSC arrays are emitted with sorted row indices per column, which is the
canonical form OsqpEigen hands to OSQP (OsqpEigen/SparseMatrixHelper.tpp:15-53).
"""
from __future__ import annotations

import dataclasses
import numpy as np

NUM_STATES = 8      # MP.h:42
NUM_CONTROLS = 5    # MP.h:43


@dataclasses.dataclass
class MpcParams:
    """Planner parameters; defaults are the intent_mpc_demo shape (SURVEY.md §8d;
    autonomous_flight/cfg/mpc_navigation/planner_param.yaml:25-39, flight_base.yaml:8-9)."""
    horizon: int = 30
    ts: float = 0.1
    max_vel: float = 5.0
    max_acc: float = 20.0
    y_min: float = -5.0
    y_max: float = 5.0
    z_min: float = 0.5
    z_max: float = 4.5
    static_safety_dist: float = 0.8
    dynamic_safety_dist: float = 1.5
    static_slack: float = 0.01
    dynamic_slack: float = 0.2
    position_weight: float = 1000.0
    velocity_weight: float = 0.0
    acceleration_weight: float = 10.0

    @property
    def N(self) -> int:           # mpcWindow, MP.cpp:382
        return self.horizon - 1

    @property
    def n(self) -> int:           # MP.cpp:450
        return NUM_STATES * (self.N + 1) + NUM_CONTROLS * self.N

    def m(self, num_obs: int, num_half_space: int = 0) -> int:   # MP.cpp:452
        return 2 * NUM_STATES * (self.N + 1) + NUM_CONTROLS * self.N + (num_half_space + num_obs) * self.N


def _f32(v):
    """`float value = ...` in MP.cpp:940,946,1003,1014 — round to binary32 and widen again."""
    return np.asarray(v, dtype=np.float64).astype(np.float32).astype(np.float64)


def dynamics_matrices(p: MpcParams):
    """MP.cpp:891-901.  Slack rows of Ad are zero; Bd maps the two slack inputs to the slack states."""
    Ad = np.zeros((8, 8)); Bd = np.zeros((8, 5))
    Ad[0:3, 0:3] = np.eye(3)
    Ad[0:3, 3:6] = np.eye(3) * p.ts
    Ad[3:6, 3:6] = np.eye(3)
    Bd[0:3, 0:3] = np.eye(3) * 1 / 2 * (p.ts ** 2)
    Bd[3:6, 0:3] = np.eye(3) * p.ts
    Bd[6:8, 3:5] = np.eye(2)
    return Ad, Bd


def box_bounds(p: MpcParams):
    """MP.cpp:904-921.  x free, slack states free (IEEE inf); slack inputs in [0, 1-(1-ratio)^2]."""
    inf = np.inf
    x_min = np.array([-inf, p.y_min, p.z_min, -p.max_vel, -p.max_vel, -p.max_vel, -inf, -inf])
    x_max = np.array([inf, p.y_max, p.z_max, p.max_vel, p.max_vel, p.max_vel, inf, inf])
    sks = 1.0 - (1 - p.static_slack) ** 2
    skd = 1.0 - (1 - p.dynamic_slack) ** 2
    u_min = np.array([-p.max_acc, -p.max_acc, -p.max_acc, 0.0, 0.0])
    u_max = np.array([p.max_acc, p.max_acc, p.max_acc, skd, sks])
    return x_min, x_max, u_min, u_max


def weight_diagonals(p: MpcParams):
    """MP.cpp:925-931."""
    Q = np.array([p.position_weight] * 3 + [p.velocity_weight] * 3 + [100.0, 1000.0])
    R = np.array([p.acceleration_weight] * 3 + [1.0, 1.0])
    return Q, R


def hessian_diagonal(p: MpcParams):
    """MP.cpp:932-951.  Returns the dense diagonal (n); zero entries are *not inserted* in the
    reference (`if (value != 0)`), so nnz(P) counts only the non-zeros.  R is indexed with the GLOBAL
    variable index modulo 5 (MP.cpp:945), which rotates R unless 8*horizon % 5 == 0."""
    Q, R = weight_diagonals(p)
    n = p.n
    idx = np.arange(n)
    nx = NUM_STATES * (p.N + 1)
    d = np.where(idx < nx, _f32(Q)[idx % NUM_STATES], _f32(R)[idx % NUM_CONTROLS])
    return d


def obstacle_param(p: MpcParams, static_obs, dyn_pos, dyn_size):
    """updateObstacleParam, MP.cpp:1148-1197, for ONE instance.

    static_obs: list of (centroid(3), size(3), yaw); dyn_pos/dyn_size: list (per obstacle) of
    arrays [steps_i, 3].  Returns oxyz[N,numObs,3], osize[N,numObs,3] (semi-axes + safety),
    yaw[N,numObs], is_dyn[N,numObs] with dynamic obstacles first.  Quirk kept: the static loop
    writes isDyamic[j][i] = 0 with i (not i+numDynamicOb), MP.cpp:1194."""
    N = p.N
    nd, ns = len(dyn_pos), len(static_obs)
    num_obs = nd + ns
    oxyz = np.zeros((N, num_obs, 3)); osize = np.zeros((N, num_obs, 3))
    yaw = np.zeros((N, num_obs)); is_dyn = np.zeros((N, num_obs), dtype=np.int32)
    for j in range(N):
        for i in range(nd):
            pos = np.asarray(dyn_pos[i]); siz = np.asarray(dyn_size[i])
            jj = j if j < len(pos) else len(pos) - 1          # MP.cpp:1166,1175-1184 (.back())
            oxyz[j, i] = pos[jj]
            osize[j, i] = siz[jj] / 2 + p.dynamic_safety_dist
            yaw[j, i] = 0.0
            is_dyn[j, i] = 1
        for i in range(ns):
            c, s, yw = static_obs[i]
            oxyz[j, i + nd] = c
            osize[j, i + nd] = np.asarray(s) / 2 + p.static_safety_dist
            yaw[j, i + nd] = yw
            is_dyn[j, i] = 0                                   # MP.cpp:1194 (sic)
    return oxyz, osize, yaw, is_dyn


def ellipsoid_linearisation(c, oxyz, osize, yaw):
    """f and its gradient at the linearisation point, exactly as spelled in MP.cpp:1052-1056 and
    :1131-1137.  c: [...,N,1,3]-broadcastable; oxyz/osize: [...,N,numObs,3]; yaw [...,N,numObs]."""
    cx, cy, cz = c[..., 0], c[..., 1], c[..., 2]
    ox, oy, oz = oxyz[..., 0], oxyz[..., 1], oxyz[..., 2]
    sx, sy, sz = osize[..., 0], osize[..., 1], osize[..., 2]
    cs, sn = np.cos(yaw), np.sin(yaw)
    xi = (cx - ox) * cs + (cy - oy) * sn
    eta = -(cx - ox) * sn + (cy - oy) * cs
    fxyz = xi ** 2 / sx ** 2 + eta ** 2 / sy ** 2 + (cz - oz) ** 2 / sz ** 2
    fxx = 2 * xi / sx ** 2 * cs + 2 * eta / sy ** 2 * (-sn)
    fyy = 2 * xi / sx ** 2 * sn + 2 * eta / sy ** 2 * cs
    fzz = 2 * (cz - oz) / sz ** 2
    low = 1 - fxyz + fxx * cx + fyy * cy + fzz * cz
    return fxx, fyy, fzz, low


_PATTERN_CACHE = {}


def _constraint_pattern(N, num_obs, nhs, is_dyn, Adf, Bdf):
    """Entry list of castMPCToQPConstraintMatrix in the reference's insertion order (MP.cpp:989-1071) and its CSC sort."""
    key = (N, num_obs, nhs, is_dyn.tobytes(), Adf.tobytes(), Bdf.tobytes())
    hit = _PATTERN_CACHE.get(key)
    if hit is not None:
        return hit
    nx, nu = NUM_STATES, NUM_CONTROLS
    n = nx * (N + 1) + nu * N
    rows, cols, kind, aux = [], [], [], []      # kind: 0 const value, 1 obstacle gradient (aux=(k,j,c)), 2 half-space
    def add(r, c, v): rows.append(r); cols.append(c); kind.append(0); aux.append(v)
    for i in range(nx * (N + 1)):
        add(i, i, -1.0)
    for i in range(N):
        for j in range(nx):
            for k in range(nx):
                if Adf[j, k] != 0: add(nx * (i + 1) + j, nx * i + k, Adf[j, k])
    for i in range(N):
        for j in range(nx):
            for k in range(nu):
                if Bdf[j, k] != 0: add(nx * (i + 1) + j, nu * i + k + nx * (N + 1), Bdf[j, k])
    for i in range(n):
        add(i + (N + 1) * nx, i, 1.0)
    base_hs = 2 * nx * (N + 1) + nu * N
    if nhs:
        for i in range(N):
            for which, comp in [(0, 0), (0, 1), (1, 0), (1, 1)]:
                rr = base_hs + nhs * i + which
                rows.append(rr); cols.append(nx * i + comp); kind.append(2); aux.append((which, comp))
    base_ob = base_hs + nhs * N
    for i in range(N):
        for j in range(num_obs):
            r = base_ob + i * num_obs + j
            for c in range(3):
                rows.append(r); cols.append(nx * i + c); kind.append(1); aux.append((i, j, c))
            sc = 3 if is_dyn[i, j] else 4
            add(r, nx * (N + 1) + nu * i + sc, -1.0)
    rows = np.array(rows, dtype=np.int64); cols = np.array(cols, dtype=np.int64)
    order = np.lexsort((rows, cols))
    A_rowidx = rows[order]
    A_colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(A_colptr, cols + 1, 1)
    A_colptr = np.cumsum(A_colptr)
    kind = np.array(kind)
    const_pos = np.nonzero(kind == 0)[0]; const_val = np.array([aux[e] for e in const_pos], dtype=np.float64)
    grad_pos = np.nonzero(kind == 1)[0]; grad_idx = np.array([aux[e] for e in grad_pos], dtype=np.int64).reshape(-1, 3)
    hs_pos = np.nonzero(kind == 2)[0]; hs_idx = np.array([aux[e] for e in hs_pos], dtype=np.int64).reshape(-1, 2)
    out = (A_colptr, A_rowidx, order, const_pos, const_val, grad_pos, grad_idx, hs_pos, hs_idx, len(rows))
    if len(_PATTERN_CACHE) < 64:
        _PATTERN_CACHE[key] = out
    return out


@dataclasses.dataclass
class QpBatch:
    """B problems sharing one CSC pattern (what OsqpEigen::Data would hold per problem)."""
    n: int
    m: int
    P_colptr: np.ndarray   # [n+1] int64, upper-triangular (diagonal) P
    P_rowidx: np.ndarray   # [nnzP]
    P_val: np.ndarray      # [B, nnzP]
    q: np.ndarray          # [B, n]
    A_colptr: np.ndarray   # [n+1]
    A_rowidx: np.ndarray   # [nnzA]  (pattern may differ per problem only via slack column: see A_rowidx_b)
    A_val: np.ndarray      # [B, nnzA]
    l: np.ndarray          # [B, m]
    u: np.ndarray          # [B, m]
    warm_x: np.ndarray     # [B, n]
    # the pattern is shared across the batch only if is_dyn is; assemble_batch asserts that.


def assemble_batch(p: MpcParams, x0, xref, oxyz, osize, yaw, is_dyn, lin_pt, warm_x=None,
                   half_space=None) -> QpBatch:
    """castMPCToQP{Hessian,Gradient,ConstraintMatrix,ConstraintVectors}, MP.cpp:932-1146.

    x0 [B,6] (pos, vel; slack states start at 0, MP.cpp:401-408); xref [B,N+1,3];
    oxyz/osize [B,N,numObs,3]; yaw [B,N,numObs]; is_dyn [N,numObs] (shared across the batch so that
    one CSC pattern serves all problems); lin_pt [B,N,3] = currentStatesSol_[i](0:3) or currPos_
    (MP.cpp:1042-1051); half_space = None or (halfMax[B,3], halfMin[B,3]) (MP.cpp:1027-1038)."""
    x0 = np.asarray(x0, dtype=np.float64)
    B = x0.shape[0]
    N = p.N
    nx, nu = NUM_STATES, NUM_CONTROLS
    n = p.n
    num_obs = oxyz.shape[2]
    nhs = 0 if half_space is None else 2
    m = p.m(num_obs, nhs)
    Ad, Bd = dynamics_matrices(p)
    Adf, Bdf = _f32(Ad), _f32(Bd)
    x_min, x_max, u_min, u_max = box_bounds(p)
    Q, _ = weight_diagonals(p)

    # ---- P (MP.cpp:932-951): diagonal, zeros skipped
    pd = hessian_diagonal(p)
    nzp = np.nonzero(pd)[0]
    P_colptr = np.zeros(n + 1, dtype=np.int64)
    P_colptr[1:] = np.cumsum(pd != 0)
    P_rowidx = nzp.astype(np.int64)
    P_val = np.broadcast_to(pd[nzp], (B, len(nzp))).copy()

    # ---- q (MP.cpp:952-966): -Q * xRef on states (full-precision Q), zero on inputs
    q = np.zeros((B, n))
    xr = np.zeros((B, N + 1, nx)); xr[:, :, 0:3] = xref
    q[:, : nx * (N + 1)] = (-(xr) * Q[None, None, :]).reshape(B, -1)

    # ---- A pattern (MP.cpp:989-1071) as COO, then sorted into CSC (cached per shape: the pattern depends only on the
    # horizon, the obstacle count, the half-space count and the slack column of every obstacle row)
    A_colptr, A_rowidx, order, const_pos, const_val, grad_pos, grad_idx, hs_pos, hs_idx, n_entries = \
        _constraint_pattern(N, num_obs, nhs, np.ascontiguousarray(is_dyn, dtype=np.int32), Adf, Bdf)

    # ---- values
    fxx, fyy, fzz, low = ellipsoid_linearisation(np.asarray(lin_pt)[:, :, None, :], oxyz, osize, yaw) \
        if num_obs else (None, None, None, np.zeros((B, N, 0)))
    vals = np.zeros((B, n_entries))
    vals[:, const_pos] = const_val[None, :]
    if num_obs:
        grads = np.stack([fxx, fyy, fzz], axis=1)                    # [B, 3, N, numObs]
        vals[:, grad_pos] = grads[:, grad_idx[:, 2], grad_idx[:, 0], grad_idx[:, 1]]
    if nhs:
        hsv = np.stack([np.asarray(half_space[0]), np.asarray(half_space[1])], axis=1)   # [B, 2, 3]
        vals[:, hs_pos] = hsv[:, hs_idx[:, 0], hs_idx[:, 1]]
    A_val = vals[:, order]

    # ---- l, u (MP.cpp:1074-1146)
    l = np.zeros((B, m)); u = np.zeros((B, m))
    x0full = np.zeros((B, nx)); x0full[:, 0:6] = x0
    l[:, 0:nx] = -x0full; u[:, 0:nx] = -x0full
    o = nx * (N + 1)
    l[:, o:o + nx * (N + 1)] = np.tile(x_min, N + 1); u[:, o:o + nx * (N + 1)] = np.tile(x_max, N + 1)
    o += nx * (N + 1)
    l[:, o:o + nu * N] = np.tile(u_min, N); u[:, o:o + nu * N] = np.tile(u_max, N)
    o += nu * N
    if nhs:
        for i in range(N):
            l[:, o + nhs * i + 0] = -np.inf; u[:, o + nhs * i + 0] = half_space[0][:, 2]
            l[:, o + nhs * i + 1] = half_space[1][:, 2]; u[:, o + nhs * i + 1] = np.inf
        o += nhs * N
    l[:, o:] = low.reshape(B, -1); u[:, o:] = np.inf

    if warm_x is None:
        warm_x = np.zeros((B, n))
    return QpBatch(n=n, m=m, P_colptr=P_colptr, P_rowidx=P_rowidx, P_val=P_val, q=q,
                   A_colptr=A_colptr, A_rowidx=A_rowidx, A_val=A_val, l=l, u=u,
                   warm_x=np.ascontiguousarray(warm_x, dtype=np.float64))
