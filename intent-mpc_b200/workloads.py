"""Synthetic mpcPlanner workloads (BASELINE.json `configs`).

Each generator returns an `MpcBatch`: the per-control-step inputs `mpcPlanner::solveTraj`
(mpcPlanner.cpp:375-541) consumes, already flattened per stage the way `updateObstacleParam`
(mpcPlanner.cpp:1148-1197) lays them out — for B independent instances.  Everything is numpy on the
host; the engine uploads these arrays and assembles/solves on the device.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class MpcParams:
    """intent_mpc_demo defaults (autonomous_flight/cfg/mpc_navigation/planner_param.yaml:25-39,
    flight_base.yaml:8-9); field meaning follows mpcPlanner::initParam (mpcPlanner.cpp:19-173)."""
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
    def N(self) -> int:
        return self.horizon - 1

    @property
    def n(self) -> int:
        return 8 * self.horizon + 5 * self.N

    def m(self, num_obs: int) -> int:
        return 16 * self.horizon + 5 * self.N + num_obs * self.N


@dataclasses.dataclass
class MpcBatch:
    params: MpcParams
    x0: np.ndarray        # [B,6]   current position, velocity (updateCurrStates)
    xref: np.ndarray      # [B,N+1,3] reference positions (getXRef)
    obs_c: np.ndarray     # [B,N,numObs,3] obstacle centre per stage
    obs_semi: np.ndarray  # [B,N,numObs,3] semi-axes = size/2 + safety distance
    obs_yaw: np.ndarray   # [B,N,numObs]
    obs_dyn: np.ndarray   # [N,numObs] int32: 1 -> slack input 3 (dynamic), 0 -> slack input 4 (static)
    lin_pt: np.ndarray    # [B,N,3] linearisation point (previous plan, unshifted, or currPos)
    warm_x: np.ndarray    # [B,n] primal warm start (previous plan or zeros); dual warm start is always 0
    nobs: np.ndarray | None = None   # [B] int32 obstacle rows per stage of each instance (<= numObs, the array stride), or None
    limits: np.ndarray | None = None  # [B,2] per-instance (max_vel, max_acc) overriding params (one launch for a whole sweep), or None

    @property
    def B(self) -> int:
        return self.x0.shape[0]

    @property
    def num_obs(self) -> int:
        return self.obs_c.shape[2]

    def slice(self, lo, hi) -> "MpcBatch":
        return MpcBatch(self.params, self.x0[lo:hi], self.xref[lo:hi], self.obs_c[lo:hi], self.obs_semi[lo:hi],
                        self.obs_yaw[lo:hi], self.obs_dyn if self.obs_dyn.ndim == 2 else self.obs_dyn[lo:hi],
                        self.lin_pt[lo:hi], self.warm_x[lo:hi], None if self.nobs is None else self.nobs[lo:hi],
                        None if self.limits is None else self.limits[lo:hi])


GOAL = np.array([105.0, 0.0, 2.0])   # end of ref_trajectory_dynus_benchmark.txt line (mpcNavigation.cpp:201-216)


def _const_vel_plan(p: MpcParams, x0):
    """Constant-velocity rollout from x0 used as 'previous plan' (states 8, controls 5) -> warm_x, lin_pt."""
    B = x0.shape[0]
    N = p.N
    t = (np.arange(N + 1) * p.ts)[None, :, None]
    pos = x0[:, None, 0:3] + x0[:, None, 3:6] * t
    states = np.zeros((B, N + 1, 8))
    states[:, :, 0:3] = pos
    states[:, :, 3:6] = x0[:, None, 3:6]
    warm = np.zeros((B, p.n))
    warm[:, : 8 * (N + 1)] = states.reshape(B, -1)
    return warm, pos[:, :N, :].copy()


def static_batch(B: int, num_obs: int = 4, params: MpcParams | None = None, seed0: int = 0,
                 warm: bool = True) -> MpcBatch:
    """BASELINE.json configs[1]: randomised default-shape QPs, static-obstacle half-spaces only
    (SURVEY.md §8d "Config 2").  One numpy Generator per instance seed so that any sub-range of the
    batch can be regenerated independently (multi-GPU shards)."""
    p = params or MpcParams()
    N = p.N
    x0 = np.zeros((B, 6)); xref = np.zeros((B, N + 1, 3))
    obs_c = np.zeros((B, N, num_obs, 3)); obs_semi = np.zeros((B, N, num_obs, 3)); obs_yaw = np.zeros((B, N, num_obs))
    for b in range(B):
        r = np.random.default_rng(seed0 + b)
        pos = np.array([r.uniform(0, 80), r.uniform(-3, 3), r.uniform(1, 4)])
        vel = r.uniform(-3, 3, size=3)
        speed = r.uniform(1, 5)
        d = GOAL - pos
        d = d / np.linalg.norm(d)
        xref[b] = pos[None, :] + d[None, :] * speed * (np.arange(N + 1) * p.ts)[:, None]
        x0[b, 0:3] = pos; x0[b, 3:6] = vel
        for j in range(num_obs):
            pillar = r.uniform() < 0.35                       # dynus_obstacles_node.cpp:81-84,102-114
            size = np.array([0.4, 0.4, 4.0]) if pillar else np.array([0.4, 4.0, 0.4])
            c = np.array([r.uniform(pos[0] + 2, pos[0] + 30), r.uniform(-7, 7), r.uniform(0, 7)])
            yw = r.uniform(-np.pi / 2, np.pi / 2)
            obs_c[b, :, j, :] = c
            obs_semi[b, :, j, :] = size / 2 + p.static_safety_dist
            obs_yaw[b, :, j] = yw
    obs_dyn = np.zeros((N, num_obs), dtype=np.int32)
    warm_x, lin_pt = _const_vel_plan(p, x0)
    if not warm:
        warm_x = np.zeros_like(warm_x)
    return MpcBatch(p, x0, xref, obs_c, obs_semi, obs_yaw, obs_dyn, lin_pt, warm_x)


def snapshot(params: MpcParams | None = None) -> MpcBatch:
    """BASELINE.json configs[0]: one intent_mpc_demo-style control step (SURVEY.md §8d "Config 1"):
    UAV on the benchmark reference line at 3 m/s, 1 static box ahead-left (yaw 0.3) + 3 dynamic boxes
    with straight-line predictions; dynamic obstacles come first (updateObstacleParam ordering) and the
    isDyamic quirk (mpcPlanner.cpp:1194) flags the first min(S,D)=1 dynamic obstacle as static."""
    p = params or MpcParams()
    N = p.N
    x0 = np.array([[20.0, 0.0, 2.0, 3.0, 0.0, 0.0]])
    ref_pts = np.stack([np.arange(43) * 2.5, np.zeros(43), np.full(43, 2.0)], axis=1)
    i0 = int(np.argmin(np.linalg.norm(ref_pts - x0[0, 0:3], axis=1)))
    idx = np.minimum(np.arange(i0, i0 + p.horizon), len(ref_pts) - 1)
    xref = ref_pts[idx][None]
    robot = np.array([0.5, 0.5, 0.3])
    dyn0 = np.array([[28.0, 1.5, 2.0], [34.0, -2.0, 2.0], [26.0, -3.0, 2.0]])
    dynv = np.array([[-1.0, 0.3, 0.0], [-0.8, 0.0, 0.0], [0.0, 1.0, 0.0]])
    nd, ns = 3, 1
    num_obs = nd + ns
    t = (np.arange(N) * p.ts)
    obs_c = np.zeros((1, N, num_obs, 3)); obs_semi = np.zeros((1, N, num_obs, 3)); obs_yaw = np.zeros((1, N, num_obs))
    for i in range(nd):
        obs_c[0, :, i, :] = dyn0[i][None, :] + dynv[i][None, :] * t[:, None]
        obs_semi[0, :, i, :] = (np.array([0.8, 0.8, 0.8]) + robot) / 2 + p.dynamic_safety_dist
    obs_c[0, :, nd, :] = np.array([30.0, 2.0, 2.0])
    obs_semi[0, :, nd, :] = np.array([0.4, 0.4, 4.0]) / 2 + p.static_safety_dist
    obs_yaw[0, :, nd] = 0.3
    obs_dyn = np.ones((N, num_obs), dtype=np.int32)
    obs_dyn[:, nd:] = 0
    obs_dyn[:, : min(ns, nd)] = 0            # quirk: static loop clears flags [0, numStatic)
    warm_x, lin_pt = _const_vel_plan(p, x0)
    return MpcBatch(p, x0, xref, obs_c, obs_semi, obs_yaw, obs_dyn, lin_pt, warm_x)


def stress_batch(B: int, seed0: int = 0, num_obs: int = 2, horizon: int = 60) -> MpcBatch:
    """BASELINE.json configs[3]: doubled horizon, tight bounds and infeasible instances (SURVEY.md §8d
    "Config 4"): horizon 60, maxVel = maxAcc = 1.5, z in [1.9, 2.1].  Instance b cycles through four kinds:
    0 nominal inside the bounds, 1 start above the z box (oracle: status -2 after 4000 iterations),
    2 initial speed above maxVel (same), 3 an obstacle whose ellipsoid contains the start (slack saturation)."""
    p = MpcParams(horizon=horizon, max_vel=1.5, max_acc=1.5, z_min=1.9, z_max=2.1)
    mb = static_batch(B, num_obs=num_obs, params=p, seed0=seed0 + 100000)
    N = p.N
    for b in range(B):
        kind = b % 4
        r = np.random.default_rng(seed0 + 200000 + b)
        mb.x0[b, 2] = r.uniform(1.95, 2.05)
        mb.x0[b, 3:6] = r.uniform(-1.0, 1.0, size=3)
        mb.x0[b, 5] *= 0.05
        if kind == 1:
            mb.x0[b, 2] = 2.6
        elif kind == 2:
            mb.x0[b, 3] = 2.5
        elif kind == 3:
            mb.obs_c[b, :, 0, :] = mb.x0[b, 0:3] + np.array([0.3, 0.1, 0.0])
        d = GOAL - mb.x0[b, 0:3]
        d = d / np.linalg.norm(d)
        mb.xref[b] = mb.x0[b, None, 0:3] + d[None, :] * 1.2 * (np.arange(N + 1) * p.ts)[:, None]
    mb.warm_x, mb.lin_pt = _const_vel_plan(p, mb.x0)
    return mb


# ---- BASELINE.json configs[4]: Monte-Carlo sweep -------------------------------------------------------------------
SWEEP_OBSTACLES = (50, 100, 200)                       # docker/README.md:1004-1014 (run_mpc_benchmark grid)
SWEEP_LIMITS = ((1.5, 1.5), (3.0, 3.0), (5.0, 20.0))   # (max_vel, max_acc)
SWEEP_DYNAMIC_RATIO = 0.65
SWEEP_FIELD = ((5.0, 105.0), (-15.0, 15.0), (0.0, 7.0))  # dynus_obstacles_node.cpp:61-66
SWEEP_RADIUS = 30.0                                    # fake_detector_param.yaml:2
SWEEP_CAP = 32                                         # rows per stage kept (nearest first); reported by sweep_groups
SWEEP_CLEARANCE = 3.0                                  # obstacles closer than this to the start are dropped (the UAV is flying)
SWEEP_CHUNK = 4096
ROBOT_SIZE = np.array([0.5, 0.5, 0.3])                 # mapping_param.yaml:11


def _sweep_chunk(chunk: int, seed0: int):
    """Instances [chunk*SWEEP_CHUNK, (chunk+1)*SWEEP_CHUNK) of the sweep, vectorised; one Generator per chunk so that a
    multi-GPU shard can regenerate exactly its own index range."""
    C_ = SWEEP_CHUNK
    r = np.random.default_rng([seed0, chunk])
    idx = chunk * C_ + np.arange(C_)
    scen = idx % 9
    nobs = np.array(SWEEP_OBSTACLES)[scen % 3]
    lim = scen // 3
    vmax = np.array([l[0] for l in SWEEP_LIMITS])[lim]
    pos = np.stack([r.uniform(5, 95, C_), r.uniform(-3, 3, C_), r.uniform(1, 4, C_)], axis=1)
    vel = r.uniform(-0.6, 0.6, (C_, 3)) * vmax[:, None]
    speed = r.uniform(0.3, 1.0, C_) * vmax
    M = max(SWEEP_OBSTACLES)
    f = SWEEP_FIELD
    oc = np.stack([r.uniform(f[0][0], f[0][1], (C_, M)), r.uniform(f[1][0], f[1][1], (C_, M)), r.uniform(f[2][0], f[2][1], (C_, M))], axis=2)
    dyn = r.uniform(size=(C_, M)) < SWEEP_DYNAMIC_RATIO
    ang = r.uniform(-np.pi, np.pi, (C_, M)); spd = r.uniform(0.5, 2.0, (C_, M))
    ov = np.stack([np.cos(ang) * spd, np.sin(ang) * spd, np.zeros((C_, M))], axis=2) * dyn[:, :, None]
    pillar = r.uniform(size=(C_, M)) < 0.35
    yaw = r.uniform(-np.pi / 2, np.pi / 2, (C_, M)) * (~dyn)        # updateObstacleParam: yaw 0 for dynamic obstacles
    dist = np.linalg.norm(oc - pos[:, None, :], axis=2)
    ok = (np.arange(M)[None, :] < nobs[:, None]) & (dist <= SWEEP_RADIUS) & (dist >= SWEEP_CLEARANCE)
    return dict(idx=idx, lim=lim, pos=pos, vel=vel, speed=speed, oc=oc, ov=ov, dyn=dyn, pillar=pillar, yaw=yaw,
                dist=np.where(ok, dist, np.inf), in_range=ok.sum(axis=1))


def sweep_groups(lo: int, hi: int, seed0: int = 0, params: MpcParams | None = None):
    """BASELINE.json configs[4] (SURVEY.md §8d "Config 5"): instances [lo, hi) of the Monte-Carlo sweep over the
    run_mpc_benchmark scenario grid — {50,100,200} obstacles in the 100 x 30 x 7 m field, 65 % of them moving, velocity /
    acceleration limits (1.5,1.5), (3,3), (5,20) — one control step each.  Obstacles within 30 m of the UAV become
    constraint rows, nearest first, capped at SWEEP_CAP per stage.  Every instance therefore has its own obstacle
    count and its own mix of dynamic / static rows (updateObstacleParam order: dynamic first; the isDyamic quirk,
    mpcPlanner.cpp:1194, flags the first min(S, D) dynamic ones static).  Moving obstacles are predicted at constant
    velocity over the horizon (the FORWARD intent of dynamicPredictor.cpp:351-501, closed form).

    Returns (groups, meta): groups = list of (index array, MpcBatch) with one batch per (limits, obstacle count) pair —
    the QP dimensions are batch-uniform — whose obs_dyn is [B,N,R] (per instance); meta counts capped instances."""
    base = params or MpcParams()
    N = base.N
    t = np.arange(N) * base.ts
    parts = []
    for chunk in range(lo // SWEEP_CHUNK, (hi - 1) // SWEEP_CHUNK + 1):
        c = _sweep_chunk(chunk, seed0)
        sel = (c["idx"] >= lo) & (c["idx"] < hi)
        parts.append({k: v[sel] for k, v in c.items()})
    c = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    Bt = len(c["idx"])
    order = np.argsort(c["dist"], axis=1, kind="stable")[:, :SWEEP_CAP]
    take = lambda a: np.take_along_axis(a, order if a.ndim == 2 else order[:, :, None], axis=1)
    dist = take(c["dist"]); valid = np.isfinite(dist)
    R_i = valid.sum(axis=1)
    dyn = take(c["dyn"]) & valid
    # dynamic rows first, then static, each nearest first (stable sort on the key "not dynamic", invalid last)
    key = np.where(valid, np.where(dyn, 0, 1), 2)
    perm = np.argsort(key, axis=1, kind="stable")
    take2 = lambda a: np.take_along_axis(a, perm if a.ndim == 2 else perm[:, :, None], axis=1)
    oc = take2(take(c["oc"])); ov = take2(take(c["ov"])); dyn = take2(dyn); pillar = take2(take(c["pillar"])); yaw = take2(take(c["yaw"]))
    D_i = dyn.sum(axis=1); S_i = R_i - D_i
    groups = []
    for li, (vm, am) in enumerate(SWEEP_LIMITS):
        for R in range(SWEEP_CAP + 1):
            g = np.nonzero((c["lim"] == li) & (R_i == R))[0]
            if len(g) == 0:
                continue
            p = dataclasses.replace(base, max_vel=vm, max_acc=am)
            B = len(g)
            x0 = np.concatenate([c["pos"][g], c["vel"][g]], axis=1)
            d = GOAL[None, :] - c["pos"][g]
            d = d / np.linalg.norm(d, axis=1, keepdims=True)
            xref = c["pos"][g][:, None, :] + d[:, None, :] * (c["speed"][g][:, None, None] * (np.arange(N + 1) * p.ts)[None, :, None])
            o_c = oc[g][:, None, :R, :] + ov[g][:, None, :R, :] * t[None, :, None, None]
            dy = dyn[g][:, :R]
            static_size = np.where(pillar[g][:, :R, None], np.array([0.4, 0.4, 4.0])[None, None, :], np.array([0.4, 4.0, 0.4])[None, None, :])
            semi = np.where(dy[:, :, None], (np.array([0.8, 0.8, 0.8]) + ROBOT_SIZE)[None, None, :] / 2 + p.dynamic_safety_dist,
                            static_size / 2 + p.static_safety_dist)
            flags = dy.astype(np.int32)
            quirk = np.arange(R)[None, :] < np.minimum(S_i[g], D_i[g])[:, None]     # isDyamic[j][i] = 0 for i < numStatic
            flags[quirk] = 0
            warm_x, lin_pt = _const_vel_plan(p, x0)
            mb = MpcBatch(p, x0, xref, np.ascontiguousarray(o_c), np.ascontiguousarray(np.broadcast_to(semi[:, None], (B, N, R, 3))),
                          np.ascontiguousarray(np.broadcast_to(yaw[g][:, None, :R], (B, N, R))),
                          np.ascontiguousarray(np.broadcast_to(flags[:, None, :], (B, N, R))), lin_pt, warm_x)
            groups.append((c["idx"][g], mb))
    meta = dict(instances=Bt, capped=int((c["in_range"] > SWEEP_CAP).sum()), cap=SWEEP_CAP,
                obstacle_rows_hist=np.bincount(R_i, minlength=SWEEP_CAP + 1))
    return groups, meta


def sweep_batches(lo: int, hi: int, seed0: int = 0, params: MpcParams | None = None, one_launch: bool = False):
    """The same instances as sweep_groups(lo, hi), packed for the engine's per-instance obstacle counts: ONE batch per
    (max_vel, max_acc) pair, obstacle arrays padded to SWEEP_CAP rows per stage, `nobs` = rows each instance really has.
    one_launch=True: a single batch for the whole range, with per-instance limits (`MpcBatch.limits`).
    Returns (batches, meta): batches = list of (index array, MpcBatch)."""
    groups, meta = sweep_groups(lo, hi, seed0, params)
    out = []
    for li, (vm, am) in enumerate(SWEEP_LIMITS):
        gs = [(idx, mb) for idx, mb in groups if mb.params.max_vel == vm and mb.params.max_acc == am]
        if not gs:
            continue
        p = gs[0][1].params
        N, Rm = p.N, SWEEP_CAP
        B = sum(mb.B for _, mb in gs)
        idx = np.concatenate([i for i, _ in gs])
        cat = lambda name: np.concatenate([getattr(mb, name) for _, mb in gs])
        obs_c = np.zeros((B, N, Rm, 3)); obs_semi = np.ones((B, N, Rm, 3)); obs_yaw = np.zeros((B, N, Rm))
        obs_dyn = np.zeros((B, N, Rm), dtype=np.int32); nobs = np.zeros(B, dtype=np.int32)
        at = 0
        for _, mb in gs:
            R = mb.num_obs
            obs_c[at:at + mb.B, :, :R] = mb.obs_c; obs_semi[at:at + mb.B, :, :R] = mb.obs_semi
            obs_yaw[at:at + mb.B, :, :R] = mb.obs_yaw; obs_dyn[at:at + mb.B, :, :R] = mb.obs_dyn
            nobs[at:at + mb.B] = R
            at += mb.B
        order = np.argsort(idx, kind="stable")
        mbp = MpcBatch(p, cat("x0")[order], cat("xref")[order], obs_c[order], obs_semi[order], obs_yaw[order], obs_dyn[order],
                       cat("lin_pt")[order], cat("warm_x")[order], nobs[order])
        out.append((idx[order], mbp))
    if one_launch and out:
        # per-instance limits: the whole range as a single batch (index order), params of the first group for the rest
        idx = np.concatenate([i for i, _ in out]); order = np.argsort(idx, kind="stable")
        cat = lambda name: np.concatenate([getattr(mb, name) for _, mb in out])[order]
        lim = np.concatenate([np.tile([[mb.params.max_vel, mb.params.max_acc]], (mb.B, 1)) for _, mb in out])[order]
        mb0 = out[0][1]
        one = MpcBatch(mb0.params, cat("x0"), cat("xref"), cat("obs_c"), cat("obs_semi"), cat("obs_yaw"), cat("obs_dyn"), cat("lin_pt"),
                       cat("warm_x"), cat("nobs"), lim)
        return [(idx[order], one)], meta
    return out, meta
