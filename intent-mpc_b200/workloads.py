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

    @property
    def B(self) -> int:
        return self.x0.shape[0]

    @property
    def num_obs(self) -> int:
        return self.obs_c.shape[2]

    def slice(self, lo, hi) -> "MpcBatch":
        return MpcBatch(self.params, self.x0[lo:hi], self.xref[lo:hi], self.obs_c[lo:hi], self.obs_semi[lo:hi],
                        self.obs_yaw[lo:hi], self.obs_dyn, self.lin_pt[lo:hi], self.warm_x[lo:hi])


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


def stress_batch(B: int, seed0: int = 0, num_obs: int = 2) -> MpcBatch:
    """BASELINE.json configs[3]: doubled horizon, tight bounds and infeasible instances (SURVEY.md §8d
    "Config 4"): horizon 60, maxVel = maxAcc = 1.5, z in [1.9, 2.1].  Instance b cycles through four kinds:
    0 nominal inside the bounds, 1 start above the z box (oracle: status -2 after 4000 iterations),
    2 initial speed above maxVel (same), 3 an obstacle whose ellipsoid contains the start (slack saturation)."""
    p = MpcParams(horizon=60, max_vel=1.5, max_acc=1.5, z_min=1.9, z_max=2.1)
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
