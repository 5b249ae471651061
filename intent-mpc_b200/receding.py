"""Scenario generator for BASELINE.json configs[2]: S independent scenarios, each a UAV following the benchmark reference
line among D dynamic obstacles with four intent-conditioned predictions per obstacle.  This module only produces INPUTS
(numpy, host): obstacle motion, the per-intent predictions, the reference window and the first control step's batch.  The
control step itself — makePlanWithPred (mpcPlanner.cpp:571-661): candidate enumeration, the six solves, scoring, selection,
carry-over of the winner as warm start and linearisation point — is `receding_device.DeviceIntentSweep`, a chain of engine
calls; its CPU restatement for the tests is oracle/mpc_planner.py.

Predictions are a simplified closed form of dynamic_predictor (PRED.cpp:351-501 mean paths): FORWARD = constant
velocity, LEFT / RIGHT = velocity rotated by ±0.6 rad/s·t about z, STOP = standing still; box sizes grow by 2 cm per
prediction step (stand-in for the 2·sqrt(var)·0.674 inflation of PRED.cpp:503-538).  Obstacle paths are the trefoil
knots of dynus_obstacles_node.cpp:13-25 (scale U[2,4], slow-down U[4,6], phase offset U[0,3])."""
from __future__ import annotations

import numpy as np

from .workloads import MpcBatch, MpcParams

FORWARD, LEFT, RIGHT, STOP = 0, 1, 2, 3          # dynamic_predictor/utils.h:15-20
# candidate c -> intents of the closest obstacle (mpcPlanner.cpp:730-741), before sorting by weight
COMBOS = [(STOP,), (LEFT,), (RIGHT,), (FORWARD,), (LEFT, FORWARD), (RIGHT, FORWARD)]
ROBOT = np.array([0.5, 0.5, 0.3])                # mapping_param.yaml:11, added to every box (fakeDetector.cpp:525-553)


def _trefoil(t, scale, slow, off, centre):
    u = t[..., None] / slow[None] + off[None]
    x = scale[None] * (np.sin(u) + 2 * np.sin(2 * u)) / 3.0
    y = scale[None] * (np.cos(u) - 2 * np.cos(2 * u)) / 3.0
    z = scale[None] * (-np.sin(3 * u)) / 3.0 * 0.3
    return centre[None] + np.stack([x, y, z], axis=-1)


class IntentSweep:
    def __init__(self, S: int, D: int = 4, seed0: int = 0, params: MpcParams | None = None):
        self.p = params or MpcParams()
        self.S, self.D = S, D
        p = self.p
        r = np.random.default_rng(seed0)
        self.pos = np.stack([r.uniform(0, 40, S), r.uniform(-1, 1, S), r.uniform(1.5, 2.5, S)], axis=1)
        self.vel = np.stack([r.uniform(1.5, 3.0, S), np.zeros(S), np.zeros(S)], axis=1)
        self.speed = r.uniform(2.0, 3.0, S)
        self.scale = r.uniform(2, 4, (S, D)); self.slow = r.uniform(4, 6, (S, D)); self.off = r.uniform(0, 3, (S, D))
        self.centre = np.stack([self.pos[:, None, 0] + r.uniform(6, 25, (S, D)), r.uniform(-4, 4, (S, D)), np.full((S, D), 2.0)], axis=-1)
        self.size = np.broadcast_to(np.array([0.8, 0.8, 0.8]) + ROBOT, (S, D, 3)).copy()
        self.prob = r.dirichlet(np.ones(4) * 1.5, size=(S, D))          # intentProb[ob](4)
        self.step_idx = 0

    # ---- obstacles and predictions -------------------------------------------------------------
    def obstacle_state(self, step):
        t = np.array([step * self.p.ts])
        pos = _trefoil(t, self.scale.reshape(-1), self.slow.reshape(-1), self.off.reshape(-1), self.centre.reshape(-1, 3))[0]
        pos2 = _trefoil(t + 1e-3, self.scale.reshape(-1), self.slow.reshape(-1), self.off.reshape(-1), self.centre.reshape(-1, 3))[0]
        vel = (pos2 - pos) / 1e-3
        return pos.reshape(self.S, self.D, 3), vel.reshape(self.S, self.D, 3)

    def history(self, step, H: int = 10, dt_hist: float = 0.1):
        """Obstacle histories as the detector hands them to the predictor (fakeDetector::getDynamicObstaclesHist,
        fakeDetector.cpp:525-553): posHist / velHist [S, D, H, 3], index 0 = newest, velocities (Vx, Vy, 0)."""
        t = step * self.p.ts - dt_hist * np.arange(H)
        flat = lambda a: a.reshape(-1)
        pos = _trefoil(t, flat(self.scale), flat(self.slow), flat(self.off), self.centre.reshape(-1, 3))            # [H, S*D, 3]
        pos2 = _trefoil(t + 1e-3, flat(self.scale), flat(self.slow), flat(self.off), self.centre.reshape(-1, 3))
        vel = (pos2 - pos) / 1e-3
        vel[..., 2] = 0.0
        sh = (self.S, self.D, H, 3)
        return np.ascontiguousarray(pos.transpose(1, 0, 2)).reshape(sh), np.ascontiguousarray(vel.transpose(1, 0, 2)).reshape(sh)

    def predictions(self, step):
        """predPos [S, D, 4, NS+1? -> NS+0: 31 steps, 3], predSize likewise."""
        T = 31
        pos, vel = self.obstacle_state(step)
        t = np.arange(T) * self.p.ts
        pp = np.zeros((self.S, self.D, 4, T, 3)); ps = np.zeros_like(pp)
        for it, om in ((FORWARD, 0.0), (LEFT, 0.6), (RIGHT, -0.6)):
            if om == 0.0:
                disp = vel[:, :, None, :] * t[None, None, :, None]
            else:
                a = om * t
                # integral of the rotated velocity:  R(om s) v ds
                sx = np.sin(a) / om; cx = (1 - np.cos(a)) / om
                disp = np.stack([vel[:, :, None, 0] * sx - vel[:, :, None, 1] * cx,
                                 vel[:, :, None, 0] * cx + vel[:, :, None, 1] * sx,
                                 vel[:, :, None, 2] * t], axis=-1)
            pp[:, :, it] = pos[:, :, None, :] + disp
        pp[:, :, STOP] = pos[:, :, None, :]
        ps[:] = self.size[:, :, None, None, :] + 0.02 * np.arange(T)[None, None, None, :, None]
        return pp, ps

    def reference(self):
        """getReferenceTraj on the benchmark line (0,0,2) -> (105,0,2) resampled at the scenario's cruise speed."""
        p = self.p
        k = np.arange(p.horizon)
        x = np.minimum(self.pos[:, None, 0] + self.speed[:, None] * p.ts * k[None, :], 105.0)
        return np.stack([x, np.zeros_like(x), np.full_like(x, 2.0)], axis=-1)

    def first_step_batch(self):
        """First control step: no predictions are used, one obstacle-free QP per scenario (mpcPlanner.cpp:598-602, 645-659)."""
        p, S = self.p, self.S
        N = p.N
        x0 = np.concatenate([self.pos, self.vel], axis=1)
        z = np.zeros((S, N, 0, 3))
        return MpcBatch(p, x0, self.reference(), z, z.copy(), np.zeros((S, N, 0)), np.zeros((N, 0), dtype=np.int32),
                        np.broadcast_to(self.pos[:, None, :], (S, N, 3)).copy(), np.zeros((S, p.n)))
