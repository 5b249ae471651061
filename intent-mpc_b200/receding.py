"""Batched receding-horizon driver for BASELINE.json configs[2]: S independent scenarios, each a UAV following the
benchmark reference line among D dynamic obstacles with four intent-conditioned predictions per obstacle; every control
step solves the six intent candidates of every scenario (mpcPlanner::makePlanWithPred, mpcPlanner.cpp:571-661), scores
them (getTrajectoryScore / evaluateTraj, :771-887), keeps the best as warm start and linearisation point of the next
step (unshifted, :485-509, :1042-1051) and rolls the state forward by perfect tracking (mpc_node.cpp:223-224).

Host logic only (numpy, vectorised over scenarios): obstacle motion, predictions, candidate enumeration, scoring.  The
QPs go to a `solve(batch) -> dict(x, status, iter, ...)` callable — the CUDA engine in production and in the GPU tests,
the oracle in the CPU tests.  Candidates 4 and 5 carry the closest obstacle twice (mpcPlanner.cpp:737-741), so each step
is two batches: 4·S QPs with D obstacle rows per stage and 2·S with D+1.

Predictions are a simplified closed form of dynamic_predictor (PRED.cpp:351-501 mean paths): FORWARD = constant
velocity, LEFT / RIGHT = velocity rotated by ±0.6 rad/s·t about z, STOP = standing still; box sizes grow by 2 cm per
prediction step (stand-in for the 2·sqrt(var)·0.674 inflation of PRED.cpp:503-538).  Obstacle paths are the trefoil knots
of dynus_obstacles_node.cpp:13-25 (scale U[2,4], slow-down U[4,6], phase offset U[0,3])."""
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
        self.states = None            # [S, NS, 8] previous accepted plan
        self.controls = None          # [S, N, 5]
        self.first = True
        self.last = {}

    # ---- obstacles and predictions -------------------------------------------------------------
    def obstacle_state(self, step):
        t = np.array([step * self.p.ts])
        pos = _trefoil(t, self.scale.reshape(-1), self.slow.reshape(-1), self.off.reshape(-1), self.centre.reshape(-1, 3))[0]
        pos2 = _trefoil(t + 1e-3, self.scale.reshape(-1), self.slow.reshape(-1), self.off.reshape(-1), self.centre.reshape(-1, 3))[0]
        vel = (pos2 - pos) / 1e-3
        return pos.reshape(self.S, self.D, 3), vel.reshape(self.S, self.D, 3)

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

    # ---- candidate enumeration (getIntentComb, findClosestObstacle) ---------------------------------
    def closest(self, ob_pos):
        S = self.S
        if self.first or self.states is None:
            d = np.linalg.norm(self.pos[:, None, :] - ob_pos, axis=-1)
            return d.argmin(axis=1)
        s0 = self.states[:, 0, 0:3]; s1 = self.states[:, 1, 0:3]
        ta = np.arctan2(s1[:, 1] - s0[:, 1], s1[:, 0] - s0[:, 0])
        oa = np.arctan2(ob_pos[:, :, 1] - s0[:, None, 1], ob_pos[:, :, 0] - s0[:, None, 0])
        d = np.linalg.norm(s0[:, None, :] - ob_pos, axis=-1)
        # sum_j exp(-j) d (3 - cos) over j < len/3 with the SAME state each term (mpcPlanner.cpp:690-699): a constant factor,
        # so the argmin over obstacles is that of d (3 - cos(.)); the early `break` cannot change the argmin either
        w = d * (3.0 - np.cos(ta[:, None] - oa))
        return w.argmin(axis=1)

    def candidates(self):
        """Returns (batches, meta): batches = [MpcBatch with R = D (4 S QPs), MpcBatch with R = D + 1 (2 S QPs)];
        meta maps each QP back to (scenario, sorted candidate position)."""
        p, S, D = self.p, self.S, self.D
        N, NS = p.N, p.N + 1
        pp, ps = self.predictions(self.step_idx)
        ob_now = pp[:, :, FORWARD, 0, :]
        ob = self.closest(ob_now)                                   # [S]
        pr = self.prob[np.arange(S), ob]                            # [S, 4]
        w = np.stack([pr[:, STOP], pr[:, LEFT], pr[:, RIGHT], pr[:, FORWARD], np.maximum(pr[:, LEFT], pr[:, FORWARD]),
                      np.maximum(pr[:, RIGHT], pr[:, FORWARD])], axis=1)          # original combo order
        # std::sort on (weight, index) pairs ascending, candidates taken from the back (mpcPlanner.cpp:728, 753-756)
        order = np.lexsort((np.broadcast_to(np.arange(6), (S, 6)), w), axis=1)[:, ::-1]    # [S, 6] combo id per sorted position
        maxint = self.prob.argmax(axis=2)                           # [S, D]
        xref = self.reference()
        lin = self.states[:, :N, 0:3] if (not self.first and self.states is not None) else np.broadcast_to(self.pos[:, None, :], (S, N, 3))
        warm = np.zeros((S, p.n))
        if not self.first and self.states is not None:
            warm[:, : 8 * NS] = self.states.reshape(S, -1); warm[:, 8 * NS:] = self.controls.reshape(S, -1)
        x0 = np.concatenate([self.pos, self.vel], axis=1)
        # vectorised over scenarios: candidate (s, sorted position) -> obstacle list [closest with each intent of the combo,
        # then every other obstacle (ascending index) with its most likely intent]
        others = np.array([[j for j in range(D) if j != o] for o in range(D)], dtype=np.int64)[ob]      # [S, D-1]
        c_int = np.full((6, 2), -1, dtype=np.int64)
        for ci, combo in enumerate(COMBOS):
            c_int[ci, : len(combo)] = combo
        ncomb = np.array([len(c) for c in COMBOS])[order]                                            # [S, 6] intents per sorted candidate
        groups = {}
        batches, meta = [], []
        sN = np.arange(N)
        for R, nc in ((D, 1), (D + 1, 2)):
            s_i, pos_i = np.nonzero(ncomb == nc)                                                      # row-major: scenario, then sorted position
            B = len(s_i)
            cid = order[s_i, pos_i]
            oj = np.concatenate([np.repeat(ob[s_i, None], nc, axis=1), others[s_i]], axis=1)           # [B, R] obstacle index of each row
            it = np.concatenate([c_int[cid, :nc], maxint[s_i[:, None], others[s_i]]], axis=1)          # [B, R] intent of each row
            oc = pp[s_i[:, None, None], oj[:, None, :], it[:, None, :], sN[None, :, None], :]          # [B, N, R, 3]
            osz = ps[s_i[:, None, None], oj[:, None, :], it[:, None, :], sN[None, :, None], :] / 2 + p.dynamic_safety_dist
            od = np.ones((N, R), dtype=np.int32)                    # all dynamic, no static obstacles: isDyamic = 1
            batches.append(MpcBatch(p, x0[s_i], xref[s_i], np.ascontiguousarray(oc), np.ascontiguousarray(osz), np.zeros((B, N, R)), od,
                                    np.ascontiguousarray(lin[s_i]), warm[s_i]))
            meta.append(np.stack([s_i, pos_i], axis=1).astype(np.int64))
            groups[R] = (s_i, pos_i)
        self.last = dict(pp=pp, ps=ps, ob=ob, w=w, order=order, xref=xref, groups=groups)
        return batches, meta

    def reference(self):
        """getReferenceTraj on the benchmark line (0,0,2) -> (105,0,2) resampled at the scenario's cruise speed."""
        p = self.p
        k = np.arange(p.horizon)
        x = np.minimum(self.pos[:, None, 0] + self.speed[:, None] * p.ts * k[None, :], 105.0)
        return np.stack([x, np.zeros_like(x), np.full_like(x, 2.0)], axis=-1)

    # ---- scoring / selection (getTrajectoryScore, evaluateTraj) -------------------------------------
    def select(self, batches, meta, outs):
        p, S = self.p, self.S
        NS = p.N + 1
        cand_x = np.zeros((S, 6, p.n)); score = np.zeros((S, 6, 3)); status = np.zeros((S, 6), dtype=np.int64); iters = np.zeros((S, 6), dtype=np.int64)
        for mb, mt, out in zip(batches, meta, outs):
            st = out["x"][:, : 8 * NS].reshape(-1, NS, 8)
            s_i, c_i = mt[:, 0], mt[:, 1]
            cand_x[s_i, c_i] = out["x"]; status[s_i, c_i] = out["status"]; iters[s_i, c_i] = out["iter"]
            pos = st[:, :, 0:3]
            if self.first or self.states is None:
                cons = np.zeros(len(st))
            else:
                cons = np.maximum(np.linalg.norm(self.states[s_i, :10, 0:3] - pos[:, :10], axis=-1).mean(axis=1), 0.1)
            det = np.maximum(np.linalg.norm(mb.xref - pos, axis=-1).mean(axis=1), 0.1)
            # safety (mpcPlanner.cpp:815-852): obstacle trajectories of this candidate, xy distance, tanh weights
            oc = np.concatenate([mb.obs_c, mb.obs_c[:, -1:, :, :]], axis=1)       # stage N reuses the last prediction we hold
            osz = (np.concatenate([mb.obs_semi, mb.obs_semi[:, -1:, :, :]], axis=1) - p.dynamic_safety_dist) * 2
            d = np.linalg.norm(pos[:, :, None, 0:2] - oc[:, :, :, 0:2], axis=-1)
            ms = np.sqrt(osz[..., 0] ** 2 + osz[..., 1] ** 2)
            wgt = 1 - np.tanh(np.arctanh(0.5) / (p.dynamic_safety_dist + ms) * d)
            saf = ((d * wgt).sum(axis=2) / wgt.sum(axis=2)).mean(axis=1)
            score[s_i, c_i] = np.stack([cons, det, saf], axis=1)
        avg = score.mean(axis=1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            rem = np.stack([avg[:, :, 0] / score[:, :, 0], avg[:, :, 1] / score[:, :, 1], score[:, :, 2] / avg[:, :, 2]], axis=-1)
        # weight(intentType[i]) with intentType[i] = sorted position i and `weight` in ORIGINAL combo order (:866-880)
        weighted = self.last["w"] * rem.sum(axis=-1)
        weighted = np.where(np.isnan(weighted), -np.inf, weighted)
        best = weighted.argmax(axis=1)
        return cand_x, status, iters, weighted, best

    def advance(self, cand_x, best):
        p, S = self.p, self.S
        NS, N = p.N + 1, p.N
        x = cand_x[np.arange(S), best]
        self.states = x[:, : 8 * NS].reshape(S, NS, 8).copy(); self.controls = x[:, 8 * NS:].reshape(S, N, 5).copy()
        # perfect tracking: state <- plan at t = ts (getPos(dt), getVel(dt) interpolate to stage 1)
        self.pos = self.states[:, 1, 0:3].copy(); self.vel = self.states[:, 1, 3:6].copy()
        self.first = False
        self.step_idx += 1

    def first_step_batch(self):
        """First control step: no predictions are used, one obstacle-free QP per scenario (mpcPlanner.cpp:598-602, 645-659)."""
        p, S = self.p, self.S
        N = p.N
        x0 = np.concatenate([self.pos, self.vel], axis=1)
        z = np.zeros((S, N, 0, 3))
        return MpcBatch(p, x0, self.reference(), z, z.copy(), np.zeros((S, N, 0)), np.zeros((N, 0), dtype=np.int32),
                        np.broadcast_to(self.pos[:, None, :], (S, N, 3)).copy(), np.zeros((S, p.n)))

    def step(self, solve):
        """One control step for all scenarios.  Returns a dict with the batches solved and their outputs."""
        if self.first:
            mb = self.first_step_batch()
            out = solve(mb)
            NS = self.p.N + 1
            self.states = out["x"][:, : 8 * NS].reshape(self.S, NS, 8).copy(); self.controls = out["x"][:, 8 * NS:].reshape(self.S, self.p.N, 5).copy()
            self.pos = self.states[:, 1, 0:3].copy(); self.vel = self.states[:, 1, 3:6].copy()
            self.first = False; self.step_idx += 1
            return dict(batches=[mb], outs=[out], best=None)
        batches, meta = self.candidates()
        outs = [solve(mb) for mb in batches]
        cand_x, status, iters, weighted, best = self.select(batches, meta, outs)
        self.advance(cand_x, best)
        return dict(batches=batches, outs=outs, meta=meta, best=best, status=status, iters=iters, weighted=weighted)
