"""Device-resident control step for BASELINE.json configs[2]: S scenarios advance together, with every
array of a step on the GPU and `makePlanWithPred` (mpcPlanner.cpp:571-661) chained from the engine's own calls —
enumeration (`mpcqp_intent_candidates_device`), replication of the scenario-level inputs (`mpcqp_gather_rows_device`), the
two solves (`mpcqp_solve_mpc_batch_device`), scoring (`mpcqp_score_candidates_device`) and choice
(`mpcqp_select_candidates_device`).  torch only holds the buffers and evaluates the closed-form obstacle predictions and
the reference window (elementwise arithmetic); no QP data crosses PCIe between steps.  With `host_predictions=True` the
predictions are evaluated on the host (numpy, `IntentSweep.predictions`) and uploaded every step and the chosen plan is
downloaded — the end-to-end form with host buffers that bench.py times for configs[2].

The scenario state starts from a host `IntentSweep` (scenario generator: seeds, obstacles).  The tests compare every piece of a
step with the literal restatement of the reference planner in oracle/mpc_planner.py (tests/test_planner_parity.py)."""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E
from .receding import IntentSweep, FORWARD, LEFT, RIGHT, STOP


class DeviceIntentSweep:
    def __init__(self, eng: E.Engine, host: IntentSweep, device: int = 0, host_predictions: bool = False, predictor: str = "closed_form",
                 num_hist: int = 10):
        """predictor: "closed_form" = the scenario generator's stand-in predictions and fixed intent probabilities;
        "sampled" = the reference predictor itself on the device (mpcqp_predict_device, dynamicPredictor.cpp:197-541) from
        the obstacles' histories — then only the histories come from the host when host_predictions is set."""
        assert predictor in ("closed_form", "sampled")
        self.eng, self.p, self.S, self.D = eng, host.p, host.S, host.D
        self.host, self.host_predictions, self.predictor, self.num_hist = host, host_predictions, predictor, num_hist
        self._staged = None; self._pinned = None
        self.pparams = E.default_predictor_params()
        self.h2d_bytes = 0; self.d2h_bytes = 0
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(eng.stream, device=self.dev)
        self.st = E.default_settings()
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.pos, self.vel, self.speed = d(host.pos), d(host.vel), d(host.speed)
        self.scale, self.slow, self.off, self.centre = d(host.scale), d(host.slow), d(host.off), d(host.centre)
        self.size, self.prob = d(host.size), d(host.prob)
        self.step_idx = host.step_idx
        self.plan = None                                   # [S, n] accepted plan (states, then controls)
        S, D, p = self.S, self.D, self.p
        N, n = p.N, p.n
        f64 = dict(dtype=torch.float64, device=self.dev); i32 = dict(dtype=torch.int32, device=self.dev)
        self.buf = dict(
            scen_a=torch.empty(4 * S, **i32), scen_b=torch.empty(2 * S, **i32), weight=torch.empty((S, 6), **f64), cand=torch.empty((S, 6), **i32),
            obs_c_a=torch.empty((4 * S, N, D, 3), **f64), obs_semi_a=torch.empty((4 * S, N, D, 3), **f64), obs_yaw_a=torch.zeros((4 * S, N, D), **f64),
            obs_c_b=torch.empty((2 * S, N, D + 1, 3), **f64), obs_semi_b=torch.empty((2 * S, N, D + 1, 3), **f64), obs_yaw_b=torch.zeros((2 * S, N, D + 1), **f64),
            # stage N of every candidate's obstacle rows: only the safety score reads it (mpcPlanner.cpp:818-826)
            obs_c_last_a=torch.empty((4 * S, D, 3), **f64), obs_semi_last_a=torch.empty((4 * S, D, 3), **f64),
            obs_c_last_b=torch.empty((2 * S, D + 1, 3), **f64), obs_semi_last_b=torch.empty((2 * S, D + 1, 3), **f64),
            x0=torch.empty((6 * S, 6), **f64), xref=torch.empty((6 * S, N + 1, 3), **f64), lin=torch.empty((6 * S, N, 3), **f64),
            warm=torch.empty((6 * S, n), **f64), x=torch.empty((6 * S, n), **f64), status=torch.empty(6 * S, **i32), iter=torch.empty(6 * S, **i32),
            rho_updates=torch.empty(6 * S, **i32), obj=torch.empty(6 * S, **f64), pri_res=torch.empty(6 * S, **f64), dua_res=torch.empty(6 * S, **f64),
            score=torch.empty((6 * S, 3), **f64), best=torch.empty(S, **i32), weighted=torch.empty((S, 6), **f64), plan=torch.empty((S, n), **f64))
        self.kernel_ms = 0.0

    # ---- closed forms of receding.IntentSweep (obstacle_state, predictions, reference), elementwise on the device
    def _trefoil(self, t):
        u = t / self.slow + self.off
        x = self.scale * (torch.sin(u) + 2 * torch.sin(2 * u)) / 3.0
        y = self.scale * (torch.cos(u) - 2 * torch.cos(2 * u)) / 3.0
        z = self.scale * (-torch.sin(3 * u)) / 3.0 * 0.3
        return self.centre + torch.stack([x, y, z], dim=-1)

    def _history(self):
        """Obstacle histories [S, D, H, 3] (newest first; velocities (Vx, Vy, 0)) on the device."""
        H = self.num_hist
        if self.host_predictions:
            ph, vh = self.host.history(self.step_idx, H=H)
            self.h2d_bytes += ph.nbytes + vh.nbytes
            return torch.from_numpy(ph).to(self.dev), torch.from_numpy(vh).to(self.dev)
        t = self.step_idx * self.p.ts - 0.1 * torch.arange(H, dtype=torch.float64, device=self.dev)
        pos = torch.stack([self._trefoil(tt) for tt in t], dim=2)
        vel = (torch.stack([self._trefoil(tt + 1e-3) for tt in t], dim=2) - pos) / 1e-3
        vel[..., 2] = 0.0
        return pos.contiguous(), vel.contiguous()

    def stage_host_predictions(self):
        """Host form: the predictor's output of the coming step, in pinned host memory (what a host-side predictor hands over).
        The synthetic generator (numpy) that stands in for it is not part of a control step: callers that time steps call this
        before they start the clock; step() then only uploads."""
        pp, ps = self.host.predictions(self.step_idx)
        if self._pinned is None:
            self._pinned = (torch.empty(pp.shape, dtype=torch.float64).pin_memory(), torch.empty(ps.shape, dtype=torch.float64).pin_memory())
        self._pinned[0].copy_(torch.from_numpy(pp)); self._pinned[1].copy_(torch.from_numpy(ps))
        self._staged = (self.step_idx, self._pinned[0], self._pinned[1])

    def predictions(self):
        if self.predictor == "sampled":                    # dynamicPredictor on the device
            ph, vh = self._history()
            T = self.pparams.prediction_size + 1
            pp = torch.empty((self.S, self.D, 4, T, 3), dtype=torch.float64, device=self.dev); ps = torch.empty_like(pp)
            self.eng.predict_ptr(self.pparams, self.S * self.D, self.num_hist, {"pos_hist": ph.data_ptr(), "vel_hist": vh.data_ptr(), "size": self.size.data_ptr(),
                                                                                "pred_pos": pp.data_ptr(), "pred_size": ps.data_ptr(), "intent_prob": self.prob.data_ptr()})
            self._keep = (ph, vh)                          # alive until the stream has run the kernel
            return pp, ps
        if self.host_predictions:                          # produced on the host, uploaded (the e2e form)
            if self._staged is None or self._staged[0] != self.step_idx:
                self.stage_host_predictions()
            _, hp, hs = self._staged
            self.h2d_bytes += hp.numel() * 8 + hs.numel() * 8
            return hp.to(self.dev, non_blocking=False), hs.to(self.dev, non_blocking=False)     # (pinned source: no staging copy)
        T = 31
        t0 = self.step_idx * self.p.ts
        pos = self._trefoil(t0); vel = (self._trefoil(t0 + 1e-3) - pos) / 1e-3
        t = torch.arange(T, dtype=torch.float64, device=self.dev) * self.p.ts
        pp = torch.empty((self.S, self.D, 4, T, 3), dtype=torch.float64, device=self.dev)
        for it, om in ((FORWARD, 0.0), (LEFT, 0.6), (RIGHT, -0.6)):
            if om == 0.0:
                disp = vel[:, :, None, :] * t[None, None, :, None]
            else:
                a = om * t
                sx = torch.sin(a) / om; cx = (1 - torch.cos(a)) / om
                disp = torch.stack([vel[:, :, None, 0] * sx - vel[:, :, None, 1] * cx, vel[:, :, None, 0] * cx + vel[:, :, None, 1] * sx,
                                    vel[:, :, None, 2] * t], dim=-1)
            pp[:, :, it] = pos[:, :, None, :] + disp
        pp[:, :, STOP] = pos[:, :, None, :]
        ps = (self.size[:, :, None, None, :] + 0.02 * torch.arange(T, dtype=torch.float64, device=self.dev)[None, None, None, :, None]).expand(-1, -1, 4, -1, -1).contiguous()
        return pp, ps

    def reference(self):
        k = torch.arange(self.p.horizon, dtype=torch.float64, device=self.dev)
        x = torch.clamp(self.pos[:, None, 0] + self.speed[:, None] * self.p.ts * k[None, :], max=105.0)
        return torch.stack([x, torch.zeros_like(x), torch.full_like(x, 2.0)], dim=-1).contiguous()

    def _solve(self, B, R, off, obs):
        b = self.buf
        sl = lambda name: b[name][off:off + B].data_ptr()
        ptrs = {"x0": sl("x0"), "xref": sl("xref"), "lin_pt": sl("lin"), "warm_x": sl("warm"), "x": sl("x"), "status": sl("status"), "iter": sl("iter"),
                "rho_updates": sl("rho_updates"), "obj": sl("obj"), "pri_res": sl("pri_res"), "dua_res": sl("dua_res")}
        if R:
            ptrs.update(obs_c=obs[0].data_ptr(), obs_semi=obs[1].data_ptr(), obs_yaw=obs[2].data_ptr())
        self.eng.solve_mpc_batch_ptr(self.p, self.st, B, R, ptrs, np.ones((self.p.N, max(R, 1)), dtype=np.int32)[:, :R], device=True)
        self.eng.sync(); self.kernel_ms += self.eng.last_kernel_ms

    def step(self):
        """One control step for all scenarios, on the device.  Returns best [S] (int32 tensor) or None on the first step."""
        eng, p, S, D, b = self.eng, self.p, self.S, self.D, self.buf
        N, n = p.N, p.n
        with torch.cuda.stream(self.stream):
            x0 = torch.cat([self.pos, self.vel], dim=1).contiguous()
            xref = self.reference()
            if self.plan is None:                          # first step: one obstacle-free QP per scenario (mpcPlanner.cpp:598-602)
                b["x0"][:S] = x0; b["xref"][:S] = xref; b["lin"][:S] = self.pos[:, None, :].expand(-1, N, -1); b["warm"][:S] = 0.0
                self._solve(S, 0, 0, None)
                self.plan = b["x"][:S].clone(); best = None
            else:
                pp, ps = self.predictions()
                lin = self.plan[:, : 8 * (N + 1)].reshape(S, N + 1, 8)[:, :N, 0:3].contiguous()
                ptrs = {k: b[k].data_ptr() for k in ("scen_a", "scen_b", "obs_c_a", "obs_semi_a", "obs_c_b", "obs_semi_b", "obs_c_last_a",
                                                     "obs_semi_last_a", "obs_c_last_b", "obs_semi_last_b", "weight", "cand")}
                ptrs.update(pred_pos=pp.data_ptr(), pred_size=ps.data_ptr(), prob=self.prob.data_ptr(), pos=self.pos.data_ptr(), prev_plan=self.plan.data_ptr())
                eng.intent_candidates_ptr(p, S, D, pp.shape[3], ptrs)
                for src, name, w in ((x0, "x0", 6), (xref, "xref", 3 * (N + 1)), (lin, "lin", 3 * N), (self.plan, "warm", n)):
                    eng.gather_rows_ptr(4 * S, w, b["scen_a"].data_ptr(), src.data_ptr(), b[name][: 4 * S].data_ptr())
                    eng.gather_rows_ptr(2 * S, w, b["scen_b"].data_ptr(), src.data_ptr(), b[name][4 * S:].data_ptr())
                self._solve(4 * S, D, 0, (b["obs_c_a"], b["obs_semi_a"], b["obs_yaw_a"]))
                self._solve(2 * S, D + 1, 4 * S, (b["obs_c_b"], b["obs_semi_b"], b["obs_yaw_b"]))
                for B_, R_, off, sfx in ((4 * S, D, 0, "a"), (2 * S, D + 1, 4 * S, "b")):
                    eng.score_candidates_ptr(p, B_, R_, R_, {"x": b["x"][off:].data_ptr(), "prev_plan": b["warm"][off:].data_ptr(), "xref": b["xref"][off:].data_ptr(),
                                                            "obs_c": b["obs_c_" + sfx].data_ptr(), "obs_semi": b["obs_semi_" + sfx].data_ptr(),
                                                            "obs_c_last": b["obs_c_last_" + sfx].data_ptr(), "obs_semi_last": b["obs_semi_last_" + sfx].data_ptr(),
                                                            "score": b["score"][off:].data_ptr()})
                eng.select_candidates_ptr(S, 6, n, {"cand": b["cand"].data_ptr(), "weight": b["weight"].data_ptr(), "score": b["score"].data_ptr(),
                                                    "x_all": b["x"].data_ptr(), "best": b["best"].data_ptr(), "weighted": b["weighted"].data_ptr(),
                                                    "plan": b["plan"].data_ptr()})
                eng.sync()
                self.plan = b["plan"].clone(); best = b["best"]
            st = self.plan[:, : 8 * (N + 1)].reshape(S, N + 1, 8)
            self.pos = st[:, 1, 0:3].contiguous(); self.vel = st[:, 1, 3:6].contiguous()   # perfect tracking (mpc_node.cpp:223-224)
            self.step_idx += 1
            if self.host_predictions:                      # what a host-side caller reads back every step: the accepted plan
                self.plan_host = self.plan.cpu(); self.d2h_bytes += self.plan.numel() * 8
        self.stream.synchronize()
        self.had_candidates = best is not None
        return best

    def batches_host(self):
        """The solve batches of the LAST step as host `MpcBatch`es with their outputs (for oracle checks): one obstacle-free
        batch after the first step, else [4S rows with D obstacle rows per stage, 2S rows with D + 1]."""
        from .workloads import MpcBatch
        p, S, D, b = self.p, self.S, self.D, self.buf
        N = p.N
        h = lambda t: t.detach().cpu().numpy()
        outs = lambda lo, hi: {k: h(b[k][lo:hi]) for k in ("x", "status", "iter", "rho_updates", "obj", "pri_res", "dua_res")}
        if not getattr(self, "had_candidates", False):
            z = np.zeros((S, N, 0, 3))
            mb = MpcBatch(p, h(b["x0"][:S]), h(b["xref"][:S]), z, z.copy(), np.zeros((S, N, 0)), np.zeros((N, 0), dtype=np.int32), h(b["lin"][:S]), h(b["warm"][:S]))
            return [(mb, outs(0, S))]
        res = []
        for lo, hi, R, sfx in ((0, 4 * S, D, "a"), (4 * S, 6 * S, D + 1, "b")):
            mb = MpcBatch(p, h(b["x0"][lo:hi]), h(b["xref"][lo:hi]), h(b["obs_c_" + sfx]), h(b["obs_semi_" + sfx]), h(b["obs_yaw_" + sfx]),
                          np.ones((N, R), dtype=np.int32), h(b["lin"][lo:hi]), h(b["warm"][lo:hi]))
            res.append((mb, outs(lo, hi)))
        return res
