"""Device-resident control step for BASELINE.json configs[2]: the same loop as `receding.IntentSweep.step`, with every
array of a step on the GPU and `makePlanWithPred` (mpcPlanner.cpp:571-661) chained from the engine's own calls —
enumeration (`mpcqp_intent_candidates_device`), replication of the scenario-level inputs (`mpcqp_gather_rows_device`), the
two solves (`mpcqp_solve_mpc_batch_device`), scoring (`mpcqp_score_candidates_device`) and choice
(`mpcqp_select_candidates_device`).  torch only holds the buffers and evaluates the closed-form obstacle predictions and
the reference window (elementwise arithmetic); no QP data crosses PCIe between steps.

The scenario state starts from a host `IntentSweep` (same seeds, same obstacles), so the two drivers can be compared
step by step (tests/test_receding.py)."""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E
from .receding import IntentSweep, FORWARD, LEFT, RIGHT, STOP


class DeviceIntentSweep:
    def __init__(self, eng: E.Engine, host: IntentSweep, device: int = 0):
        self.eng, self.p, self.S, self.D = eng, host.p, host.S, host.D
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

    def predictions(self):
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
                ptrs = {k: b[k].data_ptr() for k in ("scen_a", "scen_b", "obs_c_a", "obs_semi_a", "obs_c_b", "obs_semi_b", "weight", "cand")}
                ptrs.update(pred_pos=pp.data_ptr(), pred_size=ps.data_ptr(), prob=self.prob.data_ptr(), pos=self.pos.data_ptr(), prev_plan=self.plan.data_ptr())
                eng.intent_candidates_ptr(p, S, D, pp.shape[3], ptrs)
                for src, name, w in ((x0, "x0", 6), (xref, "xref", 3 * (N + 1)), (lin, "lin", 3 * N), (self.plan, "warm", n)):
                    eng.gather_rows_ptr(4 * S, w, b["scen_a"].data_ptr(), src.data_ptr(), b[name][: 4 * S].data_ptr())
                    eng.gather_rows_ptr(2 * S, w, b["scen_b"].data_ptr(), src.data_ptr(), b[name][4 * S:].data_ptr())
                self._solve(4 * S, D, 0, (b["obs_c_a"], b["obs_semi_a"], b["obs_yaw_a"]))
                self._solve(2 * S, D + 1, 4 * S, (b["obs_c_b"], b["obs_semi_b"], b["obs_yaw_b"]))
                for B_, R_, off, oc, osz in ((4 * S, D, 0, b["obs_c_a"], b["obs_semi_a"]), (2 * S, D + 1, 4 * S, b["obs_c_b"], b["obs_semi_b"])):
                    eng.score_candidates_ptr(p, B_, R_, R_, {"x": b["x"][off:].data_ptr(), "prev_plan": b["warm"][off:].data_ptr(), "xref": b["xref"][off:].data_ptr(),
                                                            "obs_c": oc.data_ptr(), "obs_semi": osz.data_ptr(), "score": b["score"][off:].data_ptr()})
                eng.select_candidates_ptr(S, 6, n, {"cand": b["cand"].data_ptr(), "weight": b["weight"].data_ptr(), "score": b["score"].data_ptr(),
                                                    "x_all": b["x"].data_ptr(), "best": b["best"].data_ptr(), "weighted": b["weighted"].data_ptr(),
                                                    "plan": b["plan"].data_ptr()})
                eng.sync()
                self.plan = b["plan"].clone(); best = b["best"]
            st = self.plan[:, : 8 * (N + 1)].reshape(S, N + 1, 8)
            self.pos = st[:, 1, 0:3].contiguous(); self.vel = st[:, 1, 3:6].contiguous()   # perfect tracking (mpc_node.cpp:223-224)
            self.step_idx += 1
        self.stream.synchronize()
        return best
