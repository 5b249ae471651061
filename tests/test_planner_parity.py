"""Orchestration parity: three implementations of mpcPlanner's control step run
the same scenarios in lockstep and must agree —

  * the ORACLE: oracle/mpc_planner.py, a literal restatement of mpcPlanner.cpp:543-887, 1148-1327 whose QPs are solved by the
    reference's own OSQP binary (oracle/_ref);
  * the C++ MIRROR: intent-mpc_b200/host/MpcPlannerB200.hpp (host logic in C++, QPs on the GPU through
    mpcqp_solve_mpc_batch_host), driven through the test-only C shim tests/cpp/planner_capi.cpp;
  * the DEVICE KERNELS: mpcqp_intent_candidates_device / mpcqp_score_candidates_device / mpcqp_select_candidates_device and
    the device-resident loop intent-mpc_b200/receding_device.py.

Bars: closest obstacle, hypothesis order, candidate rows and the chosen candidate bit-exact; obstacle arrays bit-exact; scores
within 1e-12 relative (the three differ only in the rounding of tanh / cos / atan2 / exp between CUDA's libm and glibc's);
every QP: identical OSQP status and iteration count, x within 1e-5 relative of the reference binary's (BASELINE.json
north_star).  Tie rule for `best`: the maximum of six weighted scores is compared for equality of the INDEX; where the two
largest values differ by less than 1e-9 relative the index may legitimately depend on the last bit of tanh and the case is
counted instead of asserted (expected and observed: none).
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from intent_mpc_b200 import engine, receding
from oracle import bindings as OB
from oracle import mpc_assembly as MA
from oracle import mpc_planner as MP
from tests.helpers import rel_inf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
SHIM = os.path.join(CPP, "libplanner_capi.so")
TOL = 1e-5            # BASELINE.json north_star: primal solution within 1e-5 relative
TOL_SCORE = 1e-12     # scores: same formula, same summation order, libm roundings only
D3 = C.c_double * 3


def _oracle():
    return OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()


def build_shim():
    import __graft_entry__ as G
    G.build()
    src = os.path.join(CPP, "planner_capi.cpp")
    deps = [src, os.path.join(ROOT, "intent-mpc_b200", "host", "MpcPlannerB200.hpp"), os.path.join(ROOT, "include", "mpcqp_b200.h")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(d) > os.path.getmtime(SHIM) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", SHIM, src, "-L" + os.path.join(ROOT, "intent-mpc_b200"),
                        "-lmpcqp_b200", "-Wl,-rpath," + os.path.join(ROOT, "intent-mpc_b200")], check=True)
    lib = C.CDLL(SHIM)
    lib.pl_create.restype = C.c_void_p
    lib.pl_last_error.restype = C.c_char_p
    lib.pl_get_ts.restype = C.c_double; lib.pl_get_horizon.restype = C.c_double
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Mirror:
    """ctypes handle on one trajPlannerB200::mpcPlanner."""

    def __init__(self, lib, p: MA.MpcParams, device=0):
        self.lib, self.p = lib, p
        self.h = C.c_void_p(lib.pl_create(C.c_int(device)))
        assert lib.pl_ready(self.h) == 1, lib.pl_last_error(self.h)
        lib.pl_set_params(self.h, C.byref(engine.params_to_c(p)))
        self.NS, self.N = p.horizon, p.horizon - 1

    def close(self):
        if self.h:
            self.lib.pl_destroy(self.h); self.h = None

    def updatePath(self, path, ts):
        a = np.ascontiguousarray(path, dtype=np.float64); self.lib.pl_update_path(self.h, _dp(a), C.c_int(len(a)), C.c_double(ts))

    def updateCurrStates(self, pos, vel):
        a, b = np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray(vel, dtype=np.float64)
        self.lib.pl_update_curr_states(self.h, _dp(a), _dp(b))

    def updateStaticObstacles(self, obs):
        a = np.ascontiguousarray([list(c) + list(s) + [y] for c, s, y in obs], dtype=np.float64).reshape(-1, 7)
        self.lib.pl_update_static(self.h, _dp(a), C.c_int(len(obs)))

    def updateDynamicObstacles(self, pos, vel, size):
        a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (pos, vel, size))
        self.lib.pl_update_dynamic(self.h, _dp(a), _dp(b), _dp(c), C.c_int(len(a)))

    def updatePredObstacles(self, pp, ps, prob):
        a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (pp, ps, prob))
        self.lib.pl_update_pred(self.h, _dp(a), _dp(b), _dp(c), C.c_int(a.shape[0]), C.c_int(a.shape[2]))

    def makePlan(self):
        return bool(self.lib.pl_make_plan(self.h))

    def makePlanWithPred(self):
        return bool(self.lib.pl_make_plan_with_pred(self.h))

    def plan(self):
        st = np.zeros((self.NS, 8)); ct = np.zeros((self.N, 5))
        assert self.lib.pl_get_plan(self.h, _dp(st), _dp(ct)) == self.NS
        return st, ct

    def candidates(self):
        nc = self.lib.pl_num_candidates(self.h)
        out = []
        for c in range(nc):
            st = np.zeros((self.NS, 8)); ct = np.zeros((self.N, 5))
            self.lib.pl_get_candidate(self.h, C.c_int(c), _dp(st), _dp(ct)); out.append((st, ct))
        return out

    def scores(self):
        sc = np.zeros((6, 3)); wd = np.zeros(6)
        k = self.lib.pl_get_scores(self.h, _dp(sc), _dp(wd))
        return sc[:k], wd[:k]

    def best(self):
        return int(self.lib.pl_best(self.h))

    def closest(self):
        return int(self.lib.pl_closest_obstacle(self.h))

    def status(self):
        st = np.zeros(8, dtype=np.int32); it = np.zeros(8, dtype=np.int32)
        k = self.lib.pl_last_status(self.h, st.ctypes.data_as(C.POINTER(C.c_int)), it.ctypes.data_as(C.POINTER(C.c_int)))
        return st[:k].copy(), it[:k].copy()

    def _v3(self, fn, t):
        o = np.zeros(3); fn(self.h, C.c_double(t), _dp(o)); return o

    def getPos(self, t): return self._v3(self.lib.pl_get_pos, t)
    def getVel(self, t): return self._v3(self.lib.pl_get_vel, t)
    def getAcc(self, t): return self._v3(self.lib.pl_get_acc, t)
    def getRef(self, t): return self._v3(self.lib.pl_get_ref, t)

    def getTrajectory(self):
        o = np.zeros((self.NS, 3)); k = self.lib.pl_get_trajectory(self.h, _dp(o)); return o[:k]

    def getReferenceTraj(self):
        o = np.zeros((self.NS, 3)); k = self.lib.pl_get_reference_traj(self.h, _dp(o)); return o[:k]


def _plan_vec(st, ct):
    return np.concatenate([np.asarray(st).reshape(-1), np.asarray(ct).reshape(-1)])


def _force_state(rp: MP.RefPlanner, st, ct):
    """Lockstep: the oracle continues from the plan of the implementation under test, so that every step compares the two
    on IDENTICAL inputs (their own plans agree to 1e-5; feeding each its own would let iteration counts drift apart)."""
    rp.currentStatesSol_ = [[float(v) for v in r] for r in st]
    rp.currentControlsSol_ = [[float(v) for v in r] for r in ct]


def _path_for(host, s, npts=400):
    """Reference path of scenario s as a list of points spaced speed * ts apart along the benchmark line (mpcNavigation.cpp:
    201-216 resamples its file the same way), so that getReferenceTraj's windowed nearest search is exercised."""
    k = np.arange(npts)
    x = np.minimum(host.pos[s, 0] + host.speed[s] * host.p.ts * k, 105.0)
    return np.stack([x, np.zeros(npts), np.full(npts, 2.0)], axis=1)


GETTER_TIMES = (0.0, 0.05, 0.1, 0.137, 1.0, 2.85, 2.9, 2.95, 3.5, 7.0)


def _check_getters(pl: Mirror, rp: MP.RefPlanner):
    """a16: getPos / getVel / getAcc / getRef / getTrajectory — the same IEEE operations in C++ and in the restatement
    (linear interpolation with the index clamped AFTER dt is formed, mpcPlanner.cpp:1257-1327): bit-exact."""
    for t in GETTER_TIMES:
        assert np.array_equal(pl.getPos(t), np.array(rp.getPos(t))), t
        assert np.array_equal(pl.getVel(t), np.array(rp.getVel(t))), t
        assert np.array_equal(pl.getAcc(t), np.array(rp.getAcc(t))), t
        assert np.array_equal(pl.getRef(t), np.array(rp.getRef(t))), t
    assert np.array_equal(pl.getTrajectory(), np.array(rp.getTrajectory()))
    # closed form on the plan itself: at stage times the interpolation returns the stage, halfway the midpoint
    st, ct = pl.plan()
    ts = pl.p.ts
    assert np.abs(pl.getPos(3 * ts) - st[3, 0:3]).max() < 1e-12 and np.abs(pl.getVel(5 * ts) - st[5, 3:6]).max() < 1e-12
    assert np.abs(pl.getPos(3.5 * ts) - 0.5 * (st[3, 0:3] + st[4, 0:3])).max() < 1e-9
    assert np.abs(pl.getAcc(2.5 * ts) - 0.5 * (ct[2, 0:3] + ct[3, 0:3])).max() < 1e-9


# =====================================================================================================================
# CPU tier: the oracle planner on the reference binary — it must behave like the reference's control loop
# =====================================================================================================================
def test_oracle_planner_receding_loop_on_reference_binary():
    orc = _oracle()
    host = receding.IntentSweep(S=3, D=3, seed0=3)
    p = MA.MpcParams()
    for s in range(host.S):
        rp = MP.RefPlanner(MA.MpcParams(), lambda qb: orc.solve_batch(qb, want_y=False))
        rp.updatePath(_path_for(host, s), p.ts)
        pos, vel = host.pos[s].copy(), host.vel[s].copy()
        x_start = pos[0]
        for step in range(4):
            pp, ps = host.predictions(step)
            rp.updateCurrStates(pos, vel)
            rp.updatePredObstacles(pp[s], ps[s], host.prob[s])
            assert rp.makePlanWithPred()
            if step == 0:
                assert len(rp.lastQps) == 1 and rp.lastQps[0]["num_obs"] == 0 and len(rp.candidateStates_) == 0
            else:
                assert [q["num_obs"] for q in rp.lastQps].count(3) == 4 and [q["num_obs"] for q in rp.lastQps].count(4) == 2
                assert len(rp.trajWeightedScore_) == 6 and 0 <= rp.bestTrajIdx_ < 6
                # evaluateTraj indexes the ORIGINAL-order weights with the SORTED position (mpcPlanner.cpp:866-880)
                w = rp.intentWeights(rp.obIdx_); sc = rp.trajScore_
                def acc(i):                                              # std::accumulate: plain left-to-right additions (Python's sum() compensates)
                    a = 0.0
                    for v in sc:
                        a += v[i]
                    return a / 6
                ca, da, sa = acc(0), acc(1), acc(2)
                for i in range(6):
                    assert rp.trajWeightedScore_[i] == w[i] * (1.0 * (ca / sc[i][0]) + 1.0 * (da / sc[i][1]) + 1.0 * (sc[i][2] / sa))
                assert sorted(rp.sortedCombo_) == list(range(6))
                ws = [w[c] for c in rp.sortedCombo_]
                assert all(ws[i] >= ws[i + 1] for i in range(5))
            assert all(int(q["out"]["status"][0]) in (1, 2, -2) for q in rp.lastQps)
            pos, vel = np.array(rp.getPos(p.ts)), np.array(rp.getVel(p.ts))
        assert pos[0] > x_start and np.isfinite(np.array(rp.currentStatesSol_)).all()
        # interpolation getters: stage values at stage times, clamped beyond the horizon
        assert rp.getPos(0.0) == rp.currentStatesSol_[0][0:3] and rp.getPos(100.0) == rp.currentStatesSol_[-1][0:3]
        assert rp.getRef(0.0) == rp.ref_[0][0:3] and rp.getAcc(100.0) == rp.currentControlsSol_[-1][0:3]


def test_oracle_obstacle_param_quirk_and_short_lists():
    """a4: updateObstacleParam (mpcPlanner.cpp:1148-1197): dynamic lists shorter than the window repeat their last element;
    the static loop clears isDyamic[j][i] with i, not i + numDynamicOb (:1194)."""
    p = MA.MpcParams()
    dyn_pos = [np.tile([1.0, 2.0, 3.0], (5, 1)) + np.arange(5)[:, None], np.tile([4.0, 5.0, 6.0], (40, 1))]
    dyn_size = [np.ones((5, 3)), 2 * np.ones((40, 3))]
    stat = [([7.0, 8.0, 9.0], [0.4, 0.4, 4.0], 0.3)]
    oxyz, osize, yaw, is_dyn = MA.obstacle_param(p, stat, dyn_pos, dyn_size)
    assert oxyz.shape == (29, 3, 3)
    assert np.array_equal(oxyz[10, 0], dyn_pos[0][-1]) and np.array_equal(oxyz[3, 0], dyn_pos[0][3])
    assert np.array_equal(osize[0, 2], np.array([0.4, 0.4, 4.0]) / 2 + p.static_safety_dist) and yaw[0, 2] == 0.3
    assert is_dyn[:, 0].tolist() == [0] * 29 and is_dyn[:, 1].tolist() == [1] * 29 and is_dyn[:, 2].tolist() == [0] * 29


# =====================================================================================================================
# GPU tier
# =====================================================================================================================
def _device_step_arrays(torch, dev, eng, p, S, D, pp, ps, prob, pos, plans):
    """getIntentComb on the device for S scenarios; returns the host copies of everything the kernels wrote + the tensors."""
    N = p.N
    d = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    t = dict(pp=d(pp), ps=d(ps), prob=d(prob), pos=d(pos), plan=None if plans is None else d(plans))
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    out = dict(scen_a=torch.empty(4 * S, **i32), scen_b=torch.empty(2 * S, **i32), weight=torch.empty((S, 6), **f64), cand=torch.empty((S, 6), **i32),
               obs_c_a=torch.empty((4 * S, N, D, 3), **f64), obs_semi_a=torch.empty((4 * S, N, D, 3), **f64),
               obs_c_b=torch.empty((2 * S, N, D + 1, 3), **f64), obs_semi_b=torch.empty((2 * S, N, D + 1, 3), **f64),
               obs_c_last_a=torch.empty((4 * S, D, 3), **f64), obs_semi_last_a=torch.empty((4 * S, D, 3), **f64),
               obs_c_last_b=torch.empty((2 * S, D + 1, 3), **f64), obs_semi_last_b=torch.empty((2 * S, D + 1, 3), **f64))
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    ptrs.update(pred_pos=t["pp"].data_ptr(), pred_size=t["ps"].data_ptr(), prob=t["prob"].data_ptr(), pos=t["pos"].data_ptr(),
                prev_plan=0 if plans is None else t["plan"].data_ptr())
    eng.intent_candidates_ptr(p, S, D, pp.shape[3], ptrs)
    eng.sync()
    return out, t


def _expected_rows(rp: MP.RefPlanner, p, D):
    """What updateObstacleParam makes of the oracle's six hypotheses: per sorted candidate the [N+1][rows][3] centres / semi-axes
    (stage N included: getSafetyScore reads it)."""
    ob, combP, combS = rp.getIntentComb()
    rows = []
    for c in range(6):
        cen = np.array([[q[k] for q in combP[c]] for k in range(p.N + 1)])                       # [N+1][rows][3]
        semi = np.array([[q[k] for q in combS[c]] for k in range(p.N + 1)]) / 2 + p.dynamic_safety_dist
        rows.append((cen, semi))
    return ob, rows, (ob, combP, combS)


@pytest.mark.gpu
def test_lockstep_mirror_device_oracle_100_scenarios_20_steps():
    """a1, a3, a15, a16, f1, f2 — see the module docstring for the bars."""
    import torch
    lib = build_shim()
    orc = _oracle()
    eng = engine.Engine(0)
    dev = torch.device("cuda", 0)
    S, D, STEPS = 100, 4, 20
    host = receding.IntentSweep(S=S, D=D, seed0=2024)
    p = MA.MpcParams()
    wp = host.p
    N, NS, n = p.N, p.horizon, p.n
    mirrors = [Mirror(lib, p) for _ in range(S)]
    oracles = [MP.RefPlanner(MA.MpcParams(), lambda qb: orc.solve_batch(qb, want_y=False)) for _ in range(S)]
    for s in range(S):
        path = _path_for(host, s)
        mirrors[s].updatePath(path, p.ts); oracles[s].updatePath(path, p.ts)
    pos, vel = host.pos.copy(), host.vel.copy()
    pool = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1))
    st_i32 = engine.default_settings()
    near_ties = 0; n_qps = 0
    try:
        for step in range(STEPS):
            pp, ps = host.predictions(step)
            for s in range(S):
                mirrors[s].updateCurrStates(pos[s], vel[s]); oracles[s].updateCurrStates(pos[s], vel[s])
                mirrors[s].updatePredObstacles(pp[s], ps[s], host.prob[s]); oracles[s].updatePredObstacles(pp[s], ps[s], host.prob[s])
            plans_prev = None if step == 0 else np.stack([_plan_vec(*mirrors[s].plan()) for s in range(S)])
            # ---- a3: enumeration, three ways (the oracle call below does not change any planner state) -------------
            if step > 0:
                dout, dten = _device_step_arrays(torch, dev, eng, wp, S, D, pp, ps, host.prob, pos, plans_prev)
                cand = dout["cand"].cpu().numpy(); weight = dout["weight"].cpu().numpy()
                scen = np.concatenate([dout["scen_a"].cpu().numpy(), dout["scen_b"].cpu().numpy()])
                oc = {"a": dout["obs_c_a"].cpu().numpy(), "b": dout["obs_c_b"].cpu().numpy()}
                om = {"a": dout["obs_semi_a"].cpu().numpy(), "b": dout["obs_semi_b"].cpu().numpy()}
                ocl = {"a": dout["obs_c_last_a"].cpu().numpy(), "b": dout["obs_c_last_b"].cpu().numpy()}
                oml = {"a": dout["obs_semi_last_a"].cpu().numpy(), "b": dout["obs_semi_last_b"].cpu().numpy()}
                exp_ob = np.zeros(S, dtype=int); combs = [None] * S
                for s in range(S):
                    ob, rows, combs[s] = _expected_rows(oracles[s], p, D)
                    exp_ob[s] = ob
                    assert np.array_equal(weight[s], np.array(oracles[s].intentWeights(ob)))       # original order, bit-exact
                    na = nb = 0
                    for c in range(6):
                        cen, semi = rows[c]
                        two = cen.shape[1] == D + 1
                        row = cand[s, c]
                        assert row == (4 * S + 2 * s + nb if two else 4 * s + na), (step, s, c)      # sorted hypotheses -> batch rows
                        assert scen[row] == s
                        key, r = ("b", row - 4 * S) if two else ("a", row)
                        assert np.array_equal(oc[key][r], cen[:N]) and np.array_equal(om[key][r], semi[:N]), (step, s, c)
                        assert np.array_equal(ocl[key][r], cen[N]) and np.array_equal(oml[key][r], semi[N]), (step, s, c)
                        nb += two; na += not two
            # ---- the mirror's control step (QPs on the GPU through the host entry point) --------------------------------
            for s in range(S):
                assert mirrors[s].makePlanWithPred(), lib.pl_last_error(mirrors[s].h)
            # ---- the oracle's control step (QPs on the reference binary), all scenarios in parallel threads -----------
            ok = list(pool.map(lambda rp: rp.makePlanWithPred(), oracles))
            assert all(ok)
            own_wd = [np.array(rp.trajWeightedScore_) for rp in oracles]
            for s in range(S):
                pl, rp = mirrors[s], oracles[s]
                st, it = pl.status()
                rst = np.array([int(q["out"]["status"][0]) for q in rp.lastQps]); rit = np.array([int(q["out"]["iter"][0]) for q in rp.lastQps])
                assert np.array_equal(st, rst) and np.array_equal(it, rit), (step, s, st, rst, it, rit)
                n_qps += len(rst)
                if step == 0:
                    mst, mct = pl.plan()
                    assert rel_inf(_plan_vec(mst, mct)[None], rp.lastQps[0]["out"]["x"]).max() < TOL
                    continue
                assert pl.closest() == rp.obIdx_ == exp_ob[s], (step, s)
                cands = pl.candidates()
                assert len(cands) == 6 and len(rp.candidateStates_) == 6
                for c in range(6):
                    xm = _plan_vec(*cands[c])
                    assert rel_inf(xm[None], rp.lastQps[c]["out"]["x"]).max() < TOL, (step, s, c)
                # a15 on identical x: the oracle's scoring functions applied to the MIRROR's candidates
                ob, combP, combS = combs[s]
                xref = rp.ref_
                msc, mwd = pl.scores()
                osc = []
                keep = (rp.currentStatesSol_, rp.currentControlsSol_)
                _force_state(rp, plans_prev[s][: 8 * NS].reshape(NS, 8), plans_prev[s][8 * NS:].reshape(N, 5))   # the plan the candidates were scored against
                for c in range(6):
                    stc = [list(map(float, r)) for r in cands[c][0]]
                    osc.append(rp.getTrajectoryScore(stc, None, [], combP[c], combS[c], xref))
                obest = rp.evaluateTraj(osc, ob, list(range(6))); owd = np.array(rp.trajWeightedScore_)
                rp.currentStatesSol_, rp.currentControlsSol_ = keep
                osc = np.array(osc)
                assert np.abs(msc - osc).max() <= TOL_SCORE * np.abs(osc).max(), (step, s)
                assert np.abs(mwd - owd).max() <= TOL_SCORE * np.abs(owd).max(), (step, s)
                top = np.sort(owd)[::-1]
                if (top[0] - top[1]) > 1e-9 * abs(top[0]):
                    assert pl.best() == obest, (step, s, mwd, owd)
                    # the oracle's OWN step (its own x, up to 1e-5 away) picks the same candidate whenever the margin exceeds twice
                    # what that difference moved the weighted scores by (an argmax cannot change under less)
                    if (top[0] - top[1]) > 4 * np.abs(own_wd[s] - owd).max():
                        assert rp.bestTrajIdx_ == obest, (step, s)
                else:
                    near_ties += 1
            # ---- f1: the device kernels on the same step: solves (bit-identical to the mirror's), scores, choice ---------
            if step > 0:
                xs = np.zeros((6 * S, n))
                for s in range(S):
                    cands = mirrors[s].candidates()
                    for c in range(6):
                        xs[cand[s, c]] = _plan_vec(*cands[c])
                d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                xref_s = np.stack([np.array(oracles[s].ref_)[:, 0:3] for s in range(S)])
                x0_s = np.concatenate([pos, vel], axis=1); lin_s = plans_prev[:, : 8 * NS].reshape(S, NS, 8)[:, :N, 0:3]
                g = {k: d(np.ascontiguousarray(v[scen])) for k, v in (("x0", x0_s), ("xref", xref_s), ("lin", lin_s), ("warm", plans_prev))}
                bufs = dict(x=torch.empty((6 * S, n), dtype=torch.float64, device=dev), score=torch.empty((6 * S, 3), dtype=torch.float64, device=dev))
                for k in ("status", "iter", "rho_updates"):
                    bufs[k] = torch.empty(6 * S, dtype=torch.int32, device=dev)
                for k in ("obj", "pri_res", "dua_res"):
                    bufs[k] = torch.empty(6 * S, dtype=torch.float64, device=dev)
                yaw = {"a": torch.zeros((4 * S, N, D), dtype=torch.float64, device=dev), "b": torch.zeros((2 * S, N, D + 1), dtype=torch.float64, device=dev)}
                for lo, B_, R_, key in ((0, 4 * S, D, "a"), (4 * S, 2 * S, D + 1, "b")):
                    sl = lambda t_: t_[lo:lo + B_].data_ptr()
                    ptrs = {"x0": sl(g["x0"]), "xref": sl(g["xref"]), "lin_pt": sl(g["lin"]), "warm_x": sl(g["warm"]), "x": sl(bufs["x"]), "status": sl(bufs["status"]),
                            "iter": sl(bufs["iter"]), "rho_updates": sl(bufs["rho_updates"]), "obj": sl(bufs["obj"]), "pri_res": sl(bufs["pri_res"]), "dua_res": sl(bufs["dua_res"]),
                            "obs_c": dout["obs_c_" + key].data_ptr(), "obs_semi": dout["obs_semi_" + key].data_ptr(), "obs_yaw": yaw[key].data_ptr()}
                    eng.solve_mpc_batch_ptr(wp, st_i32, B_, R_, ptrs, np.ones((N, R_), dtype=np.int32), device=True)
                    eng.sync()
                    eng.score_candidates_ptr(wp, B_, R_, R_, {"x": sl(bufs["x"]), "prev_plan": sl(g["warm"]), "xref": sl(g["xref"]), "obs_c": dout["obs_c_" + key].data_ptr(),
                                                             "obs_semi": dout["obs_semi_" + key].data_ptr(), "obs_c_last": dout["obs_c_last_" + key].data_ptr(),
                                                             "obs_semi_last": dout["obs_semi_last_" + key].data_ptr(), "score": sl(bufs["score"])})
                d_best = torch.empty(S, dtype=torch.int32, device=dev); d_wd = torch.empty((S, 6), dtype=torch.float64, device=dev)
                d_plan = torch.empty((S, n), dtype=torch.float64, device=dev)
                eng.select_candidates_ptr(S, 6, n, {"cand": dout["cand"].data_ptr(), "weight": dout["weight"].data_ptr(), "score": bufs["score"].data_ptr(),
                                                    "x_all": bufs["x"].data_ptr(), "best": d_best.data_ptr(), "weighted": d_wd.data_ptr(), "plan": d_plan.data_ptr()})
                eng.sync()
                assert np.array_equal(bufs["x"].cpu().numpy(), xs)             # device entry == host entry, whatever the batch composition
                dsc = bufs["score"].cpu().numpy(); dwd = d_wd.cpu().numpy(); dbest = d_best.cpu().numpy(); dplan = d_plan.cpu().numpy()
                for s in range(S):
                    msc, mwd = mirrors[s].scores()
                    assert np.abs(dsc[cand[s]] - msc).max() <= TOL_SCORE * np.abs(msc).max(), (step, s)
                    assert np.abs(dwd[s] - mwd).max() <= TOL_SCORE * np.abs(mwd).max(), (step, s)
                    top = np.sort(mwd)[::-1]
                    if (top[0] - top[1]) > 1e-9 * abs(top[0]):
                        assert dbest[s] == mirrors[s].best(), (step, s)
                        assert np.array_equal(dplan[s], _plan_vec(*mirrors[s].plan()))
            # ---- lockstep + a16 + roll forward by perfect tracking (mpc_node.cpp:223-224) ------------------------------
            for s in range(S):
                mst, mct = mirrors[s].plan()
                _force_state(oracles[s], mst, mct)
                if step in (0, 1, 7, STEPS - 1):
                    _check_getters(mirrors[s], oracles[s])
                assert np.array_equal(mirrors[s].getReferenceTraj(), np.array(oracles[s].getReferenceTraj()))
                pos[s] = mirrors[s].getPos(p.ts); vel[s] = mirrors[s].getVel(p.ts)
        assert n_qps == S + (STEPS - 1) * 6 * S
        assert near_ties == 0, near_ties
        assert (pos[:, 0] > host.pos[:, 0]).all()
    finally:
        pool.shutdown()
        for m in mirrors:
            m.close()
        eng.close()


@pytest.mark.gpu
def test_mirror_with_static_obstacles_and_make_plan_match_oracle():
    """a2 (makePlan, mpcPlanner.cpp:543-569: obstacles held at their current position) and a4 (static obstacles next to dynamic
    ones, incl. the isDyamic index quirk) through the C++ mirror against the oracle planner on the reference binary."""
    lib = build_shim()
    orc = _oracle()
    S, STEPS = 8, 5
    host = receding.IntentSweep(S=S, D=3, seed0=77)
    p = MA.MpcParams()
    r = np.random.default_rng(5)
    try:
        mirrors = [Mirror(lib, p) for _ in range(2 * S)]
        for mode in ("makePlan", "makePlanWithPred"):
            for s in range(S):
                pl = mirrors[s + (S if mode == "makePlan" else 0)]
                rp = MP.RefPlanner(MA.MpcParams(), lambda qb: orc.solve_batch(qb, want_y=False))
                path = _path_for(host, s)
                pl.updatePath(path, p.ts); rp.updatePath(path, p.ts)
                nstat = 1 + s % 3                                                # 1..3 static boxes; with 3 dynamic ones min(S, D) flags flip
                stat = [([host.pos[s, 0] + r.uniform(4, 18), r.uniform(-3, 3), 2.0], [0.4, 0.4, 4.0] if j % 2 == 0 else [0.4, 4.0, 0.4], r.uniform(-1.2, 1.2))
                        for j in range(nstat)]
                pl.updateStaticObstacles(stat); rp.updateStaticObstacles(stat)
                pos, vel = host.pos[s].copy(), host.vel[s].copy()
                for step in range(STEPS):
                    pp, ps = host.predictions(step)
                    pl.updateCurrStates(pos, vel); rp.updateCurrStates(pos, vel)
                    if mode == "makePlan":
                        op, ov = host.obstacle_state(step)
                        pl.updateDynamicObstacles(op[s], ov[s], host.size[s]); rp.updateDynamicObstacles(op[s], ov[s], host.size[s])
                        assert pl.makePlan() and rp.makePlan()
                        assert len(rp.lastQps) == 1 and rp.lastQps[0]["num_obs"] == (0 if step == 0 else 3 + nstat)
                    else:
                        pl.updatePredObstacles(pp[s], ps[s], host.prob[s]); rp.updatePredObstacles(pp[s], ps[s], host.prob[s])
                        assert pl.makePlanWithPred() and rp.makePlanWithPred()
                    st, it = pl.status()
                    rst = np.array([int(q["out"]["status"][0]) for q in rp.lastQps]); rit = np.array([int(q["out"]["iter"][0]) for q in rp.lastQps])
                    assert np.array_equal(st, rst) and np.array_equal(it, rit), (mode, s, step, st, rst, it, rit)
                    cands = pl.candidates() if (mode == "makePlanWithPred" and step > 0) else [pl.plan()]
                    for c, (cs, cc) in enumerate(cands):
                        assert rel_inf(_plan_vec(cs, cc)[None], rp.lastQps[c]["out"]["x"]).max() < TOL, (mode, s, step, c)
                    if mode == "makePlanWithPred" and step > 0:
                        msc, mwd = pl.scores()
                        assert np.abs(msc - np.array(rp.trajScore_)).max() <= 1e-3 * np.abs(msc).max()      # own x on each side: up to 1e-5 of |x| apart
                        top = np.sort(mwd)[::-1]
                        if top[0] - top[1] > 4 * np.abs(mwd - np.array(rp.trajWeightedScore_)).max():
                            assert pl.best() == rp.bestTrajIdx_
                    mst, mct = pl.plan()
                    _force_state(rp, mst, mct)
                    _check_getters(pl, rp)
                    pos, vel = pl.getPos(p.ts), pl.getVel(p.ts)
    finally:
        for m in mirrors:
            m.close()
