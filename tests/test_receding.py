"""BASELINE.json configs[2] at test size: intent-predicted dynamic obstacles, six candidates per scenario per control
step, warm-started receding-horizon loop (intent-mpc_b200/receding.py).  CPU tier: the host logic runs end to end on the
oracle.  GPU tier: the loop is driven by the CUDA engine and every QP of every step is checked against the oracle on the
identical inputs (status, iterations, 1e-5 on x and objective); then the full-size batch (65,536 QPs) for two steps,
checked through size-independent properties."""
import os

import numpy as np
import pytest

from intent_mpc_b200 import receding
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf

TOL = 1e-5


def _oracle():
    return OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()


def _oracle_solve(mb):
    return _oracle().solve_batch(to_qp_batch(mb), want_y=False, nthreads=os.cpu_count() or 1)


def test_receding_horizon_host_logic_on_oracle():
    sw = receding.IntentSweep(S=6, D=3, seed0=3)
    x_start = sw.pos[:, 0].copy()
    for step in range(4):
        r = sw.step(_oracle_solve)
        if step == 0:
            assert len(r["batches"]) == 1 and r["batches"][0].num_obs == 0
        else:
            assert [b.num_obs for b in r["batches"]] == [3, 4] and [b.B for b in r["batches"]] == [24, 12]
            assert set(np.unique(r["status"])) <= {1, 2, -2}
            assert (r["best"] >= 0).all() and (r["best"] < 6).all()
    assert (sw.pos[:, 0] > x_start).all() and np.isfinite(sw.states).all()


@pytest.mark.gpu
def test_receding_horizon_gpu_matches_oracle_every_step():
    from intent_mpc_b200 import engine
    eng = engine.Engine(0)
    sw = receding.IntentSweep(S=16, D=4, seed0=11)
    n_checked = 0
    for step in range(6):
        r = sw.step(lambda mb: eng.solve_mpc_batch(mb))
        for mb, out in zip(r["batches"], r["outs"]):
            ref = _oracle_solve(mb)
            assert (out["status"] == ref["status"]).all(), f"step {step}"
            assert (out["iter"] == ref["iter"]).all(), f"step {step}"
            assert (out["rho_updates"] == ref["rho_updates"]).all()
            assert rel_inf(out["x"], ref["x"]).max() < TOL
            assert np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL
            n_checked += mb.B
    assert n_checked == 16 + 5 * 96
    eng.close()


@pytest.mark.gpu
def test_receding_horizon_full_size_properties():
    """65,536 candidate QPs per control step (10,923 scenarios x 6, BASELINE.json configs[2]); two control steps."""
    from intent_mpc_b200 import engine
    eng = engine.Engine(0)
    S = 10923
    sw = receding.IntentSweep(S=S, D=4, seed0=5)
    sw.step(lambda mb: eng.solve_mpc_batch(mb))
    r = sw.step(lambda mb: eng.solve_mpc_batch(mb))
    assert sum(b.B for b in r["batches"]) == 6 * S >= 65536
    p = sw.p
    ts = float(np.float32(p.ts)); h = float(np.float32(0.5 * p.ts ** 2))
    for mb, out in zip(r["batches"], r["outs"]):
        again = eng.solve_mpc_batch(mb)
        assert np.array_equal(again["x"], out["x"])                     # deterministic at full size
        ok = out["status"] == 1
        assert ok.mean() > 0.8
        X = out["x"][ok]; NS = p.N + 1
        st = X[:, : 8 * NS].reshape(-1, NS, 8); u = X[:, 8 * NS:].reshape(-1, p.N, 5)
        scale = 1e-2 * (1 + np.abs(st[:, :, 0:3]).max())
        assert np.abs(st[:, :-1, 0:3] + ts * st[:, :-1, 3:6] + h * u[:, :, 0:3] - st[:, 1:, 0:3]).max() < scale
        assert np.abs(st[:, 0, 0:6] - np.concatenate([mb.x0[ok, 0:3], mb.x0[ok, 3:6]], axis=1)).max() < scale
        assert np.abs(u[:, :, 0:3]).max() <= p.max_acc + 0.2              # input box to the solver tolerance
        assert (out["pri_res"][ok] < 1e-3 + 1e-3 * 1e3).all()                # solved => primal residual within eps_abs + eps_rel*|Ax|
    eng.close()


@pytest.mark.gpu
def test_device_scoring_and_selection_match_host_logic():
    """§8(f) row 1: getTrajectoryScore / evaluateTraj (mpcPlanner.cpp:771-887) as device kernels
    (mpcqp_score_candidates_device, mpcqp_select_candidates_device) against the numpy restatement that drives the
    receding-horizon tests above: same scores, same weighted values, same chosen candidate and plan, every step."""
    import torch
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding import IntentSweep
    eng = engine.Engine(0)
    dev = torch.device("cuda", 0)
    try:
        sw = IntentSweep(48, seed0=123)
        sw.step(eng.solve_mpc_batch)                       # first control step: one obstacle-free QP per scenario
        for step in range(4):
            prev_first = sw.first
            batches, meta = sw.candidates()
            outs = [eng.solve_mpc_batch(mb) for mb in batches]
            cand_x, status, iters, weighted, best = sw.select(batches, meta, outs)
            p = sw.p; S = sw.S; n = p.n
            xs = torch.from_numpy(np.concatenate([o["x"] for o in outs])).to(dev)
            score = torch.empty((xs.shape[0], 3), dtype=torch.float64, device=dev)
            off = 0
            keep = []
            for mb, out in zip(batches, outs):
                B, R = mb.B, mb.num_obs
                t = {k: torch.from_numpy(np.ascontiguousarray(getattr(mb, k))).to(dev) for k in ("xref", "obs_c", "obs_semi", "warm_x")}
                keep.append(t)
                eng.score_candidates_ptr(p, B, R, R, {"x": xs[off:off + B].data_ptr(), "prev_plan": 0 if prev_first else t["warm_x"].data_ptr(),
                                                      "xref": t["xref"].data_ptr(), "obs_c": t["obs_c"].data_ptr(), "obs_semi": t["obs_semi"].data_ptr(),
                                                      "score": score[off:off + B].data_ptr()})
                off += B
            cand = np.zeros((S, 6), dtype=np.int32)
            off = 0
            for mt in meta:
                cand[mt[:, 0], mt[:, 1]] = off + np.arange(len(mt))
                off += len(mt)
            d_cand = torch.from_numpy(cand).to(dev); d_w = torch.from_numpy(np.ascontiguousarray(sw.last["w"])).to(dev)
            d_best = torch.empty(S, dtype=torch.int32, device=dev); d_wd = torch.empty((S, 6), dtype=torch.float64, device=dev)
            d_plan = torch.empty((S, n), dtype=torch.float64, device=dev)
            eng.select_candidates_ptr(S, 6, n, {"cand": d_cand.data_ptr(), "weight": d_w.data_ptr(), "score": score.data_ptr(), "x_all": xs.data_ptr(),
                                                "best": d_best.data_ptr(), "weighted": d_wd.data_ptr(), "plan": d_plan.data_ptr()})
            eng.sync()
            wd = d_wd.cpu().numpy()
            fin = np.isfinite(weighted)
            assert np.array_equal(np.isfinite(wd), fin)
            assert np.abs(wd[fin] - weighted[fin]).max() <= 1e-9 * np.abs(weighted[fin]).max()
            assert np.array_equal(d_best.cpu().numpy(), best.astype(np.int32))
            assert np.array_equal(d_plan.cpu().numpy(), cand_x[np.arange(S), best])
            sw.advance(cand_x, best)
    finally:
        eng.close()


@pytest.mark.gpu
def test_receding_horizon_100_steps_warm_started():
    """BASELINE.json configs[2]'s loop length: 100 control steps, every candidate QP warm-started from the plan chosen one
    step earlier.  64 scenarios x 6 candidates per step on the GPU; every 10th step is re-solved by the oracle on the
    identical inputs (status, iterations, 1e-5 on x); the scenarios must keep advancing along the reference line."""
    from intent_mpc_b200 import engine
    eng = engine.Engine(0)
    try:
        sw = receding.IntentSweep(S=64, D=4, seed0=77)
        x_start = sw.pos[:, 0].copy()
        total_q = 0; it_sum = 0
        for step in range(100):
            r = sw.step(lambda mb: eng.solve_mpc_batch(mb))
            for mb, out in zip(r["batches"], r["outs"]):
                assert np.isin(out["status"], [1, 2, -2]).all()
                total_q += mb.B; it_sum += int(out["iter"].sum())
                if step % 10 == 9:
                    ref = _oracle_solve(mb)
                    assert (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all(), f"step {step}"
                    assert rel_inf(out["x"], ref["x"]).max() < TOL
        assert total_q == 64 + 99 * 384
        assert np.isfinite(sw.states).all()
        assert (sw.pos[:, 0] >= x_start - 1e-6).all() and (sw.pos[:, 0] - x_start).mean() > 5.0     # >= 5 m of progress in 10 s
        # warm start pays: a cold solve of the last step's QPs needs more iterations than the warm one did
        mb = r["batches"][0]
        warm_it = int(r["outs"][0]["iter"].sum())
        cold = receding.MpcBatch(mb.params, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.obs_dyn, mb.lin_pt, np.zeros_like(mb.warm_x))
        assert int(eng.solve_mpc_batch(cold)["iter"].sum()) > warm_it
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_candidate_enumeration_matches_host_logic():
    """§8(f) row 1, first half: getIntentComb / findClosestObstacle (mpcPlanner.cpp:663-769) as device kernels
    (mpcqp_intent_candidates_device + mpcqp_gather_rows_device) against the numpy restatement: same closest obstacle, same
    sorted hypotheses, same rows in the two solve batches, same weights — on the first step and with a previous plan."""
    import torch
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding import IntentSweep
    eng = engine.Engine(0)
    dev = torch.device("cuda", 0)
    try:
        sw = IntentSweep(300, seed0=321)
        for step in range(3):
            if step == 0:
                sw.first = False                              # exercise the enumeration on the very first step too (no plan yet)
            first = sw.states is None
            batches, meta = sw.candidates()
            p, S, D = sw.p, sw.S, sw.D
            N, n = p.N, p.n
            pp, ps = sw.last["pp"], sw.last["ps"]
            d = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
            t_pp, t_ps, t_prob, t_pos = d(pp), d(ps), d(sw.prob), d(sw.pos)
            t_plan = None if first else d(batches[0].warm_x[np.unique(meta[0][:, 0], return_index=True)[1]])
            out = {"scen_a": torch.empty(4 * S, dtype=torch.int32, device=dev), "scen_b": torch.empty(2 * S, dtype=torch.int32, device=dev),
                   "obs_c_a": torch.empty((4 * S, N, D, 3), dtype=torch.float64, device=dev), "obs_semi_a": torch.empty((4 * S, N, D, 3), dtype=torch.float64, device=dev),
                   "obs_c_b": torch.empty((2 * S, N, D + 1, 3), dtype=torch.float64, device=dev), "obs_semi_b": torch.empty((2 * S, N, D + 1, 3), dtype=torch.float64, device=dev),
                   "weight": torch.empty((S, 6), dtype=torch.float64, device=dev), "cand": torch.empty((S, 6), dtype=torch.int32, device=dev)}
            ptrs = {k: v.data_ptr() for k, v in out.items()}
            ptrs.update(pred_pos=t_pp.data_ptr(), pred_size=t_ps.data_ptr(), prob=t_prob.data_ptr(), pos=t_pos.data_ptr(),
                        prev_plan=0 if first else t_plan.data_ptr())
            eng.intent_candidates_ptr(p, S, D, pp.shape[3], ptrs)
            # scenario-level arrays replicated per row of batch a
            t_x0 = d(np.concatenate([sw.pos, sw.vel], axis=1)); g_x0 = torch.empty((4 * S, 6), dtype=torch.float64, device=dev)
            eng.gather_rows_ptr(4 * S, 6, out["scen_a"].data_ptr(), t_x0.data_ptr(), g_x0.data_ptr())
            eng.sync()
            assert np.array_equal(out["scen_a"].cpu().numpy(), meta[0][:, 0]) and np.array_equal(out["scen_b"].cpu().numpy(), meta[1][:, 0])
            assert np.array_equal(out["obs_c_a"].cpu().numpy(), batches[0].obs_c) and np.array_equal(out["obs_semi_a"].cpu().numpy(), batches[0].obs_semi)
            assert np.array_equal(out["obs_c_b"].cpu().numpy(), batches[1].obs_c) and np.array_equal(out["obs_semi_b"].cpu().numpy(), batches[1].obs_semi)
            assert np.array_equal(out["weight"].cpu().numpy(), sw.last["w"])
            cand = np.zeros((S, 6), dtype=np.int32); off = 0
            for mt in meta:
                cand[mt[:, 0], mt[:, 1]] = off + np.arange(len(mt)); off += len(mt)
            assert np.array_equal(out["cand"].cpu().numpy(), cand)
            assert np.array_equal(g_x0.cpu().numpy(), batches[0].x0)
            if step == 0:
                sw.first = True
                sw.step(eng.solve_mpc_batch)                 # the real first step (obstacle-free QPs), then continue with plans
            else:
                outs = [eng.solve_mpc_batch(mb) for mb in batches]
                cand_x, status, iters, weighted, best = sw.select(batches, meta, outs)
                sw.advance(cand_x, best)
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_resident_control_loop_tracks_the_host_loop():
    """makePlanWithPred as a chain of engine calls with every array on the GPU (intent-mpc_b200/receding_device.py):
    enumeration -> gather -> two solves -> scoring -> choice -> next state.  Run beside the host-logic loop from the same
    initial scenarios: same chosen candidates and the same UAV states step after step (predictions are evaluated with torch
    on one side and numpy on the other, so inputs agree to rounding, not bitwise)."""
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding import IntentSweep
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        host = IntentSweep(256, seed0=4242)
        devs = DeviceIntentSweep(eng, IntentSweep(256, seed0=4242))
        for step in range(6):
            r = host.step(eng.solve_mpc_batch)
            best_d = devs.step()
            pos_d = devs.pos.cpu().numpy()
            close = np.abs(pos_d - host.pos).max(axis=1) < 1e-6
            assert close.mean() >= 0.98, (step, close.mean())
            if r["best"] is not None:
                same = best_d.cpu().numpy() == r["best"]
                assert same.mean() >= 0.98, (step, same.mean())
                st_d = devs.buf["status"].cpu().numpy()
                assert np.isin(st_d, [1, 2, -2]).all()
        assert devs.kernel_ms > 0
    finally:
        eng.close()
