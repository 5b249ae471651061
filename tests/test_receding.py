"""BASELINE.json configs[2] at test size: intent-predicted dynamic obstacles, six candidates per scenario per control
step, warm-started receding-horizon loop with every array on the GPU (intent-mpc_b200/receding_device.py: enumeration,
gather, two solves, scoring, choice are engine calls).  Every QP of every checked step is re-solved by the oracle (the
reference's OSQP binary) on the identical inputs (status, iterations, rho updates, 1e-5 on x and objective); the full-size batch
(65,538 QPs per step) is checked through size-independent properties.  The orchestration around the QPs (which obstacle, which
hypotheses, scores, choice) is pinned in tests/test_planner_parity.py."""
import os

import numpy as np
import pytest

from intent_mpc_b200 import receding
from intent_mpc_b200.workloads import MpcBatch
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf

TOL = 1e-5


def _oracle():
    return OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()


def _oracle_solve(mb):
    return _oracle().solve_batch(to_qp_batch(mb), want_y=False, nthreads=os.cpu_count() or 1)


def _check_against_oracle(mb, out, tag):
    ref = _oracle_solve(mb)
    assert (out["status"] == ref["status"]).all(), tag
    assert (out["iter"] == ref["iter"]).all(), tag
    assert (out["rho_updates"] == ref["rho_updates"]).all(), tag
    assert rel_inf(out["x"], ref["x"]).max() < TOL, tag
    assert np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL, tag


def test_scenario_generator_shapes_and_first_batch():
    sw = receding.IntentSweep(S=5, D=3, seed0=1)
    pp, ps = sw.predictions(0)
    assert pp.shape == ps.shape == (5, 3, 4, 31, 3) and sw.prob.shape == (5, 3, 4)
    assert np.allclose(sw.prob.sum(axis=-1), 1.0)
    assert np.array_equal(pp[:, :, receding.STOP, 0], pp[:, :, receding.STOP, 30])          # STOP: standing still
    assert np.array_equal(pp[:, :, receding.FORWARD, 0], pp[:, :, receding.LEFT, 0])         # all intents start at the obstacle
    mb = sw.first_step_batch()
    assert mb.B == 5 and mb.num_obs == 0 and mb.xref.shape == (5, 30, 3) and not mb.warm_x.any()


@pytest.mark.gpu
def test_device_loop_every_qp_matches_oracle():
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        ds = DeviceIntentSweep(eng, receding.IntentSweep(S=16, D=4, seed0=11))
        n_checked = 0
        for step in range(6):
            best = ds.step()
            assert (best is None) == (step == 0)
            for mb, out in ds.batches_host():
                _check_against_oracle(mb, out, f"step {step}")
                n_checked += mb.B
        assert n_checked == 16 + 5 * 96
        assert ds.kernel_ms > 0
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_loop_with_host_predictions_counts_its_traffic():
    """The end-to-end form bench.py times for configs[2]: predictions come from the host every step, the plan goes back."""
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        ds = DeviceIntentSweep(eng, receding.IntentSweep(S=32, D=4, seed0=21), host_predictions=True)
        for step in range(3):
            ds.step()
        assert ds.h2d_bytes == 2 * 2 * 32 * 4 * 4 * 31 * 3 * 8 and ds.d2h_bytes == 3 * 32 * ds.p.n * 8
        for mb, out in ds.batches_host():
            _check_against_oracle(mb, out, "host predictions")
        assert np.array_equal(ds.plan_host.numpy(), ds.plan.cpu().numpy())
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_loop_with_the_sampled_predictor():
    """SURVEY.md 8(f) row 4 wired in: predictions and intent probabilities come from mpcqp_predict_device (the reference
    predictor's sampling, dynamicPredictor.cpp:197-541) instead of the closed-form stand-in; the candidates it leads to are
    still solved exactly like the reference would (oracle on the identical QPs)."""
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        for hostp in (False, True):
            host = receding.IntentSweep(S=24, D=4, seed0=61)
            ds = DeviceIntentSweep(eng, host, predictor="sampled", host_predictions=hostp)
            x_start = host.pos[:, 0].copy()
            for step in range(5):
                ds.step()
            prob = ds.prob.cpu().numpy()
            assert np.isfinite(prob).all() and (prob >= 0).all() and not np.array_equal(prob, host.prob)      # the HMM's, not the generator's
            for mb, out in ds.batches_host():
                _check_against_oracle(mb, out, f"sampled predictor, host histories {hostp}")
            assert (ds.pos.cpu().numpy()[:, 0] > x_start).all()
            if hostp:
                assert ds.h2d_bytes == 4 * 2 * 24 * 4 * 10 * 3 * 8                                             # histories only
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_loop_full_size_properties():
    """65,538 candidate QPs per control step (10,923 scenarios x 6, BASELINE.json configs[2]); two control steps."""
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        S = 10923
        ds = DeviceIntentSweep(eng, receding.IntentSweep(S=S, D=4, seed0=5))
        ds.step(); best = ds.step()
        bt = ds.batches_host()
        assert sum(mb.B for mb, _ in bt) == 6 * S >= 65536
        b = best.cpu().numpy()
        assert b.min() >= 0 and b.max() < 6
        p = ds.p
        ts = float(np.float32(p.ts)); h = float(np.float32(0.5 * p.ts ** 2))
        for mb, out in bt:
            again = eng.solve_mpc_batch(mb)                                     # host entry point, same inputs
            assert np.array_equal(again["x"], out["x"])                         # deterministic at full size, device == host entry
            ok = out["status"] == 1
            assert ok.mean() > 0.8
            X = out["x"][ok]; NS = p.N + 1
            st = X[:, : 8 * NS].reshape(-1, NS, 8); u = X[:, 8 * NS:].reshape(-1, p.N, 5)
            scale = 1e-2 * (1 + np.abs(st[:, :, 0:3]).max())
            assert np.abs(st[:, :-1, 0:3] + ts * st[:, :-1, 3:6] + h * u[:, :, 0:3] - st[:, 1:, 0:3]).max() < scale
            assert np.abs(st[:, 0, 0:6] - np.concatenate([mb.x0[ok, 0:3], mb.x0[ok, 3:6]], axis=1)).max() < scale
            assert np.abs(u[:, :, 0:3]).max() <= p.max_acc + 0.2                  # input box to the solver tolerance
            assert (out["pri_res"][ok] < 1e-3 + 1e-3 * 1e3).all()                # solved => primal residual within eps_abs + eps_rel*|Ax|
        # the chosen plan of every scenario is one of its six candidates
        cand = ds.buf["cand"].cpu().numpy(); xs = ds.buf["x"].cpu().numpy(); plan = ds.plan.cpu().numpy()
        assert np.array_equal(plan, xs[cand[np.arange(S), b]])
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_loop_100_steps_warm_started():
    """BASELINE.json configs[2]'s loop length: 100 control steps, every candidate QP warm-started from the plan chosen one
    step earlier.  64 scenarios x 6 candidates per step; every 10th step is re-solved by the oracle on the identical inputs;
    the scenarios must keep advancing along the reference line."""
    from intent_mpc_b200 import engine
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    eng = engine.Engine(0)
    try:
        host = receding.IntentSweep(S=64, D=4, seed0=77)
        ds = DeviceIntentSweep(eng, host)
        x_start = host.pos[:, 0].copy()
        total_q = 0
        for step in range(100):
            ds.step()
            st = ds.buf["status"][: (64 if step == 0 else 384)].cpu().numpy()
            assert np.isin(st, [1, 2, -2]).all()
            total_q += len(st)
            if step % 10 == 9:
                for mb, out in ds.batches_host():
                    _check_against_oracle(mb, out, f"step {step}")
        assert total_q == 64 + 99 * 384
        pos = ds.pos.cpu().numpy()
        assert np.isfinite(ds.plan.cpu().numpy()).all()
        assert (pos[:, 0] >= x_start - 1e-6).all() and (pos[:, 0] - x_start).mean() > 5.0     # >= 5 m of progress in 10 s
        # warm start pays: a cold solve of the last step's QPs needs more iterations than the warm one did
        mb, out = ds.batches_host()[0]
        cold = MpcBatch(mb.params, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.obs_dyn, mb.lin_pt, np.zeros_like(mb.warm_x))
        assert int(eng.solve_mpc_batch(cold)["iter"].sum()) > int(out["iter"].sum())
    finally:
        eng.close()
