"""BASELINE.json configs[4]: the Monte-Carlo sweep over the run_mpc_benchmark scenario grid (SURVEY.md §8d "Config 5").
Every instance has its own obstacle count and its own mix of dynamic / static rows, so the batched entry point is
driven with per-instance obs_dyn flags (mpcqp_engine_obs_dyn_per_instance) and one call per (limits, obstacle count)
group.  CPU tier: the generator.  GPU tier: parity with the oracle on a subset, properties at a larger size."""
import numpy as np
import pytest

from intent_mpc_b200 import workloads as W
from intent_mpc_b200 import sharding

TOL = 1e-5   # BASELINE.json north_star: primal solution and objective within 1e-5 relative in FP64


def test_sweep_generator_is_index_addressable():
    """Any sub-range regenerates exactly the same instances (multi-GPU shards build their own range)."""
    whole, meta = W.sweep_groups(0, 600)
    part, _ = W.sweep_groups(200, 450)
    ref = {}
    for idx, mb in whole:
        for j, i in enumerate(idx):
            ref[int(i)] = (mb.num_obs, mb.params.max_vel, mb.x0[j], mb.obs_c[j], mb.obs_dyn[j])
    seen = 0
    for idx, mb in part:
        for j, i in enumerate(idx):
            R, vm, x0, oc, od = ref[int(i)]
            assert R == mb.num_obs and vm == mb.params.max_vel
            assert np.array_equal(x0, mb.x0[j]) and np.array_equal(oc, mb.obs_c[j]) and np.array_equal(od, mb.obs_dyn[j])
            seen += 1
    assert seen == 250
    assert meta["instances"] == 600 and meta["cap"] == W.SWEEP_CAP
    assert meta["obstacle_rows_hist"].sum() == 600


def test_sweep_rows_follow_update_obstacle_param():
    """Dynamic rows first, static after; flags carry the isDyamic quirk (mpcPlanner.cpp:1194): the first min(S, D) dynamic
    rows are flagged static; static rows have yaw, dynamic rows move and have yaw 0."""
    groups, _ = W.sweep_groups(0, 400)
    checked = 0
    for idx, mb in groups:
        p = mb.params
        moving = np.abs(mb.obs_c[:, -1] - mb.obs_c[:, 0]).max(axis=2) > 0            # [B,R]
        dyn_semi = np.isclose(mb.obs_semi[:, 0, :, 2], (0.8 + 0.3) / 2 + p.dynamic_safety_dist)
        for b in range(mb.B):
            D = int(dyn_semi[b].sum()); S = mb.num_obs - D
            assert dyn_semi[b, :D].all() and not dyn_semi[b, D:].any()
            assert not moving[b, D:].any() and (mb.obs_yaw[b, 0, :D] == 0).all()
            want = np.zeros(mb.num_obs, dtype=np.int32); want[min(S, D):D] = 1
            assert np.array_equal(mb.obs_dyn[b, 0], want) and np.array_equal(mb.obs_dyn[b, -1], want)
            d0 = np.linalg.norm(mb.obs_c[b, 0] - mb.x0[b, None, 0:3], axis=1)
            assert (d0 <= W.SWEEP_RADIUS + 1e-9).all() and (d0 >= W.SWEEP_CLEARANCE - 1e-9).all()
            assert (np.diff(d0[:D]) >= 0).all() and (np.diff(d0[D:]) >= 0).all()     # nearest first within each kind
            checked += 1
    assert checked == 400


def test_sweep_shards_partition_the_index_range():
    B, world = 1000, 4
    got = []
    for rank in range(world):
        lo, hi = sharding.shard_bounds(B, world)[rank]
        groups, _ = W.sweep_groups(lo, hi)
        got += [int(i) for idx, _ in groups for i in idx]
    assert sorted(got) == list(range(B))


@pytest.mark.gpu
def test_sweep_gpu_matches_oracle():
    from intent_mpc_b200 import engine
    from oracle import bindings as OB
    from tests.helpers import oracle_solve, rel_inf
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    eng = engine.Engine(0)
    try:
        groups, meta = W.sweep_groups(0, 384)
        n = 0
        for idx, mb in groups:
            out = eng.solve_mpc_batch(mb)
            ref = oracle_solve(orc, mb)
            assert (out["status"] == ref["status"]).all(), (mb.num_obs, out["status"], ref["status"])
            assert (out["iter"] == ref["iter"]).all(), (mb.num_obs, out["iter"], ref["iter"])
            assert (out["rho_updates"] == ref["rho_updates"]).all()
            assert rel_inf(out["x"], ref["x"]).max() < TOL
            assert np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL
            n += mb.B
        assert n == 384
        # the same instances packed three launches wide (per-instance obstacle counts, arrays padded to the cap)
        refs = {}
        for idx, mb in groups:
            r = oracle_solve(orc, mb)
            for j, i in enumerate(idx):
                refs[int(i)] = {k: r[k][j] for k in ("status", "iter", "rho_updates", "x", "obj")}
        batches, _ = W.sweep_batches(0, 384)
        assert len(batches) == 3
        for idx, mb in batches:
            out = eng.solve_mpc_batch(mb, want_y=True)
            assert eng.last_path == "cta"
            for j, i in enumerate(idx):
                r = refs[int(i)]
                assert out["status"][j] == r["status"] and out["iter"][j] == r["iter"] and out["rho_updates"][j] == r["rho_updates"]
                assert rel_inf(out["x"][j], r["x"]) < TOL and abs((out["obj"][j] - r["obj"]) / r["obj"]) < TOL
        # ... and as ONE launch with per-instance velocity / acceleration limits
        (idx, mb), = W.sweep_batches(0, 384, one_launch=True)[0]
        assert mb.B == 384 and mb.limits.shape == (384, 2)
        out = eng.solve_mpc_batch(mb)
        for j, i in enumerate(idx):
            r = refs[int(i)]
            assert out["status"][j] == r["status"] and out["iter"][j] == r["iter"] and out["rho_updates"][j] == r["rho_updates"]
            assert rel_inf(out["x"][j], r["x"]) < TOL and abs((out["obj"][j] - r["obj"]) / r["obj"]) < TOL
        # per-instance flags and a shared pattern are the same thing when the patterns agree
        mb = W.static_batch(32, num_obs=4, seed0=300)
        a = eng.solve_mpc_batch(mb)
        mb3 = W.MpcBatch(mb.params, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw,
                         np.ascontiguousarray(np.broadcast_to(mb.obs_dyn[None], (mb.B,) + mb.obs_dyn.shape)), mb.lin_pt, mb.warm_x)
        b = eng.solve_mpc_batch(mb3)
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["iter"], b["iter"])
    finally:
        eng.close()


@pytest.mark.gpu
def test_sweep_properties_at_larger_size():
    """8,192 instances: every instance comes back with a status OSQP can return for this problem class, solved ones obey
    the dynamics to solver tolerance and their limits, repeated solves are bit-identical."""
    from intent_mpc_b200 import engine
    eng = engine.Engine(0)
    try:
        groups, meta = W.sweep_batches(0, 8192)
        total = 0; solved = 0
        for idx, mb in groups:
            a = eng.solve_mpc_batch(mb)
            b = eng.solve_mpc_batch(mb)
            assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["iter"], b["iter"])
            assert np.isin(a["status"], [1, 2, -2]).all()
            assert ((a["iter"] % 25 == 0) & (a["iter"] > 0) & (a["iter"] <= 4000)).all()
            p = mb.params; N, NS = p.N, p.N + 1
            ok = a["status"] == 1
            total += mb.B; solved += int(ok.sum())
            if not ok.any():
                continue
            X = a["x"][ok]
            st = X[:, :8 * NS].reshape(-1, NS, 8); u = X[:, 8 * NS:].reshape(-1, N, 5)
            ts = float(np.float32(p.ts)); h = float(np.float32(0.5 * p.ts ** 2))
            pos_gap = np.abs(st[:, :-1, 0:3] + ts * st[:, :-1, 3:6] + h * u[:, :, 0:3] - st[:, 1:, 0:3]).max(axis=(1, 2))
            assert (pos_gap <= a["pri_res"][ok] + 1e-9).all()
            # OSQP accepts |Ax - z|_inf <= eps_abs + eps_rel * max(|Ax|, |z|), and |Ax| is large here (obstacle gradients
            # times positions), so "solved" trajectories overshoot their limits by up to the reported primal residual
            pr = a["pri_res"][ok]
            assert (np.abs(u[:, :, 0:3]).max(axis=(1, 2)) <= p.max_acc + pr + 1e-9).all()
            assert (np.abs(st[:, :, 3:6]).max(axis=(1, 2)) <= p.max_vel + pr + 1e-9).all()
            dyn_gap = np.abs(st[:, :-1, 3:6] + ts * u[:, :, 0:3] - st[:, 1:, 3:6]).max(axis=(1, 2))
            assert (dyn_gap <= pr + 1e-9).all()
        assert total == 8192 and solved / total > 0.85
    finally:
        eng.close()
