"""GPU tier: the CUDA path, called through the C ABI, against (a) golden vectors from the reference's own
solver binary, (b) the oracle run on the same seeded inputs.  Bar (BASELINE.json north_star): identical solver
status, primal solution and objective within 1e-5 relative in FP64."""
import os

import numpy as np
import pytest

from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.golden.make_golden import cases
from tests.helpers import to_qp_batch, rel_inf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "osqp_ref_golden.npz")
TOL = 1e-5


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("name", ["snapshot", "static4", "static0", "static4_256", "static8", "h60", "stress"])
def test_gpu_matches_reference_golden(eng, name):
    g = np.load(GOLD)
    out = eng.solve_mpc_batch(cases()[name])
    assert (out["status"] == g[name + "_status"]).all()
    assert (out["iter"] == g[name + "_iter"]).all()
    assert (out["rho_updates"] == g[name + "_rho_updates"]).all()
    assert rel_inf(out["x"], g[name + "_x"]).max() < TOL
    assert np.abs((out["obj"] - g[name + "_obj"]) / g[name + "_obj"]).max() < TOL


@pytest.mark.parametrize("path", ["generic", "fast", "cta_plain"])
@pytest.mark.parametrize("name", ["snapshot", "static4", "static0", "static8"])
def test_other_kernels_match_reference_golden(eng, name, path):
    """The one-warp kernels — generic shared-memory (used for horizons / obstacle counts without a compiled
    specialisation) and register-resident — and the CTA kernel without its assistant warps on the same golden cases;
    the default dispatch above runs them on the CTA kernel (with assistants where a CTA has an SM to itself)."""
    g = np.load(GOLD)
    eng.force_generic(path)
    try:
        out = eng.solve_mpc_batch(cases()[name])
        assert eng.last_path == ("cta" if path == "cta_plain" else path)
    finally:
        eng.force_generic(False)
    assert (out["status"] == g[name + "_status"]).all()
    assert (out["iter"] == g[name + "_iter"]).all()
    assert (out["rho_updates"] == g[name + "_rho_updates"]).all()
    assert rel_inf(out["x"], g[name + "_x"]).max() < TOL


def test_dispatch_uses_fast_path_for_default_shape(eng):
    eng.solve_mpc_batch(W.static_batch(4, num_obs=4))
    assert eng.last_path == "cta"
    eng.solve_mpc_batch(W.static_batch(4, num_obs=2, params=W.MpcParams(horizon=60)))
    assert eng.last_path == "generic"


def _oracle():
    return OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()


@pytest.mark.parametrize("num_obs,B,seed0", [(4, 1024, 0), (0, 128, 5000), (16, 64, 7000), (1, 64, 9000), (32, 24, 11000), (40, 8, 13000)])
def test_gpu_matches_oracle_on_seeded_batches(eng, num_obs, B, seed0):
    """configs[1] (B=1024, static obstacles only) at full size, plus an obstacle-count sweep."""
    mb = W.static_batch(B, num_obs=num_obs, seed0=seed0)
    out = eng.solve_mpc_batch(mb, want_y=True)
    ref = _oracle().solve_batch(to_qp_batch(mb), nthreads=os.cpu_count() or 1)
    assert (out["status"] == ref["status"]).all()
    same = out["iter"] == ref["iter"]
    assert same.all(), f"iteration mismatch on {np.where(~same)[0][:8]}"
    assert (out["rho_updates"] == ref["rho_updates"]).all()
    assert rel_inf(out["x"], ref["x"]).max() < TOL
    assert np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL
    # duals: same tolerance relative to the dual's own scale (not part of the north_star bar, reported)
    assert rel_inf(out["y"], ref["y"]).max() < 1e-4


def test_cold_start_and_settings(eng):
    """warm_x = NULL is OSQP's cold start; non-default settings reach the kernel (max_iter=30 -> inaccurate)."""
    mb = W.static_batch(32, num_obs=4, seed0=300, warm=False)
    o = _oracle()
    ref = o.solve_batch(to_qp_batch(mb), want_y=False)
    mb.warm_x = None
    out = eng.solve_mpc_batch(mb)
    assert (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all()
    assert rel_inf(out["x"], ref["x"]).max() < TOL
    mb2 = W.static_batch(32, num_obs=4, seed0=300)
    s = engine.default_settings(max_iter=30, eps_abs=1e-5, eps_rel=1e-5)
    out2 = eng.solve_mpc_batch(mb2, settings=s)
    ref2 = o.solve_batch(to_qp_batch(mb2), want_y=False, max_iter=30, eps_abs=1e-5, eps_rel=1e-5)
    assert (out2["status"] == ref2["status"]).all() and (out2["iter"] == ref2["iter"]).all()
    assert rel_inf(out2["x"], ref2["x"]).max() < TOL


def test_solution_properties_at_full_size(eng):
    """Size-independent checks at B = 4096: solved instances satisfy the dynamics equalities and box bounds
    to the solver tolerance, and solving is deterministic (bit-identical on a second run)."""
    mb = W.static_batch(4096, num_obs=4, seed0=20000)
    a = eng.solve_mpc_batch(mb)
    b = eng.solve_mpc_batch(mb)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["iter"], b["iter"])
    p = mb.params
    N, NS = p.N, p.N + 1
    ok = a["status"] == 1
    assert ok.mean() > 0.9
    X = a["x"][ok]
    st = X[:, :8 * NS].reshape(-1, NS, 8); u = X[:, 8 * NS:].reshape(-1, N, 5)
    ts = float(np.float32(p.ts)); h = float(np.float32(0.5 * p.ts ** 2))
    pos_pred = st[:, :-1, 0:3] + ts * st[:, :-1, 3:6] + h * u[:, :, 0:3]
    vel_pred = st[:, :-1, 3:6] + ts * u[:, :, 0:3]
    scale = 1e-3 * (1 + np.abs(st[:, :, 0:3]).max())
    assert np.abs(pos_pred - st[:, 1:, 0:3]).max() < 10 * scale
    assert np.abs(vel_pred - st[:, 1:, 3:6]).max() < 10 * scale
    assert np.abs(st[:, 0, 0:6] - mb.x0[ok]).max() < 10 * scale
    assert np.abs(u[:, :, 0:3]).max() <= p.max_acc + 0.1


def test_bitwise_determinism_under_cold_caches_and_scheduling(eng):
    """Results must not depend on timing or on the scheduling hints: the same batch solved repeatedly — cold L2
    (256 MiB flush on the engine's stream before each call), history hint on and off, and single-QP launches — is
    bit-identical.  (Guards the shared-memory hand-offs between the four warps of a CTA.)"""
    import torch
    mb = W.static_batch(512, num_obs=4, seed0=0)
    eng.use_history(False)
    ref = eng.solve_mpc_batch(mb)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", 0))
    eng.use_history(True)
    try:
        for rep in range(12):
            with torch.cuda.stream(stream):
                flush.zero_()
            out = eng.solve_mpc_batch(mb)
            assert np.array_equal(out["iter"], ref["iter"]) and np.array_equal(out["x"], ref["x"]), f"repeat {rep} differs"
    finally:
        eng.use_history(True)
    # migration of long-running instances to a follow-up launch (default on) against everything solved where it started
    for hist in (False, True):
        eng.use_history(hist); eng.use_migration(False)
        off = eng.solve_mpc_batch(mb)
        eng.use_migration(True)
        on = eng.solve_mpc_batch(mb)
        for k in ("x", "iter", "status", "rho_updates", "obj", "pri_res", "dua_res"):
            assert np.array_equal(off[k], ref[k]) and np.array_equal(on[k], ref[k]), (hist, k)
    eng.use_history(True)
    long_ones = np.argsort(ref["iter"])[-3:]
    for i in list(long_ones) + [0, 1]:
        one = eng.solve_mpc_batch(mb.slice(int(i), int(i) + 1))
        assert np.array_equal(one["x"][0], ref["x"][i])


@pytest.mark.parametrize("num_obs", [4, 5, 8])
def test_row_helper_blocks_match_plain_blocks_bitwise(eng, num_obs):
    """One-per-SM blocks with four or more obstacle rows per stage carry an eighth warp that runs the slack warp's rows from
    shared memory (Qp::helper_role; two rows at eight per stage).  Same arithmetic as a register row: iterates, duals, residuals
    and iteration counts must be bit-identical to the plain 4-warp blocks, across bursts, rho updates (rescaled rows) and
    the residual checks that read the helper's rows — and identical from call to call."""
    mb = W.static_batch(96, num_obs=num_obs, seed0=7000 + num_obs)
    eng.use_history(False)
    try:
        eng.force_generic("cta_plain")
        plain = eng.solve_mpc_batch(mb, want_y=True)
        eng.force_generic("cta")
        for rep in range(3):
            out = eng.solve_mpc_batch(mb, want_y=True)
            assert eng.last_path == "cta"
            for k in ("x", "y", "iter", "status", "rho_updates", "obj", "pri_res", "dua_res"):
                assert np.array_equal(out[k], plain[k]), (rep, k)
        assert plain["rho_updates"].max() >= 1 and plain["iter"].max() > 100      # the case exercises what it claims
    finally:
        eng.force_generic("cta"); eng.use_history(True)
    ref = _oracle().solve_batch(to_qp_batch(mb), want_y=False)
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["iter"], ref["iter"])
    assert rel_inf(out["x"], ref["x"]).max() < 1e-5


@pytest.mark.parametrize("horizon,num_obs,kw", [
    (20, 3, {}),                                       # the code default of mpcPlanner::initParam (mpcPlanner.cpp:19-173)
    (25, 2, {}),                                       # 8*horizon % 5 != 0: the R[i % numControls] rotation of castMPCToQPHessian (:945)
    (45, 1, {"velocity_weight": 2.5}),                 # non-zero velocity weights: P gains entries (:941-942)
    (30, 4, {"max_vel": 1.5, "max_acc": 1.5, "z_min": 1.0, "z_max": 3.0, "acceleration_weight": 1.0, "position_weight": 50.0}),
])
def test_shapes_weights_and_hessian_quirk(eng, horizon, num_obs, kw):
    """Horizons / weights other than the demo's, through whichever kernel the dispatch picks, against the oracle."""
    p = W.MpcParams(horizon=horizon, **kw)
    mb = W.static_batch(24, num_obs=num_obs, params=p, seed0=4000 + horizon)
    out = eng.solve_mpc_batch(mb)
    assert eng.last_path == ("cta" if horizon <= 30 else "generic")      # horizons 20, 25, 30 have CTA kernels
    ref = _oracle().solve_batch(to_qp_batch(mb), want_y=False)
    assert (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all()
    assert (out["rho_updates"] == ref["rho_updates"]).all()
    assert rel_inf(out["x"], ref["x"]).max() < TOL
    assert np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL


@pytest.mark.parametrize("horizon", [20, 25])
def test_other_horizons_on_the_cta_kernel_with_migration(eng, horizon):
    """Horizons 20 (the code default of mpcPlanner::initParam) and 25 run the 4-warp CTA / PCR kernel built for that stage count
    (chains of the composed final operator are shorter; no assistant variant): a batch large enough for the two-launch regime
    (hard list, two per SM, parked instances resumed one per SM) against the oracle and against the generic kernel."""
    p = W.MpcParams(horizon=horizon)
    mb = W.static_batch(400, num_obs=4, params=p, seed0=9000 + horizon)
    eng.use_history(False)
    try:
        out = eng.solve_mpc_batch(mb, want_y=True)
        assert eng.last_path == "cta"
        again = eng.solve_mpc_batch(mb, want_y=True)
        eng.force_generic("generic")
        gen = eng.solve_mpc_batch(mb)
        assert eng.last_path == "generic"
    finally:
        eng.force_generic("cta"); eng.use_history(True)
    for k in ("x", "y", "iter", "status", "obj"):
        assert np.array_equal(out[k], again[k]), k
    ref = _oracle().solve_batch(to_qp_batch(mb), want_y=True)
    assert (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all() and (out["rho_updates"] == ref["rho_updates"]).all()
    assert (gen["status"] == ref["status"]).all() and (gen["iter"] == ref["iter"]).all()
    assert rel_inf(out["x"], ref["x"]).max() < TOL and np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max() < TOL
    assert out["iter"].max() > 100                      # something was parked and resumed


def test_ragged_batches_and_mixed_rows(eng):
    """Edge cases of the batched entry point: a single QP, a batch that is not a multiple of anything, instances with
    zero obstacle rows inside a padded batch, mixed dynamic / static flags per instance — all against the oracle."""
    from tests.helpers import oracle_solve
    orc = _oracle()
    for B in (1, 3, 149):
        mb = W.static_batch(B, num_obs=4, seed0=6000 + B)
        out = eng.solve_mpc_batch(mb)
        ref = orc.solve_batch(to_qp_batch(mb), want_y=False)
        assert (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all()
        assert rel_inf(out["x"], ref["x"]).max() < TOL
    # per-instance obstacle counts 0..6 in one call (stride 6), per-instance dynamic/static flags
    rng = np.random.default_rng(5)
    base = W.static_batch(40, num_obs=6, seed0=7000)
    nobs = rng.integers(0, 7, size=40).astype(np.int32); nobs[:3] = (0, 6, 1)
    flags = (rng.uniform(size=(40, 1, 6)) < 0.5).astype(np.int32) * np.ones((1, base.params.N, 1), dtype=np.int32)
    mb = W.MpcBatch(base.params, base.x0, base.xref, base.obs_c, base.obs_semi, base.obs_yaw, np.ascontiguousarray(flags),
                    base.lin_pt, base.warm_x, nobs)
    out = eng.solve_mpc_batch(mb, want_y=True)
    for b in range(40):
        R = int(nobs[b])
        one = W.MpcBatch(base.params, base.x0[b:b + 1], base.xref[b:b + 1], base.obs_c[b:b + 1, :, :R], base.obs_semi[b:b + 1, :, :R],
                         base.obs_yaw[b:b + 1, :, :R], np.ascontiguousarray(flags[b, :, :R]), base.lin_pt[b:b + 1], base.warm_x[b:b + 1])
        ref = orc.solve_batch(to_qp_batch(one), want_y=True)
        assert out["status"][b] == ref["status"][0] and out["iter"][b] == ref["iter"][0], (b, R)
        assert rel_inf(out["x"][b], ref["x"][0]) < TOL
        m_b = base.params.m(R)
        assert np.abs(out["y"][b, :m_b] - ref["y"][0]).max() <= 1e-4 * max(1.0, np.abs(ref["y"][0]).max())


def test_migration_with_duals_and_other_settings(eng):
    """Instances parked after 300 iterations and resumed by the follow-up launch: duals requested, a check interval other
    than the default, a batch in the two-launch regime — bit-identical to the run without migration, and equal to the oracle."""
    st = engine.default_settings(max_iter=1500, check_termination=50)
    mb = W.static_batch(400, num_obs=4, seed0=12000)
    eng.use_history(False)
    try:
        eng.use_migration(False)
        off = eng.solve_mpc_batch(mb, settings=st, want_y=True)
        eng.use_migration(True)
        on = eng.solve_mpc_batch(mb, settings=st, want_y=True)
    finally:
        eng.use_history(True); eng.use_migration(True)
    assert (off["iter"] > 300).sum() >= 5                      # something was there to migrate
    for k in ("x", "y", "iter", "status", "rho_updates", "obj", "pri_res", "dua_res"):
        assert np.array_equal(off[k], on[k]), k
    ref = _oracle().solve_batch(to_qp_batch(mb), want_y=True, max_iter=1500, check_termination=50)
    assert (on["status"] == ref["status"]).all() and (on["iter"] == ref["iter"]).all()
    assert rel_inf(on["x"], ref["x"]).max() < TOL
    assert np.abs(on["y"] - ref["y"]).max() <= 1e-4 * max(1.0, np.abs(ref["y"]).max())


def test_per_instance_limits_on_the_register_row_kernels(eng):
    """mpcqp_engine_limits_per_instance with a compiled obstacle count (R = 4: CTA kernels with the rows in registers,
    two-launch regime incl. migration): every instance against the oracle solving it with its own max_vel / max_acc."""
    import dataclasses
    B = 300
    base = W.static_batch(B, num_obs=4, seed0=15000)
    lims = np.array([(1.5, 1.5), (3.0, 3.0), (5.0, 20.0)])[np.arange(B) % 3]
    base.x0[:, 3:6] *= (lims[:, 0:1] / 5.0)                       # keep the start inside each instance's velocity box
    warm_x, lin_pt = W._const_vel_plan(base.params, base.x0)
    mb = W.MpcBatch(base.params, base.x0, base.xref, base.obs_c, base.obs_semi, base.obs_yaw, base.obs_dyn, lin_pt, warm_x, None, lims)
    eng.use_history(False)
    try:
        out = eng.solve_mpc_batch(mb)
    finally:
        eng.use_history(True)
    assert eng.last_path == "cta"
    orc = _oracle()
    for g, (vm, am) in enumerate([(1.5, 1.5), (3.0, 3.0), (5.0, 20.0)]):
        sel = np.nonzero(np.arange(B) % 3 == g)[0]
        p = dataclasses.replace(base.params, max_vel=vm, max_acc=am)
        sub = W.MpcBatch(p, mb.x0[sel], mb.xref[sel], mb.obs_c[sel], mb.obs_semi[sel], mb.obs_yaw[sel], mb.obs_dyn, mb.lin_pt[sel], mb.warm_x[sel])
        ref = orc.solve_batch(to_qp_batch(sub), want_y=False)
        assert (out["status"][sel] == ref["status"]).all() and (out["iter"][sel] == ref["iter"]).all(), (vm, am)
        assert (out["rho_updates"][sel] == ref["rho_updates"]).all()
        assert rel_inf(out["x"][sel], ref["x"]).max() < TOL
