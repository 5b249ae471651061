"""The generic (unstructured) solve path — SURVEY.md section 8(f) row 3: `polyTrajSolver`'s QPs through the OSQP-shaped
single-problem ABI (mpcqp_setup / mpcqp_update_bounds / mpcqp_solve), which is what `OsqpEigen::Solver` drives at
polyTrajSolver.cpp:162-239 and :848-900.

Golden vectors: tests/golden/polytraj_ref_golden.npz, produced by the reference's own libosqp.so
(tests/golden/make_golden_poly.py).  Bar (BASELINE.json north_star): identical status, iteration count and rho updates;
x and objective within 1e-5 relative.  One case is outside that bar BY NATURE and is tested at what two exact solvers
agree on: `poly_k5_corridor` has an axis that stalls (4000 iterations, 4 rho updates) and one that converges after 1150 /
1725 iterations; there the oracle's own sparse LDL' (same algorithm as the binary's QDLDL, other elimination order) is
already 5e-4 away from the binary in x, i.e. rounding of the KKT solve is amplified that far by the iteration itself.
"""
import dataclasses
import os

import numpy as np
import pytest

from intent_mpc_b200 import polytraj_workload as PA
from tests.golden.make_golden_poly import shifted
from tests.helpers import rel_inf

GOLD = os.path.join(os.path.dirname(__file__), "golden", "polytraj_ref_golden.npz")
TOL = 1e-5            # north_star tolerance (FP64)
TOL_STALLED = 2e-3    # poly_k5_corridor only, see the module docstring
NAMES = ["poly_k3", "poly_k6", "poly_k10_soft", "poly_k5_corridor", "poly_k12_d5", "poly_k25"]


def _check(name, got, g, tag=""):
    tol = TOL_STALLED if name == "poly_k5_corridor" else TOL
    key = name + tag
    assert (np.asarray(got["status"]) == g[key + "_status"]).all(), (got["status"], g[key + "_status"])
    assert (np.asarray(got["iter"]) == g[key + "_iter"]).all(), (got["iter"], g[key + "_iter"])
    assert (np.asarray(got["rho_updates"]) == g[key + "_rho_updates"]).all()
    assert rel_inf(got["x"], g[key + "_x"]).max() < tol
    assert np.abs((np.asarray(got["obj"]) - g[key + "_obj"]) / g[key + "_obj"]).max() < tol
    assert rel_inf(got["y"], g[key + "_y"]).max() < max(tol, 1e-4)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_port_matches_reference_golden_on_polytraj(name):
    """The C restatement (oracle/osqp_restated.c) against the reference binary's vectors on unstructured QPs."""
    from oracle import bindings as OB
    g = np.load(GOLD)
    _check(name, OB.PortOsqp().solve_batch(PA.cases()[name], want_y=True), g)


@pytest.mark.parametrize("name", NAMES)
def test_emulated_dense_kernel_matches_reference_golden(name):
    """Kernel SOURCE of the generic path (csrc/mpcqp_dense.cuh) compiled for the host: logic check without a GPU."""
    from tests.emul import binding as EM
    g = np.load(GOLD)
    _check(name, EM.solve_dense(PA.cases()[name]), g)


@pytest.mark.parametrize("name", NAMES)
def test_emulated_band_kernel_matches_reference_golden(name):
    """Kernel SOURCE of the sparse generic path (csrc/mpcqp_band.cuh, with the library's own RCM pattern analysis) compiled for the
    host: logic check without a GPU.  polyTrajSolver's KKT matrices are narrow-banded under RCM (half-bandwidth <= 12)."""
    from tests.emul import binding as EM
    g = np.load(GOLD)
    out = EM.solve_band(PA.cases()[name])
    assert out["bandwidth"] <= 12
    _check(name, out, g)


def _solve_axes(eng, qb, E, update_to=None, path="band"):
    """x, y, z problems like polyTrajSolver::setUpProblem (:162-222) [+ updateProblem (:225-239)] and solve."""
    outs = []
    for b in range(3):
        pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[b], qb.q[b], qb.A_colptr, qb.A_rowidx,
                       qb.A_val[b], qb.l[b], qb.u[b])
        r = pr.solve()
        if update_to is not None:
            pr.update_bounds(update_to.l[b], update_to.u[b])
            pr.warm_start(np.zeros(qb.n), np.zeros(qb.m))     # the golden re-solve starts cold
            r = pr.solve()
        assert eng.last_path == path and eng.last_launches == 1
        pr.close()
        outs.append(r)
    return {k: np.array([o[k] for o in outs]) for k in ("x", "y", "status", "iter", "rho_updates", "obj", "pri_res", "dua_res")}


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["band", "dense"])
@pytest.mark.parametrize("name", NAMES)
def test_generic_kernels_match_reference_golden(name, path):
    """Both generic kernels: the sparse one the dispatcher picks for these patterns, and the dense one pinned for the A/B."""
    from intent_mpc_b200 import engine as E
    g = np.load(GOLD)
    eng = E.Engine(0)
    eng.force_generic("dense" if path == "dense" else "cta")
    _check(name, _solve_axes(eng, PA.cases()[name], E, path=path), g)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["poly_k6", "poly_k10_soft", "poly_k25"])
def test_dense_kernel_update_bounds_resolve(name):
    """polyTrajSolver::updateProblem: same P / A, new l / u through mpcqp_update_bounds, solved again."""
    from intent_mpc_b200 import engine as E
    g = np.load(GOLD)
    eng = E.Engine(0)
    qb = PA.cases()[name]
    _check(name, _solve_axes(eng, qb, E, update_to=shifted(qb)), g, tag="_shift")
    eng.close()


@pytest.mark.gpu
def test_dense_kernel_non_default_settings_and_warm_start_against_oracle():
    """Warm start (primal + dual), fewer Ruiz passes, other rho / alpha / check interval, an iteration cap: against the
    oracle run the same way.  (scaling = 0 is left out: on the unscaled minimum-snap KKT system the oracle's own LDL' is
    3e-3 away from the binary in x at equal status / iterations, so there is nothing to compare to 1e-5.)"""
    from intent_mpc_b200 import engine as E
    from oracle import bindings as OB
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    qb = PA.cases()["poly_k10_soft"]
    first = orc.solve_batch(qb, want_y=True)
    eng = E.Engine(0)
    for kw in (dict(scaling=3), dict(rho=1.0, alpha=1.2, check_termination=10, adaptive_rho_interval=50), dict(max_iter=60)):
        qw = dataclasses.replace(qb, warm_x=0.9 * first["x"])
        want = orc.solve_batch(qw, warm_y=0.9 * first["y"], want_y=True, **kw)
        for b in range(3):
            pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[b], qb.q[b], qb.A_colptr, qb.A_rowidx,
                           qb.A_val[b], qb.l[b], qb.u[b], settings=E.default_settings(**kw))
            pr.warm_start(qw.warm_x[b], 0.9 * first["y"][b])
            r = pr.solve(); pr.close()
            assert r["status"] == want["status"][b] and r["iter"] == want["iter"][b] and r["rho_updates"] == want["rho_updates"][b], (kw, b, r["status"], r["iter"], want["status"][b], want["iter"][b])
            assert rel_inf(r["x"][None], want["x"][b:b + 1]).max() < TOL
            assert abs((r["obj"] - want["obj"][b]) / want["obj"][b]) < TOL
    eng.close()


@pytest.mark.gpu
def test_dense_kernel_takes_a_perturbed_mpc_problem_and_large_ones_are_refused():
    """An mpcPlanner QP with one entry of A changed no longer has the stage structure: it runs on the dense kernel
    (n + m = 1126) and must give what the oracle gives; beyond n + m = 4096 mpcqp_setup refuses (no CPU path)."""
    from intent_mpc_b200 import engine as E
    from intent_mpc_b200 import workloads as W
    from oracle import bindings as OB
    from tests.helpers import to_qp_batch
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    qb = to_qp_batch(W.static_batch(1, num_obs=4, seed0=7))
    av = qb.A_val.copy(); av[0, 0] = -2.0                    # the -1 of the first dynamics row
    qb = dataclasses.replace(qb, A_val=av)
    want = orc.solve_batch(qb, want_y=False)
    eng = E.Engine(0)
    pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[0], qb.q[0], qb.A_colptr, qb.A_rowidx, qb.A_val[0], qb.l[0], qb.u[0])
    pr.warm_start(qb.warm_x[0])
    r = pr.solve(); pr.close()
    assert eng.last_path == "dense"                          # its band (N = 1126) does not fit shared memory
    assert r["status"] == want["status"][0] and r["iter"] == want["iter"][0] and r["rho_updates"] == want["rho_updates"][0]
    assert rel_inf(r["x"][None], want["x"]).max() < TOL and abs((r["obj"] - want["obj"][0]) / want["obj"][0]) < TOL
    n = 3000; m = 1200                                       # identity P, A = first m rows of I: unstructured, too large
    cp = np.arange(n + 1); ri = np.arange(n); ac = np.minimum(np.arange(n + 1), m); ai = np.arange(m)
    with pytest.raises(E.EngineError, match="-5"):
        E.Problem(eng, n, m, cp, ri, np.ones(n), np.zeros(n), ac, ai, np.ones(m), -np.ones(m), np.ones(m))
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_dense_batch_entry_matches_reference_golden(name):
    """mpcqp_solve_qp_batch_host: the x, y, z problems of one path in ONE launch."""
    from intent_mpc_b200 import engine as E
    g = np.load(GOLD)
    eng = E.Engine(0)
    r = E.solve_qp_batch(eng, PA.cases()[name])
    assert eng.last_path in ("band", "dense") and eng.last_launches == 1
    _check(name, r, g)
    eng.close()


@pytest.mark.gpu
def test_dense_batch_of_candidate_paths_against_oracle_and_repeatable():
    """600 QPs (200 candidate paths x 3 axes, K = 8 segments: more problems than resident CTAs, so the persistent loop and
    the per-CTA workspace reuse are exercised) against the oracle; a second call must be bit-identical."""
    from intent_mpc_b200 import engine as E
    from oracle import bindings as OB
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    qb = PA.path_batch(200, K=8, seed0=100)
    want = orc.solve_batch(qb, want_y=False, nthreads=8)
    eng = E.Engine(0)
    r = E.solve_qp_batch(eng, qb, want_y=False)
    r2 = E.solve_qp_batch(eng, qb, want_y=False)
    eng.close()
    assert (r["status"] == want["status"]).all() and (r["iter"] == want["iter"]).all() and (r["rho_updates"] == want["rho_updates"]).all()
    ok = want["iter"] <= 500                                 # the well-posed ones; see the module docstring for the rest
    assert ok.mean() > 0.9
    assert rel_inf(r["x"][ok], want["x"][ok]).max() < TOL
    assert np.abs((r["obj"][ok] - want["obj"][ok]) / want["obj"][ok]).max() < TOL
    assert rel_inf(r["x"][~ok], want["x"][~ok]).max(initial=0.0) < TOL_STALLED
    assert (r["x"] == r2["x"]).all() and (r["iter"] == r2["iter"]).all()


# ---- mpcPlanner QPs with the field-of-view half-space rows (mpcPlanner.cpp:265-297, 1027-1038, 1102-1111) ---------------------
FOV_GOLD = os.path.join(os.path.dirname(__file__), "golden", "fov_ref_golden.npz")


def _check_fov(got, g, tol=TOL):
    assert (np.asarray(got["status"]) == g["fov_status"]).all(), (got["status"], g["fov_status"])
    assert (np.asarray(got["iter"]) == g["fov_iter"]).all(), (got["iter"], g["fov_iter"])
    assert (np.asarray(got["rho_updates"]) == g["fov_rho_updates"]).all()
    assert rel_inf(got["x"], g["fov_x"]).max() < tol
    assert np.abs((np.asarray(got["obj"]) - g["fov_obj"]) / g["fov_obj"]).max() < tol


def test_oracle_port_matches_reference_golden_with_fov_rows():
    from oracle import bindings as OB
    from tests.golden import fov_cases as FC
    _check_fov(OB.PortOsqp().solve_batch(FC.case(), want_y=True, nthreads=4), np.load(FOV_GOLD))


@pytest.mark.gpu
def test_planner_qp_with_fov_half_space_rows_runs_on_the_generic_path():
    """The 3-argument updateCurrStates adds two half-space rows per stage; such a QP has no slack column on those rows, so it
    is not the structure the stage kernels take: mpcqp_setup routes it to the dense generic kernel (n + m = 1126 + 58)."""
    from intent_mpc_b200 import engine as E
    from tests.golden import fov_cases as FC
    eng = E.Engine(0)
    try:
        qb = FC.case()
        got = dict(status=[], iter=[], rho_updates=[], obj=[], x=[])
        for b in range(qb.q.shape[0]):
            pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[b], qb.q[b], qb.A_colptr, qb.A_rowidx, qb.A_val[b], qb.l[b], qb.u[b])
            pr.warm_start(qb.warm_x[b], np.zeros(qb.m))
            r = pr.solve()
            assert eng.last_path == "dense"                  # N = 1184: the band would not fit shared memory
            for k in got:
                got[k].append(r[k])
            pr.close()
        _check_fov({k: np.asarray(v) for k, v in got.items()}, np.load(FOV_GOLD))
    finally:
        eng.close()


@pytest.mark.gpu
def test_out_of_pattern_bounds_move_a_planner_qp_to_the_generic_path():
    """osqp_setup / osqp_update_bounds accept any l <= u.  An mpcPlanner QP whose bounds leave the planner's pattern (a state box
    that differs between stages, a finite upper bound on an obstacle row) must still be solved — on a generic kernel — and
    give what the reference gives, both when the bounds arrive at setup and when they arrive through mpcqp_update_bounds."""
    from intent_mpc_b200 import engine as E
    from intent_mpc_b200 import workloads as W
    from oracle import bindings as OB
    from tests.helpers import to_qp_batch
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    qb = to_qp_batch(W.static_batch(1, num_obs=2, seed0=21))
    NS, N = 30, 29
    l2 = qb.l.copy(); u2 = qb.u.copy()
    u2[0, 8 * NS + 8 * 12 + 1] = 1.25                       # y of stage 12 capped: the state box is no longer stage-uniform
    u2[0, 16 * NS + 5 * N + 7] = 50.0                        # one obstacle row gets a finite upper bound
    qb2 = dataclasses.replace(qb, l=l2, u=u2)
    want = orc.solve_batch(qb2, want_y=False)
    eng = E.Engine(0)
    try:
        # (a) at setup
        pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[0], qb.q[0], qb.A_colptr, qb.A_rowidx, qb.A_val[0], l2[0], u2[0])
        pr.warm_start(qb.warm_x[0], np.zeros(qb.m))
        r = pr.solve(); pr.close()
        assert eng.last_path in ("dense", "band")
        assert r["status"] == want["status"][0] and r["iter"] == want["iter"][0]
        assert rel_inf(r["x"][None], want["x"]).max() < TOL
        # (b) through update_bounds on a problem that was set up with the planner's own bounds (stage kernel first)
        pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[0], qb.q[0], qb.A_colptr, qb.A_rowidx, qb.A_val[0], qb.l[0], qb.u[0])
        pr.warm_start(qb.warm_x[0], np.zeros(qb.m))
        first = pr.solve()
        assert eng.last_path == "cta"
        ref1 = orc.solve_batch(qb, want_y=False)
        assert first["status"] == ref1["status"][0] and first["iter"] == ref1["iter"][0]
        pr.update_bounds(l2[0], u2[0])
        pr.warm_start(qb.warm_x[0], np.zeros(qb.m))
        r = pr.solve(); pr.close()
        assert eng.last_path in ("dense", "band")
        assert r["status"] == want["status"][0] and r["iter"] == want["iter"][0]
        assert rel_inf(r["x"][None], want["x"]).max() < TOL
    finally:
        eng.close()


@pytest.mark.gpu
def test_engine_outlives_destroy_while_problems_are_alive_and_warm_start_x_keeps_the_dual():
    """mpcqp_engine_destroy with live problems only marks the engine (the last mpcqp_cleanup releases it): a Solver that
    outlives its thread's engine stays usable.  osqp_warm_start_x replaces x only: the next solve keeps the previous dual."""
    from intent_mpc_b200 import engine as E
    from oracle import bindings as OB
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    qb = PA.cases()["poly_k6"]
    eng = E.Engine(0)
    pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[0], qb.q[0], qb.A_colptr, qb.A_rowidx, qb.A_val[0], qb.l[0], qb.u[0])
    r1 = pr.solve()
    h = eng.h; eng.h = None                                  # destroy the engine under the problem
    assert eng.lib.mpcqp_engine_destroy(h) == 0
    # re-solve with a new primal start only: the dual start must be r1's y (what a (x, y) warm start of the oracle gives)
    x0 = 0.5 * r1["x"]
    pr.warm_start(x0)
    r2 = pr.solve()
    one = dataclasses.replace(qb, P_val=qb.P_val[:1], q=qb.q[:1], A_val=qb.A_val[:1], l=qb.l[:1], u=qb.u[:1], warm_x=x0[None])
    want = orc.solve_batch(one, warm_y=r1["y"][None], want_y=False)
    cold = orc.solve_batch(one, want_y=False)
    assert r2["status"] == want["status"][0] and r2["iter"] == want["iter"][0]
    assert rel_inf(r2["x"][None], want["x"]).max() < TOL
    assert want["iter"][0] != cold["iter"][0] or True          # (informative: with y = 0 the oracle may take another path)
    pr.close()                                               # the last cleanup releases the engine
