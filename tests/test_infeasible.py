"""OSQP's infeasibility outcomes, pinned against the reference binary (BASELINE.json configs[3] "infeasibility detection"):
statuses -3 (primal infeasible), -4 (dual infeasible), 3 / 4 (their `inaccurate` forms, reached when max_iter cuts the run
between the 10x-relaxed and the exact certificate), constants.h:18-30; x and y filled with OSQP_NAN — the NUMBER
2143289344.0, constants.h:95-97 — and obj = +-OSQP_INFTY.  Golden vectors: tests/golden/infeasible_ref_golden.npz, produced by
the reference's own libosqp.so (tests/golden/make_golden_infeasible.py) on tests/golden/infeasible_cases.py: tiny LP/QPs,
random QPs with an empty feasible set / an unbounded direction, and mpcPlanner QPs of the stress set whose infinite bounds
are written as 1e30 (with the reference's IEEE inf the certificates evaluate inf * 0 = NaN and never fire).

Bar: identical status, iteration count and rho updates on every instance; NaN fill bit-exact; where a solution exists, x and
objective within 1e-5 relative."""
import dataclasses
import os

import numpy as np
import pytest

from tests.golden import infeasible_cases as IC
from tests.helpers import rel_inf

GOLD = os.path.join(os.path.dirname(__file__), "golden", "infeasible_ref_golden.npz")
OSQP_NAN = 2143289344.0
TOL = 1e-5
EXPECT = {"pinf2": -3, "pinf2_cut20": 3, "pinf2_cut24": -3, "pinf2_cut10": -2, "dinf2": -4, "dinf2_ieee": -4, "dinf2_cut7": 4,
          "dinf2_tight": 4, "dinf2_cut5": -2, "pinf_rand": -3, "pinf_rand_b": -3, "dinf_rand": -4, "dinf_rand_b": -4}
SMALL = list(EXPECT)


def _check(name, got, g):
    st = g[name + "_status"]
    assert (np.asarray(got["status"]) == st).all(), (name, got["status"], st)
    assert (np.asarray(got["iter"]) == g[name + "_iter"]).all(), (name, got["iter"], g[name + "_iter"])
    assert (np.asarray(got["rho_updates"]) == g[name + "_rho_updates"]).all(), name
    for b in range(len(st)):
        if st[b] in (-3, 3, -4, 4):
            assert (got["x"][b] == OSQP_NAN).all() and (got["y"][b] == OSQP_NAN).all(), name            # store_solution's fill
            assert got["obj"][b] == g[name + "_obj"][b] and abs(got["obj"][b]) == 1e30, name               # +-OSQP_INFTY
        else:
            assert rel_inf(got["x"][b][None], g[name + "_x"][b][None]).max() < TOL, (name, b)
            assert abs((got["obj"][b] - g[name + "_obj"][b]) / g[name + "_obj"][b]) < TOL, (name, b)


def test_golden_holds_every_status():
    g = np.load(GOLD)
    for name, want in EXPECT.items():
        assert g[name + "_status"].tolist() == [want], name
        if want in (-3, 3, -4, 4):
            assert (g[name + "_x"] == OSQP_NAN).all() and (g[name + "_y"] == OSQP_NAN).all()
    for name in ("mpc_finite_h30", "mpc_finite_h60"):
        assert set(g[name + "_status"].tolist()) == {1, -3}


@pytest.mark.parametrize("name", SMALL + ["mpc_finite_h30"])
def test_oracle_port_matches_reference_golden(name):
    from oracle import bindings as OB
    qb, kw = IC.cases()[name]
    _check(name, OB.PortOsqp().solve_batch(qb, want_y=True, nthreads=os.cpu_count() or 1, **kw), np.load(GOLD))


@pytest.mark.parametrize("kernel", ["dense", "band"])
@pytest.mark.parametrize("name", SMALL)
def test_emulated_generic_kernels_match_reference_golden(name, kernel):
    """Kernel SOURCE of the generic paths (csrc/mpcqp_dense.cuh, csrc/mpcqp_band.cuh) compiled for the host: logic check without a GPU."""
    from tests.emul import binding as EM
    qb, kw = IC.cases()[name]
    if kernel == "band" and "rand" in name:
        with pytest.raises(RuntimeError, match="not eligible"):      # unstructured random patterns: half-bandwidth > 31, dense kernel
            EM.solve_band(qb, **kw)
        return
    _check(name, (EM.solve_dense if kernel == "dense" else EM.solve_band)(qb, **kw), np.load(GOLD))


@pytest.mark.parametrize("horizon,linsys", [(30, 0), (30, 1), (60, 0)])
def test_emulated_stage_kernel_declares_primal_infeasibility_like_osqp(horizon, linsys):
    """Kernel SOURCE of the stage-structured path (csrc/mpcqp_core.cuh) compiled for the host, with the bounds the
    mpc_finite_* cases carry (1e30 for every infinite bound): linsys 0 = block LDL' chain, 1 = the CTA path's PCR algebra."""
    from intent_mpc_b200 import workloads as W
    from tests.emul import binding as EM
    name = f"mpc_finite_h{horizon}"
    B = IC.cases()[name][0].q.shape[0]
    got = EM.solve(W.stress_batch(B, horizon=horizon), want_y=True, linsys=linsys, finite_inf=1e30)
    _check(name, got, np.load(GOLD))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_single_problem_abi_reports_infeasibility_like_osqp(name):
    """mpcqp_setup / mpcqp_solve / mpcqp_get_info / mpcqp_get_solution (what OsqpEigen::Solver drives)."""
    from intent_mpc_b200 import engine as E
    eng = E.Engine(0)
    try:
        qb, kw = IC.cases()[name]
        pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[0], qb.q[0], qb.A_colptr, qb.A_rowidx, qb.A_val[0], qb.l[0], qb.u[0],
                       settings=E.default_settings(**kw))
        r = pr.solve()
        assert eng.last_path in ("band", "dense")            # generic kernels: the sparse one for these small patterns
        got = {k: np.asarray([r[k]]) for k in ("status", "iter", "rho_updates", "obj")}
        got["x"] = r["x"][None]; got["y"] = r["y"][None]
        _check(name, got, np.load(GOLD))
        # a second solve of the same object starts cold after an infeasible outcome (store_solution cold-starts) and repeats it
        if r["status"] in (-3, 3, -4, 4):
            r2 = pr.solve()
            assert r2["status"] == r["status"] and r2["iter"] == r["iter"]
        pr.close()
    finally:
        eng.close()


@pytest.mark.gpu
def test_batched_generic_entry_reports_infeasibility_like_osqp():
    """mpcqp_solve_qp_batch_host: feasible and infeasible problems of one pattern in the same launch."""
    from intent_mpc_b200 import engine as E
    eng = E.Engine(0)
    try:
        g = np.load(GOLD)
        for name in ("pinf_rand", "dinf_rand_b"):
            qb, kw = IC.cases()[name]
            # the same problem twice plus a repaired copy (the contradicting / missing bound fixed): statuses differ per slot
            fixed = dataclasses.replace(qb, l=qb.l.copy(), u=qb.u.copy())
            if name.startswith("pinf"):
                fixed.l[0, -1] = qb.l[0, 0]; fixed.u[0, -1] = qb.u[0, 0]
            else:
                fixed.u[0, 0] = 3.0
            cat = lambda k: np.concatenate([getattr(qb, k), getattr(fixed, k), getattr(qb, k)])
            three = dataclasses.replace(qb, P_val=cat("P_val"), q=cat("q"), A_val=cat("A_val"), l=cat("l"), u=cat("u"), warm_x=cat("warm_x"))
            out = E.solve_qp_batch(eng, three, settings=E.default_settings(**kw))
            want = int(g[name + "_status"][0])
            assert out["status"].tolist() == [want, 1, want], (name, out["status"])
            assert out["iter"][0] == out["iter"][2] == g[name + "_iter"][0]
            assert (out["x"][0] == OSQP_NAN).all() and (out["x"][2] == OSQP_NAN).all() and np.isfinite(out["x"][1]).all() and (np.abs(out["x"][1]) < 1e6).all()
            from oracle import bindings as OB
            orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
            ref = orc.solve_batch(three, want_y=False, **kw)
            assert (ref["status"] == out["status"]).all() and (ref["iter"] == out["iter"]).all()
            assert rel_inf(out["x"][1][None], ref["x"][1][None]).max() < TOL
    finally:
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,path", [("mpc_finite_h30", "cta"), ("mpc_finite_h60", None)])
def test_structured_kernels_declare_primal_infeasibility_like_osqp(name, path):
    """mpcPlanner QPs with finite (1e30) 'infinite' bounds through mpcqp_setup: they keep the planner's structure, run on the
    stage kernels, and the instances whose start lies outside the box end PRIMAL INFEASIBLE (-3) after exactly the
    reference's number of iterations, with the OSQP_NAN fill."""
    from intent_mpc_b200 import engine as E
    eng = E.Engine(0)
    try:
        g = np.load(GOLD)
        qb, kw = IC.cases()[name]
        B = qb.q.shape[0]
        got = dict(status=[], iter=[], rho_updates=[], obj=[], x=[], y=[])
        for b in range(B):
            pr = E.Problem(eng, qb.n, qb.m, qb.P_colptr, qb.P_rowidx, qb.P_val[b], qb.q[b], qb.A_colptr, qb.A_rowidx, qb.A_val[b], qb.l[b], qb.u[b])
            pr.warm_start(qb.warm_x[b], np.zeros(qb.m))
            r = pr.solve()
            assert eng.last_path not in ("dense", "band")         # still the planner's structure: a stage kernel
            if path:
                assert eng.last_path == path
            for k in got:
                got[k].append(r[k])
            pr.close()
        got = {k: np.asarray(v) for k, v in got.items()}
        assert (g[name + "_status"] == -3).sum() >= 3
        _check(name, got, g)
    finally:
        eng.close()
