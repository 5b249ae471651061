"""CPU tier: the oracle's C restatement of OSQP 0.6.2 (oracle/osqp_restated.c) pinned against the reference's
own solver binary (oracle/_ref/libosqp.so, driven by oracle/ref_driver.c) and against committed golden
vectors generated from that binary (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from intent_mpc_b200 import workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf

GOLD = os.path.join(os.path.dirname(__file__), "golden", "osqp_ref_golden.npz")


def _cases():
    return {"snapshot": W.snapshot(), "static4": W.static_batch(24, num_obs=4), "static0": W.static_batch(8, num_obs=0)}


@pytest.mark.skipif(not OB.RefOsqp.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["snapshot", "static4", "static0"])
def test_port_matches_reference_binary(name):
    qb = to_qp_batch(_cases()[name])
    r = OB.RefOsqp().solve_batch(qb)
    p = OB.PortOsqp().solve_batch(qb)
    assert (r["status"] == p["status"]).all()
    assert (r["iter"] == p["iter"]).all()
    assert (r["rho_updates"] == p["rho_updates"]).all()
    assert rel_inf(p["x"], r["x"]).max() < 1e-6
    assert np.abs((p["obj"] - r["obj"]) / r["obj"]).max() < 1e-8


def test_port_matches_golden():
    g = np.load(GOLD)
    for name, mb in _cases().items():
        p = OB.PortOsqp().solve_batch(to_qp_batch(mb))
        assert (p["status"] == g[name + "_status"]).all(), name
        assert (p["iter"] == g[name + "_iter"]).all(), name
        assert (p["rho_updates"] == g[name + "_rho_updates"]).all(), name
        assert rel_inf(p["x"], g[name + "_x"]).max() < 1e-6, name
        assert np.abs((p["obj"] - g[name + "_obj"]) / g[name + "_obj"]).max() < 1e-8, name


def test_reference_behaviours_probed_in_survey():
    """x0 outside the z box -> status -2 after 4000 iterations, exit flag 0; max_iter=30 -> 2."""
    mb = W.static_batch(2, num_obs=0)
    mb.x0[0, 2] = 6.0
    mb.warm_x[:] = 0
    p = OB.PortOsqp().solve_batch(to_qp_batch(mb))
    assert p["status"][0] == -2 and p["iter"][0] == 4000 and p["exitflag"][0] == 0
    p2 = OB.PortOsqp().solve_batch(to_qp_batch(W.static_batch(2, num_obs=0, warm=False)), max_iter=30)
    assert set(p2["status"]) <= {2, -2, 1}
