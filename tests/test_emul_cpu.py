"""CPU tier: the kernel SOURCE (intent-mpc_b200/csrc/mpcqp_core.cuh) compiled for the host (tests/emul) must
reproduce the reference solver's golden vectors: identical status / iterations / rho updates, x and objective
within 1e-5 relative (BASELINE.json north_star tolerance).  This checks kernel logic without a GPU; the
shipped library has no host solve path."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import cases
from tests.helpers import rel_inf
from tests.emul import binding as EM

GOLD = os.path.join(os.path.dirname(__file__), "golden", "osqp_ref_golden.npz")
TOL = 1e-5


@pytest.mark.parametrize("name", ["snapshot", "static4", "static0", "static8", "h60", "stress"])
def test_emulated_kernel_matches_reference_golden(name):
    g = np.load(GOLD)
    mb = cases()[name]
    e = EM.solve(mb, want_y=False)
    assert (e["status"] == g[name + "_status"]).all()
    assert (e["iter"] == g[name + "_iter"]).all()
    assert (e["rho_updates"] == g[name + "_rho_updates"]).all()
    assert rel_inf(e["x"], g[name + "_x"]).max() < TOL
    assert np.abs((e["obj"] - g[name + "_obj"]) / g[name + "_obj"]).max() < TOL


@pytest.mark.parametrize("name", ["snapshot", "static4", "static0", "static8"])
def test_emulated_pcr_linear_algebra_matches_reference_golden(name):
    """The CTA kernel's linear algebra — leaf elimination + block parallel cyclic reduction (pcr_factor, with the
    plain-loop pcr_solve_ref standing in for the three axis warps) — inside the same ADMM iteration."""
    g = np.load(GOLD)
    e = EM.solve(cases()[name], want_y=False, linsys=1)
    assert (e["status"] == g[name + "_status"]).all()
    assert (e["iter"] == g[name + "_iter"]).all()
    assert (e["rho_updates"] == g[name + "_rho_updates"]).all()
    assert rel_inf(e["x"], g[name + "_x"]).max() < TOL
    assert np.abs((e["obj"] - g[name + "_obj"]) / g[name + "_obj"]).max() < TOL
