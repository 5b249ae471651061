"""Generate tests/golden/polytraj_ref_golden.npz: the REFERENCE's own solver binary (oracle/_ref/libosqp.so) on the
polyTrajSolver-shaped QPs of intent-mpc_b200/polytraj_workload.py (cases()), with the determinism pins of SURVEY.md section 8(c)
(adaptive_rho_interval=25, time_limit=0), plus each case re-solved with shifted bounds (polyTrajSolver::updateProblem,
polyTrajSolver.cpp:225-239).  Run in the build container: python tests/golden/make_golden_poly.py"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as OB  # noqa: E402
from intent_mpc_b200 import polytraj_workload as PA  # noqa: E402


def shifted(qb):
    """The bounds of a later call: the whole path moved by (0.4, -0.3, 0.1) - only l / u change, as in updateProblem."""
    d = np.array([0.4, -0.3, 0.1])[:, None]
    pos_rows = (np.abs(qb.l) + np.abs(qb.u)) > 0
    return dataclasses.replace(qb, l=np.where(pos_rows, qb.l + d, qb.l), u=np.where(pos_rows, qb.u + d, qb.u))


if __name__ == "__main__":
    ref = OB.RefOsqp()
    out = {}
    for name, qb in PA.cases().items():
        for tag, q in (("", qb), ("_shift", shifted(qb))):
            r = ref.solve_batch(q, want_y=True)
            for k in ("status", "iter", "rho_updates", "obj", "pri_res", "dua_res", "x", "y"):
                out[f"{name}{tag}_{k}"] = r[k]
            print(name + tag, qb.n, qb.m, r["status"], r["iter"], r["rho_updates"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "polytraj_ref_golden.npz"), **out)
