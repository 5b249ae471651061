"""Generate tests/golden/fov_ref_golden.npz from the REFERENCE's own solver binary (oracle/_ref/libosqp.so) on the
field-of-view case of tests/golden/fov_cases.py, with the determinism pins of SURVEY.md section 8(c).
Run in the build container: python tests/golden/make_golden_fov.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as OB  # noqa: E402
from tests.golden import fov_cases as FC  # noqa: E402

if __name__ == "__main__":
    qb = FC.case()
    r = OB.RefOsqp().solve_batch(qb, want_y=True, nthreads=4)
    out = {f"fov_{k}": r[k] for k in ("status", "iter", "rho_updates", "obj", "pri_res", "dua_res", "x", "y")}
    print("fov", qb.n, qb.m, r["status"], r["iter"], r["rho_updates"])
    # how much the half-space rows matter: multipliers on the FOV rows
    base = 16 * 30 + 5 * 29
    print("max |y| on the FOV rows:", np.abs(r["y"][:, base:base + 58]).max(axis=1))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fov_ref_golden.npz"), **out)
