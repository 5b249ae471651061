"""Generate tests/golden/infeasible_ref_golden.npz from the REFERENCE's own solver binary (oracle/_ref/libosqp.so) on the
cases of tests/golden/infeasible_cases.py, with the determinism pins of SURVEY.md section 8(c).
Run in the build container: python tests/golden/make_golden_infeasible.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as OB  # noqa: E402
from tests.golden import infeasible_cases as IC  # noqa: E402

if __name__ == "__main__":
    ref = OB.RefOsqp()
    out = {}
    for name, (qb, kw) in IC.cases().items():
        r = ref.solve_batch(qb, want_y=True, **kw)
        for k in ("status", "iter", "rho_updates", "obj", "pri_res", "dua_res", "x", "y"):
            out[f"{name}_{k}"] = r[k]
        print(f"{name:16s} n={qb.n} m={qb.m} status={r['status'].tolist()} iter={r['iter'].tolist()} rho_updates={r['rho_updates'].tolist()}")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "infeasible_ref_golden.npz"), **out)
