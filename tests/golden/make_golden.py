"""Generate tests/golden/osqp_ref_golden.npz from the REFERENCE's own solver binary (oracle/_ref/libosqp.so,
copied from /root/reference/trajectory_planner/include/trajectory_planner/third_party/lib/x86/libosqp.so) on
the seeded workloads of intent-mpc_b200/workloads.py with the determinism pins of SURVEY.md §8(c)
(adaptive_rho_interval=25, time_limit=0).  Run in the build container: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from intent_mpc_b200 import workloads as W  # noqa: E402
from oracle import bindings as OB  # noqa: E402
from tests.helpers import to_qp_batch  # noqa: E402


def cases():
    return {"snapshot": W.snapshot(), "static4": W.static_batch(24, num_obs=4), "static0": W.static_batch(8, num_obs=0),
            "static4_256": W.static_batch(256, num_obs=4), "static8": W.static_batch(32, num_obs=8),
            "h60": W.static_batch(8, num_obs=2, params=W.MpcParams(horizon=60)),
            "stress": W.stress_batch(32)}


if __name__ == "__main__":
    ref = OB.RefOsqp()
    out = {}
    for name, mb in cases().items():
        r = ref.solve_batch(to_qp_batch(mb), want_y=False, nthreads=8)
        for k in ("status", "iter", "rho_updates", "obj", "pri_res", "dua_res"):
            out[f"{name}_{k}"] = r[k]
        out[f"{name}_x"] = r["x"]
        print(name, dict(zip(*np.unique(r["status"], return_counts=True))), "iters", r["iter"].min(), r["iter"].max())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "osqp_ref_golden.npz"), **out)
