"""Shared helpers for the tests: workload -> oracle CSC batch conversion."""
import dataclasses

import numpy as np

from intent_mpc_b200 import workloads as W
from oracle import mpc_assembly as MA


def to_qp_batch(mb):
    """Assemble the explicit CSC QPs (what OsqpEigen would hand to OSQP) for an MpcBatch via the
    oracle's numpy restatement of mpcPlanner.cpp:932-1146."""
    p = MA.MpcParams(**dataclasses.asdict(mb.params))
    return MA.assemble_batch(p, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.obs_dyn, mb.lin_pt,
                             mb.warm_x)


def rel_inf(a, b):
    """max_i |a-b|_inf / |b|_inf per row."""
    a = np.asarray(a); b = np.asarray(b)
    den = np.maximum(np.abs(b).max(axis=-1), 1e-300)
    return np.abs(a - b).max(axis=-1) / den


def oracle_solve(orc, mb, want_y=False, nthreads=1):
    """Oracle results for an MpcBatch whose obs_dyn may be per instance ([B,N,R]): instances are grouped by flag
    pattern (one CSC pattern per group, as the oracle's assembly needs) and the results scattered back."""
    if mb.obs_dyn.ndim == 2:
        return orc.solve_batch(to_qp_batch(mb), want_y=want_y, nthreads=nthreads)
    B = mb.B
    pat, inv = np.unique(mb.obs_dyn.reshape(B, -1), axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    out = None
    for g in range(len(pat)):
        sel = np.nonzero(inv == g)[0]
        sub = W.MpcBatch(mb.params, mb.x0[sel], mb.xref[sel], mb.obs_c[sel], mb.obs_semi[sel], mb.obs_yaw[sel],
                         mb.obs_dyn[sel[0]], mb.lin_pt[sel], mb.warm_x[sel])
        r = orc.solve_batch(to_qp_batch(sub), want_y=want_y, nthreads=nthreads)
        if out is None:
            out = {k: (None if v is None else np.empty((B,) + np.asarray(v).shape[1:], dtype=np.asarray(v).dtype)) for k, v in r.items()}
        for k, v in r.items():
            if v is not None:
                out[k][sel] = v
    return out
