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
