"""mpcPlanner QPs WITH the field-of-view half-space rows (3-argument updateCurrStates -> updateFovParam, mpcPlanner.cpp:
265-297; two rows per stage on (x_k, y_k), castMPCToQPConstraintMatrix :1027-1038 and ...ConstraintVectors :1102-1111).  Off in
the reference's production path (mpcNavigation calls the 2-argument form), so these are a stand-alone case: B instances of
the static workload, each with its own yaw.  Shared by make_golden_fov.py and tests/test_dense_generic.py."""
import dataclasses
import math

import numpy as np

from intent_mpc_b200 import workloads as W
from oracle import mpc_assembly as MA


def fov_half_spaces(pos, yaw):
    """updateFovParam, mpcPlanner.cpp:276-297 (87/2 is an integer division: 43 degrees)."""
    hmax = np.zeros((len(yaw), 3)); hmin = np.zeros((len(yaw), 3))
    for b, (p, y) in enumerate(zip(pos, yaw)):
        a_max = y - (87 // 2) * math.pi / 180.0; a_min = y + (87 // 2) * math.pi / 180.0
        a1, b1 = math.sin(a_max), -math.cos(a_max); a2, b2 = math.sin(a_min), -math.cos(a_min)
        hmax[b] = (a1, b1, a1 * p[0] + b1 * p[1]); hmin[b] = (a2, b2, a2 * p[0] + b2 * p[1])
    return hmax, hmin


def case(B=4, num_obs=2, seed0=500):
    mb = W.static_batch(B, num_obs=num_obs, seed0=seed0)
    yaw = np.random.default_rng(seed0).uniform(-0.5, 0.5, B)
    p = MA.MpcParams(**dataclasses.asdict(mb.params))
    hs = fov_half_spaces(mb.x0[:, 0:3], yaw)
    return MA.assemble_batch(p, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.obs_dyn, mb.lin_pt, mb.warm_x, half_space=hs)
