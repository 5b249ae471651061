"""the predictor rollouts upstream of the planner (dynamicPredictor.cpp:197-541) as a device kernel
(mpcqp_predict_device) against the literal restatement oracle/predictor.py.

Bars: the number of sampled trajectories per intent is exact by construction (the sampling loops run on the same accumulating
double counters on both sides; checked through the variance, which scales with it); mean paths and inflated box sizes within
1e-9 relative — the two sides differ in the rounding of cos / sin / atan2 (CUDA's libm vs glibc's), in fused multiply-adds
and in the order in which the samples are summed, nothing else.  Intent probabilities within 1e-12."""
import math

import numpy as np
import pytest

from intent_mpc_b200 import receding
from oracle import predictor as OP

TOL = 1e-9


def test_oracle_predictor_shapes_and_properties():
    p = OP.PredictorParams()
    c = p.derived()
    assert c["paraml"] == 1.0 and abs(c["frontAngle"] - 25 * math.pi / 180) < 1e-15 and abs(c["params"] - math.atanh(0.5) / 0.1) < 1e-12
    host = receding.IntentSweep(S=2, D=2, seed0=9)
    ph, vh = host.history(7, H=8)
    pp, ps, prob, ns = OP.predict_obstacle(p, ph[0, 0], vh[0, 0], host.size[0, 0])
    assert np.array(pp).shape == (4, 31, 3) and np.array(ps).shape == (4, 31, 3) and len(prob) == 4
    v = math.hypot(vh[0, 0, 0, 0], vh[0, 0, 0, 1])
    assert ns[OP.STOP] == 1 and ns[OP.FORWARD] == 9 * math.ceil(2 * v / 0.1 - 1e-12)        # 9 headings x speeds 0, 0.1, ... < 2 v
    assert ns[OP.LEFT] == ns[OP.RIGHT] > 50
    pp = np.array(pp); ps = np.array(ps)
    assert np.array_equal(pp[OP.STOP], np.tile(ph[0, 0, 0], (31, 1)))                          # standing still
    assert np.allclose(pp[:, 0], ph[0, 0, 0])                                                  # every intent starts at the obstacle
    assert (np.diff(ps[OP.FORWARD][:, 0]) >= -1e-12).all() and ps[OP.FORWARD][-1, 0] > ps[OP.FORWARD][0, 0]   # the box grows with the spread
    assert ps[OP.STOP][1, 0] == pytest.approx(host.size[0, 0, 0] + 2 * 0.1 * 0.1)             # min(v, stopVel) * 2 dt per step
    # forward mean path = constant velocity at the MEAN sampled speed along the current heading (symmetric fan of headings)
    head = math.atan2(vh[0, 0, 0, 1], vh[0, 0, 0, 0])
    d = pp[OP.FORWARD][-1, :2] - pp[OP.FORWARD][0, :2]
    assert abs(math.atan2(d[1], d[0]) - head) < 0.06
    # left turns left of the heading, right turns right
    cross = lambda a: math.cos(head) * a[1] - math.sin(head) * a[0]
    assert cross(pp[OP.LEFT][-1, :2] - pp[OP.LEFT][0, :2]) > 0 > cross(pp[OP.RIGHT][-1, :2] - pp[OP.RIGHT][0, :2])
    assert abs(sum(prob)) > 0 and all(q >= 0 for q in prob)
    # a slow obstacle: every intent is the stop model
    slow = vh[0, 0].copy() * 1e-3
    pp2, ps2, _, ns2 = OP.predict_obstacle(p, ph[0, 0], slow, host.size[0, 0])
    assert ns2 == [1, 1, 1, 1] and np.array_equal(np.array(pp2)[OP.LEFT], np.array(pp2)[OP.STOP])


@pytest.mark.gpu
def test_device_predictor_matches_oracle():
    import torch
    from intent_mpc_b200 import engine
    eng = engine.Engine(0)
    dev = torch.device("cuda", 0)
    try:
        host = receding.IntentSweep(S=6, D=4, seed0=31)
        p = OP.PredictorParams()
        for step, H in ((0, 10), (13, 6), (40, 3)):
            ph, vh = host.history(step, H=H)
            ph = ph.reshape(-1, H, 3).copy(); vh = vh.reshape(-1, H, 3).copy(); sz = host.size.reshape(-1, 3).copy()
            vh[5] *= 1e-3                                             # one obstacle below the stop threshold
            NOB = ph.shape[0]
            d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            t = dict(pos_hist=d(ph), vel_hist=d(vh), size=d(sz), pred_pos=torch.empty((NOB, 4, 31, 3), dtype=torch.float64, device=dev),
                     pred_size=torch.empty((NOB, 4, 31, 3), dtype=torch.float64, device=dev), intent_prob=torch.empty((NOB, 4), dtype=torch.float64, device=dev))
            eng.predict_ptr(engine.default_predictor_params(), NOB, H, {k: v.data_ptr() for k, v in t.items()})
            eng.sync()
            gp, gs, gq = t["pred_pos"].cpu().numpy(), t["pred_size"].cpu().numpy(), t["intent_prob"].cpu().numpy()
            for ob in range(NOB):
                pp, ps, prob, _ = OP.predict_obstacle(p, ph[ob], vh[ob], sz[ob])
                pp = np.array(pp); ps = np.array(ps)
                assert np.abs(gp[ob] - pp).max() <= TOL * np.abs(pp).max(), (step, ob)
                assert np.abs(gs[ob] - ps).max() <= TOL * np.abs(ps).max(), (step, ob)
                assert np.abs(gq[ob] - np.array(prob)).max() <= 1e-12, (step, ob)
    finally:
        eng.close()
