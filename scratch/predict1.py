import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch
orc = OB.RefOsqp()
mb = W.static_batch(2048, num_obs=4)
r = orc.solve_batch(to_qp_batch(mb), want_y=False, nthreads=8)
its = r["iter"]
c = mb.lin_pt[:, :, None, :]; d = c - mb.obs_c
cs, sn = np.cos(mb.obs_yaw), np.sin(mb.obs_yaw)
xi = d[..., 0] * cs + d[..., 1] * sn; eta = -d[..., 0] * sn + d[..., 1] * cs
f = xi ** 2 / mb.obs_semi[..., 0] ** 2 + eta ** 2 / mb.obs_semi[..., 1] ** 2 + d[..., 2] ** 2 / mb.obs_semi[..., 2] ** 2
inside = f < 1.0
nin = inside.sum(axis=(1, 2)); nin0 = inside[:, 0].any(axis=1)
hard = its >= 1000
print("hard", hard.sum(), "of", len(its), "status", dict(zip(*np.unique(r["status"], return_counts=True))))
print("stage-0 violated (already flagged):", nin0.sum(), "hard among them", (nin0 & hard).sum())
for thr in (1, 3, 5, 10, 20):
    pred = nin >= thr
    print(f"rows inside >= {thr}: flagged {pred.sum()}, hard among flagged {(pred & hard).sum()} / {hard.sum()}; mean iters flagged {its[pred].mean():.0f} vs rest {its[~pred].mean():.0f}")
fmin = f.min(axis=(1, 2))
for thr in (0.2, 0.5, 0.8):
    pred = fmin < thr
    print(f"min f < {thr}: flagged {pred.sum()}, hard among flagged {(pred & hard).sum()} / {hard.sum()}")
