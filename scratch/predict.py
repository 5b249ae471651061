import sys; sys.path.insert(0, ".")
import numpy as np, os
from intent_mpc_b200 import workloads as W
from oracle import bindings as OB
from tests.helpers import oracle_solve
orc = OB.RefOsqp()
groups, _ = W.sweep_groups(0, 1500)
its = []; feats = []
for idx, mb in groups:
    r = oracle_solve(orc, mb, nthreads=8)
    c = mb.lin_pt[:, :, None, :]                       # [B,N,1,3]
    d = c - mb.obs_c
    cs, sn = np.cos(mb.obs_yaw), np.sin(mb.obs_yaw)
    xi = d[..., 0] * cs + d[..., 1] * sn; eta = -d[..., 0] * sn + d[..., 1] * cs
    f = xi ** 2 / mb.obs_semi[..., 0] ** 2 + eta ** 2 / mb.obs_semi[..., 1] ** 2 + d[..., 2] ** 2 / mb.obs_semi[..., 2] ** 2
    inside = (f < 1.0)
    nin = inside.sum(axis=(1, 2)); nin0 = inside[:, 0].sum(axis=1); fmin = f.min(axis=(1, 2)); depth = np.maximum(1 - f, 0).sum(axis=(1, 2))
    for b in range(mb.B):
        its.append(r["iter"][b]); feats.append((nin[b], nin0[b], fmin[b], depth[b], mb.num_obs, mb.params.max_vel, r["status"][b]))
its = np.array(its); F = np.array(feats)
print("n", len(its), "4000-iter:", (its == 4000).sum(), "status hist", dict(zip(*np.unique(F[:, 6], return_counts=True))))
for name, col in (("rows inside (all stages)", 0), ("rows inside at stage 0", 1), ("min f", 2), ("depth sum", 3)):
    x = F[:, col]
    print(name, "corr with iters %.3f" % np.corrcoef(x, its)[0, 1], "corr with log iters %.3f" % np.corrcoef(x, np.log(its))[0, 1])
hard = its >= 1000
for thr in (1, 5, 20, 50):
    pred = F[:, 0] >= thr
    print(f"nin>={thr}: flagged {pred.sum()}, hard among flagged {(pred & hard).sum()} / hard total {hard.sum()}")
for thr in (0.5, 2, 5, 10):
    pred = F[:, 3] >= thr
    print(f"depth>={thr}: flagged {pred.sum()}, hard among flagged {(pred & hard).sum()} / hard total {hard.sum()}")
