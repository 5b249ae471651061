import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.helpers import oracle_solve, rel_inf
orc = OB.RefOsqp()
eng = engine.Engine(0)
batches, _ = W.sweep_batches(0, 8192)
for idx, mb in batches:
    a = eng.solve_mpc_batch(mb)
    p = mb.params; N, NS = p.N, p.N + 1
    u = a["x"][:, 8 * NS:].reshape(-1, N, 5)
    viol = (np.abs(u[:, :, 0:3]).max(axis=(1, 2)) > p.max_acc + 0.2) & (a["status"] == 1)
    print(p.max_vel, mb.B, "violators", int(viol.sum()))
    for j in np.nonzero(viol)[0][:3]:
        R = int(mb.nobs[j])
        sub = W.MpcBatch(p, mb.x0[j:j+1], mb.xref[j:j+1], mb.obs_c[j:j+1, :, :R], mb.obs_semi[j:j+1, :, :R], mb.obs_yaw[j:j+1, :, :R],
                         mb.obs_dyn[j:j+1, :, :R], mb.lin_pt[j:j+1], mb.warm_x[j:j+1])
        r = oracle_solve(orc, sub)
        one = eng.solve_mpc_batch(sub)
        print("  inst", int(idx[j]), "R", R, "gpu(padded) it/status", a["iter"][j], a["status"][j], "gpu(single)", one["iter"][0], one["status"][0], "oracle", r["iter"][0], r["status"][0],
              "xerr padded", rel_inf(a["x"][j], r["x"][0]), "xerr single", rel_inf(one["x"][0], r["x"][0]), "max|u| oracle", np.abs(r["x"][0][8*NS:].reshape(N,5)[:, :3]).max(), "pri_res", a["pri_res"][j])
