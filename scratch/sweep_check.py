import sys, os, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
if len(sys.argv) > 2: engine.LIB_PATH = os.path.abspath(sys.argv[2])
from oracle import bindings as OB
from tests.helpers import oracle_solve, rel_inf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
orc = OB.RefOsqp()
eng = engine.Engine(0)
groups, meta = W.sweep_groups(0, n)
bad = 0
for idx, mb in groups:
    out = eng.solve_mpc_batch(mb)
    ref = oracle_solve(orc, mb)
    ok = (out["status"] == ref["status"]).all() and (out["iter"] == ref["iter"]).all() and (out["rho_updates"] == ref["rho_updates"]).all()
    ex = rel_inf(out["x"], ref["x"]).max(); eo = np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max()
    if not ok or ex > 1e-5 or eo > 1e-5:
        bad += 1
        print("MISMATCH", mb.params.max_vel, mb.num_obs, mb.B, eng.last_path, ok, ex, eo, out["iter"][:8], ref["iter"][:8])
print("groups", len(groups), "bad", bad, "path", eng.last_path)
