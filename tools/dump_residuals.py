"""Residuals after K iterations (max_iter = K runs) next to the iteration counts of the full solve: how well do they predict
which instances run long?  (offline study for the order in which parked instances are resumed)"""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = engine.Engine(0)
mb = W.static_batch(B, num_obs=4)
full = eng.solve_mpc_batch(mb)
d = {"iter": full["iter"], "status": full["status"]}
for K in (25, 50, 100):
    s = engine.default_settings(); s.max_iter = K
    o = eng.solve_mpc_batch(mb, settings=s)
    d[f"pri{K}"] = o["pri_res"]; d[f"dua{K}"] = o["dua_res"]; d[f"it{K}"] = o["iter"]; d[f"st{K}"] = o["status"]
np.savez_compressed("gpurun_out/residuals8k.npz", **d)
print("saved")
