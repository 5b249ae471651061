import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
mb = W.static_batch(8, num_obs=4)
for mi in (25, 50, 100):
    s = engine.default_settings(max_iter=mi)
    res = {}
    for path in ("cta", "fast"):
        eng.force_generic(path)
        res[path] = eng.solve_mpc_batch(mb, settings=s)
    a, b = res["cta"], res["fast"]
    print("max_iter", mi)
    for key in ("status", "iter", "rho_updates", "obj", "pri_res", "dua_res"):
        print(" ", key, a[key][:4], b[key][:4])
    print("  xdiff", np.abs(a["x"] - b["x"]).max(axis=1) / np.abs(b["x"]).max(axis=1))
