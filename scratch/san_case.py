import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
st = engine.default_settings(max_iter=200)
mode = sys.argv[1]
if mode == "assist":
    mb = W.static_batch(6, num_obs=4); out = eng.solve_mpc_batch(mb, settings=st)
elif mode == "plain":
    eng.force_generic("cta_plain"); mb = W.static_batch(6, num_obs=4); out = eng.solve_mpc_batch(mb, settings=st)
elif mode == "r0":
    mb = W.static_batch(4, num_obs=0); out = eng.solve_mpc_batch(mb, settings=st)
elif mode == "wide":
    b, _ = W.sweep_batches(0, 12); out = None
    for idx, mb in b: out = eng.solve_mpc_batch(mb, settings=st)
print(mode, "iters", out["iter"], "path", eng.last_path)
