import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
mb = W.static_batch(1024, num_obs=4)
o = eng.solve_mpc_batch(mb)
idx = np.nonzero(o["iter"] == 4000)[0]
print("hard", idx.tolist())
print("status", o["status"][idx].tolist())
print("rho_updates", o["rho_updates"][idx].tolist())
