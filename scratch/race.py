import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mb = W.static_batch(3, num_obs=R)
o = eng.solve_mpc_batch(mb)
print("iters", o["iter"], "status", o["status"], "path", eng.last_path)
