import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
idx = int(sys.argv[1]); path = sys.argv[2] if len(sys.argv) > 2 else "cta"
if len(sys.argv) > 3: engine.LIB_PATH = os.path.abspath(sys.argv[3])
eng = engine.Engine(0); eng.force_generic(path)
mb = W.static_batch(1024, num_obs=4).slice(idx, idx + 1)
for _ in range(2):
    o = eng.solve_mpc_batch(mb)
print("iters", o["iter"], "ms", eng.last_solve_kernel_ms)
