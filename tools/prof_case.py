import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]); R = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
path = sys.argv[4] if len(sys.argv) > 4 else "cta"
eng = engine.Engine(0)
eng.force_generic(path)
mb = W.static_batch(B, num_obs=R)
for _ in range(reps):
    out = eng.solve_mpc_batch(mb)
print("path", eng.last_path, "kernel ms", eng.last_kernel_ms, "iters", out["iter"].sum())
