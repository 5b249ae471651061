"""Iteration counts of a static batch next to its inputs (for offline studies of which instances run long)."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = engine.Engine(0)
mb = W.static_batch(B, num_obs=4)
out = eng.solve_mpc_batch(mb)
np.savez_compressed(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/iters.npz", iter=out["iter"], status=out["status"], rho_updates=out["rho_updates"])
print("saved", B, int((out["iter"] == 4000).sum()))
