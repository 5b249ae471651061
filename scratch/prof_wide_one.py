import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
if len(sys.argv) > 1: engine.LIB_PATH = os.path.abspath(sys.argv[1])
eng = engine.Engine(0)
b, _ = W.sweep_batches(0, 1024)
idx, mb = b[-1]
if os.path.exists("/tmp/hard_idx"):
    j = int(open("/tmp/hard_idx").read())
else:
    o = eng.solve_mpc_batch(mb)
    j = int(np.nonzero((o["iter"] == 4000) & (mb.nobs == 32))[0][0]); open("/tmp/hard_idx", "w").write(str(j))
one = mb.slice(j, j + 1)
for _ in range(2):
    o = eng.solve_mpc_batch(one)
print("instance", j, "R", one.nobs, "iters", o["iter"], "ms", eng.last_solve_kernel_ms)
