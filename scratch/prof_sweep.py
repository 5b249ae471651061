import sys; sys.path.insert(0, ".")
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
b, _ = W.sweep_batches(0, 4096)
for _ in range(2):
    for idx, mb in b[-1:]:
        o = eng.solve_mpc_batch(mb)
print("sweep slice", b[-1][1].B, "ms", eng.last_kernel_ms, "iters", o["iter"].sum())
