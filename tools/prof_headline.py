"""Profile target for the headline: batch j of bench.py's rotation (1,024 QPs, 4 obstacles), scheduling hint off, as the timed loop runs it."""
import sys; sys.path.insert(0, ".")
from intent_mpc_b200 import engine, workloads as W
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = engine.Engine(0)
eng.use_history(False)
for j in range(reps):
    mb = W.static_batch(1024, num_obs=4, seed0=bench.batch_seed(0, 1024, j))
    out = eng.solve_mpc_batch(mb)
    print("batch", j, "path", eng.last_path, "kernel ms", eng.last_kernel_ms, "solve ms", eng.last_solve_kernel_ms, "launches", eng.last_launches, "iters", int(out["iter"].sum()))
