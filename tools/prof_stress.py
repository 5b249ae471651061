"""Profile target for the generic one-warp kernel: B stress-set instances (configs[3]: horizon 60, 2 obstacles), max_iter capped."""
import sys; sys.path.insert(0, ".")
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
eng = engine.Engine(0)
s = engine.default_settings(); s.max_iter = iters; s.eps_abs = 1e-12; s.eps_rel = 1e-12
mb = W.stress_batch(B)
for _ in range(2):
    out = eng.solve_mpc_batch(mb, settings=s)
    print("path", eng.last_path, "solve kernel ms", eng.last_solve_kernel_ms, "us/iter/QP", eng.last_solve_kernel_ms * 1e3 / iters, "iters", int(out["iter"].sum()))
