"""Profile target for the one-CTA-per-SM regime: B (<= 148) QPs that never reach the tolerances and run `iters` iterations each
(eps 1e-12): the launch is iterations + residual checks + the occasional re-factorisation only."""
import sys; sys.path.insert(0, ".")
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
path = sys.argv[3] if len(sys.argv) > 3 else "cta"
eng = engine.Engine(0); eng.force_generic(path); eng.use_history(False); eng.use_migration(False)
s = engine.default_settings(); s.eps_abs = 1e-12; s.eps_rel = 1e-12; s.max_iter = iters
mb = W.static_batch(B, num_obs=4)
for _ in range(2):
    out = eng.solve_mpc_batch(mb, settings=s)
    print("path", eng.last_path, "solve kernel ms", eng.last_solve_kernel_ms, "us/iter", eng.last_solve_kernel_ms * 1e3 / iters, "rho updates", out["rho_updates"].mean())
