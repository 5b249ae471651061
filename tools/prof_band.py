"""Profile target for the sparse generic path: `paths` candidate paths x 3 axes of K-segment minimum-snap QPs, one launch."""
import sys; sys.path.insert(0, ".")
from intent_mpc_b200 import engine, polytraj_workload as PA
K = int(sys.argv[1]) if len(sys.argv) > 1 else 25
paths = int(sys.argv[2]) if len(sys.argv) > 2 else 320
eng = engine.Engine(0)
qb = PA.path_batch(paths, K=K, seed0=100)
for _ in range(2):
    out = engine.solve_qp_batch(eng, qb, want_y=False)
print("path", eng.last_path, "QPs", qb.q.shape[0], "n", qb.n, "m", qb.m, "kernel ms", eng.last_kernel_ms, "iters", int(out["iter"].sum()))
