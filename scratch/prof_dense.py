"""ncu target: the dense generic kernel on (a) 600 polyTrajSolver QPs of K = 8 segments (throughput regime, more QPs than CTAs),
(b) 3 QPs of K = 25 segments (one path's x, y, z).  Launch 0 / 1 of the capture (-k regex:mpcqp_dense)."""
import sys
sys.path.insert(0, ".")
from intent_mpc_b200 import engine as E
from oracle import polytraj_assembly as PA
eng = E.Engine(0)
for paths, K in ((200, 8), (1, 25)):
    qb = PA.path_batch(paths, K=K, seed0=100)
    r = E.solve_qp_batch(eng, qb, want_y=False)
    print("paths", paths, "K", K, "kernel ms", eng.last_kernel_ms, "iters", int(r["iter"].sum()))
