import sys, os
sys.path.insert(0, ".")
from intent_mpc_b200 import engine as E
from oracle import polytraj_assembly as PA
eng = E.Engine(0)
for paths, K in ((1000, 8), (400, 16), (200, 25)):
    qb = PA.path_batch(paths, K=K, seed0=100)
    E.solve_qp_batch(eng, qb, want_y=False)
    ms = []
    for _ in range(3):
        E.solve_qp_batch(eng, qb, want_y=False); ms.append(eng.last_kernel_ms)
    print(os.environ.get("MPCQP_DENSE_L2_MB", "none"), "K", K, "B", qb.q.shape[0], f"{min(ms):.2f} ms {qb.q.shape[0] / min(ms) * 1e3:.0f} QPs/s", flush=True)
