import sys, dataclasses
sys.path.insert(0, ".")
from intent_mpc_b200 import engine as E
from oracle import polytraj_assembly as PA
K = int(sys.argv[1]) if len(sys.argv) > 1 else 25
eng = E.Engine(0)
qb = PA.path_batch(1, K=K, seed0=100)
one = dataclasses.replace(qb, P_val=qb.P_val[:1], q=qb.q[:1], A_val=qb.A_val[:1], l=qb.l[:1], u=qb.u[:1], warm_x=qb.warm_x[:1])
r = E.solve_qp_batch(eng, one, want_y=False)
print("kernel ms", eng.last_kernel_ms, r["iter"], r["status"])
