"""Where the dense generic kernel's time goes: one QP, settings variants (development script)."""
import sys, dataclasses
import numpy as np
sys.path.insert(0, ".")
from intent_mpc_b200 import engine as E
from oracle import polytraj_assembly as PA
eng = E.Engine(0)
for K in (8, 25):
    qb = PA.path_batch(1, K=K, seed0=100)
    one = dataclasses.replace(qb, P_val=qb.P_val[:1], q=qb.q[:1], A_val=qb.A_val[:1], l=qb.l[:1], u=qb.u[:1], warm_x=qb.warm_x[:1])
    for name, kw in (("default", {}), ("max_iter=25", dict(max_iter=25)), ("max_iter=50,no adapt", dict(max_iter=50, adaptive_rho=0)), ("max_iter=100,no adapt", dict(max_iter=100, adaptive_rho=0)),
                     ("max_iter=25,no adapt", dict(max_iter=25, adaptive_rho=0)), ("max_iter=25,no adapt,scaling=0", dict(max_iter=25, adaptive_rho=0, scaling=0)),
                     ("max_iter=25,no adapt,scaling=1", dict(max_iter=25, adaptive_rho=0, scaling=1)), ("max_iter=100,no adapt,check=0", dict(max_iter=100, adaptive_rho=0, check_termination=0))):
        s = E.default_settings(**kw)
        E.solve_qp_batch(eng, one, settings=s, want_y=False)
        ms = []
        for _ in range(3):
            r = E.solve_qp_batch(eng, one, settings=s, want_y=False); ms.append(eng.last_kernel_ms)
        print(f"K {K} N {qb.n + qb.m} {name}: {min(ms):.3f} ms iter {r['iter'][0]} rho_updates {r['rho_updates'][0]}", flush=True)
