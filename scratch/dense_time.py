"""Timing of the dense generic kernel on polyTrajSolver-shaped batches (development script)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from intent_mpc_b200 import engine as E
from oracle import polytraj_assembly as PA
from oracle import bindings as OB
eng = E.Engine(0)
ref = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
for paths, K in ((1, 8), (50, 8), (200, 8), (1000, 8), (100, 25), (400, 25)):
    qb = PA.path_batch(paths, K=K, seed0=100)
    E.solve_qp_batch(eng, qb, want_y=False)
    t = time.time(); r = E.solve_qp_batch(eng, qb, want_y=False); wall = time.time() - t
    ms = eng.last_kernel_ms
    B = qb.q.shape[0]
    c = ref.solve_batch(qb, want_y=False, nthreads=16)
    print(f"paths {paths} K {K} n {qb.n} m {qb.m} B {B} kernel {ms:.3f} ms ({B / ms * 1e3:.0f} QPs/s) wall {wall * 1e3:.2f} ms iters {r['iter'].sum()} max {r['iter'].max()} "
          f"cpu {c['wall'] * 1e3:.1f} ms ({B / c['wall']:.0f} QPs/s) same_iter {(r['iter'] == c['iter']).all()}", flush=True)
