"""history x migration on a development build: python scratch/qt3.py LIB B,R [...]"""
import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
lib = sys.argv[1]
if lib != "-": engine.LIB_PATH = os.path.abspath(lib)
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf
eng = engine.Engine(0)
orc = OB.RefOsqp()
for a in sys.argv[2:]:
    B, R = map(int, a.split(","))
    mb = W.static_batch(B, num_obs=R)
    ref = None
    for hist in (0, 1):
        for mig in (0, 1):
            eng.use_history(bool(hist)); eng.use_migration(bool(mig))
            ms = []
            for rep in range(3):
                out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
            if ref is None: ref = {k: out[k].copy() for k in ("x", "iter", "status", "obj", "rho_updates")}
            same = all(np.array_equal(out[k], ref[k]) for k in ref)
            print(f"B={B} R={R} hist={hist} migrate={mig}: {min(ms):.3f} ms -> {B/(min(ms)*1e-3):.0f} QP/s; launches {eng.last_launches}; bitwise same as first config: {same}", flush=True)
    n = min(B, 256)
    r = orc.solve_batch(to_qp_batch(mb.slice(0, n)), want_y=False, nthreads=8)
    ok = (out["status"][:n] == r["status"]).all() and (out["iter"][:n] == r["iter"]).all() and (out["rho_updates"][:n] == r["rho_updates"]).all()
    print(f"   parity vs reference on first {n}: {ok} x_err={rel_inf(out['x'][:n], r['x']).max():.2e}")
