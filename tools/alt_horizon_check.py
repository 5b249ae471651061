"""Horizons other than 30 on the CTA kernel (where built) against the oracle and against the generic kernel's time."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf
eng = engine.Engine(0)
orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
for H, R, B in [tuple(map(int, a.split(","))) for a in sys.argv[1:]] or [(20, 3, 512)]:
    p = W.MpcParams(horizon=H)
    mb = W.static_batch(B, num_obs=R, params=p, seed0=4000 + H)
    res = {}
    for path in ("cta", "generic"):
        eng.force_generic(path)
        ms = []
        for _ in range(2):
            out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
        res[path] = (eng.last_path, min(ms), out)
    eng.force_generic("cta")
    ref = orc.solve_batch(to_qp_batch(mb.slice(0, min(B, 256))), want_y=False)
    o = res["cta"][2]; nb = min(B, 256)
    print(f"horizon {H} R {R} B {B}: path {res['cta'][0]} {res['cta'][1]:.3f} ms ({B/res['cta'][1]*1e3:.0f} QPs/s) vs generic {res['generic'][1]:.3f} ms; "
          f"status equal {bool((o['status'][:nb] == ref['status']).all())} iter equal {bool((o['iter'][:nb] == ref['iter']).all())} "
          f"x err {rel_inf(o['x'][:nb], ref['x']).max():.2e}; cta == generic iters {bool((o['iter'] == res['generic'][2]['iter']).all())}", flush=True)
