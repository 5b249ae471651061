"""One-off wide parity check against the reference binary: B instances for each obstacle count (default dispatch: two-launch regime
with migration, row helper on the one-per-SM launches), status / iterations / rho updates equal, x and objective within 1e-5."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = engine.Engine(0); eng.use_history(False)
orc = OB.RefOsqp()
bad = 0
for R in [int(r) for r in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,4,5,6,7,8".split(","))]:
    mb = W.static_batch(B, num_obs=R, seed0=50000 + 1000 * R)
    out = eng.solve_mpc_batch(mb)
    ref = orc.solve_batch(to_qp_batch(mb), want_y=False)
    ok = (out["status"] == ref["status"]) & (out["iter"] == ref["iter"]) & (out["rho_updates"] == ref["rho_updates"])
    ex = rel_inf(out["x"], ref["x"]); eo = np.abs((out["obj"] - ref["obj"]) / ref["obj"])
    bad += int((~ok).sum()) + int((ex >= 1e-5).sum())
    print(f"R={R}: path {eng.last_path} launches {eng.last_launches}; status/iter/rho equal {int(ok.sum())}/{B}; x err max {ex.max():.2e}; obj err max {eo.max():.2e}; "
          f"iters max {int(out['iter'].max())}, 4000s {int((out['iter'] == 4000).sum())}", flush=True)
print("MISMATCHES", bad)
