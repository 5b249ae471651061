"""Small target for compute-sanitizer (one tool per run): a one-per-SM launch with the row helper (B = 6), a two-launch batch with
migration (B = 200: hard list + two per SM + resume) and a wide-kernel batch, a few iterations each."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0); eng.use_history(False)
s = engine.default_settings(); s.max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for B, R in ((6, 4), (200, 4), (6, 8), (6, 2)):
    out = eng.solve_mpc_batch(W.static_batch(B, num_obs=R), settings=s)
    print("B", B, "R", R, "path", eng.last_path, "launches", eng.last_launches, "status", np.bincount(out["status"] + 10)[np.bincount(out["status"] + 10) > 0], flush=True)
sb, _ = W.sweep_batches(0, 16, one_launch=True)
for _, smb in sb:
    out = eng.solve_mpc_batch(smb, settings=s)
    print("sweep slice", smb.B, "path", eng.last_path, "iters", int(out["iter"].sum()), flush=True)
