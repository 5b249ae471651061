import sys, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
if len(sys.argv) > 2:
    import os; engine.LIB_PATH = os.path.abspath(sys.argv[2])
groups, meta = W.sweep_groups(0, n)
eng = engine.Engine(0)
for rep in range(2):
    tot_ms = 0.0; iters = 0; t0 = time.time(); per = []
    for idx, mb in groups:
        o = eng.solve_mpc_batch(mb); tot_ms += eng.last_kernel_ms; iters += int(o["iter"].sum())
        per.append((mb.num_obs, mb.B, eng.last_kernel_ms, eng.last_path, int(o["iter"].max())))
    print(f"rep {rep}: {n} instances, {len(groups)} groups, device ms {tot_ms:.1f} -> {n/tot_ms*1e3:.0f} QPs/s (kernels only), wall {time.time()-t0:.2f} s, iters {iters}")
big = sorted(per, key=lambda t: -t[2])[:6]
print("slowest groups (R, B, ms, path, max iter):", big)
