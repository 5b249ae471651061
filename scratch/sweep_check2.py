import sys, os, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
if len(sys.argv) > 2: engine.LIB_PATH = os.path.abspath(sys.argv[2])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
eng = engine.Engine(0)
groups, meta = W.sweep_groups(0, n)
batches, _ = W.sweep_batches(0, n)
ref = {}
for idx, mb in groups:
    out = eng.solve_mpc_batch(mb)
    for j, i in enumerate(idx): ref[int(i)] = (out["x"][j], out["iter"][j], out["status"][j], out["obj"][j])
bad = 0; tot = 0.0
for rep in range(2):
    tot = 0.0
    for idx, mb in batches:
        out = eng.solve_mpc_batch(mb); tot += eng.last_kernel_ms
        if rep == 0:
            for j, i in enumerate(idx):
                x, it, stt, ob = ref[int(i)]
                if it != out["iter"][j] or stt != out["status"][j] or not np.array_equal(x, out["x"][j]): bad += 1
    print(f"rep {rep}: padded path {n} instances in {len(batches)} launches: {tot:.1f} ms -> {n/tot*1e3:.0f} QPs/s")
print("instances differing from the per-group path:", bad, "of", n)
