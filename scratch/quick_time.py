import sys, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
cases = [(1024, 4), (1024, 0), (8192, 4)] if len(sys.argv) < 2 else [tuple(map(int, a.split(","))) for a in sys.argv[1:]]
for B, R in cases:
    mb = W.static_batch(B, num_obs=R)
    ms = []
    for rep in range(4):
        out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
    it = out["iter"]
    print(f"B={B} R={R} path={eng.last_path} solve kernel ms {min(ms):.3f} -> {B/(min(ms)*1e-3):.0f} QP/s; iters sum {it.sum()} max {it.max()}; us/iter(straggler) {min(ms)*1e3/it.max():.2f}")
