import sys, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
paths = ["cta", "fast"]
cases = [(1024, 4), (1024, 0), (16384, 4)] if len(sys.argv) < 2 else [tuple(map(int, a.split(","))) for a in sys.argv[1:]]
for B, R in cases:
    mb = W.static_batch(B, num_obs=R)
    ref = None
    for path in paths:
        eng.force_generic(path)
        ms = []
        for rep in range(3):
            out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
        it = out["iter"]
        if ref is None: ref = out
        same = (out["iter"] == ref["iter"]).all() and (out["status"] == ref["status"]).all()
        err = np.abs(out["x"] - ref["x"]).max() / np.abs(ref["x"]).max()
        print(f"B={B} R={R} path={eng.last_path} solve kernel ms {min(ms):.3f} -> {B/(min(ms)*1e-3):.0f} QP/s; iters sum {it.sum()} max {it.max()}; us/iter(straggler) {min(ms)*1e3/it.max():.3f}; same_as_first={same} xdiff={err:.2e}", flush=True)
