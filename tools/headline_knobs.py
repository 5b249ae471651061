"""Headline step (1,024 QPs, 4 obstacles, batches in rotation, scheduling hint off) under development knobs of the engine
(environment, read at engine creation): python tools/headline_knobs.py KEY=V[,V..] ..."""
import os, sys, itertools; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
import bench
knobs = [a.split("=") for a in sys.argv[1:]]
keys = [k for k, _ in knobs]; vals = [v.split(",") for _, v in knobs]
batches = [W.static_batch(1024, num_obs=4, seed0=bench.batch_seed(0, 1024, j)) for j in range(8)]
for combo in itertools.product(*vals) if knobs else [()]:
    for k, v in zip(keys, combo): os.environ[k] = v
    eng = engine.Engine(0); eng.use_history(False)
    for mb in batches[:3]: eng.solve_mpc_batch(mb)
    ms = []
    for rep in range(2):
        for mb in batches:
            out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
    ms = np.array(ms)
    print(dict(zip(keys, combo)), f"solve kernel ms: mean {ms.mean():.3f} min {ms.min():.3f} max {ms.max():.3f} -> {1024/ms.mean()*1e3:.0f} QPs/s", flush=True)
    eng.close()
