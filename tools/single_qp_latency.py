import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
mb = W.static_batch(1024, num_obs=4)
out = eng.solve_mpc_batch(mb)
it = out["iter"]; ru = out["rho_updates"]
strag = np.where(it == 4000)[0]
print("stragglers", strag, "rho_updates", ru[strag], "status", out["status"][strag])
print("rho_updates hist", np.bincount(ru)[:10], "iters mean", it.mean())
for idx in [int(strag[0]), int(np.argsort(it)[len(it)//2])]:
    m1 = mb.slice(idx, idx + 1)
    for path in ("cta", "fast"):
        eng.force_generic(path)
        ms = []
        for _ in range(3):
            o = eng.solve_mpc_batch(m1); ms.append(eng.last_solve_kernel_ms)
        print(f"instance {idx}: iters {o['iter'][0]} rho_updates {o['rho_updates'][0]} path {path}: {min(ms):.3f} ms -> {min(ms)*1e3/o['iter'][0]:.3f} us/iter")
