"""configs[3] timing: the stress set (horizon 60) on whatever kernel the dispatcher picks; prints kernel ms, QPs/s, path."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H = int(sys.argv[2]) if len(sys.argv) > 2 else 60
eng = engine.Engine(0)
mb = W.stress_batch(B, horizon=H)
out = eng.solve_mpc_batch(mb)
ms = []
for _ in range(2):
    out = eng.solve_mpc_batch(mb); ms.append(eng.last_kernel_ms)
it = out["iter"]
print(f"stress B={B} horizon={H} path={eng.last_path} kernel ms {min(ms):.2f} -> {B / (min(ms) * 1e-3):.0f} QPs/s; iterations total {it.sum()} max {it.max()}; "
      f"us per iteration per SM {min(ms) * 1e3 * 148 / it.sum():.2f}; status hist {dict(zip(*np.unique(out['status'], return_counts=True)))}")
one = mb.slice(3, 4)
o1 = eng.solve_mpc_batch(one); o1 = eng.solve_mpc_batch(one)
print(f"single instance: iterations {o1['iter'][0]} kernel ms {eng.last_kernel_ms:.3f} -> {eng.last_kernel_ms * 1e3 / o1['iter'][0]:.2f} us per iteration")
