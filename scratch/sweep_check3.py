import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
if len(sys.argv) > 2: engine.LIB_PATH = os.path.abspath(sys.argv[2])
n = int(sys.argv[1])
eng = engine.Engine(0)
b3, _ = W.sweep_batches(0, n)
b1, _ = W.sweep_batches(0, n, one_launch=True)
ref = {}
t3 = 0.0
for rep in range(2):
    t3 = 0.0
    for idx, mb in b3:
        o = eng.solve_mpc_batch(mb); t3 += eng.last_kernel_ms
        for j, i in enumerate(idx): ref[int(i)] = (o["x"][j], o["iter"][j], o["status"][j])
idx, mb = b1[0]
for rep in range(2):
    o = eng.solve_mpc_batch(mb); t1 = eng.last_kernel_ms
bad = sum(1 for j, i in enumerate(idx) if ref[int(i)][1] != o["iter"][j] or ref[int(i)][2] != o["status"][j] or not np.array_equal(ref[int(i)][0], o["x"][j]))
print(f"{n} instances: three launches {t3:.1f} ms ({n/t3*1e3:.0f} QPs/s), one launch {t1:.1f} ms ({n/t1*1e3:.0f} QPs/s); differing: {bad}")
