"""16,384-QP batch: solve kernel time without / with the iteration-count history of a previous call of the same slots (the hint
puts the instances that ran long on the hard list, which starts first: the bound that perfect knowledge of the long instances gives)."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = engine.Engine(0)
mb = W.static_batch(B, num_obs=4)
for hist in (False, True, False, True):
    eng.use_history(hist)
    ms = []
    for _ in range(3):
        out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
    it = out["iter"]
    print(f"B={B} history={hist}: solve kernels {ms[0]:.2f} {ms[1]:.2f} {ms[2]:.2f} ms -> {B/min(ms)*1e3:.0f} QPs/s; iterations {int(it.sum())}, >=500: {int((it>=500).sum())}, ==4000: {int((it==4000).sum())}; launches {eng.last_launches}", flush=True)
