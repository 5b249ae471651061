"""configs[4] slice on the wide kernel: kernel time without / with the iteration-count history of the same slots."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = engine.Engine(0)
sb, _ = W.sweep_batches(0, B, one_launch=True)
for hist in (False, True, True):
    eng.use_history(hist)
    ms = 0.0; it = 0
    for _, smb in sb:
        o = eng.solve_mpc_batch(smb); ms += eng.last_solve_kernel_ms; it += int(o["iter"].sum())
    print(f"B={B} history={hist}: {ms:.2f} ms -> {B/ms*1e3:.0f} QPs/s; iterations {it}; launches {eng.last_launches}", flush=True)
