import sys, time; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
import ctypes as C
tf = C.c_double(); eng.lib.mpcqp_fp64_fma_peak(eng.h, C.byref(tf)); print("fp64 fma peak TFLOP/s", tf.value)
for B, R in ((1024, 4), (1024, 0), (8192, 4)):
    t = time.time(); mb = W.static_batch(B, num_obs=R); tg = time.time() - t
    for rep in range(3):
        t = time.time(); out = eng.solve_mpc_batch(mb); dt = time.time() - t
        print(f"B={B} R={R} gen {tg:.1f}s  e2e {dt*1e3:.2f} ms  kernel {eng.last_kernel_ms:.3f} ms  -> {B/ (eng.last_kernel_ms*1e-3):.0f} QP/s kernel; iters sum {out['iter'].sum()} max {out['iter'].max()} status {dict(zip(*np.unique(out['status'], return_counts=True)))}")
