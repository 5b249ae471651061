"""Development check of a library build (MPCQP_B200_LIB=devlibs/x.so python tools/dev_check.py [big]): iteration time with an
SM per QP, the headline step, parity of one headline batch against the reference binary, optionally the 16k batch."""
import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf
import bench
eng = engine.Engine(0); eng.use_history(False)
s = engine.default_settings(); s.eps_abs = 1e-12; s.eps_rel = 1e-12; s.max_iter = 1000
mb = W.static_batch(64, num_obs=4)
ms = []
for _ in range(3):
    out = eng.solve_mpc_batch(mb, settings=s); ms.append(eng.last_solve_kernel_ms)
print(f"solo: 64 QPs x 1000 iterations: {min(ms):.4f} ms -> {min(ms):.4f} us/iter (incl. setup, 3 factorisations, 40 checks)", flush=True)
batches = [W.static_batch(1024, num_obs=4, seed0=bench.batch_seed(0, 1024, j)) for j in range(8)]
for mbj in batches[:3]: eng.solve_mpc_batch(mbj)
ms = []
for rep in range(2):
    for mbj in batches:
        out = eng.solve_mpc_batch(mbj); ms.append(eng.last_solve_kernel_ms)
ms = np.array(ms)
print(f"headline: mean {ms.mean():.3f} ms (min {ms.min():.3f}, max {ms.max():.3f}) -> {1024/ms.mean()*1e3:.0f} QPs/s", flush=True)
eng.use_history(True)
for _ in range(3): out = eng.solve_mpc_batch(batches[0])
print(f"with history: {eng.last_solve_kernel_ms:.3f} ms", flush=True)
eng.use_history(False)
out = eng.solve_mpc_batch(batches[0])
orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
ref = orc.solve_batch(to_qp_batch(batches[0]), want_y=False)
print("parity vs", orc.kind, ": status equal", bool((out["status"] == ref["status"]).all()), "iter equal", bool((out["iter"] == ref["iter"]).all()),
      "x err", float(rel_inf(out["x"], ref["x"]).max()), "obj err", float(np.abs((out["obj"] - ref["obj"]) / ref["obj"]).max()), flush=True)
if len(sys.argv) > 1:
    mbl = W.static_batch(16384, num_obs=4)
    ms = []
    for _ in range(3):
        o = eng.solve_mpc_batch(mbl); ms.append(eng.last_solve_kernel_ms)
    print(f"16k batch: {min(ms):.2f} ms -> {16384/min(ms)*1e3:.0f} QPs/s, iterations {int(o['iter'].sum())}", flush=True)
    sb, smeta = W.sweep_batches(0, 8192, one_launch=True)
    for _ in range(2):
        sms = 0.0
        for _, smb in sb:
            so = eng.solve_mpc_batch(smb); sms += eng.last_solve_kernel_ms
    print(f"sweep slice 8192: {sms:.2f} ms -> {8192/sms*1e3:.0f} QPs/s", flush=True)
