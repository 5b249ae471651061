"""Quick parity + timing of a development build: python scratch/qt.py LIB [B,R ...]  (compares with the golden/oracle-free
reference = the installed product library's previous results is not available, so compare against liboracle reference binary on a subset)"""
import sys, os; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
lib = sys.argv[1]
if lib != "-": engine.LIB_PATH = os.path.abspath(lib)
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf
eng = engine.Engine(0)
cases = [(1024, 4), (16384, 4)] if len(sys.argv) < 3 else [tuple(map(int, a.split(","))) for a in sys.argv[2:]]
orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
for B, R in cases:
    mb = W.static_batch(B, num_obs=R)
    for hist in (0, 1):
        eng.use_history(bool(hist))
        ms = []
        for rep in range(3):
            out = eng.solve_mpc_batch(mb); ms.append(eng.last_solve_kernel_ms)
        it = out["iter"]
        print(f"B={B} R={R} hist={hist} path={eng.last_path} solve ms {min(ms):.3f} -> {B/(min(ms)*1e-3):.0f} QP/s; iters sum {it.sum()} max {it.max()}; us/iter(straggler) {min(ms)*1e3/it.max():.3f}", flush=True)
    nchk = min(B, 256)
    sub = mb.slice(0, nchk)
    ref = orc.solve_batch(to_qp_batch(sub), want_y=False)
    ok = (out["status"][:nchk] == ref["status"]).all() and (out["iter"][:nchk] == ref["iter"]).all()
    ex = rel_inf(out["x"][:nchk], ref["x"]).max(); eo = np.abs((out["obj"][:nchk] - ref["obj"]) / ref["obj"]).max()
    print(f"   parity vs {orc.kind} on first {nchk}: status/iter equal={ok} x_err={ex:.2e} obj_err={eo:.2e}", flush=True)
