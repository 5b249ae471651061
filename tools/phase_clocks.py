import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
engine.LIB_PATH = "devlibs/libmpcqp_timing.so"
eng = engine.Engine(0)
names = ["setup", "leaf+rhs(warp0)", "pcr_factor", "load", "iterate", "info+check", "park+adapt", "store", "pcr:init", "pcr:invert", "pcr:products", "pcr:final", "setup:ruiz", "-", "-", "-"]
for B in (1, 1024, 16384):
    mb = W.static_batch(max(B, 1024), num_obs=4).slice(25, 26) if B == 1 else W.static_batch(B, num_obs=4)
    eng.use_history(False); eng.use_migration(False)
    out = eng.solve_mpc_batch(mb); out = eng.solve_mpc_batch(mb)
    buf = np.zeros((mb.B, 16), dtype=np.int64)
    eng.lib.mpcqp_debug_phase_clocks(eng.h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(mb.B))
    it = out["iter"]; nf = 1 + out["rho_updates"]
    tot = buf.sum(axis=0)
    print(f"B={mb.B}: iters {it.sum()} factorizations {nf.sum()} checks {np.ceil(it/25).sum():.0f}")
    for i, n in enumerate(names):
        per = {"iterate": it.sum(), "leaf+rhs(warp0)": nf.sum(), "pcr_factor": nf.sum(), "pcr:init": nf.sum(), "pcr:invert": nf.sum(), "pcr:products": nf.sum(), "pcr:final": nf.sum(), "load": nf.sum(), "info+check": np.ceil(it / 25).sum(), "park+adapt": max(out["rho_updates"].sum(), 1)}.get(n, mb.B)
        print(f"   {n:18s} total {tot[i]/1e6:9.2f} Mcycles  share {100*tot[i]/tot.sum():5.1f}%   per event {tot[i]/per:9.0f} cycles")
