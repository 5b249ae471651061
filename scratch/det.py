import sys; sys.path.insert(0, ".")
import numpy as np
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
mb = W.static_batch(1024, num_obs=4)
eng.use_history(False)
a = eng.solve_mpc_batch(mb)
a2 = eng.solve_mpc_batch(mb)
print("no-history repeat identical:", np.array_equal(a["x"], a2["x"]), np.array_equal(a["iter"], a2["iter"]))
eng.use_history(True)
b = eng.solve_mpc_batch(mb)
c = eng.solve_mpc_batch(mb)
for name, o in (("hist-1st", b), ("hist-2nd", c)):
    same = np.array_equal(a["x"], o["x"])
    di = np.where(a["iter"] != o["iter"])[0]
    dx = np.where(np.abs(a["x"] - o["x"]).max(axis=1) > 0)[0]
    print(name, "identical:", same, "iter diff at", di[:10], "x diff at", dx[:10], "count", len(dx))
    for i in dx[:5]:
        print("   inst", i, "iters", a["iter"][i], o["iter"][i], "status", a["status"][i], o["status"][i], "maxdiff", np.abs(a["x"][i] - o["x"][i]).max())
# single-instance (solo path) vs batch
for i in list(np.where(a["iter"] == 4000)[0][:3]) + [0, 1, 2]:
    o = eng.solve_mpc_batch(mb.slice(int(i), int(i) + 1))
    print("solo", i, "iters", o["iter"][0], a["iter"][i], "xdiff", np.abs(o["x"][0] - a["x"][i]).max())
