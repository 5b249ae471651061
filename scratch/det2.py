import sys; sys.path.insert(0, ".")
import numpy as np, torch
from intent_mpc_b200 import engine, workloads as W
eng = engine.Engine(0)
B, R = 1024, 4
mb = W.static_batch(B, num_obs=R)
p = mb.params; n = p.n
st = engine.default_settings()
dev = torch.device("cuda", 0)
names = ["x0", "xref", "obs_c", "obs_semi", "obs_yaw", "lin_pt", "warm_x"]
host = {k: np.ascontiguousarray(getattr(mb, k), dtype=np.float64) for k in names}
din = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
dout = {"x": torch.empty((B, n), dtype=torch.float64, device=dev), "status": torch.empty(B, dtype=torch.int32, device=dev),
        "iter": torch.empty(B, dtype=torch.int32, device=dev), "rho_updates": torch.empty(B, dtype=torch.int32, device=dev),
        "obj": torch.empty(B, dtype=torch.float64, device=dev), "pri_res": torch.empty(B, dtype=torch.float64, device=dev),
        "dua_res": torch.empty(B, dtype=torch.float64, device=dev)}
ptrs = {k: int(v.data_ptr()) for k, v in {**din, **dout}.items()}
torch.cuda.synchronize()
ref = eng.solve_mpc_batch(mb)   # host path (first call: no history)
for step in range(5):
    eng.solve_mpc_batch_ptr(p, st, B, R, ptrs, mb.obs_dyn, device=True); eng.sync()
    it = dout["iter"].cpu().numpy(); x = dout["x"].cpu().numpy()
    d = np.where(it != ref["iter"])[0]
    print("device step", step, "iter mismatches", len(d), d[:8], it[d[:8]], ref["iter"][d[:8]], "ms", eng.last_kernel_ms, "xdiff", np.abs(x - ref["x"]).max())
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
stream = torch.cuda.ExternalStream(eng.stream, device=dev)
for step in range(40):
    with torch.cuda.stream(stream):
        flush.zero_()
    eng.solve_mpc_batch_ptr(p, st, B, R, ptrs, mb.obs_dyn, device=True); eng.sync()
    it = dout["iter"].cpu().numpy(); x = dout["x"].cpu().numpy()
    d = np.where(it != ref["iter"])[0]
    print("flushed device step", step, "iter mismatches", len(d), d[:8], it[d[:8]], ref["iter"][d[:8]], "xdiff", np.abs(x - ref["x"]).max(), "at", np.abs(x - ref["x"]).max(axis=1).argmax())
for step in range(4):
    o = eng.solve_mpc_batch(mb)
    d = np.where(o["iter"] != ref["iter"])[0]
    print("host step", step, "iter mismatches", len(d), d[:8], "xdiff", np.abs(o["x"] - ref["x"]).max())
print("inst 953: iters", ref["iter"][953], "status", ref["status"][953], "rho_updates", ref["rho_updates"][953])
from oracle import bindings as OB
from tests.helpers import to_qp_batch
orc = OB.RefOsqp().solve_batch(to_qp_batch(mb.slice(953, 954)), want_y=False)
print("oracle 953:", orc["iter"], orc["status"], orc["rho_updates"])
