import sys; sys.path.insert(0, ".")
import numpy as np, torch
from intent_mpc_b200 import engine
from intent_mpc_b200.receding import IntentSweep
eng = engine.Engine(0); dev = torch.device("cuda", 0)
sw = IntentSweep(300, seed0=321)
mode = sys.argv[1]
if mode in ("enum", "both"):
    sw.first = False
    batches, meta = sw.candidates()
    p, S, D = sw.p, sw.S, sw.D; N = p.N
    pp, ps = sw.last["pp"], sw.last["ps"]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    t_pp, t_ps, t_prob, t_pos = d(pp), d(ps), d(sw.prob), d(sw.pos)
    out = {"scen_a": torch.empty(4 * S, dtype=torch.int32, device=dev), "scen_b": torch.empty(2 * S, dtype=torch.int32, device=dev),
           "obs_c_a": torch.empty((4 * S, N, D, 3), dtype=torch.float64, device=dev), "obs_semi_a": torch.empty((4 * S, N, D, 3), dtype=torch.float64, device=dev),
           "obs_c_b": torch.empty((2 * S, N, D + 1, 3), dtype=torch.float64, device=dev), "obs_semi_b": torch.empty((2 * S, N, D + 1, 3), dtype=torch.float64, device=dev),
           "weight": torch.empty((S, 6), dtype=torch.float64, device=dev), "cand": torch.empty((S, 6), dtype=torch.int32, device=dev)}
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    ptrs.update(pred_pos=t_pp.data_ptr(), pred_size=t_ps.data_ptr(), prob=t_prob.data_ptr(), pos=t_pos.data_ptr(), prev_plan=0)
    eng.intent_candidates_ptr(p, S, D, pp.shape[3], ptrs); eng.sync(); print("enum ok")
    if mode == "both":
        t_x0 = d(np.concatenate([sw.pos, sw.vel], axis=1)); g_x0 = torch.empty((4 * S, 6), dtype=torch.float64, device=dev)
        eng.gather_rows_ptr(4 * S, 6, out["scen_a"].data_ptr(), t_x0.data_ptr(), g_x0.data_ptr()); eng.sync(); print("gather ok")
    sw.first = True
mb = sw.first_step_batch()
print("solving first-step batch", mb.B, mb.num_obs)
o = eng.solve_mpc_batch(mb); print("solve ok", o["iter"][:5], eng.last_launches)
