#!/usr/bin/env python
"""bench.py — batched MPC QP solves/sec on N B200s (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own OSQP binary on the host cores

A "step" is one pass of the hot path over one batch: device-side assembly of B mpcPlanner QPs + the batched
ADMM solve.  Workload = BASELINE.json configs[1]: B = 1,024 randomised default-shape QPs per GPU (horizon 30,
4 static obstacles, warm-started from the constant-velocity plan), seeds sharded by rank (weak scaling, no
collective on the solve path; one NCCL all_gather of the solutions after the timed region for verification).
Every step solves a DIFFERENT batch (NB pre-generated batches in rotation, seeds disjoint) and the engine's scheduling hint
from the previous call is switched off: the steps are independent batches, as the metric says, and nothing about a batch
is known before it is solved.  `value` is timed with CUDA events on the engine's stream with inputs resident in HBM and
the L2 flushed between steps; `e2e` is the same metric through mpcqp_solve_mpc_batch_host with pinned host buffers (H2D
and D2H inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "batched MPC QP solves/sec"
UNIT = "QPs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="QPs per GPU per step (configs[1]: 1024)")
    ap.add_argument("--num-obs", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep-strong", action="store_true", help="skip the strong-scaling configs[4] leg of the default workload")
    ap.add_argument("--sweep-strong-instances", type=int, default=131072, help="total sweep instances of that leg (sharded over the ranks)")
    ap.add_argument("--workload", default="static", choices=["static", "sweep", "receding", "polytraj"],
                    help="static: configs[1], the bench line the driver reads (default).  sweep: configs[4], --instances Monte-Carlo "
                         "instances sharded by index over the ranks (strong scaling), one JSON line of the same shape.  receding: "
                         "configs[2], 10,923 scenarios x 6 intent candidates per rank, --steps warm-started control steps")
    ap.add_argument("--paths", type=int, default=1000, help="--workload polytraj: candidate paths per step (3 QPs each: x, y, z)")
    ap.add_argument("--segments", type=int, default=8, help="--workload polytraj: path segments K (n = 8K coefficients per axis)")
    ap.add_argument("--instances", type=int, default=1000000, help="--workload sweep: total instances over all ranks")
    ap.add_argument("--chunk", type=int, default=32768, help="--workload sweep: instances per engine call")
    ap.add_argument("--device-loop", action="store_true", help="--workload receding: evaluate the obstacle predictions on the device too instead of uploading them from the host every step "
                                                                "(intent-mpc_b200/receding_device.py); e2e is then the wall clock of the whole control steps")
    return ap.parse_args()


NB = 8            # distinct batches in rotation (step i solves batch i % NB)


def workload_name(B, R):
    return (f"configs[1]: {B} randomised default-shape QPs per GPU (horizon 30, n=385, m={625 + 29 * R}, {R} static obstacles, warm start); "
            f"{NB} distinct batches in rotation, one per step")


def batch_seed(rank, B, j):
    """Seed of the first instance of batch j on `rank`: disjoint instance ranges for every (rank, batch)."""
    return (j * 64 + rank) * B


# ---- algorithmic work per QP --------------------------------------------
def algorithmic_flops(N, R, iters, rho_updates):
    b, n, m = 13, 8 * (N + 1) + 5 * N, 16 * (N + 1) + 5 * N + R * N
    nnzA, nnzP = 240 + 17 * N + n + 4 * N * R, 295 if N == 29 else 6 * (N + 1) + 5 * N
    f_iter = 6 * b * b * N + 6 * 64 + 4 * nnzA + 2 * n + 10 * m
    f_fact = (7.0 / 3.0) * b ** 3 * N + 32 * R * N + 2 * nnzA
    f_chk = 4 * nnzA + 2 * nnzP + 6 * (n + m)
    f_scale = 10 * (6 * (nnzP + 2 * nnzA) + 4 * (n + m))
    it = np.asarray(iters, dtype=np.float64); nf = 1.0 + np.asarray(rho_updates, dtype=np.float64)
    return f_scale + nf * f_fact + it * f_iter + np.ceil(it / 25.0) * f_chk


def algorithmic_bytes(N, R):
    return (870 + 174 * R) * 8.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.rows = []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: libosqp.so (OSQP 0.6.2) through its C API exactly as
    OsqpEigen::Solver drives it (oracle/ref_driver.c), one QP per thread on all host cores, setup on the clock."""
    if rank != 0:
        return
    from intent_mpc_b200 import workloads as W
    from oracle import bindings as OB
    from tests.helpers import to_qp_batch
    kind = "reference" if OB.RefOsqp.available() else "port"
    orc = OB.RefOsqp() if kind == "reference" else OB.PortOsqp()
    cores = os.cpu_count() or 1
    B, R = args.batch, args.num_obs
    qbs = [to_qp_batch(W.static_batch(B, num_obs=R, seed0=batch_seed(0, B, j))) for j in range(min(NB, max(args.steps, 1)))]
    for i in range(max(args.warmup, 0)):
        orc.solve_batch(qbs[i % len(qbs)], want_y=False, nthreads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = orc.solve_batch(qbs[i % len(qbs)], want_y=False, nthreads=cores)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(B, R), "pins": "adaptive_rho_interval=25,time_limit=0"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"the full {B}-QP batch per step (the same rotation of batches as the GPU arm), one QP per thread from a shared work queue, osqp_setup+warm_start+solve+cleanup per QP"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "p50_latency_ms": float(np.median(out["wall_time"]) * 1e3), "status_hist": _hist(out["status"])}
    print(json.dumps(line), flush=True)


def _hist(a):
    v, c = np.unique(np.asarray(a), return_counts=True)
    return {str(int(k)): int(n) for k, n in zip(v, c)}


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from intent_mpc_b200 import engine, workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = engine.Engine(local_rank)
    B, R, K, Wm = args.batch, args.num_obs, args.steps, max(args.warmup, 3)
    nb = min(NB, max(K, 1))
    mbs = [W.static_batch(B, num_obs=R, seed0=batch_seed(rank, B, j)) for j in range(nb)]
    mb = mbs[0]
    p = mb.params
    N, n, m = p.N, p.n, p.m(R)
    st = engine.default_settings()
    dev = torch.device("cuda", local_rank)
    names = ["x0", "xref", "obs_c", "obs_semi", "obs_yaw", "lin_pt", "warm_x"]
    hosts = [{k: np.ascontiguousarray(getattr(b_, k), dtype=np.float64) for k in names} for b_ in mbs]
    dins = [{k: torch.from_numpy(v).to(dev) for k, v in h_.items()} for h_ in hosts]
    def outbufs(pin=False):
        o = {"x": torch.empty((B, n), dtype=torch.float64), "status": torch.empty(B, dtype=torch.int32), "iter": torch.empty(B, dtype=torch.int32),
             "rho_updates": torch.empty(B, dtype=torch.int32), "obj": torch.empty(B, dtype=torch.float64), "pri_res": torch.empty(B, dtype=torch.float64),
             "dua_res": torch.empty(B, dtype=torch.float64)}
        return {k: (v.pin_memory() if pin else v.to(dev)) for k, v in o.items()}
    douts = [outbufs() for _ in range(nb)]
    ptrs = [{k: int(v.data_ptr()) for k, v in {**dins[j], **douts[j]}.items()} for j in range(nb)]
    dout = douts[0]
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MiB > 126 MB L2
    torch.cuda.synchronize()
    # independent batches: what the previous call did says nothing about this one (the receding-horizon hint is measured
    # separately, extras.headline_same_batch_with_history)
    eng.use_history(False)

    def step_device(j=0):
        eng.solve_mpc_batch_ptr(p, st, B, R, ptrs[j], mbs[j].obs_dyn, device=True)

    for i in range(Wm):
        step_device(i % nb)
    eng.sync()
    peak_tf = eng.fp64_fma_peak_tflops()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    solve_ms = []
    launches = 0
    for i, (a, b) in enumerate(evs):
        with torch.cuda.stream(stream):
            flush.zero_()                                  # L2 flush, outside the timed pair
            a.record(stream)
        step_device(i % nb)
        with torch.cuda.stream(stream):
            b.record(stream)
        eng.sync()
        solve_ms.append(eng.last_solve_kernel_ms)
        launches += eng.last_launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)

    # iteration counts of exactly the batches the timed steps solved (batch j was solved ceil-or-floor(K / nb) times)
    reps = [len(range(j, K, nb)) for j in range(nb)]
    it_all = [douts[j]["iter"].cpu().numpy() for j in range(nb)]; ru_all = [douts[j]["rho_updates"].cpu().numpy() for j in range(nb)]
    iters = it_all[0]; rhou = ru_all[0]; status = douts[0]["status"].cpu().numpy()
    flops_timed = float(sum(reps[j] * algorithmic_flops(N, R, it_all[j], ru_all[j]).sum() for j in range(nb)))
    iters_timed = int(sum(reps[j] * int(it_all[j].sum()) for j in range(nb)))

    # ---- e2e: host (pinned) buffers through the public host entry point --------------------------------
    pins = [{k: torch.from_numpy(v).pin_memory() for k, v in h_.items()} for h_ in hosts]
    pouts = [outbufs(pin=True) for _ in range(nb)]
    hps = [{k: int(v.data_ptr()) for k, v in {**pins[j], **pouts[j]}.items()} for j in range(nb)]
    pin, pout = pins[0], pouts[0]
    for i in range(2):
        eng.solve_mpc_batch_ptr(p, st, B, R, hps[i % nb], mbs[i % nb].obs_dyn, device=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        eng.solve_mpc_batch_ptr(p, st, B, R, hps[i % nb], mbs[i % nb].obs_dyn, device=False)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = sum(v.numel() * v.element_size() for v in pout.values())
    for j in range(min(nb, K)):
        if not np.array_equal(pouts[j]["iter"].numpy(), it_all[j]):
            bad = np.where(pouts[j]["iter"].numpy() != it_all[j])[0]
            raise SystemExit(f"bench.py: host and device entry points disagree on {len(bad)} instances of batch {j}, e.g. {bad[:8]}: "
                             f"host {pouts[j]['iter'].numpy()[bad[:8]]} device {it_all[j][bad[:8]]}")

    # ---- configs[4] as a STRONG-scaling leg on all ranks: a fixed total of sweep instances, index-sharded over the ranks --
    sweep_strong = None
    if not args.no_sweep_strong:
        from intent_mpc_b200 import sharding
        Mtot = args.sweep_strong_instances
        lo_, hi_ = sharding.shard_bounds(Mtot, world)[rank]
        wm_ = W.sweep_batches(lo_, min(lo_ + 2048, hi_), one_launch=True)[0]
        eng.solve_mpc_batch(wm_[0][1])                 # warm-up of the wide kernel
        if world > 1:
            dist.barrier()
        sdev, swall, sit_, shist_, _ = sweep_shard(eng, lo_, hi_, 32768)
        t = torch.tensor([sdev, swall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sweep_strong = {"instances_total": Mtot, "value": Mtot / (float(t[0]) * 1e-3), "unit": UNIT, "kernel_ms_max_over_ranks": float(t[0]),
                        "e2e": {"value": Mtot / (float(t[1]) * 1e-3), "unit": UNIT, "ms_max_over_ranks": float(t[1]),
                                "note": "pinned host buffers through mpcqp_solve_mpc_batch_host, copies inside; generation untimed"},
                        "scaling": "strong", "note": "configs[4]: the same total number of Monte-Carlo sweep instances at every N, contiguous index shards, no collective"}

    # ---- verification gather (NCCL over NVLink; not on the solve path, not timed into `value`) ------------
    gather_ms = None
    if world > 1:
        allx = torch.empty((world * B, n), dtype=torch.float64, device=dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); dist.all_gather_into_tensor(allx, dout["x"]); g1.record(); torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.join(timeout=2)

    # ---- parity spot check against the oracle (untimed) and the CPU baseline ---------------------------
    parity = cpu = None
    p50_qp_ms = None
    try:
        from oracle import bindings as OB
        from tests.helpers import to_qp_batch, rel_inf
        kind = "reference" if OB.RefOsqp.available() else "port"
        orc = OB.RefOsqp() if kind == "reference" else OB.PortOsqp()
        cores = os.cpu_count() or 1
        qb = to_qp_batch(mb)
        if not args.no_cpu_baseline:
            orc.solve_batch(qb, want_y=False, nthreads=cores)
            ref = orc.solve_batch(qb, want_y=False, nthreads=cores)
            cpu = {"value": B / ref["wall"], "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"the full {B}-QP batch once (after one warm-up pass), one QP per thread; {ref['wall']:.2f} s wall",
                   "p50_latency_ms": float(np.median(ref["wall_time"]) * 1e3)}
            xg = dout["x"].cpu().numpy()
            parity = {"checked": int(B), "status_equal": bool((ref["status"] == status).all()),
                      "iter_equal": bool((ref["iter"] == iters).all()),
                      "x_rel_err_max": float(rel_inf(xg, ref["x"]).max()),
                      "obj_rel_err_max": float(np.abs((dout["obj"].cpu().numpy() - ref["obj"]) / ref["obj"]).max())}
    except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
        parity = {"error": repr(ex)}

    # ---- explanatory extras (device-resident inputs, untimed above): throughput on a batch large enough to fill the GPU,
    # and the latency of single QPs (B = 1: one CTA on an otherwise idle GPU), the "p50 per-QP latency" of the metric
    extras = {}
    try:
        # receding-horizon use: the SAME batch slots again one step later, with the engine's hint from the previous call on
        eng.use_history(True)
        msc = []
        for _ in range(min(K, 5) + 1):
            step_device(0); eng.sync(); msc.append(eng.last_kernel_ms)
        eng.use_history(False)
        extras["headline_same_batch_with_history"] = {"value": B / (float(np.mean(msc[1:])) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(msc[1:])),
                                                      "note": "batch 0 re-solved with the scheduling hint from the previous call of the same slots (what a receding-horizon loop gets); not the headline"}
        Bl = 16384
        ml = W.static_batch(Bl, num_obs=R, seed0=100000 + rank * Bl)
        outl = eng.solve_mpc_batch(ml)
        msl = []
        for _ in range(3):
            eng.solve_mpc_batch(ml, out=outl); msl.append(eng.last_kernel_ms)
        extras["large_batch"] = {"batch_per_gpu": Bl, "value": Bl / (min(msl) * 1e-3), "unit": UNIT, "ms_per_batch": min(msl),
                                 "iterations_total": int(outl["iter"].sum()),
                                 "fp64_tflops": float(algorithmic_flops(N, R, outl["iter"], outl["rho_updates"]).sum()) / (min(msl) * 1e-3) / 1e12,
                                 "note": "device kernels only (assembly + solve), one GPU, not the headline"}
        lat = []
        for i in range(48):
            m1 = mb.slice(i, i + 1)
            eng.solve_mpc_batch(m1); lat.append(eng.last_kernel_ms)
        lat = np.sort(np.array(lat))
        extras["single_qp_latency_ms"] = {"p50": float(lat[len(lat) // 2]), "p90": float(lat[int(0.9 * len(lat))]), "max": float(lat[-1]), "samples": len(lat)}
        p50_qp_ms = float(lat[len(lat) // 2])
        # BASELINE.json configs[4]: a slice of the Monte-Carlo sweep (per-instance obstacle counts up to 32 rows per stage,
        # one launch with per-instance velocity/acceleration limits), device kernels only; the CPU reference on a sample of it
        Bs = 16384
        sb, smeta = W.sweep_batches(rank * Bs, (rank + 1) * Bs, one_launch=True)
        for _ in range(2):
            sms = 0.0; sit = 0; shist = {}
            for _, smb in sb:
                so = eng.solve_mpc_batch(smb); sms += eng.last_kernel_ms; sit += int(so["iter"].sum())
                for k_, v_ in _hist(so["status"]).items():
                    shist[k_] = shist.get(k_, 0) + v_
        sw = {"instances_per_gpu": Bs, "value": Bs / (sms * 1e-3), "unit": UNIT, "ms": sms, "iterations_total": sit, "launches": len(sb),
              "rows_per_stage_cap": smeta["cap"], "capped_instances": smeta["capped"], "status_hist": shist,
              "note": "configs[4] slice; every instance has its own obstacle count (7..32), dynamic/static mix and velocity/acceleration limits; device kernels only"}
        if not args.no_cpu_baseline:
            from oracle import bindings as OB2
            from tests.helpers import oracle_solve
            if OB2.RefOsqp.available():
                o2 = OB2.RefOsqp()
                sg, _ = W.sweep_groups(rank * Bs, rank * Bs + 2048)
                sg = [g_ for g_ in sg if g_[1].num_obs == smeta["cap"]]          # the three large groups: enough QPs per call to fill the cores
                nq = sum(g_[1].B for g_ in sg)
                t0 = time.perf_counter()
                for _, g_ in sg:
                    oracle_solve(o2, g_, nthreads=os.cpu_count() or 1)
                sw["cpu_reference"] = {"value": nq / (time.perf_counter() - t0), "unit": UNIT, "cores": os.cpu_count() or 1,
                                       "sample": f"{nq} instances with {smeta['cap']} rows per stage out of the first 2048, assembly by pattern group + libosqp one QP per thread (includes the numpy assembly)"}
        # the same slice end to end: pinned host buffers through the host entry point (copies inside the timed region)
        (sidx, smb), = sb
        spin = {k: torch.from_numpy(np.ascontiguousarray(getattr(smb, k), dtype=np.float64)).pin_memory() for k in names}
        spout = {"x": torch.empty((smb.B, smb.params.n), dtype=torch.float64).pin_memory(), "status": torch.empty(smb.B, dtype=torch.int32).pin_memory(),
                 "iter": torch.empty(smb.B, dtype=torch.int32).pin_memory(), "rho_updates": torch.empty(smb.B, dtype=torch.int32).pin_memory(),
                 "obj": torch.empty(smb.B, dtype=torch.float64).pin_memory(), "pri_res": torch.empty(smb.B, dtype=torch.float64).pin_memory(),
                 "dua_res": torch.empty(smb.B, dtype=torch.float64).pin_memory()}
        shp = {k: int(v.data_ptr()) for k, v in {**spin, **spout}.items()}
        t0 = time.perf_counter()
        eng.solve_mpc_batch_ptr(smb.params, st, smb.B, smb.num_obs, shp, smb.obs_dyn, device=False, nobs=smb.nobs, limits=smb.limits)
        sw["e2e"] = {"value": smb.B / (time.perf_counter() - t0), "unit": UNIT, "h2d_bytes": int(sum(v.numel() * v.element_size() for v in spin.values())),
                     "d2h_bytes": int(sum(v.numel() * v.element_size() for v in spout.values())), "note": "one call of mpcqp_solve_mpc_batch_host with pinned host buffers"}
        extras["sweep"] = sw
        # BASELINE.json configs[3]: the stress set — doubled horizon (60: n = 775, m = 1373), tight limits, starts outside the box
        # (max_iter), an obstacle around the start; device kernels, end to end through host buffers, and the CPU reference
        Bst = 2048
        stm = W.stress_batch(Bst, seed0=rank * Bst)
        sto = eng.solve_mpc_batch(stm)
        t0 = time.perf_counter(); eng.solve_mpc_batch(stm, out=sto); st_wall = time.perf_counter() - t0
        stress = {"instances_per_gpu": Bst, "value": Bst / (eng.last_kernel_ms * 1e-3), "unit": UNIT, "ms": eng.last_kernel_ms, "kernel_path": eng.last_path,
                  "iterations_total": int(sto["iter"].sum()), "status_hist": _hist(sto["status"]),
                  "e2e": {"value": Bst / st_wall, "unit": UNIT, "note": "mpcqp_solve_mpc_batch_host with pageable numpy buffers"},
                  "note": "configs[3]: horizon 60, max_vel = max_acc = 1.5, z in [1.9, 2.1], 2 obstacles; a quarter of the starts above the box, a quarter too fast"}
        if not args.no_cpu_baseline:
            from oracle import bindings as OB3
            from tests.helpers import to_qp_batch as tq3
            if OB3.RefOsqp.available():
                ns_ = 256
                r3 = OB3.RefOsqp().solve_batch(tq3(stm.slice(0, ns_)), want_y=False, nthreads=os.cpu_count() or 1)
                stress["cpu_reference"] = {"value": ns_ / r3["wall"], "unit": UNIT, "cores": os.cpu_count() or 1, "sample": f"the first {ns_} instances, libosqp one QP per thread",
                                           "status_equal": bool((r3["status"] == sto["status"][:ns_]).all()), "iter_equal": bool((r3["iter"] == sto["iter"][:ns_]).all())}
        extras["stress_h60"] = stress
        # BASELINE.json configs[2]: 10,923 scenarios x 6 intent candidates = 65,538 QPs per control step, warm-started from the
        # plan chosen one step earlier (candidate enumeration / scoring on the host, intent-mpc_b200/receding.py; untimed)
        if rank == 0:
            from intent_mpc_b200 import receding
            from intent_mpc_b200.receding_device import DeviceIntentSweep
            per = {}
            for mode in ("device", "host"):
                ds_ = DeviceIntentSweep(eng, receding.IntentSweep(S=10923, D=4, seed0=5), device=local_rank, host_predictions=(mode == "host"))
                ds_.step(); ds_.step()                   # obstacle-free first step + one step with candidates (warm-up)
                rows = []
                for _ in range(3):
                    m0 = ds_.kernel_ms; b0 = (ds_.h2d_bytes, ds_.d2h_bytes)
                    if mode == "host":
                        ds_.stage_host_predictions()       # the synthetic predictor's numpy time is not the path's: its output waits in pinned memory
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    ds_.step()
                    torch.cuda.synchronize()
                    rows.append((1e3 * (time.perf_counter() - t0), ds_.kernel_ms - m0, int(ds_.buf["iter"].sum().item()), ds_.h2d_bytes - b0[0], ds_.d2h_bytes - b0[1]))
                per[mode] = rows
                del ds_
            nq = 6 * 10923
            extras["receding_horizon"] = {
                "qps_per_step": nq, "value": nq / (per["device"][-1][1] * 1e-3), "unit": UNIT,
                "ms_per_step": [r_[1] for r_ in per["device"]], "iterations_per_step": [r_[2] for r_ in per["device"]],
                "e2e": {"value": nq / (float(np.mean([r_[0] for r_ in per["host"]])) * 1e-3), "unit": UNIT, "ms_per_step": [r_[0] for r_ in per["host"]],
                        "h2d_bytes_per_step": per["host"][-1][3], "d2h_bytes_per_step": per["host"][-1][4],
                        "note": "whole control step by the wall clock with HOST predictions: the predictor's output (pinned host memory) -> H2D, enumeration, gather, two solves, scoring, choice on the device, chosen plan -> D2H"},
                "e2e_device_resident": {"value": nq / (float(np.mean([r_[0] for r_ in per["device"]])) * 1e-3), "unit": UNIT, "ms_per_step": [r_[0] for r_ in per["device"]]},
                "note": "configs[2] at full size, control steps 3-5 of a warm-started loop (intent-mpc_b200/receding_device.py); value = device kernels of the two solve calls per step"}
    except Exception as ex:
        extras["error"] = repr(ex)

    flops = flops_timed / K                                   # per step, averaged over exactly the batches that were timed
    solve_avg_ms = float(np.mean(solve_ms))
    ach_tf = flops / (solve_avg_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None                                   # ncu dram__bytes of one step of this workload (committed capture), if it is this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if B == 1024 and R == 4:
            traffic = int(tj["bytes_per_step"])
    except Exception:
        pass
    hbm_ach = algorithmic_bytes(N, R) * B / (solve_avg_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(B, R), "batch_per_gpu": B, "num_obs": R, "horizon": p.horizon,
                   "pins": "adaptive_rho_interval=25,time_limit=0", "l2": "flushed between steps (256 MiB write)",
                   "schedule": "independent batches: the engine's hint from the previous call is OFF; instances whose fixed start violates a stage-0 obstacle row start first; the two-per-SM launch is a probe (setup, factorisation, 25 iterations); anything still running is parked and resumed bit-identically on an SM of its own (4 solver warps + 3 PCR assistants + a row helper) by a follow-up launch, largest primal residual first; results never depend on scheduling (extras.headline_same_batch_with_history = the receding-horizon case)",
                   "kernel_path": eng.last_path, "iterations_total": iters_timed, "iterations_max": int(max(int(v.max()) for v in it_all))},
        "e2e": {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / K},
        "gpu_launches": int(launches),
        "p50_latency_ms": p50_qp_ms, "p50_latency_note": "per-QP latency: one QP alone on the GPU (B = 1), device kernels, median over 48 instances of batch 0",
        "batch_step_ms_p50": float(np.median(step_ms)),
        "status_hist": _hist(status),
        "roofline": {"bound": "fp64_fma", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per step (all solve launches; ncu flushes the caches between launches, so the parked cold blocks count), ncu capture in profiles/; algorithmic: %d" % int(algorithmic_bytes(N, R) * B),
                     "kernel": "mpcqp_solve_cta_kernel (two-per-SM launch + one-per-SM launches of the hard list and of the parked instances: 4 solver warps, 3 PCR assistants, row helper)", "kernel_ms": solve_avg_ms,
                     "peak_source": "fp64 DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                     "algorithmic_flops_per_launch": flops,
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"}},
        "cpu_baseline": cpu, "parity": parity, "clocks": sampler.summary(), "extras": extras,
    }
    if sweep_strong is not None:
        line["extras"]["sweep_strong_scaling"] = sweep_strong
    if gather_ms is not None:
        line["verification_gather_ms"] = gather_ms
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_shard(eng, lo, hi, chunk):
    """Generate (host, untimed) and solve the sweep instances [lo, hi) chunk by chunk through the host entry point with pinned
    buffers.  Returns device ms of the kernels, wall seconds of the calls (copies inside), iterations, status histogram, launches."""
    import torch
    from intent_mpc_b200 import engine, workloads as W
    dev_ms = 0.0; wall = 0.0; iters = 0; hist = {}; launches = 0
    st = engine.default_settings()
    names = ["x0", "xref", "obs_c", "obs_semi", "obs_yaw", "lin_pt", "warm_x"]
    for c0 in range(lo, hi, chunk):
        (idx, mb), = W.sweep_batches(c0, min(c0 + chunk, hi), one_launch=True)[0]
        B, R, n = mb.B, mb.num_obs, mb.params.n
        # pinned staging of the generated chunk (untimed, like the generation itself): the timed call then is what a caller
        # with page-locked buffers pays — host -> device copies, kernels, device -> host copies
        pin = {k: torch.from_numpy(np.ascontiguousarray(getattr(mb, k), dtype=np.float64)).pin_memory() for k in names}
        pout = {"x": torch.empty((B, n), dtype=torch.float64).pin_memory(), "status": torch.empty(B, dtype=torch.int32).pin_memory(),
                "iter": torch.empty(B, dtype=torch.int32).pin_memory(), "rho_updates": torch.empty(B, dtype=torch.int32).pin_memory(),
                "obj": torch.empty(B, dtype=torch.float64).pin_memory(), "pri_res": torch.empty(B, dtype=torch.float64).pin_memory(),
                "dua_res": torch.empty(B, dtype=torch.float64).pin_memory()}
        hp = {k: int(v.data_ptr()) for k, v in {**pin, **pout}.items()}
        t0 = time.perf_counter()
        eng.solve_mpc_batch_ptr(mb.params, st, B, R, hp, mb.obs_dyn, device=False, nobs=mb.nobs, limits=mb.limits)
        wall += time.perf_counter() - t0
        out = {k: v.numpy() for k, v in pout.items()}
        dev_ms += eng.last_kernel_ms; launches += eng.last_launches; iters += int(out["iter"].sum())
        for k_, v_ in _hist(out["status"]).items():
            hist[k_] = hist.get(k_, 0) + v_
    return dev_ms, wall, iters, hist, launches


def run_sweep(args, rank, world, local_rank):
    """BASELINE.json configs[4]: --instances sweep instances (intent-mpc_b200/workloads.py:sweep_batches), contiguous index
    shards over the ranks, no collective on the solve path.  Each rank generates its shard chunk by chunk on the host
    (untimed), solves every chunk through the host entry point (e2e: pinned host buffers, copies inside) and sums the
    device time of its kernels (`value`); the job's time is the max over ranks."""
    import torch
    import torch.distributed as dist
    from intent_mpc_b200 import engine, workloads as W, sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = engine.Engine(local_rank)
    eng.use_history(False)                              # independent instances: slot history means nothing here
    lo, hi = sharding.shard_bounds(args.instances, world)[rank]
    warm = W.sweep_batches(lo, min(lo + 2048, hi), one_launch=True)[0]
    eng.solve_mpc_batch(warm[0][1])                    # warm-up (kernel load, buffers)
    dev_ms, wall, iters, hist, launches = sweep_shard(eng, lo, hi, args.chunk)
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=torch.device("cuda", local_rank))
    agg = torch.tensor([float(iters), float(hi - lo)], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    if rank == 0:
        dev_ms, wall_ms = float(t[0]), float(t[1])
        n = int(agg[1])
        line = {"metric": METRIC, "value": n / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"configs[4]: {n} Monte-Carlo sweep instances (run_mpc_benchmark grid: 50/100/200 obstacles, 65 % dynamic, "
                                       f"limits (1.5,1.5)/(3,3)/(5,20)), up to {W.SWEEP_CAP} obstacle rows per stage, sharded by index",
                           "chunk": args.chunk, "pins": "adaptive_rho_interval=25,time_limit=0", "iterations_total": int(agg[0]),
                           "status_hist_rank0": hist},
                "e2e": {"value": n / (wall_ms * 1e-3), "unit": UNIT, "ms": wall_ms, "note": "pinned host buffers through mpcqp_solve_mpc_batch_host, host<->device copies inside; generation of the instances untimed"},
                "gpu_launches": int(launches)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_receding(args, rank, world, local_rank):
    """BASELINE.json configs[2]: per rank 10,923 scenarios x 6 intent candidates = 65,538 QPs per control step, --steps
    control steps of the warm-started receding-horizon loop (intent-mpc_b200/receding_device.py: enumeration, gather, two
    solves, scoring and choice are engine calls on device arrays).  `value` = QPs per second of device kernel time of the solve
    calls, summed over the steps; `e2e` = the wall clock of the whole control steps — with --device-loop the obstacle
    predictions are evaluated on the device too (nothing crosses PCIe), without it they come from the host every step and
    the chosen plan goes back (host buffers inside the timed region).  The max over ranks is the job's time."""
    import torch
    import torch.distributed as dist
    from intent_mpc_b200 import engine, receding
    from intent_mpc_b200.receding_device import DeviceIntentSweep
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = engine.Engine(local_rank)
    S = 10923
    ds = DeviceIntentSweep(eng, receding.IntentSweep(S=S, D=4, seed0=1000 * rank + 5), device=local_rank, host_predictions=not args.device_loop)
    ds.step(); ds.step()                            # first (obstacle-free) step and one warm-up step with candidates
    ds.kernel_ms = 0.0; ds.h2d_bytes = ds.d2h_bytes = 0
    per_step = []; its = 0; hist = {}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall_ms = 0.0
    for _ in range(args.steps):
        m0 = ds.kernel_ms
        if not args.device_loop:
            ds.stage_host_predictions()                # the synthetic predictor (numpy) is not part of the step: its output waits in pinned memory
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ds.step()
        torch.cuda.synchronize()
        wall_ms += (time.perf_counter() - t0) * 1e3
        per_step.append(ds.kernel_ms - m0); its += int(ds.buf["iter"].sum().item())
        for k_, v_ in _hist(ds.buf["status"].cpu().numpy()).items():
            hist[k_] = hist.get(k_, 0) + v_
    dev = torch.device("cuda", local_rank)
    t = torch.tensor([ds.kernel_ms, wall_ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(6 * S * args.steps), float(its)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    if rank == 0:
        where = "on the device (nothing crosses PCIe)" if args.device_loop else "on the host, uploaded every step; the chosen plan is downloaded every step"
        line = {"metric": METRIC, "value": float(agg[0]) / (float(t[0]) * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": 2,
                "ms_per_step": float(t[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"configs[2]: {S} scenarios x 6 intent candidates = {6 * S} QPs per control step per GPU, {args.steps} "
                                       "warm-started receding-horizon steps (4 dynamic obstacles with 4 intent predictions each; candidates 4, 5 carry the closest "
                                       f"obstacle twice); enumeration, gather, solves, scoring and choice are engine calls on device arrays; predictions {where}",
                           "pins": "adaptive_rho_interval=25,time_limit=0", "iterations_total": int(agg[1]), "status_hist_rank0": hist,
                           "ms_per_step_rank0": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)},
                           "progress_m_rank0": float(ds.pos[:, 0].mean().item())},
                "e2e": {"value": float(agg[0]) / (float(t[1]) * 1e-3), "unit": UNIT, "ms": float(t[1]),
                        "h2d_bytes_per_step": int(ds.h2d_bytes // max(args.steps, 1)), "d2h_bytes_per_step": int(ds.d2h_bytes // max(args.steps, 1)),
                        "note": "wall clock of the whole control steps (upload of the predictions from pinned host memory unless --device-loop, enumeration, gather, two solves, scoring, choice, state update)"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_polytraj(args, rank, world, local_rank):
    """SURVEY.md section 8(f) row 3 — the boundary's second consumer: `--paths` candidate paths of `--segments` segments, i.e.
    3 x paths minimum-snap QPs of polyTrajSolver's shape (intent-mpc_b200/polytraj_workload.py), through the batched generic entry
    point (mpcqp_solve_qp_batch_host: host buffers in and out, one launch of the dense kernel).  Not the driver's bench line.
    `value`: kernel time (CUDA events of the engine, inputs resident); `e2e`: wall clock of the host call.  One GPU."""
    from intent_mpc_b200 import polytraj_workload as PA
    qb = PA.path_batch(args.paths, K=args.segments, seed0=100)
    B = int(qb.q.shape[0])
    work = f"polyTrajSolver QPs: {args.paths} paths x 3 axes, K={args.segments} segments, n={qb.n}, m={qb.m} (minimum snap, degree 7, C4 continuity)"
    if args.impl == "reference":
        from oracle import bindings as OB
        orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
        cores = os.cpu_count() or 1
        for _ in range(args.warmup):
            orc.solve_batch(qb, want_y=False, nthreads=cores)
        ts = [orc.solve_batch(qb, want_y=False, nthreads=cores)["wall"] for _ in range(args.steps)]
        ms = 1e3 * float(np.mean(ts))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": work, "pins": "adaptive_rho_interval=25,time_limit=0"},
                          "cpu_baseline": {"value": B / (ms * 1e-3), "unit": UNIT, "cores": cores, "kind": orc.kind, "sample": "the whole batch per step, one QP per thread, setup on the clock"}}), flush=True)
        return
    from intent_mpc_b200 import engine
    eng = engine.Engine(local_rank)
    for _ in range(max(args.warmup, 1)):
        r = engine.solve_qp_batch(eng, qb, want_y=False)
    kms = []; wms = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); r = engine.solve_qp_batch(eng, qb, want_y=False); wms.append(1e3 * (time.perf_counter() - t0)); kms.append(eng.last_kernel_ms)
    line = {"metric": METRIC, "value": B / (np.mean(kms) * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(kms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": work, "pins": "adaptive_rho_interval=25,time_limit=0", "kernel_path": eng.last_path,
                       "iterations_total": int(r["iter"].sum()), "iterations_max": int(r["iter"].max()), "status_hist": _hist(r["status"])},
            "e2e": {"value": B / (np.mean(wms) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(wms)),
                    "note": "pageable host arrays through mpcqp_solve_qp_batch_host, host<->device copies inside"},
            "gpu_launches": int(eng.last_launches) * args.steps}
    if not args.no_cpu_baseline:
        from oracle import bindings as OB
        orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
        cores = os.cpu_count() or 1
        orc.solve_batch(qb, want_y=False, nthreads=cores)
        c = orc.solve_batch(qb, want_y=False, nthreads=cores)
        d = np.abs(c["x"]).max(axis=1)
        line["cpu_baseline"] = {"value": B / c["wall"], "unit": UNIT, "cores": cores, "kind": orc.kind, "sample": "the same batch once, one QP per thread"}
        line["parity"] = {"checked": B, "status_equal": bool((r["status"] == c["status"]).all()), "iter_equal": bool((r["iter"] == c["iter"]).all()),
                          "x_rel_err_max": float((np.abs(r["x"] - c["x"]).max(axis=1) / d).max())}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "polytraj":
        if rank == 0:
            run_polytraj(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "sweep":
        run_sweep(args, rank, world, local_rank)
    elif args.workload == "receding":
        run_receding(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
