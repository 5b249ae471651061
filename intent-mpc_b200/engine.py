"""Host-side Python mirror of the C ABI (include/mpcqp_b200.h) — ctypes only, no solve logic.

`Engine.solve_mpc_batch` is the batched counterpart of `mpcPlanner::solveTraj`
(trajectory_planner/include/trajectory_planner/mpcPlanner.cpp:375-541): same inputs (current state,
reference window, per-stage obstacle ellipsoids, linearisation point, warm start), same outputs
(stacked states/controls solution, OSQP status), for B instances at once.  There is no CPU fallback:
if the CUDA library is missing or no GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCQP_B200_LIB", os.path.join(HERE, "libmpcqp_b200.so"))   # development builds only

STATUS_NAMES = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate",
                4: "dual infeasible inaccurate", -2: "maximum iterations reached", -3: "primal infeasible",
                -4: "dual infeasible", -7: "non convex", -10: "unsolved"}


class Settings(C.Structure):
    """mpcqp_settings (mirrors OSQPSettings, third_party/osqp/types.h:139-176)."""
    _fields_ = [(k, C.c_double) for k in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf",
                                          "eps_dual_inf", "adaptive_rho_tolerance", "adaptive_rho_fraction",
                                          "delta", "time_limit")] + \
               [(k, C.c_int64) for k in ("max_iter", "scaling", "adaptive_rho", "adaptive_rho_interval",
                                         "check_termination", "warm_start", "scaled_termination", "polish",
                                         "polish_refine_iter", "verbose")]


class MpcParamsC(C.Structure):
    """mpcqp_mpc_params."""
    _fields_ = [("horizon", C.c_int32)] + \
               [(k, C.c_double) for k in ("ts", "max_vel", "max_acc", "y_min", "y_max", "z_min", "z_max",
                                          "static_safety_dist", "dynamic_safety_dist", "static_slack",
                                          "dynamic_slack", "position_weight", "velocity_weight",
                                          "acceleration_weight")]


class PredictorParamsC(C.Structure):
    """mpcqp_predictor_params (predictor_param.yaml keys)."""
    _fields_ = [("prediction_size", C.c_int32)] + \
               [(k, C.c_double) for k in ("prediction_time_step", "min_turning_time", "max_turning_time", "prediction_z_score",
                                          "max_front_prob", "front_angle_deg", "stop_velocity_threshold", "prob_scale_param")]


def default_predictor_params(**overrides) -> PredictorParamsC:
    p = PredictorParamsC()
    load_library().mpcqp_default_predictor_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"unknown predictor parameter {k}")
        setattr(p, k, v)
    return p


class Info(C.Structure):
    _fields_ = [("iter", C.c_int64), ("status_val", C.c_int64), ("rho_updates", C.c_int64),
                ("obj_val", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double),
                ("setup_time", C.c_double), ("solve_time", C.c_double)]


_lib = None


def load_library() -> C.CDLL:
    """Load the in-tree CUDA library; fail loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.mpcqp_engine_last_error.restype = C.c_char_p
        lib.mpcqp_engine_last_kernel_ms.restype = C.c_double
        lib.mpcqp_engine_last_solve_kernel_ms.restype = C.c_double
        lib.mpcqp_engine_last_launches.restype = C.c_int64
        lib.mpcqp_engine_stream.restype = C.c_void_p
        _lib = lib
    return _lib


def default_settings(**overrides) -> Settings:
    s = Settings()
    load_library().mpcqp_set_default_settings(C.byref(s))
    for k, v in overrides.items():
        if not hasattr(s, k):
            raise AttributeError(f"unknown setting {k}")
        setattr(s, k, v)
    return s


def params_to_c(p) -> MpcParamsC:
    c = MpcParamsC()
    for k, _ in MpcParamsC._fields_:
        setattr(c, k, getattr(p, k))
    return c


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


class EngineError(RuntimeError):
    pass


class Engine:
    """One engine per GPU / host thread (mpcqp_engine_create)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.mpcqp_engine_create(C.c_int(device), C.byref(self.h))
        if rc != 0 or not self.h:
            raise EngineError(f"mpcqp_engine_create(device={device}) failed with {rc}: no usable CUDA device "
                              "(this engine has no CPU fallback)")
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcqp_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError(f"mpcqp error {rc}: {self.lib.mpcqp_engine_last_error(self.h).decode()}")

    @property
    def last_kernel_ms(self) -> float:
        return float(self.lib.mpcqp_engine_last_kernel_ms(self.h))

    @property
    def last_solve_kernel_ms(self) -> float:
        return float(self.lib.mpcqp_engine_last_solve_kernel_ms(self.h))

    @property
    def last_launches(self) -> int:
        return int(self.lib.mpcqp_engine_last_launches(self.h))

    @property
    def last_path(self) -> str:
        """'cta' (4-warp CTA per QP, PCR solve), 'fast' (one warp per QP, register-resident), 'generic'
        (one warp per QP, shared-memory kernel for any stage-structured shape), 'band' (unstructured QP, banded KKT factor, one
        warp) or 'dense' (unstructured QP, dense KKT factor, one CTA)."""
        return {2: "cta", 1: "fast", 4: "dense", 5: "band"}.get(int(self.lib.mpcqp_engine_last_path(self.h)), "generic")

    def force_generic(self, on=True):
        """True / 1 / 'generic': generic kernel; 2 / 'fast': one-warp register kernel; 3 / 'cta_plain': CTA kernel without
        the assistant warps on one-per-SM launches; False / 0 / 'cta': default dispatch."""
        names = {"generic": 1, "fast": 2, "cta": 0, "cta_plain": 3, "dense": 4}
        code = names[on] if isinstance(on, str) else (1 if on is True else int(on))
        self._check(self.lib.mpcqp_engine_force_generic(self.h, C.c_int(code)))

    def use_history(self, on: bool = True):
        """Scheduling hint from the previous call's iteration counts (mpcqp_engine_use_history)."""
        self._check(self.lib.mpcqp_engine_use_history(self.h, C.c_int(1 if on else 0)))

    def use_migration(self, on: bool = True):
        self._check(self.lib.mpcqp_engine_use_migration(self.h, C.c_int(1 if on else 0)))

    def fp64_fma_peak_tflops(self) -> float:
        tf = C.c_double()
        self._check(self.lib.mpcqp_fp64_fma_peak(self.h, C.byref(tf)))
        return float(tf.value)

    @property
    def stream(self) -> int:
        return int(self.lib.mpcqp_engine_stream(self.h) or 0)

    def sync(self):
        self._check(self.lib.mpcqp_engine_sync(self.h))

    # ---- batched entry point, host buffers -------------------------------------------------------
    def solve_mpc_batch(self, mb, settings: Settings | None = None, want_y: bool = False, out: dict | None = None):
        """Solve every instance of an `MpcBatch` (workloads.py).  Returns dict(x[B,n], y[B,m]|None,
        status, iter, rho_updates [B] int32, obj, pri_res, dua_res [B])."""
        s = settings or default_settings()
        p = params_to_c(mb.params)
        B, R = mb.B, mb.num_obs
        n, m = mb.params.n, mb.params.m(R)
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x0, xref, oc, os_, oy, lp, wx = map(f, (mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.lin_pt, mb.warm_x))
        od = np.ascontiguousarray(mb.obs_dyn, dtype=np.int32)
        self._check(self.lib.mpcqp_engine_obs_dyn_per_instance(self.h, C.c_int(1 if od.ndim == 3 else 0)))
        nobs = getattr(mb, "nobs", None)
        nobs = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)
        self._check(self.lib.mpcqp_engine_num_obs_per_instance(self.h, _ip(nobs)))
        lim = getattr(mb, "limits", None)
        lim = None if lim is None else np.ascontiguousarray(lim, dtype=np.float64)
        self._check(self.lib.mpcqp_engine_limits_per_instance(self.h, _dp(lim)))
        if out is None:
            out = dict(x=np.empty((B, n)), y=np.empty((B, m)) if want_y else None, status=np.empty(B, np.int32),
                       iter=np.empty(B, np.int32), rho_updates=np.empty(B, np.int32), obj=np.empty(B),
                       pri_res=np.empty(B), dua_res=np.empty(B))
        rc = self.lib.mpcqp_solve_mpc_batch_host(self.h, C.byref(p), C.byref(s), C.c_int32(B), C.c_int32(R), _dp(x0),
                                                 _dp(xref), _dp(oc), _dp(os_), _dp(oy), _ip(od), _dp(lp), _dp(wx),
                                                 _dp(out["x"]), _dp(out["y"]), _ip(out["status"]), _ip(out["iter"]),
                                                 _ip(out["rho_updates"]), _dp(out["obj"]), _dp(out["pri_res"]),
                                                 _dp(out["dua_res"]))
        self.lib.mpcqp_engine_num_obs_per_instance(self.h, None)
        self.lib.mpcqp_engine_limits_per_instance(self.h, None)
        self._check(rc)
        return out

    # ---- batched entry point, raw pointers (device-resident or pinned host) ----------------------
    def solve_mpc_batch_ptr(self, params, settings: Settings, B: int, R: int, ptrs: dict, obs_dyn: np.ndarray,
                            device: bool, nobs: np.ndarray | None = None, limits: np.ndarray | None = None):
        """ptrs: name -> integer address for x0,xref,obs_c,obs_semi,obs_yaw,lin_pt,warm_x,x,y,status,iter,
        rho_updates,obj,pri_res,dua_res (0 / missing = NULL).  device=True: device pointers, asynchronous on the
        engine stream (call sync()); device=False: host pointers, synchronous."""
        p = params_to_c(params)
        od = np.ascontiguousarray(obs_dyn, dtype=np.int32)
        self._check(self.lib.mpcqp_engine_obs_dyn_per_instance(self.h, C.c_int(1 if od.ndim == 3 else 0)))
        nobs = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)       # host arrays [B] / [B][2], see the header
        limits = None if limits is None else np.ascontiguousarray(limits, dtype=np.float64)
        self._check(self.lib.mpcqp_engine_num_obs_per_instance(self.h, _ip(nobs)))
        self._check(self.lib.mpcqp_engine_limits_per_instance(self.h, _dp(limits)))
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        fn = self.lib.mpcqp_solve_mpc_batch_device if device else self.lib.mpcqp_solve_mpc_batch_host
        rc = fn(self.h, C.byref(p), C.byref(settings), C.c_int32(B), C.c_int32(R), g("x0"), g("xref"), g("obs_c"),
                g("obs_semi"), g("obs_yaw"), _ip(od), g("lin_pt"), g("warm_x"), g("x"), g("y"), g("status"), g("iter"),
                g("rho_updates"), g("obj"), g("pri_res"), g("dua_res"))
        self.lib.mpcqp_engine_num_obs_per_instance(self.h, None)
        self.lib.mpcqp_engine_limits_per_instance(self.h, None)
        self._check(rc)

    # ---- candidate scoring / selection on the device (getTrajectoryScore, evaluateTraj) --------------------------
    def score_candidates_ptr(self, params, B: int, R: int, n_dynamic: int, ptrs: dict):
        """ptrs: device addresses for x, prev_plan (0 = first control step), xref, obs_c, obs_semi, obs_c_last, obs_semi_last
        (stage N, [B][R][3]), score.  Asynchronous."""
        p = params_to_c(params)
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        self._check(self.lib.mpcqp_score_candidates_device(self.h, C.byref(p), C.c_int32(B), C.c_int32(R), C.c_int32(n_dynamic),
                                                           g("x"), g("prev_plan"), g("xref"), g("obs_c"), g("obs_semi"), g("obs_c_last"),
                                                           g("obs_semi_last"), g("score")))

    def select_candidates_ptr(self, S: int, Cn: int, n: int, ptrs: dict):
        """ptrs: device addresses for cand [S][C] int32, weight [S][C], score, x_all, best [S] int32, weighted (optional),
        plan (optional).  Asynchronous on the engine stream."""
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        self._check(self.lib.mpcqp_select_candidates_device(self.h, C.c_int32(S), C.c_int32(Cn), C.c_int32(n), g("cand"), g("weight"),
                                                            g("score"), g("x_all"), g("best"), g("weighted"), g("plan")))

    def intent_candidates_ptr(self, params, S: int, D: int, NP: int, ptrs: dict):
        """ptrs: device addresses for pred_pos, pred_size, prob, prev_plan (0 on the first step), pos, scen_a, scen_b, obs_c_a,
        obs_semi_a, obs_c_b, obs_semi_b, obs_c_last_a, obs_semi_last_a, obs_c_last_b, obs_semi_last_b (stage N for the scoring;
        all four or none), weight, cand.  Asynchronous on the engine stream."""
        p = params_to_c(params)
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        self._check(self.lib.mpcqp_intent_candidates_device(self.h, C.byref(p), C.c_int32(S), C.c_int32(D), C.c_int32(NP), g("pred_pos"),
                                                            g("pred_size"), g("prob"), g("prev_plan"), g("pos"), g("scen_a"), g("scen_b"),
                                                            g("obs_c_a"), g("obs_semi_a"), g("obs_c_b"), g("obs_semi_b"), g("obs_c_last_a"),
                                                            g("obs_semi_last_a"), g("obs_c_last_b"), g("obs_semi_last_b"), g("weight"), g("cand")))

    def predict_ptr(self, pparams: PredictorParamsC, num_obstacles: int, num_hist: int, ptrs: dict):
        """mpcqp_predict_device: ptrs = device addresses of pos_hist, vel_hist [NOB][H][3] (newest first), size [NOB][3], pred_pos,
        pred_size [NOB][4][prediction_size + 1][3], intent_prob [NOB][4].  Asynchronous on the engine stream."""
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        self._check(self.lib.mpcqp_predict_device(self.h, C.byref(pparams), C.c_int32(num_obstacles), C.c_int32(num_hist), g("pos_hist"), g("vel_hist"),
                                                  g("size"), g("pred_pos"), g("pred_size"), g("intent_prob")))

    def gather_rows_ptr(self, B: int, width: int, idx_ptr: int, src_ptr: int, dst_ptr: int):
        self._check(self.lib.mpcqp_gather_rows_device(self.h, C.c_int64(B), C.c_int32(width), C.c_void_p(idx_ptr), C.c_void_p(src_ptr),
                                                      C.c_void_p(dst_ptr)))



def solve_qp_batch(eng: Engine, qb, warm_y=None, settings: Settings | None = None, want_y: bool = True) -> dict:
    """mpcqp_solve_qp_batch_host: B unstructured QPs sharing one CSC pattern (attributes of `qb`: n, m, P_colptr, P_rowidx,
    P_val [B, nnzP], q [B, n], A_colptr, A_rowidx, A_val [B, nnzA], l, u [B, m], warm_x [B, n] or None) in one launch of the
    dense generic kernel — polyTrajSolver's x / y / z problems, or a set of candidate paths."""
    s = settings if settings is not None else default_settings()
    B, n, m = int(qb.q.shape[0]), int(qb.n), int(qb.m)
    I = C.POINTER(C.c_int64)
    pat = [np.ascontiguousarray(v, dtype=np.int64) for v in (qb.P_colptr, qb.P_rowidx, qb.A_colptr, qb.A_rowidx)]
    d = [np.ascontiguousarray(v, dtype=np.float64) for v in (qb.P_val, qb.q, qb.A_val, qb.l, qb.u)]
    wx = None if getattr(qb, "warm_x", None) is None else np.ascontiguousarray(qb.warm_x, dtype=np.float64)
    wy = None if warm_y is None else np.ascontiguousarray(warm_y, dtype=np.float64)
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, m)) if want_y else None, status=np.zeros(B, np.int32),
               iter=np.zeros(B, np.int32), rho_updates=np.zeros(B, np.int32), obj=np.zeros(B), pri_res=np.zeros(B),
               dua_res=np.zeros(B))
    eng._check(eng.lib.mpcqp_solve_qp_batch_host(
        eng.h, C.byref(s), C.c_int32(B), C.c_int64(n), C.c_int64(m), pat[0].ctypes.data_as(I), pat[1].ctypes.data_as(I),
        _dp(d[0]), _dp(d[1]), pat[2].ctypes.data_as(I), pat[3].ctypes.data_as(I), _dp(d[2]), _dp(d[3]), _dp(d[4]),
        _dp(wx), _dp(wy), _dp(out["x"]), _dp(out["y"]), _ip(out["status"]), _ip(out["iter"]), _ip(out["rho_updates"]),
        _dp(out["obj"]), _dp(out["pri_res"]), _dp(out["dua_res"])))
    return out


class Problem:
    """OSQP-shaped single problem (mpcqp_setup ... mpcqp_cleanup, include/mpcqp_b200.h section 3): what
    `OsqpEigen::Solver` drives for one QP.  mpcPlanner-structured problems run on the stage kernels, anything else
    (polyTrajSolver.cpp:162-239) on the dense generic kernel; `engine.last_path` tells which."""

    def __init__(self, eng: Engine, n, m, P_colptr, P_rowidx, P_val, q, A_colptr, A_rowidx, A_val, l, u,
                 settings: Settings | None = None):
        self.eng, self.lib, self.n, self.m = eng, eng.lib, int(n), int(m)
        self.h = C.c_void_p()
        s = settings if settings is not None else default_settings()
        I = C.POINTER(C.c_int64)
        a = [np.ascontiguousarray(v, dtype=np.int64) for v in (P_colptr, P_rowidx, A_colptr, A_rowidx)]
        d = [np.ascontiguousarray(v, dtype=np.float64) for v in (P_val, q, A_val, l, u)]
        rc = self.lib.mpcqp_setup(eng.h, C.byref(self.h), C.c_int64(self.n), C.c_int64(self.m), a[0].ctypes.data_as(I),
                                  a[1].ctypes.data_as(I), _dp(d[0]), _dp(d[1]), a[2].ctypes.data_as(I),
                                  a[3].ctypes.data_as(I), _dp(d[2]), _dp(d[3]), _dp(d[4]), C.byref(s))
        if rc != 0:
            self.h = None
        eng._check(rc)

    def warm_start(self, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            self.eng._check(self.lib.mpcqp_warm_start_x(self.h, _dp(x)))
        else:
            self.eng._check(self.lib.mpcqp_warm_start(self.h, _dp(x), _dp(np.ascontiguousarray(y, dtype=np.float64))))

    def update_bounds(self, l, u):
        self.eng._check(self.lib.mpcqp_update_bounds(self.h, _dp(np.ascontiguousarray(l, dtype=np.float64)),
                                                     _dp(np.ascontiguousarray(u, dtype=np.float64))))

    def update_lin_cost(self, q):
        self.eng._check(self.lib.mpcqp_update_lin_cost(self.h, _dp(np.ascontiguousarray(q, dtype=np.float64))))

    def solve(self) -> dict:
        self.eng._check(self.lib.mpcqp_solve(self.h))
        info = Info()
        self.eng._check(self.lib.mpcqp_get_info(self.h, C.byref(info)))
        x = np.zeros(self.n); y = np.zeros(max(self.m, 1))
        self.eng._check(self.lib.mpcqp_get_solution(self.h, _dp(x), _dp(y)))
        return dict(x=x, y=y[:self.m], status=int(info.status_val), iter=int(info.iter), rho_updates=int(info.rho_updates),
                    obj=info.obj_val, pri_res=info.pri_res, dua_res=info.dua_res, solve_time=info.solve_time)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcqp_cleanup(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
