"""ORACLE (test infrastructure): ctypes bindings for
  * oracle/_ref/libref_driver.so + oracle/_ref/libosqp.so  — the reference's own OSQP 0.6.2 binary
    (kind "reference"), and
  * oracle/liboracle.so — the C restatement of OSQP's algorithm (kind "port").
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# Determinism pins: the reference's default adaptive_rho_interval=0 derives the interval
# from wall-clock time and its 0.05 s time_limit makes status -6 timing dependent; both are pinned.
PINS = dict(adaptive_rho_interval=25, time_limit=0.0)

STATUS_NAMES = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate",
                4: "dual infeasible inaccurate", -2: "maximum iterations reached", -3: "primal infeasible",
                -4: "dual infeasible", -5: "interrupted", -6: "time limit reached", -7: "non convex",
                -10: "unsolved"}


class Overrides(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf",
                                          "eps_dual_inf", "time_limit", "adaptive_rho_tolerance")] + \
               [(k, C.c_longlong) for k in ("max_iter", "adaptive_rho", "adaptive_rho_interval",
                                            "check_termination", "scaling", "warm_start", "scaled_termination")]

    @classmethod
    def make(cls, **kw):
        o = cls()
        for k, t in cls._fields_:
            setattr(o, k, float("nan") if t is C.c_double else -1)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


def build(force: bool = False) -> None:
    """Compile the oracle's C pieces (and refresh oracle/_ref when /root/reference is present)."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so")) \
        or not os.path.exists(os.path.join(REF_DIR, "libref_driver.so"))
    if need or os.path.exists("/root/reference"):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _c(a, dt=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class _Solver:
    """Common batched-call wrapper; subclasses pick the shared object + entry point."""
    kind = "?"

    def _call(self, fn, qb, warm_y, ov, nthreads, want_y, dump_idx):
        B = qb.q.shape[0]
        n, m = qb.n, qb.m
        Pc, Pr = _c(qb.P_colptr, np.int64), _c(qb.P_rowidx, np.int64)
        Ac, Ar = _c(qb.A_colptr, np.int64), _c(qb.A_rowidx, np.int64)
        Pv, Av, q, l, u = _c(qb.P_val), _c(qb.A_val), _c(qb.q), _c(qb.l), _c(qb.u)
        wx = _c(qb.warm_x); wy = _c(warm_y)
        out = dict(x=np.zeros((B, n)), y=np.zeros((B, m)) if want_y else None,
                   status=np.zeros(B, np.int64), iter=np.zeros(B, np.int64), rho_updates=np.zeros(B, np.int64),
                   exitflag=np.zeros(B, np.int64), obj=np.zeros(B), pri_res=np.zeros(B), dua_res=np.zeros(B),
                   setup_time=np.zeros(B), solve_time=np.zeros(B), wall_time=np.zeros(B))
        dump = np.zeros(2 + 2 * n + 4 * m) if dump_idx is not None else None
        LL, D = C.c_longlong, C.c_double
        fn.restype = C.c_double
        wall = fn(LL(n), LL(m), LL(Pv.shape[1]), LL(Av.shape[1]), LL(B),
                  _p(Pc, LL), _p(Pr, LL), _p(Pv, D), _p(q, D), _p(Ac, LL), _p(Ar, LL), _p(Av, D),
                  _p(l, D), _p(u, D), _p(wx, D), _p(wy, D), C.byref(ov), C.c_int(nthreads),
                  _p(out["x"], D), _p(out["y"], D), _p(out["status"], LL), _p(out["iter"], LL),
                  _p(out["rho_updates"], LL), _p(out["exitflag"], LL), _p(out["obj"], D), _p(out["pri_res"], D),
                  _p(out["dua_res"], D), _p(out["setup_time"], D), _p(out["solve_time"], D),
                  _p(out["wall_time"], D), _p(dump, D), LL(-1 if dump_idx is None else dump_idx))
        out["wall"] = wall
        if dump is not None:
            o = 2
            d = dict(c=dump[0], rho=dump[1])
            for k, sz in (("D", n), ("E", m), ("rho_vec", m), ("xs", n), ("zs", m), ("ys", m)):
                d[k] = dump[o:o + sz].copy(); o += sz
            out["dump"] = d
        return out

    def solve_batch(self, qb, warm_y=None, nthreads=1, want_y=True, dump_idx=None, **settings):
        kw = dict(PINS); kw.update(settings)
        ov = Overrides.make(**kw)
        return self._call(self._fn, qb, warm_y, ov, nthreads, want_y, dump_idx)


class RefOsqp(_Solver):
    """The reference's own libosqp.so (OSQP 0.6.2), one problem per thread."""
    kind = "reference"

    def __init__(self):
        drv = os.path.join(REF_DIR, "libref_driver.so")
        lib = os.path.join(REF_DIR, "libosqp.so")
        if not (os.path.exists(drv) and os.path.exists(lib)):
            raise FileNotFoundError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(drv)
        rc = self.lib.ref_open(lib.encode())
        if rc:
            raise OSError(f"ref_open failed: {rc}")
        self._fn = self.lib.ref_solve_batch

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(REF_DIR, "libref_driver.so")) and \
            os.path.exists(os.path.join(REF_DIR, "libosqp.so"))


class PortOsqp(_Solver):
    """C restatement of OSQP 0.6.2's algorithm (oracle/osqp_restated.c)."""
    kind = "port"

    def __init__(self):
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build(force=True)
        self.lib = C.CDLL(so)
        self._fn = self.lib.port_solve_batch
