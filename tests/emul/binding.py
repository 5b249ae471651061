"""TEST INFRASTRUCTURE: ctypes driver for the host emulation of the kernel source (tests/emul/emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import mpc_assembly as MA

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libemul.so")
SRC = [os.path.join(HERE, "emul.cpp"), os.path.join(ROOT, "intent-mpc_b200", "csrc", "mpcqp_core.cuh"),
       os.path.join(ROOT, "intent-mpc_b200", "csrc", "mpcqp_dense.cuh"), os.path.join(ROOT, "intent-mpc_b200", "csrc", "mpcqp_band.cuh"),
       os.path.join(ROOT, "intent-mpc_b200", "csrc", "mpcqp_band_host.hpp")]

DEFAULT_D = dict(rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4, eps_dual_inf=1e-4,
                 adaptive_rho_tolerance=5.0)
DEFAULT_I = dict(max_iter=4000, scaling=10, adaptive_rho=1, adaptive_rho_interval=25, check_termination=25,
                 warm_start=1)


def build():
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRC):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, SRC[0]],
                       check=True)


def structured_inputs(mb):
    """Numpy computation of the kernel's structured inputs (q, x0, g, low, pd, slack, bounds) from an
    MpcBatch using the ORACLE's assembly pieces — independent of the product's device builder."""
    p = MA.MpcParams(**{k: getattr(mb.params, k) for k in MA.MpcParams.__dataclass_fields__})
    B, N, NS, R = mb.B, p.N, p.N + 1, mb.num_obs
    Q, _ = MA.weight_diagonals(p)
    n = p.n
    q = np.zeros((B, n)); xr = np.zeros((B, NS, 8)); xr[:, :, 0:3] = mb.xref
    q[:, :8 * NS] = (-(xr) * Q[None, None, :]).reshape(B, -1)
    x0 = np.zeros((B, 8)); x0[:, :6] = mb.x0
    if R:
        fxx, fyy, fzz, low = MA.ellipsoid_linearisation(mb.lin_pt[:, :, None, :], mb.obs_c, mb.obs_semi, mb.obs_yaw)
        g = np.stack([fxx, fyy, fzz], axis=-1)
    else:
        g = np.zeros((B, N, 0, 3)); low = np.zeros((B, N, 0))
    pdg = MA.hessian_diagonal(p)
    pd = np.zeros((NS, 13)); pd[:, :8] = pdg[:8 * NS].reshape(NS, 8); pd[:N, 8:] = pdg[8 * NS:].reshape(N, 5)
    x_min, x_max, u_min, u_max = MA.box_bounds(p)
    blo = np.concatenate([x_min, u_min]); bhi = np.concatenate([x_max, u_max])
    slack = np.where(mb.obs_dyn != 0, 0, 1).astype(np.uint8)
    f32 = lambda v: float(np.float32(v))
    return dict(NS=NS, R=R, B=B, n=n, m=p.m(R), a_pv=f32(p.ts), b_pa=f32(0.5 * p.ts ** 2), b_va=f32(p.ts),
                blo=blo, bhi=bhi, pd=pd, slack=slack, q=q, x0=x0, g=np.ascontiguousarray(g),
                low=np.ascontiguousarray(low), warm_x=np.ascontiguousarray(mb.warm_x))


def solve(mb, want_y=True, linsys=0, finite_inf=None, **settings):
    build()
    lib = C.CDLL(SO)
    s = structured_inputs(mb)
    obs_hi = float("inf")
    if finite_inf is not None:                         # the caller writes its infinite bounds as a finite number (OSQP_INFTY)
        s["blo"] = np.where(np.isinf(s["blo"]), -finite_inf, s["blo"]); s["bhi"] = np.where(np.isinf(s["bhi"]), finite_inf, s["bhi"]); obs_hi = float(finite_inf)
    sd = dict(DEFAULT_D); si = dict(DEFAULT_I)
    for k, v in settings.items():
        (sd if k in sd else si)[k] = v
    sdv = np.array([sd[k] for k in DEFAULT_D], dtype=np.float64)
    siv = np.array([si[k] for k in DEFAULT_I], dtype=np.int32)
    B, n, m = s["B"], s["n"], s["m"]
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, m)) if want_y else None, status=np.zeros(B, np.int32),
               iter=np.zeros(B, np.int32), rho_updates=np.zeros(B, np.int32), obj=np.zeros(B),
               pri_res=np.zeros(B), dua_res=np.zeros(B))
    D = C.POINTER(C.c_double); I = C.POINTER(C.c_int)
    def dp(a): return a.ctypes.data_as(D) if a is not None else None
    rc = lib.emul_solve_batch(C.c_int(linsys), C.c_int(s["NS"]), C.c_int(s["R"]), C.c_int(B), C.c_double(s["a_pv"]), C.c_double(s["b_pa"]),
                         C.c_double(s["b_va"]), dp(s["blo"]), dp(s["bhi"]), C.c_double(obs_hi), dp(sdv), siv.ctypes.data_as(I),
                         dp(np.ascontiguousarray(s["pd"])), s["slack"].ctypes.data_as(C.POINTER(C.c_ubyte)),
                         dp(s["q"]), dp(s["x0"]), dp(s["g"]), dp(s["low"]), dp(s["warm_x"]),
                         dp(out["x"]), dp(out["y"]), out["status"].ctypes.data_as(I), out["iter"].ctypes.data_as(I),
                         out["rho_updates"].ctypes.data_as(I), dp(out["obj"]), dp(out["pri_res"]), dp(out["dua_res"]))
    if rc != 0:
        raise RuntimeError(f"emul_solve_batch(linsys={linsys}) -> {rc}")
    return out


def solve_band(qb, warm_y=None, **settings):
    """The sparse generic-path kernel source (mpcqp_band.cuh) on the host; out["bandwidth"] = the RCM half-bandwidth."""
    return solve_dense(qb, warm_y=warm_y, _entry="emul_band_solve", **settings)


def solve_dense(qb, warm_y=None, _entry="emul_dense_solve", **settings):
    """The generic-path kernel source (mpcqp_dense.cuh) on the host, one QpBatch problem at a time."""
    build()
    lib = C.CDLL(SO)
    sd = dict(DEFAULT_D); si = dict(DEFAULT_I)
    for k, v in settings.items():
        (sd if k in sd else si)[k] = v
    sdv = np.array([sd[k] for k in DEFAULT_D], dtype=np.float64)
    siv = np.array([si[k] for k in DEFAULT_I], dtype=np.int32)
    B, n, m = qb.q.shape[0], qb.n, qb.m
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, m)), status=np.zeros(B, np.int32), iter=np.zeros(B, np.int32),
               rho_updates=np.zeros(B, np.int32), obj=np.zeros(B), pri_res=np.zeros(B), dua_res=np.zeros(B))
    D = C.POINTER(C.c_double); I = C.POINTER(C.c_int); LL = C.POINTER(C.c_longlong)
    def dp(a): return a.ctypes.data_as(D) if a is not None else None
    pat = [np.ascontiguousarray(v, dtype=np.int64) for v in (qb.P_colptr, qb.P_rowidx, qb.A_colptr, qb.A_rowidx)]
    for b in range(B):
        ii = np.zeros(3, np.int32); dd = np.zeros(3)
        wx = np.ascontiguousarray(qb.warm_x[b]) if qb.warm_x is not None else None
        wy = np.ascontiguousarray(warm_y[b]) if warm_y is not None else None
        rc = getattr(lib, _entry)(C.c_int(n), C.c_int(m), pat[0].ctypes.data_as(LL), pat[1].ctypes.data_as(LL),
                             dp(np.ascontiguousarray(qb.P_val[b])), pat[2].ctypes.data_as(LL), pat[3].ctypes.data_as(LL),
                             dp(np.ascontiguousarray(qb.A_val[b])), dp(np.ascontiguousarray(qb.q[b])),
                             dp(np.ascontiguousarray(qb.l[b])), dp(np.ascontiguousarray(qb.u[b])), dp(wx), dp(wy), dp(sdv),
                             siv.ctypes.data_as(I), dp(out["x"][b]), dp(out["y"][b]), ii.ctypes.data_as(I), dp(dd))
        if _entry == "emul_band_solve":
            if rc < 0:
                raise RuntimeError(f"pattern not eligible for the band path (half-bandwidth {-rc})")
            out["bandwidth"] = rc
        out["status"][b], out["iter"][b], out["rho_updates"][b] = ii
        out["obj"][b], out["pri_res"][b], out["dua_res"][b] = dd
    return out
