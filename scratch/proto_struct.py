"""Prototype: structured (leaf-eliminated 6x6 block tridiagonal) solve of the reduced KKT vs dense."""
import sys, dataclasses; sys.path.insert(0, "/root/repo")
import numpy as np
from intent_mpc_b200 import workloads as W
from oracle import mpc_assembly as MA, bindings as OB
np.set_printoptions(linewidth=200, precision=4)
def to_qb(mb):
    p = MA.MpcParams(**dataclasses.asdict(mb.params))
    return MA.assemble_batch(p, mb.x0, mb.xref, mb.obs_c, mb.obs_semi, mb.obs_yaw, mb.obs_dyn, mb.lin_pt, mb.warm_x)
mb = W.snapshot(); qb = to_qb(mb)
ref = OB.RefOsqp()
r = ref.solve_batch(qb, dump_idx=0)
d = r["dump"]; n, m = qb.n, qb.m
N = mb.params.N; R = mb.num_obs
# dense scaled matrices
A = np.zeros((m, n)); 
for j in range(n):
    for k in range(qb.A_colptr[j], qb.A_colptr[j+1]): A[qb.A_rowidx[k], j] = qb.A_val[0, k]
Pd = np.zeros(n); Pd[qb.P_rowidx] = qb.P_val[0]
D, E, c, rho = d["D"], d["E"], d["c"], d["rho_vec"]
Ab = E[:, None] * A * D[None, :]; Pb = c * D * Pd * D
sigma = 1e-6
H = np.diag(Pb + sigma) + Ab.T @ (rho[:, None] * Ab)
rng = np.random.default_rng(0); rhs = rng.standard_normal(n)
xd = np.linalg.solve(H, rhs)

# ---- structured representation
def gidx(k, j):  # local var j of stage k -> global index
    return 8*k + j if j < 8 else 8*(N+1) + 5*k + (j-8)
def row_dy(k, r): return 8*k + r
def row_bx(k, j): return 8*(N+1) + 8*k + j if j < 8 else 16*(N+1) + 5*k + (j-8)
base = 16*(N+1) + 5*N
def row_ex(k, o): return base + k*R + o
ts_f = float(np.float32(0.1)); h_f = float(np.float32(0.5*0.1*0.1))
a_pp, a_pv, b_pa, a_vv, b_va, b_ss = 1.0, ts_f, h_f, 1.0, ts_f, 1.0
nv = lambda k: 13 if k < N else 8
Dk = np.zeros((N+1, 13)); Pk = np.zeros((N+1, 13)); Edy = np.zeros((N+1, 8)); Ebx = np.zeros((N+1, 13)); Eex = np.zeros((N, R))
rdy = np.zeros((N+1, 8)); rbx = np.zeros((N+1, 13)); rex = np.zeros((N, R)); g = np.zeros((N, R, 3)); tsl = np.zeros((N, R), int)
for k in range(N+1):
    for j in range(nv(k)):
        Dk[k, j] = D[gidx(k, j)]; Pk[k, j] = Pb[gidx(k, j)]; Ebx[k, j] = E[row_bx(k, j)]; rbx[k, j] = rho[row_bx(k, j)]
    for r_ in range(8): Edy[k, r_] = E[row_dy(k, r_)]; rdy[k, r_] = rho[row_dy(k, r_)]
for k in range(N):
    for o in range(R):
        Eex[k, o] = E[row_ex(k, o)]; rex[k, o] = rho[row_ex(k, o)]
        g[k, o] = A[row_ex(k, o), 8*k:8*k+3]; tsl[k, o] = 0 if mb.obs_dyn[k, o] else 1

def factor():
    F = {}
    hd = Pk + sigma + rbx * (Ebx * Dk) ** 2
    hd[:, :8] += rdy * (Edy * Dk[:, :8]) ** 2
    Tkk = np.zeros((N+1, 6, 6)); Tnk = np.zeros((N, 6, 6))   # Tnk[k] = T[k+1,k]
    ds = hd[:, 6:8].copy()                      # s pivots
    es = np.zeros((N+1, 2)); dsg = np.zeros((N, 2)); fsg = np.zeros((N, 2, 3)); da = np.zeros((N, 3)); ca = np.zeros((N, 3, 4))
    for k in range(N+1):
        for i in range(6): Tkk[k, i, i] += hd[k, i]
    for k in range(N):
        Ep = Edy[k+1]; rp = rdy[k+1]
        for t in range(2):
            gam = Ep[6+t] * b_ss * Dk[k, 11+t]; om = Ep[6+t] * Dk[k+1, 6+t]
            es[k+1, t] = -rp[6+t] * om * gam
            dsg[k, t] = hd[k, 11+t] + rp[6+t] * gam * gam - es[k+1, t] ** 2 / ds[k+1, t]
        for o in range(R):
            t = tsl[k, o]; eo = Eex[k, o]; cs = -eo * Dk[k, 11+t]; cp = eo * g[k, o] * Dk[k, 0:3]
            dsg[k, t] += rex[k, o] * cs * cs
            fsg[k, t] += rex[k, o] * cs * cp
            Tkk[k, 0:3, 0:3] += rex[k, o] * np.outer(cp, cp)
        for t in range(2):
            Tkk[k, 0:3, 0:3] -= np.outer(fsg[k, t], fsg[k, t]) / dsg[k, t]
        for c_ in range(3):
            al_p = Ep[c_] * a_pp * Dk[k, c_]; al_v = Ep[c_] * a_pv * Dk[k, 3+c_]; al_a = Ep[c_] * b_pa * Dk[k, 8+c_]; om_p = Ep[c_] * Dk[k+1, c_]
            be_v = Ep[3+c_] * a_vv * Dk[k, 3+c_]; be_a = Ep[3+c_] * b_va * Dk[k, 8+c_]; om_v = Ep[3+c_] * Dk[k+1, 3+c_]
            r1, r2 = rp[c_], rp[3+c_]
            Tkk[k, c_, c_] += r1 * al_p ** 2; Tkk[k, c_, 3+c_] += r1 * al_p * al_v; Tkk[k, 3+c_, c_] += r1 * al_p * al_v
            Tkk[k, 3+c_, 3+c_] += r1 * al_v ** 2 + r2 * be_v ** 2
            Tnk[k, c_, c_] += -r1 * om_p * al_p; Tnk[k, c_, 3+c_] += -r1 * om_p * al_v; Tnk[k, 3+c_, 3+c_] += -r2 * om_v * be_v
            da[k, c_] = hd[k, 8+c_] + r1 * al_a ** 2 + r2 * be_a ** 2
            cv = np.array([r1 * al_a * al_p, r1 * al_a * al_v + r2 * be_a * be_v, -r1 * al_a * om_p, -r2 * be_a * om_v])  # p_c, v_c, p'_c, v'_c
            ca[k, c_] = cv
            idx_k = [c_, 3+c_]
            for a_i in range(2):
                for b_i in range(2):
                    Tkk[k, idx_k[a_i], idx_k[b_i]] -= cv[a_i] * cv[b_i] / da[k, c_]
                    Tkk[k+1, idx_k[a_i], idx_k[b_i]] -= cv[2+a_i] * cv[2+b_i] / da[k, c_]
                    Tnk[k, idx_k[a_i], idx_k[b_i]] -= cv[2+a_i] * cv[b_i] / da[k, c_]
    # block LDL: S_0 = T_00 ; G_k = Tnk[k] S_k^-1 ; S_{k+1} = T_{k+1,k+1} - G_k Tnk[k]^T
    Sinv = np.zeros((N+1, 6, 6)); G = np.zeros((N, 6, 6)); S = Tkk[0].copy()
    for k in range(N+1):
        Sinv[k] = np.linalg.inv(S)
        if k < N:
            G[k] = Tnk[k] @ Sinv[k]
            S = Tkk[k+1] - G[k] @ Tnk[k].T
    F.update(ds=ds, es=es, dsg=dsg, fsg=fsg, da=da, ca=ca, Sinv=Sinv, G=G)
    return F

def solve(F, rhs):
    r = np.zeros((N+1, 13))
    for k in range(N+1):
        for j in range(nv(k)): r[k, j] = rhs[gidx(k, j)]
    # forward leaves
    for k in range(N):
        for t in range(2): r[k, 11+t] -= F["es"][k+1, t] / F["ds"][k+1, t] * r[k+1, 6+t]
    for k in range(N):
        for t in range(2): r[k, 0:3] -= F["fsg"][k, t] / F["dsg"][k, t] * r[k, 11+t]
    for k in range(N):
        for c_ in range(3):
            cv = F["ca"][k, c_] / F["da"][k, c_] * r[k, 8+c_]
            r[k, c_] -= cv[0]; r[k, 3+c_] -= cv[1]; r[k+1, c_] -= cv[2]; r[k+1, 3+c_] -= cv[3]
    w = np.zeros((N+1, 6)); w[0] = r[0, :6]
    for k in range(1, N+1): w[k] = r[k, :6] - F["G"][k-1] @ w[k-1]
    y = np.zeros((N+1, 6))
    for k in range(N+1): y[k] = F["Sinv"][k] @ w[k]
    for k in range(N-1, -1, -1): y[k] -= F["G"][k].T @ y[k+1]
    x = np.zeros((N+1, 13)); x[:, :6] = y
    for k in range(N):
        for c_ in range(3):
            cv = F["ca"][k, c_]
            x[k, 8+c_] = (r[k, 8+c_] - cv[0]*y[k, c_] - cv[1]*y[k, 3+c_] - cv[2]*y[k+1, c_] - cv[3]*y[k+1, 3+c_]) / F["da"][k, c_]
        for t in range(2):
            x[k, 11+t] = (r[k, 11+t] - F["fsg"][k, t] @ y[k, 0:3]) / F["dsg"][k, t]
    for k in range(N+1):
        for t in range(2):
            x[k, 6+t] = (r[k, 6+t] - (F["es"][k, t] * x[k-1, 11+t] if k >= 1 else 0.0)) / F["ds"][k, t]
    out = np.zeros(n)
    for k in range(N+1):
        for j in range(nv(k)): out[gidx(k, j)] = x[k, j]
    return out
F = factor(); xs = solve(F, rhs)
print("rel err struct vs dense:", np.abs(xs - xd).max() / np.abs(xd).max(), " cond(H)=%.2e" % np.linalg.cond(H))
print("residual dense", np.abs(H @ xd - rhs).max(), "struct", np.abs(H @ xs - rhs).max())
