"""The C++ host layer above the C ABI: the OsqpEigen::Solver-shaped facade (intent-mpc_b200/host/OsqpEigenB200.hpp),
the mpcPlanner mirror (MpcPlannerB200.hpp) and the OSQP-shaped single-problem entry points of include/mpcqp_b200.h.
CPU tier: everything compiles and links against the library and fails loudly without a GPU.  GPU tier: one QP driven
exactly like mpcPlanner::solveTraj drives OsqpEigen (mpcPlanner.cpp:436-527) matches the oracle; a receding-horizon
loop like mpc_node.cpp:209-236 runs through makePlanWithPred."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from intent_mpc_b200 import engine, workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "host_test")


def build_host_test():
    import __graft_entry__ as G
    G.build()
    src = os.path.join(CPP, "host_test.cpp")
    deps = [src, os.path.join(ROOT, "intent-mpc_b200", "host", "OsqpEigenB200.hpp"),
            os.path.join(ROOT, "intent-mpc_b200", "host", "MpcPlannerB200.hpp"), os.path.join(ROOT, "include", "mpcqp_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, src, "-L" + os.path.join(ROOT, "intent-mpc_b200"),
                        "-lmpcqp_b200", "-Wl,-rpath," + os.path.join(ROOT, "intent-mpc_b200")], check=True)
    return EXE


def _oracle():
    return OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()


def _csc_to_triplets(colptr, rowidx, val):
    cols = np.repeat(np.arange(len(colptr) - 1), np.diff(colptr))
    return rowidx.astype(np.int64), cols.astype(np.int64), val.astype(np.float64)


def test_host_layer_compiles_and_refuses_without_gpu():
    exe = build_host_test()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tier")
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_facade_solves_like_osqpeigen(tmp_path):
    exe = build_host_test()
    qb = to_qp_batch(W.static_batch(4, num_obs=4, seed0=40))
    b = 2
    pr, pc, pv = _csc_to_triplets(qb.P_colptr, qb.P_rowidx, qb.P_val[b])
    ar, ac, av = _csc_to_triplets(qb.A_colptr, qb.A_rowidx, qb.A_val[b])
    prob = tmp_path / "problem.bin"; out = tmp_path / "out.bin"
    with open(prob, "wb") as f:
        f.write(np.array([qb.n, qb.m, len(pv), len(av)], dtype=np.int64).tobytes())
        for a in (pr, pc, pv, ar, ac, av, qb.q[b], qb.l[b], qb.u[b], qb.warm_x[b]):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([exe, "facade", str(prob), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    raw = np.fromfile(out, dtype=np.float64)
    head, x, y, x2 = raw[:6], raw[6:6 + qb.n], raw[6 + qb.n:6 + qb.n + qb.m], raw[6 + qb.n + qb.m:]
    import dataclasses
    one = dataclasses.replace(qb, P_val=qb.P_val[b:b + 1], q=qb.q[b:b + 1], A_val=qb.A_val[b:b + 1], l=qb.l[b:b + 1], u=qb.u[b:b + 1], warm_x=qb.warm_x[b:b + 1])
    ref = _oracle().solve_batch(one)
    assert int(head[0]) == ref["status"][0] and int(head[1]) == ref["iter"][0] and int(head[2]) == ref["rho_updates"][0]
    assert rel_inf(x[None], ref["x"]).max() < 1e-5
    assert abs((head[3] - ref["obj"][0]) / ref["obj"][0]) < 1e-5
    assert rel_inf(y[None], ref["y"]).max() < 1e-4
    # second solve after updateGradient (0.5 q) + updateBounds + a fresh warm start
    two = dataclasses.replace(one, q=0.5 * one.q)
    ref2 = _oracle().solve_batch(two, want_y=False)
    assert int(head[4]) == ref2["status"][0] and int(head[5]) == ref2["iter"][0]
    assert rel_inf(x2[None], ref2["x"]).max() < 1e-5


@pytest.mark.gpu
def test_setup_rejects_what_it_cannot_solve():
    """mpcqp_setup: OSQP's data validation (l > u); unstructured problems go to the dense generic kernel (refused only
    beyond n + m = 4096, tests/test_dense_generic.py); never a silent CPU solve."""
    lib = engine.load_library()
    eng = engine.Engine(0)
    qb = to_qp_batch(W.static_batch(1, num_obs=4, seed0=7))
    I = C.POINTER(C.c_int64); D = C.POINTER(C.c_double)
    def call(A_val=None, l=None):
        pc = np.ascontiguousarray(qb.P_colptr); pi = np.ascontiguousarray(qb.P_rowidx); pv = np.ascontiguousarray(qb.P_val[0])
        ac = np.ascontiguousarray(qb.A_colptr); ai = np.ascontiguousarray(qb.A_rowidx)
        av = np.ascontiguousarray(qb.A_val[0] if A_val is None else A_val)
        q = np.ascontiguousarray(qb.q[0]); ll = np.ascontiguousarray(qb.l[0] if l is None else l); uu = np.ascontiguousarray(qb.u[0])
        h = C.c_void_p(); s = engine.default_settings()
        rc = lib.mpcqp_setup(eng.h, C.byref(h), C.c_int64(qb.n), C.c_int64(qb.m), pc.ctypes.data_as(I), pi.ctypes.data_as(I), pv.ctypes.data_as(D),
                             q.ctypes.data_as(D), ac.ctypes.data_as(I), ai.ctypes.data_as(I), av.ctypes.data_as(D), ll.ctypes.data_as(D), uu.ctypes.data_as(D), C.byref(s))
        if rc == 0:
            lib.mpcqp_cleanup(h)
        return rc
    assert call() == 0
    bad = qb.A_val[0].copy(); bad[0] = -2.0                 # the -1 of the first dynamics row: no stage structure any more
    assert call(A_val=bad) == 0                             # taken by the dense generic kernel (tests/test_dense_generic.py)
    l = qb.l[0].copy(); l[300] = qb.u[0][300] + 1.0
    assert call(l=l) == -3                                  # MPCQP_ERR_DATA
    eng.close()


@pytest.mark.gpu
def test_planner_mirror_receding_horizon(tmp_path):
    exe = build_host_test()
    out = tmp_path / "plan.bin"
    steps = 12
    r = subprocess.run([exe, "planner", str(steps), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    rows = np.fromfile(out, dtype=np.float64).reshape(steps, 10)
    assert np.isfinite(rows).all()
    assert (np.diff(rows[:, 1]) > 0).all()                  # perfect tracking moves along the reference
    assert np.abs(rows[:, 2]).max() < 5.0 and (rows[:, 3] > 0.5).all() and (rows[:, 3] < 4.5).all()
    assert set(rows[:, 6].astype(int)) <= {1, 2, -2}        # OSQP statuses the reference would consume as plans
    assert rows[0, 4] == 1 and (rows[1:, 4] >= 2).all()     # first step: one obstacle-free QP; then the intent candidates
    assert (rows[1:, 5] >= 0).all()
