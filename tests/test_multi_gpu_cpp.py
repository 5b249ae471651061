"""One process, one host thread and one engine per worker (intent-mpc_b200/host/MultiGpuB200.hpp; SURVEY.md section 8(e),
BASELINE.json north_star "one host thread per GPU") through include/mpcqp_b200.h only.  CPU tier: it compiles, links and
refuses to run without a GPU.  GPU tier: 1, 2, 3 and 8 workers (device g mod device_count: several engines on one GPU on a
one-GPU box, one per GPU on a larger one) return bit-identical results, equal to the single Python engine's and — on a sample
— to the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from intent_mpc_b200 import workloads as W
from oracle import bindings as OB
from tests.helpers import to_qp_batch, rel_inf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "multi_gpu_test")


def build():
    import __graft_entry__ as G
    G.build()
    src = os.path.join(CPP, "multi_gpu_test.cpp")
    deps = [src, os.path.join(ROOT, "intent-mpc_b200", "host", "MultiGpuB200.hpp"), os.path.join(ROOT, "include", "mpcqp_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, src, "-L" + os.path.join(ROOT, "intent-mpc_b200"), "-lmpcqp_b200", "-lpthread",
                        "-Wl,-rpath," + os.path.join(ROOT, "intent-mpc_b200")], check=True)
    return EXE


def _write(mb, path):
    with open(path, "wb") as f:
        f.write(np.array([mb.B, mb.num_obs, mb.params.horizon], dtype=np.int64).tobytes())
        for k in ("x0", "xref", "obs_c", "obs_semi", "obs_yaw", "lin_pt", "warm_x"):
            f.write(np.ascontiguousarray(getattr(mb, k), dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(mb.obs_dyn, dtype=np.int32).tobytes())


def _read(path, B, n, G):
    raw = np.fromfile(path, dtype=np.float64)
    o = 0
    x = raw[o:o + B * n].reshape(B, n); o += B * n
    st, it, ru, obj = (raw[o + i * B:o + (i + 1) * B] for i in range(4)); o += 4 * B
    return dict(x=x, status=st.astype(np.int64), iter=it.astype(np.int64), rho_updates=ru.astype(np.int64), obj=obj, kernel_ms=raw[o:o + G])


def test_multi_engine_driver_compiles_and_refuses_without_gpu(tmp_path):
    exe = build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tier")
    mb = W.static_batch(4, num_obs=2)
    _write(mb, tmp_path / "in.bin")
    r = subprocess.run([exe, "2", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_thread_per_engine_results_do_not_depend_on_the_worker_count(tmp_path):
    from intent_mpc_b200 import engine
    exe = build()
    mb = W.static_batch(601, num_obs=4, seed0=9000)          # not divisible by the worker counts: ragged shards
    _write(mb, tmp_path / "in.bin")
    outs = {}
    for G in (1, 2, 3, 8):
        r = subprocess.run([exe, str(G), str(tmp_path / "in.bin"), str(tmp_path / f"out{G}.bin"), "2"], capture_output=True, text=True)
        assert r.returncode == 0, (G, r.returncode, r.stderr)
        outs[G] = _read(tmp_path / f"out{G}.bin", mb.B, mb.params.n, G)
        assert (outs[G]["kernel_ms"] > 0).all()
    for G in (2, 3, 8):
        for k in ("x", "status", "iter", "rho_updates", "obj"):
            assert np.array_equal(outs[G][k], outs[1][k]), (G, k)
    eng = engine.Engine(0)
    try:
        one = eng.solve_mpc_batch(mb)
        assert np.array_equal(one["x"], outs[3]["x"]) and np.array_equal(one["iter"], outs[3]["iter"])
    finally:
        eng.close()
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    ref = orc.solve_batch(to_qp_batch(mb.slice(280, 330)), want_y=False, nthreads=os.cpu_count() or 1)      # straddles shard borders
    assert (ref["status"] == outs[8]["status"][280:330]).all() and (ref["iter"] == outs[8]["iter"][280:330]).all()
    assert rel_inf(outs[8]["x"][280:330], ref["x"]).max() < 1e-5
