"""N > 1 path on CPU (gloo, world size 2): contiguous batch-index sharding, independent solves per rank with no
collective on the solve path, one all_gather for verification — the gathered result must equal the unsharded solve.
The per-rank solver here is the oracle (there is no GPU in this tier); on the GPU box bench.py runs the same sharding
with the CUDA engine and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intent_mpc_b200 import sharding, workloads as W


def test_shard_bounds_cover_exactly():
    for B, G, unit in [(1024, 8, 1), (65536, 8, 6 * 0 + 1), (10, 4, 1), (66, 4, 6), (7, 2, 1), (6, 8, 6)]:
        b = sharding.shard_bounds(B, G, unit)
        assert b[0][0] == 0 and b[-1][1] == B
        assert all(b[g][1] == b[g + 1][0] for g in range(G - 1))
        assert all(lo % unit == 0 and hi % unit == 0 and hi >= lo for lo, hi in b)
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= unit


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings as OB
    from tests.helpers import to_qp_batch
    mb = W.static_batch(B, num_obs=2, seed0=900)
    mine, (lo, hi) = sharding.shard(mb, rank, world)
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    out = orc.solve_batch(to_qp_batch(mine), want_y=False)
    x = sharding.gather_rows(out["x"], B, world)
    it = sharding.gather_rows(out["iter"].astype(np.int64), B, world)
    if rank == 0:
        q.put((x.numpy(), it.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather_matches_unsharded():
    from oracle import bindings as OB
    from tests.helpers import to_qp_batch
    B, world = 11, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    x, it = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = OB.RefOsqp() if OB.RefOsqp.available() else OB.PortOsqp()
    ref = orc.solve_batch(to_qp_batch(W.static_batch(B, num_obs=2, seed0=900)), want_y=False)
    assert np.array_equal(it, ref["iter"])
    assert np.array_equal(x, ref["x"])
