"""world_size-2 gloo test (CPU) of the data-parallel contract: sharding a reader batch by rows,
scaling by 1/B_global, keying dropout by the global row and sum-all-reducing the gradients
reproduces the single-process gradients of the whole batch.  The per-rank compute is
the oracle here (no GPU in this container); the GPU path uses the same shard_coo / allreduce_grads."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dae_oracle as O
from spotify_recsys_challenge_2018_b200.dp import allreduce_grads, shard_coo


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _batch(B, N, seed):
    rng = np.random.default_rng(seed)
    n = B * 12
    x = np.stack([np.sort(rng.integers(0, B, n)), rng.integers(0, N, n)], 1)
    y = np.stack([np.sort(rng.integers(0, B, 2 * n)), rng.integers(0, N, 2 * n)], 1)
    return x, np.ones(n, np.float32), y, np.ones(2 * n, np.float32)


def _worker(rank, world, port, tied, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, N, H = 16, 90, 8
    b_local = B // world
    m = O.DAEOracle(N, H, 0.01, tied=tied, seed=4)
    x, xv, y, yv = _batch(B, N, 0)
    xs, xvs = shard_coo(x, xv, rank, b_local)
    ys, yvs = shard_coo(y, yv, rank, b_local)
    cost, g, _ = m.loss_and_grads(xs, xvs, ys, yvs, b_local, 0.8, 0.7, seed=21, step=3,
                                  row_offset=rank * b_local, global_batch=B)
    ts = [torch.tensor(g[k]) for k in ("W_enc", "W_dec", "b_enc", "b_dec")] + [torch.tensor([cost])]
    flags = torch.tensor((np.abs(g["W_enc"]).sum(1) > 0).astype(np.uint8))
    allreduce_grads(ts, None, flags)
    if rank == 0:
        out.put([t.numpy() for t in ts] + [flags.numpy()])
    dist.destroy_process_group()


@pytest.mark.parametrize("tied", [False, True])
def test_dp_sharded_equals_single(tied):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tied, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, N, H = 16, 90, 8
    m = O.DAEOracle(N, H, 0.01, tied=tied, seed=4)
    x, xv, y, yv = _batch(B, N, 0)
    cost, g, _ = m.loss_and_grads(x, xv, y, yv, B, 0.8, 0.7, seed=21, step=3)
    for a, k in zip(got[:4], ("W_enc", "W_dec", "b_enc", "b_dec")):
        np.testing.assert_allclose(a, g[k], rtol=1e-4, atol=1e-7)
    assert abs(got[4][0] - cost) < 1e-5 * abs(cost)
    assert np.array_equal(got[5].astype(bool), np.abs(g["W_enc"]).sum(1) > 0)


def test_shard_coo_partitions_and_rebases():
    x, xv, _, _ = _batch(16, 50, 1)
    parts = [shard_coo(x, xv, r, 4) for r in range(4)]
    assert sum(len(p[0]) for p in parts) == len(x)
    for r, (p, v) in enumerate(parts):
        assert p[:, 0].min() >= 0 and p[:, 0].max() < 4
        sel = (x[:, 0] // 4) == r
        assert np.array_equal(p[:, 1], x[sel, 1])          # order preserved inside the shard
