"""world_size-2 gloo test (CPU) of the data-parallel contract and its host logic.

The GPU path shards the batch by playlist and the catalogue rows tile-cyclically (dp.py,
include/dae_b200.h "data parallelism"): rank r computes dz / h_d / da for its own rows, the OWNER of
an item tile contracts that tile's dz columns from every rank with every rank's h_d and applies
Adam to its rows, and the updated rows are handed back to everybody.  Here the per-rank compute is
the oracle and the exchange is gloo; the result must equal one oracle step on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dae_oracle as O
from spotify_recsys_challenge_2018_b200.dp import exchange_handles, shard_coo, tile_local_row, tile_owner

TILE = 8          # small tile so a 90-item catalogue spans several owners


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _batch(B, N, seed):
    rng = np.random.default_rng(seed)
    n = B * 12
    x = np.stack([np.sort(rng.integers(0, B, n)), rng.integers(0, N, n)], 1)
    y = np.stack([np.sort(rng.integers(0, B, 2 * n)), rng.integers(0, N, 2 * n)], 1)
    return x, np.ones(n, np.float32), y, np.ones(2 * n, np.float32)


def _gather(t):
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def _worker(rank, world, port, tied, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = exchange_handles(bytes([rank]) * 64)
    assert hs == [bytes([r]) * 64 for r in range(world)]
    B, N, H = 16, 90, 8
    b_local = B // world
    m = O.DAEOracle(N, H, 0.01, tied=tied, seed=4)
    x, xv, y, yv = _batch(B, N, 0)
    xs, xvs = shard_coo(x, xv, rank, b_local)
    ys, yvs = shard_coo(y, yv, rank, b_local)
    cost, g, f = m.loss_and_grads(xs, xvs, ys, yvs, b_local, 0.8, 0.7, seed=21, step=3,
                                  row_offset=rank * b_local, global_batch=B)
    mine = tile_owner(np.arange(N), world, TILE) == rank
    # decode side: the owner contracts dz columns of every rank with every rank's h_d
    dz_all = np.concatenate(_gather(torch.tensor(f["dz"])), 0)            # [B, N]
    hd_all = np.concatenate(_gather(torch.tensor(f["h_d"])), 0)           # [B, H]
    dW_dec = np.zeros((N, H), np.float32)
    dW_dec[mine] = dz_all[:, mine].T @ hd_all
    # encode side: every rank publishes (col, x_n, row) and da; the owner keeps the entries of its rows
    rows = O.csr_rows(f["row_ptr"])
    pub = np.full((b_local * 40, 3), -1.0, np.float64)
    pub[:len(rows)] = np.stack([f["col"], f["x_n"], rows], 1)
    pubs = _gather(torch.tensor(pub))
    das = _gather(torch.tensor(f["da"]))
    dW_enc = np.zeros((N, H), np.float32)
    for s in range(world):
        for c, xn, r in pubs[s]:
            if c >= 0 and tile_owner(int(c), world, TILE) == rank:
                dW_enc[int(c)] += np.float32(xn) * das[s][int(r)]
    # biases and cost: fixed-order sum of every rank's partial
    db_dec = sum(_gather(torch.tensor(g["b_dec"])))
    db_enc = sum(_gather(torch.tensor(g["b_enc"])))
    cost_all = sum(c[0] for c in _gather(torch.tensor([cost])))
    # the owner's dense Adam on its rows, then the rows go back to everybody
    grads = dict(W_enc=dW_enc, W_dec=dW_dec, b_enc=db_enc, b_dec=db_dec)
    m.apply_grads(grads)
    W_enc = sum(w * (tile_owner(np.arange(N), world, TILE) == s)[:, None] for s, w in enumerate(_gather(torch.tensor(m.W_enc))))
    W_dec = sum(w * (tile_owner(np.arange(N), world, TILE) == s)[:, None] for s, w in enumerate(_gather(torch.tensor(m.W_dec))))
    if rank == 0:
        out.put([W_enc, W_dec, m.b_enc, m.b_dec, cost_all])
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
    m.step = 3
    cost = m.train_step(x, xv, y, yv, B, 0.8, 0.7, seed=21)
    for a, b, name in zip(got[:4], m.params(), ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        # first Adam step moves every element by ~lr*sign(g): summation-order noise may flip near-zero gradients
        assert (d > 1e-5).mean() < 0.02, name
    assert abs(got[4] - cost) < 1e-5 * abs(cost)


def test_shard_coo_partitions_and_rebases():
    x, xv, _, _ = _batch(16, 50, 1)
    parts = [shard_coo(x, xv, r, 4) for r in range(4)]
    assert sum(len(p[0]) for p in parts) == len(x)
    for r, (p, v) in enumerate(parts):
        assert p[:, 0].min() >= 0 and p[:, 0].max() < 4
        sel = (x[:, 0] // 4) == r
        assert np.array_equal(p[:, 1], x[sel, 1])          # order preserved inside the shard


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_cyclic_ownership_is_a_bijection(world):
    N = 128 * 11 + 37
    item = np.arange(N)
    own, loc = tile_owner(item, world), tile_local_row(item, world)
    assert own.min() >= 0 and own.max() < world
    seen = set(zip(own.tolist(), loc.tolist()))
    assert len(seen) == N                                   # (owner, local row) identifies the item
    back = ((loc // 128) * world + own) * 128 + loc % 128   # csrc/kernels.h item_global
    assert np.array_equal(back, item)
    n_local = -(-(-(-N // 128)) // world) * 128             # rows every rank allocates
    assert loc.max() < n_local


@pytest.mark.parametrize("T,world", [(250000, 8), (2000000, 8), (1000, 3), (129, 2), (5, 4)])
def test_item_shard_partitions_the_tracks(T, world):
    from spotify_recsys_challenge_2018_b200.dp import item_shard
    r = [item_shard(T, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == T
    for a, b in zip(r, r[1:]):
        assert a[1] == b[0] and a[0] <= a[1]
    assert all(lo % 128 == 0 for lo, hi in r if lo < T)


def test_merge_topk_lists_equals_unsharded_ranking():
    """The merge rule of item-sharded inference (score desc, id asc) over per-shard top-k lists == the ranking of the
    whole catalogue (oracle/ranking.py), ties and short shards included."""
    from oracle import ranking
    from spotify_recsys_challenge_2018_b200.dp import item_shard, merge_topk_lists
    rng = np.random.default_rng(0)
    T, B, k, world = 3000, 6, 50, 4
    scores = np.round(rng.random((B, T)).astype(np.float32), 2)            # heavy ties
    idx_l, sc_l = [], []
    for r in range(world):
        lo, hi = item_shard(T, r, world)
        ii = np.full((B, k), -1, np.int32); ss = np.full((B, k), -np.inf, np.float32)
        for b in range(B):
            loc = ranking.topk_excluding_seeds(scores[b, lo:hi], [], k)
            ii[b, :len(loc)] = loc + lo; ss[b, :len(loc)] = scores[b, loc + lo]
        idx_l.append(ii); sc_l.append(ss)
    got_i, got_s = merge_topk_lists(idx_l, sc_l, k)
    for b in range(B):
        assert np.array_equal(got_i[b], ranking.topk_excluding_seeds(scores[b], [], k))
