"""GPU tests of the data-parallel train step: N ranks x B_local rows must reproduce one
rank x N*B_local rows.  `world` models attached through dae_model_attach_local share ONE GPU in this
process (the kernels, peer addressing and flag barriers are exactly those of the multi-GPU path);
the multi-process CUDA-IPC variant needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import dae_oracle as O
from spotify_recsys_challenge_2018_b200.dp import shard_coo
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE, DAE_tied
from tests.gpu_util import Conf, random_batch

pytestmark = pytest.mark.gpu


def _params(N, H, seed=5):
    ora = O.DAEOracle(N, H, 0.005, tied=False, seed=seed)
    ora.b_enc[:] = np.random.default_rng(1).normal(0, 0.1, H)
    ora.b_dec[:] = np.random.default_rng(2).normal(0, 0.1, N)
    return ora.params()


def _run_single(cls, N, T, H, B, params, batches, lam=0.0):
    m = cls(Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=11, reg_lambda=lam)).fit()
    m.set_params(params)
    costs = [m.train_step(x, xv, y, yv, 0.8, 0.75) for x, xv, y, yv in batches]
    out = m.get_params()
    m.close()
    return out, costs


@pytest.mark.parametrize("tied,world,N,T,H,b_local,lam", [(False, 2, 3001, 2500, 64, 64, 0.0),
                                                         (True, 2, 3001, 2500, 128, 128, 0.0),
                                                         (False, 4, 20000, 17000, 256, 64, 1e-4),
                                                         (False, 3, 1500, 1200, 64, 64, 0.0)])
def test_dp_local_equals_single(tied, world, N, T, H, b_local, lam):
    cls = DAE_tied if tied else DAE
    B = world * b_local
    params = _params(N, H)
    rng = np.random.default_rng(N + world)
    batches = []
    for i in range(3):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=25, empty_rows=(1,))
        x = trk if i % 2 == 0 else art
        batches.append((x, np.ones(len(x), np.float32), y, np.ones(len(y), np.float32)))
    want, want_costs = _run_single(cls, N, T, H, B, params, batches, lam)

    ms = [cls(Conf(batch=b_local, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=11, reg_lambda=lam,
                   world=world, rank=r)).fit() for r in range(world)]
    for m in ms:
        m.attach_local(ms)
        m.set_params(params)
    costs = []
    for x, xv, y, yv in batches:
        for r, m in enumerate(ms):
            m.stage_batch(0, *shard_coo(x, xv, r, b_local), *shard_coo(y, yv, r, b_local))
        for m in ms:
            m.backward_staged(0, 0.8, 0.75)
        for m in ms:
            m.apply_adam()
        cs = [m.sync_cost() for m in ms]
        assert max(cs) == min(cs)                              # every rank sums the partial costs in the same order
        costs.append(cs[0])
    got = [m.get_params() for m in ms]
    for g in got[1:]:
        for a, b in zip(g, got[0]):
            assert np.array_equal(a, b)                        # gathered parameters are identical on every rank
    for c, w in zip(costs, want_costs):
        assert abs(c - w) <= 1e-5 * abs(w)
    W_enc, W_dec, b_enc, b_dec = got[0]
    # decoder: same dz, same h_d, same contraction order over the batch columns -> bit-exact
    # (not asserted bit-exact: the fp32 atomics of the dW_enc scatter commute differently per layout, which
    #  reaches W_dec through h_d from the second step on)
    for a, b, name in zip(got[0], want, ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        assert (d > 1e-6).mean() < 2e-3 and d.max() <= 3 * 2.001 * 0.005, (name, d.max(), (d > 1e-6).mean())
    # inference on any rank scores every item: the gathered operand copy == bf16 of the gathered master
    x, xv, y, yv = batches[0]
    xs = shard_coo(x, xv, 0, b_local)
    ps = [m.predict(*xs) for m in ms]
    for m in ms:
        p, n, _ = m.buffer("W_dec_bf16_full")
        from tests.gpu_util import dev_view
        sh = dev_view(p, n, torch.bfloat16).float().cpu().numpy().reshape(N, H)
        assert np.array_equal(sh, O.bf16_round(W_dec))
    for p_ in ps[1:]:
        assert np.array_equal(p_, ps[0])
    for m in ms:
        m.close()


@pytest.mark.parametrize("world,b_local", [(4, 128), (3, 256)])
def test_dp_streamed_operand_vs_oracle(world, b_local):
    """world x bpad = 512 / 768 batch columns: the dW contraction streams BOTH operands (K > 256); K > 512 switches the
    fused dW + Adam kernel to its 3 + 3 ring layout."""
    N, T, H = 6007, 5000, 256
    B = world * b_local
    ora = O.DAEOracle(N, H, 0.005, tied=False, seed=5, mode="b200")
    ora.b_dec[:] = np.random.default_rng(2).normal(0, 0.1, N)
    ms = [DAE(Conf(batch=b_local, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=11, world=world, rank=r)).fit()
          for r in range(world)]
    for m in ms:
        m.attach_local(ms)
        m.set_params(ora.params())
    rng = np.random.default_rng(9)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=25)
    xv, yv = np.ones(len(trk), np.float32), np.ones(len(y), np.float32)
    for r, m in enumerate(ms):
        m.stage_batch(0, *shard_coo(trk, xv, r, b_local), *shard_coo(y, yv, r, b_local))
    for m in ms:
        m.backward_staged(0, 0.8, 0.75)
    for m in ms:
        m.apply_adam()
    cost = ms[0].sync_cost()
    for m in ms[1:]:
        m.sync_cost()
    c_ora = ora.train_step(trk, xv, y, yv, B, 0.8, 0.75, seed=11)
    assert abs(cost - c_ora) <= 1e-3 * abs(c_ora)
    got = ms[0].get_params()
    for a, b, name in zip(got, ora.params(), ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        assert (d > 1e-3 * 0.005).mean() < 5e-3 and d.max() <= 2.001 * 0.005, name
    for m in ms:
        m.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _ipc_worker(rank, world, port, q):
    import torch.distributed as dist
    from spotify_recsys_challenge_2018_b200.dp import DataParallelDAE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, T, H, b_local = 6007, 5000, 256, 128
    B = b_local * world
    m = DAE(Conf(batch=b_local, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=11, world=world, rank=rank,
                 device=rank)).fit()
    dp = DataParallelDAE(m)
    m.set_params(_params(N, H))
    rng = np.random.default_rng(3)
    costs = []
    for i in range(4):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=25)
        dp.stage_global_batch(i & 1, trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32))
        dp.train_step_staged(i & 1, 0.8, 0.75)
        costs.append(m.sync_cost())
    params = dp.get_params()
    if rank == 0:
        q.put((params, costs))
    dist.barrier()
    m.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dp_ipc_n_gpus_equals_single(world):
    """One PROCESS per GPU over CUDA IPC (the production path: peer stores through NVLink, three flag barriers per step):
    `world` ranks x 128 rows must reproduce one rank x (world * 128)... rows -- at world = 2 against ONE world = 1 model fed the
    whole batch (<= 256 rows fit a single training tile); at 4 and 8 against the same ranks run as `world` models in one
    process on one GPU (attach_local: the same kernels and peer addressing, no NVLink, no IPC)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, costs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    N, T, H, b_local = 6007, 5000, 256, 128
    B = b_local * world
    rng = np.random.default_rng(3)
    batches = []
    for i in range(4):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=25)
        batches.append((trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32)))
    if world == 2:
        want, want_costs = _run_single(DAE, N, T, H, B, _params(N, H), batches)
    else:
        ms = [DAE(Conf(batch=b_local, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=11, world=world, rank=r)).fit()
              for r in range(world)]
        for mm in ms:
            mm.attach_local(ms)
            mm.set_params(_params(N, H))
        want_costs = []
        for x, xv, y, yv in batches:
            for r, mm in enumerate(ms):
                mm.stage_batch(0, *shard_coo(x, xv, r, b_local), *shard_coo(y, yv, r, b_local))
            for mm in ms:
                mm.backward_staged(0, 0.8, 0.75)
            for mm in ms:
                mm.apply_adam()
            want_costs.append([mm.sync_cost() for mm in ms][0])
        want = ms[0].get_params()
        for mm in ms:
            mm.close()
    for c, w in zip(costs, want_costs):
        assert abs(c - w) <= 1e-5 * abs(w)
    # (not bit-exact over several steps: the fp32 atomics of the dW_enc scatter commute differently per layout)
    for a, b, name in zip(got, want, ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        assert (d > 1e-6).mean() < 2e-3, (name, d.max())


def test_sharded_recommender_single_rank_group():
    """dp.ShardedRecommender's device path (lists left on the device, peer stores, device merge) on a 1-rank group:
    must return exactly what the plain recommend call returns."""
    import torch.distributed as dist
    from spotify_recsys_challenge_2018_b200.dp import ShardedRecommender
    N, T, H, B, k = 40000, 36000, 128, 300, 500
    port = _free_port()
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, world_size=1, rank=0,
                            device_id=torch.device("cuda", 0))
    try:
        conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.01, DAEval=None)
        m = DAE(conf)
        m.trainable = False
        m.fit()
        rng = np.random.default_rng(5)
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=30)
        xv = np.ones(len(trk), np.float32)
        seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
        for flags in (32, 16):                       # dense ranking, fused decode + top-K
            m.set_debug(flags)
            want_i, want_s = m.recommend(trk, xv, seeds, k=k, return_scores=True)
            got_i, got_s = ShardedRecommender(m).recommend(trk, xv, seeds, k=k, return_scores=True)
            assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s), flags
        m.close()
    finally:
        dist.destroy_process_group()


# (ranks sharing ONE process and GPU: keep `world` small -- more streams than hardware queues (CUDA_DEVICE_MAX_CONNECTIONS,
# 8 by default) alias, and a rank's stores can end up queued behind a peer's spinning barrier kernel.  One process per GPU,
# the production layout, has no such coupling: bench.py --workload cfg5 --gpus N asserts the same equality there.)
@pytest.mark.parametrize("world,N,T,H,B,k", [(3, 40000, 36000, 128, 300, 500), (2, 300000, 270000, 64, 256, 500)])
def test_sharded_recommender_peer_store_merge_local(world, N, T, H, B, k):
    """`world` ShardedRecommenders in ONE process on one GPU (dae_exchange_attach_local: the same kernels, peer addressing
    and flag barrier as the one-process-per-GPU path): each rank ranks its item slice, stores its lists into every peer's
    merge buffer, and every rank's merged list must be exactly the unsharded ranking (main_challenge.py:26-36)."""
    from spotify_recsys_challenge_2018_b200.dp import ShardedRecommender
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.01, DAEval=None)
    ms = []
    for r in range(world):
        m = DAE(conf)
        m.trainable = False
        m.fit()
        ms.append(m)
    params = ms[0].get_params()
    for m in ms[1:]:
        m.set_params(params)
    recs = [ShardedRecommender(m, rank=r, world=world, max_k=k) for r, m in enumerate(ms)]
    for rc in recs:
        rc.attach_local(recs)
    rng = np.random.default_rng(N)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=30, empty_rows=(2,))
    xv = np.ones(len(trk), np.float32)
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    want_i, want_s = ms[0].recommend(trk, xv, seeds, k=k, return_scores=True)
    # the call is collective (every rank's barrier kernel waits for all the others): ranks of one process call it from
    # threads; several calls in a row exercise the call-parity double buffering of the merge buffer
    import threading
    res = [None] * world
    for rc in recs:                      # first call allocates the models' list buffers: cudaMalloc synchronises the DEVICE,
        # which ranks sharing one GPU must not do while a peer's barrier kernel is spinning (rank_shard itself is
        # collective -- the shards exchange their filter thresholds -- so the warm-up goes through the model)
        rc.model.recommend(trk, xv, seeds, k=k, item_range=rc.range, on_device=True)

    def run(r):
        for call in range(3):
            recs[r].share_thresholds = call != 1          # per-shard thresholds (call 1) and shared ones: the same lists
            recs[r].rank_shard(trk, xv, seeds, k)
            res[r] = recs[r].merge(k, return_scores=True)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t_ in ts:
        t_.start()
    for t_ in ts:
        t_.join(timeout=300)
    for r in range(world):
        got_i, got_s = res[r]
        assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s), r

    # the merge sharded by playlist (rows="own"): every rank returns exactly its rows of the unsharded ranking, the rows
    # of all ranks tile the batch, and the calls keep alternating the merge buffers with the "all" calls before them
    def run_own(r):
        for call in range(3):
            recs[r].rank_shard(trk, xv, seeds, k)
            res[r] = recs[r].merge(k, return_scores=True, rows="own")
    ts = [threading.Thread(target=run_own, args=(r,)) for r in range(world)]
    for t_ in ts:
        t_.start()
    for t_ in ts:
        t_.join(timeout=300)
    nxt = 0
    for r in range(world):
        r0, r1 = recs[r].rows
        assert r0 == nxt and r1 >= r0
        nxt = r1
        got_i, got_s = res[r]
        assert got_i.shape == (r1 - r0, k)
        assert np.array_equal(got_i, want_i[r0:r1]) and np.array_equal(got_s, want_s[r0:r1]), r
    assert nxt == B
    for rc in recs:
        rc.close()
    for m in ms:
        m.close()
