"""Phase times of the fused decode + top-K over item ranges of different sizes on ONE GPU (what a shard of an N-way
item-sharded challenge inference runs), with the candidate-list lengths the filter passes produce.
  gpurun -- python tools/gpu_cfg5_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE
from tools.synth_mpd import SynthMPD

T, A, H, B, tied = bench.WORKLOADS["cfg5"]


class Conf:
    pass


conf = Conf()
conf.save = "/tmp/bench_w"; conf.n_input = T + A; conf.n_tracks = T; conf.n_output = T + A; conf.hidden = H
conf.lr = 0.01; conf.reg_lambda = 0.0; conf.initval = "NULL"; conf.DAEval = "NULL"; conf.seed = 0; conf.device = 0
conf.batch = B
m = DAE(conf)
m.trainable = False
m.fit()
g = SynthMPD(T, max(A, 1), n_clusters=64, seed=180610)
rng = np.random.default_rng(7)
trk, art, y, titles, tv, av = g.coo_batch(B, rng)
trk = np.ascontiguousarray(trk); tv = tv.astype(np.float32)
order = np.argsort(trk[:, 0], kind="stable")
bounds = np.searchsorted(trk[order, 0], np.arange(B + 1))
seeds = (bounds.astype(np.int32), trk[order, 1].astype(np.int32))
for div in (1, 2, 4, 8):
    hi = T // div // 128 * 128
    for i in range(2):
        m.recommend(trk, tv, seeds, k=500, item_range=(0, hi), on_device=True)
    m.set_profiling(True)
    for i in range(3):
        m.recommend(trk, tv, seeds, k=500, item_range=(0, hi), on_device=True)
    torch.cuda.synchronize()
    ph = {k: round(ms_ / max(n, 1), 3) for k, (ms_, n) in m.phase_times().items() if n}
    m.set_profiling(False)
    from tests.gpu_util import model_buf
    c = model_buf(m, "cand_cnt", torch.int32).cpu().numpy().reshape(3, -1)[:, :B]
    extra = " mean list length per pass: %s" % [int(v) for v in c.mean(1)]
    print("items %8d: %s%s" % (hi, ph, extra), flush=True)
m.close()
