"""Stability run of the fused decode + top-K: N recommend calls of the cfg5 shape (4096 playlists x 2 M items) on one model,
every call's ids and scores compared with the first call's: the lists are exact, so they must not depend on the order in
which warps append candidates (fused == dense path is what tests/test_gpu_model.py checks, at sizes whose score matrix fits).  A protocol bug of the FILTER epilogue (queue, staging
halves, appends in flight, mbarrier pipeline) shows up here as a device trap, a hang (run it under `timeout`) or a mismatch.
  gpurun -- timeout 300 python tools/gpu_cfg5_soak.py [calls]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE
from tools.synth_mpd import SynthMPD

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
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
batches = []
for i in range(3):
    trk, art, y, titles, tv, av = g.coo_batch(B, rng)
    trk = np.ascontiguousarray(trk); tv = tv.astype(np.float32)
    order = np.argsort(trk[:, 0], kind="stable")
    bounds = np.searchsorted(trk[order, 0], np.arange(B + 1))
    batches.append((trk, tv, (bounds.astype(np.int32), trk[order, 1].astype(np.int32))))
first = [m.recommend(t, v, s, k=500, return_scores=True) for t, v, s in batches]
first = [(i.copy(), sc.copy()) for i, sc in first]
t0 = time.perf_counter()
bad = 0
for c in range(calls):
    t, v, s = batches[c % 3]
    i, sc = m.recommend(t, v, s, k=500, return_scores=True)
    if not (np.array_equal(i, first[c % 3][0]) and np.array_equal(sc, first[c % 3][1])):
        bad += 1
        print("call %d differs from the first call on batch %d: %d rows" % (c, c % 3, int((i != first[c % 3][0]).any(1).sum())), flush=True)
dt = time.perf_counter() - t0
print("cfg5 soak: %d calls, %.2f ms per call, %d mismatches" % (calls, 1e3 * dt / calls, bad), flush=True)
m.close()
sys.exit(0 if bad == 0 else 1)
