"""Host / device split of the pipelined e2e loop (train_step_async): wall time per step, and the host time of a call when
the device is idle (= pure enqueue + staging cost).  GPU box only."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE

T, A, H, B, tied = bench.WORKLOADS["cfg2"]


class C:
    pass
for flags in [int(a) for a in sys.argv[1:]] or [0]:
    c = C()
    c.save = "/tmp/w"; c.batch = B; c.n_input = T + A; c.n_tracks = T; c.hidden = H; c.lr = bench.LR; c.reg_lambda = 0.0
    c.initval = "NULL"; c.seed = 0
    m = DAE(c).fit()
    m.set_debug(flags)
    batches = bench.make_batches("cfg2", 8, seed=7)
    for i in range(10):
        m.train_step_async(*batches[i % 8], bench.KP, bench.KP_IN)
    m.flush(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        m.train_step_async(*batches[i % 8], bench.KP, bench.KP_IN)
    m.flush(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 200
    # host cost of one call with an idle device: flush first, so nothing is waited for
    host = []
    for i in range(20):
        m.flush(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        m.train_step_async(*batches[i % 8], bench.KP, bench.KP_IN)
        host.append(time.perf_counter() - t1)
    m.flush()
    per_kind = [np.mean([host[i] for i in range(20) if i % 2 == k]) for k in (0, 1)]
    print("flags %5d: e2e %.4f ms/step; host call (idle device) median %.4f ms, tracks-batches %.4f, artist-batches %.4f; launches/step %.1f"
          % (flags, wall * 1e3, np.median(host) * 1e3, per_kind[0] * 1e3, per_kind[1] * 1e3, m.launch_count() / 230.0))
    m.close()
