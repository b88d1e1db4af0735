"""Timeline of the LAST step of a pipelined train_step_async run (8 host batches, H2D every step), for several run lengths
so that every batch kind ends a run once.  GPU box only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from tools.gpu_trace import NAMES
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE
from tests.gpu_util import model_buf

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
T, A, H, B, tied = bench.WORKLOADS["cfg2"]


class C:
    pass
c = C()
c.save = "/tmp/w"; c.batch = B; c.n_input = T + A; c.n_tracks = T; c.hidden = H; c.lr = bench.LR; c.reg_lambda = 0.0
c.initval = "NULL"; c.seed = 0
m = DAE(c).fit()
m.set_debug(8192 | flags)
batches = bench.make_batches("cfg2", 8, seed=7)
print("flags", flags)
for n in (40, 41):
    for i in range(n):
        m.train_step_async(*batches[i % 8], bench.KP, bench.KP_IN)
    m.flush()
    tt = model_buf(m, "trace", torch.int64).cpu().numpy().astype(np.float64)
    last, prev = (n - 1) & 1, n & 1                       # step index of the run's last step is n_total - 1: parity by total count
    tot = getattr(m, "_tot", 0) + n; m._tot = tot
    last, prev = (tot - 1) & 1, tot & 1
    a, b = tt[16 * prev:16 * prev + len(NAMES)], tt[16 * last:16 * last + len(NAMES)]
    print(" run of %d: prev step: %s" % (n, " ".join("%s=%.0f" % (k, v) for k, v in zip(NAMES, (a - a[0]) / 1e3) if abs(v) < 1e7)))
    print("            last step (rel. to prev start): %s" % " ".join("%s=%.0f" % (k, v) for k, v in zip(NAMES, (b - a[0]) / 1e3) if abs(v) < 1e7))
m.close()
