import numpy as np, torch, sys
sys.path.insert(0, ".")
from tests.test_gpu_title import _setup, _titles, _tbuf
from tests.gpu_util import random_batch
N, T, H, B, fn, fs = 6007, 5000, 256, 250, 100, (3, 5, 7, 9)
rng = np.random.default_rng(N)
steps = []
for i in range(1):
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=20, empty_rows=(1,))
    steps.append((y, np.ones(len(y), np.float32), _titles(rng, B, 25, 41)))
out = []
for flags in (0, 16384):
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, fn, fs)
    m.set_debug(flags)
    costs = [tm.train_step(m, y, yv, titles, 0.8, 0.7, 0.3) for y, yv, titles in steps]
    snap = {k: _tbuf(tm, k, torch.float32).clone().cpu().numpy() for k in ("W_out", "m_W_out", "v_W_out")}
    out.append(snap)
    tm.close(); m.close()
Np = (N + 127) // 128 * 128
for k in out[0]:
    a, b = out[0][k], out[1][k]
    d = np.nonzero(a.view(np.int32) != b.view(np.int32))[0]
    print(k, "differ", len(d), "of", a.size)
    if len(d):
        blk = d >= Np * 256
        i0 = d[~blk]; i1 = d[blk] - Np * 256
        print("  block0:", len(i0), "rows", np.unique(i0 // 256)[:10], "cols", np.unique(i0 % 256)[:10])
        print("  block1:", len(i1), "rows", np.unique(i1 // 192)[:10], "cols", np.unique(i1 % 192)[:20])
        j = d[:5]; print("  vals", a[j], b[j])
