"""GPU parity tests of the title branch (Char_CNN.py:16-75, DAEs.py:153-201) behind the C ABI against
oracle/title_oracle.py on the same seeded inputs: CNN features, mixed scores, loss, every gradient,
and the parameters after Adam."""
import numpy as np
import pytest
import torch

from oracle import dae_oracle as O
from oracle import ranking
from oracle import title_oracle as TO
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE_title
from spotify_recsys_challenge_2018_b200.models.title_get import get_model
from tests.gpu_util import Conf, dev_view, random_batch

pytestmark = pytest.mark.gpu


def _tbuf(tm, name, dtype):
    p, n, _ = tm.buffer(name)
    return dev_view(p, n, dtype)


def _setup(N, T, H, B, filter_num=100, filter_size=(3, 5, 7, 9), char_emb=50, L=25, C=41, lr=0.005):
    conf = Conf(batch=B, n_input=N, n_tracks=T, n_output=N, hidden=H, lr=lr, seed=11, DAEval=None, charsize=C,
                strmaxlen=L, char_emb=char_emb, char_model="Char_CNN", filter_num=filter_num,
                filter_size=list(filter_size))
    dae_o = O.DAEOracle(N, H, lr, tied=False, seed=5, mode="b200")
    dae_o.b_dec[:] = np.random.default_rng(2).normal(0, 0.2, N)
    cnn_o = TO.CharCNNOracle(C, L, char_emb, filter_num, list(filter_size), N, seed=7, mode="b200")
    tm = get_model(conf)
    m = DAE_title(conf, tm)
    m._create()
    m.set_params(dae_o.params())
    tm.fit(m)
    tm.set_params(cnn_o.params())
    return conf, dae_o, cnn_o, m, tm


def _titles(rng, B, L, C):
    t = rng.integers(0, C, (B, L))
    lens = rng.integers(0, L + 1, B)
    for r in range(B):
        t[r, lens[r]:] = -1
    return t


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.mark.parametrize("N,T,H,B,fn,fs", [(1500, 1200, 64, 64, 16, (3, 5)), (6007, 5000, 256, 250, 100, (3, 5, 7, 9)),
                                          (3001, 2500, 128, 150, 100, (3, 5, 7, 9))])
def test_title_predict_and_recommend(N, T, H, B, fn, fs):
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, fn, fs)
    rng = np.random.default_rng(N)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=20, empty_rows=(2,))
    xv = np.concatenate([np.ones(len(trk)), 0.5 * np.ones(len(art))]).astype(np.float32)
    titles = _titles(rng, B, 25, 41)
    use = (rng.random(B) < 0.8).astype(np.float32)
    ora = TO.DAETitleOracle(dae_o, cnn_o, 0.005)
    f = ora.forward(y, xv, titles, use, B)
    q_gpu = m.predict(y, xv, titles=titles, titles_use=use)
    D = fn * len(fs)
    feat = _tbuf(tm, "feat", torch.float32).cpu().numpy()[:B * D].reshape(B, D)
    np.testing.assert_allclose(feat, f["title"]["feat"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(_tbuf(tm, "w_t", torch.float32).cpu().numpy()[:B], f["w_t"][:, 0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(_tbuf(tm, "w_p", torch.float32).cpu().numpy()[:B], f["w_p"][:, 0], rtol=1e-5, atol=1e-7)
    assert _rel(q_gpu, f["q"], 1e-6).max() < 1e-3                          # north_star: 1e-3 relative on scores
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    idx = m.recommend(y, xv, seeds, titles=titles, titles_use=use, k=500)
    for r in range(0, B, max(B // 8, 1)):
        want = ranking.topk_excluding_seeds(q_gpu[r, :T], seeds[r], 500)
        assert np.array_equal(idx[r], want), r
    tm.close(); m.close()


@pytest.mark.parametrize("N,T,H,B,fn,fs", [(1500, 1200, 64, 64, 16, (3, 5)), (6007, 5000, 256, 250, 100, (3, 5, 7, 9))])
def test_title_train_step_stage_by_stage(N, T, H, B, fn, fs):
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, fn, fs)
    rng = np.random.default_rng(N + 1)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=20, empty_rows=(1,))
    yv = np.ones(len(y), np.float32)
    titles = _titles(rng, B, 25, 41)
    kp, kp_in, kp_t = 0.8, 0.5, 0.7
    ora = TO.DAETitleOracle(dae_o, cnn_o, 0.005)
    c_ora, grads, f = ora.loss_and_grads(y, yv, y, yv, titles, np.ones(B, np.float32), B, kp, kp_in, kp_t, seed=11, step=0)
    before = [p.copy() for p in cnn_o.params()]
    m.set_debug(16384)                                  # dW_out through HBM ("g_W_out"); the default keeps it in tensor memory
    cost = tm.train_step(m, y, yv, titles, kp, kp_t, kp_in)
    assert abs(cost - c_ora) <= 1e-3 * abs(c_ora), (cost, c_ora)
    D = fn * len(fs)
    bpad = (B + 63) // 64 * 64
    feat_d = _tbuf(tm, "feat_d", torch.bfloat16).float().cpu().numpy()[:bpad * 512].reshape(bpad, 512)
    assert np.array_equal(feat_d[:B, :D] != 0, (f["title"]["keep"] & (f["title"]["feat"] > 0)))   # identical dropout mask
    assert np.all(feat_d[B:] == 0) and np.all(feat_d[:, D:] == 0)
    dzT = _tbuf(tm, "dzT", torch.bfloat16).float().cpu().numpy()[:N * bpad].reshape(N, bpad)
    dz_o = f["dz_t"].T
    err = _rel(dzT[:, :B], dz_o, 1e-3 * np.abs(dz_o).max())
    assert err.max() < 3e-2 and err.mean() < 3e-3
    names = ["g_emb", "g_conv_W", "g_conv_b"]
    g_emb = _tbuf(tm, "g_emb", torch.float32).cpu().numpy().reshape(41, 50)
    np.testing.assert_allclose(g_emb, grads[0], rtol=0, atol=3e-2 * np.abs(grads[0]).max())
    g_cw = _tbuf(tm, "g_conv_W", torch.float32).cpu().numpy()
    off = 0
    for i, w in enumerate(fs):
        want = grads[1 + 2 * i]
        got = g_cw[off:off + want.size].reshape(want.shape)
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-2 * np.abs(want).max())
        off += want.size
    g_cb = _tbuf(tm, "g_conv_b", torch.float32).cpu().numpy().reshape(len(fs), fn)
    for i in range(len(fs)):
        np.testing.assert_allclose(g_cb[i], grads[2 + 2 * i], rtol=0, atol=3e-2 * np.abs(grads[2 + 2 * i]).max())
    # dW_out: two dense column blocks [N, h0] ++ [N, h1] (include/dae_b200.h)
    h0 = 256 if D >= 256 else (D + 63) // 64 * 64
    h1 = (D - 256 + 63) // 64 * 64 if D > 256 else 0
    g_raw = _tbuf(tm, "g_W_out", torch.float32).cpu().numpy()
    Np = (N + 127) // 128 * 128
    g_wo = np.concatenate([g_raw[:N * h0].reshape(N, h0), g_raw[Np * h0:Np * h0 + N * h1].reshape(N, h1)], 1)
    live = np.r_[0:min(D, 256), h0:h0 + max(D - 256, 0)]
    dead = np.setdiff1d(np.arange(h0 + h1), live)
    assert np.all(g_wo[:, dead] == 0)
    np.testing.assert_allclose(g_wo[:, live].T, grads[-2], rtol=0, atol=2e-2 * np.abs(grads[-2]).max())
    np.testing.assert_allclose(_tbuf(tm, "g_b_out", torch.float32).cpu().numpy(), grads[-1], rtol=5e-3,
                               atol=1e-3 * np.abs(grads[-1]).max())
    # Adam: first step moves every element by ~lr * sign(g)
    for i, (p, g) in enumerate(zip(cnn_o.params(), grads)):
        ora.adam.apply("t%d" % i, p, g)
    got = tm.get_params()
    for a, b, b0 in zip(got, cnn_o.params(), before):
        d = np.abs(a - b)
        assert (d > 1e-3 * 0.005).mean() < 1e-2 and d.max() <= 2.001 * 0.005
        assert np.abs(a - b0).max() > 0                                     # every variable moved
    shadow = _tbuf(tm, "W_out_bf16", torch.bfloat16).float().cpu().numpy().reshape(N, 512)[:, :D]
    assert np.array_equal(shadow.T, O.bf16_round(got[-2]))
    tm.close(); m.close()


@pytest.mark.parametrize("N,T,H,B,fn,fs", [(6007, 5000, 256, 250, 100, (3, 5, 7, 9)), (3001, 2500, 128, 150, 40, (3, 5, 7)),
                                          (20000, 17000, 256, 256, 128, (3, 5, 7, 9))])
def test_title_fused_dw_adam_equals_two_kernel_path(N, T, H, B, fn, fs):
    """Default: the dW_out tile of each column block stays in tensor memory and the dense TF1 Adam is applied from there
    (k_dw_adam_fused over 2-D TMA boxes of the [N, 512] layout; D = 400 -> blocks of 256 + 192 columns).  Debug bit 14:
    dW_out through HBM + k_adam_rows_vec4.  Same MMA order, same rounded Adam ops: master, moments and the bf16 operand
    copy agree bit for bit after several steps (Char_CNN.py:62-75 trained by DAEs.py:198)."""
    rng = np.random.default_rng(N)
    steps = []
    for i in range(3):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=20, empty_rows=(1,))
        steps.append((y, np.ones(len(y), np.float32), _titles(rng, B, 25, 41)))
    out = []
    for flags in (0, 16384):
        conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, fn, fs)
        m.set_debug(flags)
        costs = [tm.train_step(m, y, yv, titles, 0.8, 0.7, 0.3) for y, yv, titles in steps]
        snap = {k: _tbuf(tm, k, torch.bfloat16 if k == "W_out_bf16" else torch.float32).view(torch.int16 if k == "W_out_bf16" else torch.int32).clone()
                for k in ("W_out", "m_W_out", "v_W_out", "W_out_bf16")}
        out.append((costs, snap, [p.copy() for p in tm.get_params()]))
        tm.close(); m.close()
    assert out[0][0] == out[1][0]
    for k in out[0][1]:
        assert torch.equal(out[0][1][k], out[1][1][k]), k
    for a, b in zip(out[0][2], out[1][2]):
        assert np.array_equal(a, b)


def test_title_training_trajectory_matches_oracle():
    N, T, H, B = 3000, 2500, 64, 128
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, 32, (3, 5, 7), lr=0.01)
    ora = TO.DAETitleOracle(dae_o, cnn_o, 0.01)
    rng = np.random.default_rng(4)
    for step in range(8):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=15)
        yv = np.ones(len(y), np.float32)
        titles = _titles(rng, B, 25, 41)
        c_gpu = tm.train_step(m, y, yv, titles, 0.8, 0.7, 0.3)
        c_ora = ora.train_step(y, yv, y, yv, titles, np.ones(B, np.float32), B, 0.8, 0.3, 0.7, seed=11)
        tol = 1e-3 if step < 3 else 1e-2
        assert abs(c_gpu - c_ora) <= tol * abs(c_ora), (step, c_gpu, c_ora)
    tm.close(); m.close()


def test_title_pipelined_step_equals_synchronous_step():
    """train_step_async (the runner's call in --title mode) returns the same costs as train_step, one call late, and leaves
    the same parameters (the title step has no atomics: bit for bit)."""
    N, T, H, B = 3000, 2500, 64, 128
    rng = np.random.default_rng(6)
    steps = []
    for i in range(5):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=15)
        steps.append((y, np.ones(len(y), np.float32), _titles(rng, B, 25, 41)))
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, 32, (3, 5, 7), lr=0.01)
    sync_costs = [tm.train_step(m, y, yv, t, 0.8, 0.7, 0.3) for y, yv, t in steps]
    p1 = tm.get_params(); tm.close(); m.close()
    conf, dae_o, cnn_o, m, tm = _setup(N, T, H, B, 32, (3, 5, 7), lr=0.01)
    got = [tm.train_step_async(m, y, yv, t, 0.8, 0.7, 0.3) for y, yv, t in steps]
    assert got[0] is None
    got = got[1:] + [tm.flush()]
    assert tm.flush() is None
    p2 = tm.get_params(); tm.close(); m.close()
    assert got == sync_costs
    for a, b in zip(p1, p2):
        assert np.array_equal(a, b)
