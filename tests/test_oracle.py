"""CPU tests: the oracle against its pins (golden vectors from the reference, known-answer cases,
independent derivations).  No GPU, no /root/reference at run time."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dae_oracle as O
from oracle import philox, ranking
from oracle.tf1_graph_cpu import TF1GraphCPU

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- Philox: Random123 known answers
@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_kat(ctr, key, out):
    r = philox.philox4x32_10(*ctr, *key)
    assert tuple(int(x) for x in r) == out


def test_keep_mask_rate_and_identity():
    rows, cols = np.meshgrid(np.arange(200, dtype=np.uint32), np.arange(500, dtype=np.uint32), indexing="ij")
    m = philox.keep_mask(7, philox.STREAM_INPUT, 3, rows, cols, 0.75)
    assert abs(m.mean() - 0.75) < 0.01
    assert philox.keep_mask(7, 0, 3, rows, cols, 1.0).all()
    m2 = philox.keep_mask(7, philox.STREAM_INPUT, 4, rows, cols, 0.75)
    assert (m != m2).any()


# ---------------------------------------------------------------- a2: densify, last wins
def test_densify_last_wins_and_csr_agree():
    rng = np.random.default_rng(0)
    B, N = 6, 15
    pos = np.stack([rng.integers(0, B, 80), rng.integers(0, N, 80)], 1)
    val = rng.integers(0, 3, 80).astype(np.float32)
    dense = O.densify_last_wins(pos, val, B, N)
    # hand check of one duplicate
    pos2 = np.array([[0, 2], [0, 2], [1, 1]]); val2 = np.array([1.0, 0.0, 5.0], np.float32)
    d2 = O.densify_last_wins(pos2, val2, 2, 3)
    assert d2[0, 2] == 0.0 and d2[1, 1] == 5.0 and d2.sum() == 5.0
    row_ptr, col, v = O.coo_to_csr_last_wins(pos, val, B, N)
    rebuilt = np.zeros((B, N), np.float32)
    rebuilt[O.csr_rows(row_ptr), col] = v
    assert np.array_equal(rebuilt, dense)
    for r in range(B):      # sorted unique columns per row
        c = col[row_ptr[r]:row_ptr[r + 1]]
        assert np.all(np.diff(c) > 0)


def test_empty_rows_and_empty_batch():
    row_ptr, col, v = O.coo_to_csr_last_wins(np.zeros((0, 2)), [], 4, 10)
    assert row_ptr.tolist() == [0] * 5 and col.size == 0
    # float64 positions as the reference emits for batches with an empty row
    row_ptr, col, v = O.coo_to_csr_last_wins(np.array([[2.0, 3.0]]), [1.0], 4, 10)
    assert row_ptr.tolist() == [0, 0, 0, 1, 1] and col.tolist() == [3]


# ---------------------------------------------------------------- known-answer: tiny forward / loss (B=2,N=5,H=2)
def test_known_answer_tiny_forward_loss():
    W = np.array([[0.1, -0.2], [0.3, 0.4], [-0.5, 0.6], [0.7, -0.8], [0.9, 1.0]], np.float32)
    b_enc = np.array([0.05, -0.05], np.float32)
    b_dec = np.array([0.0, 0.1, -0.1, 0.2, -0.2], np.float32)
    m = O.DAEOracle(5, 2, lr=0.01, tied=True, params=[W, W, b_enc, b_dec])
    x_pos = np.array([[0, 0], [0, 2], [1, 4]]); x_val = [1, 1, 1]
    f = m.forward(x_pos, x_val, 2)
    # row 0: x_n = [.5, 0, .5, 0, 0]; row 1: [0,0,0,0,1]
    a0 = 0.5 * W[0] + 0.5 * W[2] + b_enc
    a1 = W[4] + b_enc
    sig = lambda t: 1 / (1 + np.exp(-t))
    h = np.stack([sig(a0), sig(a1)])
    np.testing.assert_allclose(f["h"], h, rtol=1e-6)
    p = sig(h @ W.T + b_dec)
    np.testing.assert_allclose(f["p"], p, rtol=1e-5)
    y = O.densify_last_wins(np.array([[0, 0], [0, 2], [0, 3], [1, 4], [1, 1]]), np.ones(5), 2, 5)
    L = -(y * np.log(p + 1e-10) + 0.55 * (1 - y) * np.log(1 - p + 1e-10)).sum(1)
    np.testing.assert_allclose(O.bce_rows(f["p"], y), L, rtol=1e-5)


def test_sigmoid_saturation_fp32():
    # sigma(17) == 1.0f, so log(1-p+eps) clamps at log(1e-10) and p(1-p) == 0
    p = O.sigmoid32(np.array([17.0], np.float32))
    assert p[0] == np.float32(1.0)
    dz = O.bce_dz(p, np.zeros(1, np.float32), 1.0)
    assert dz[0] == 0.0
    L = O.bce_rows(p.reshape(1, 1), np.zeros((1, 1), np.float32))
    np.testing.assert_allclose(L, 0.55 * 23.02585, rtol=1e-5)


# ---------------------------------------------------------------- a7: closed-form backward vs autograd (fp64)
@pytest.mark.parametrize("tied", [True, False])
def test_backward_matches_autograd(tied):
    rng = np.random.default_rng(1)
    B, N, H = 4, 12, 3
    m = O.DAEOracle(N, H, lr=0.01, tied=tied, seed=3)
    m.b_enc[:] = rng.normal(0, 0.1, H); m.b_dec[:] = rng.normal(0, 0.1, N)
    x_pos = np.stack([rng.integers(0, B, 14), rng.integers(0, N, 14)], 1); x_val = np.ones(14, np.float32)
    y_pos = np.stack([rng.integers(0, B, 20), rng.integers(0, N, 20)], 1); y_val = np.ones(20, np.float32)
    kp, kp_in = 0.8, 0.75
    cost, g, f = m.loss_and_grads(x_pos, x_val, y_pos, y_val, B, kp, kp_in, seed=5, step=2)
    # independent: torch autograd in float64 on the dense graph with the same masks
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    W_enc = t(m.W_enc); W_dec = W_enc if tied else t(m.W_dec); b_enc = t(m.b_enc); b_dec = t(m.b_dec)
    x = torch.tensor(O.densify_last_wins(x_pos, x_val, B, N).astype(np.float64))
    keep = np.zeros((B, N)); keep[O.csr_rows(f["row_ptr"]), f["col"]] = f["keep_in"]
    x_d = x / kp_in * torch.tensor(keep)
    x_n = x_d / (x_d.sum(1, keepdim=True) + 1e-10)
    h = torch.sigmoid(x_n @ W_enc + b_enc)
    h_d = h * torch.tensor(f["keep_h"].astype(np.float64)) / kp
    p = torch.sigmoid(h_d @ W_dec.T + b_dec)
    y = torch.tensor(f["y"].astype(np.float64))
    L = -(y * torch.log(p + 1e-10) + 0.55 * (1 - y) * torch.log(1 - p + 1e-10)).sum(1)
    c = L.mean()
    c.backward()
    assert abs(float(c) - cost) < 1e-4 * abs(cost)
    gW = g["W_enc"] + g["W_dec"] if tied else g["W_enc"]
    np.testing.assert_allclose(gW, W_enc.grad.numpy(), rtol=2e-3, atol=2e-6)
    if not tied:
        np.testing.assert_allclose(g["W_dec"], W_dec.grad.numpy(), rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(g["b_enc"], b_enc.grad.numpy(), rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(g["b_dec"], b_dec.grad.numpy(), rtol=2e-3, atol=2e-6)


# ---------------------------------------------------------------- a8: TF1 Adam known answer
def test_adam_tf1_first_steps():
    a = O.AdamTF1(0.01)
    w = np.array([1.0, -2.0], np.float32); g = np.array([0.5, -0.25], np.float32)
    a.apply("w", w, g); a.finish_step()
    # step 1: m = .1 g, v = .001 g^2, alpha = lr*sqrt(1-.999)/(1-.9); update = alpha*m/(sqrt(v)+eps) ~ lr*sign(g)
    alpha = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = np.array([1.0, -2.0]) - alpha * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
    np.testing.assert_allclose(w, exp, rtol=1e-6)
    np.testing.assert_allclose(w, [1.0 - 0.01, -2.0 + 0.01], rtol=1e-5)
    # a zero gradient still moves the weight on the next step (dense Adam, m decays)
    w0 = w.copy()
    a.apply("w", w, np.zeros(2, np.float32)); a.finish_step()
    assert np.all(w != w0)


# ---------------------------------------------------------------- dense TF1 graph port == sparse oracle
@pytest.mark.parametrize("tied", [True, False])
def test_tf1_graph_port_matches_oracle(tied):
    rng = np.random.default_rng(2)
    B, N, H = 8, 60, 16
    ref = O.DAEOracle(N, H, lr=0.01, tied=tied, seed=1)
    port = TF1GraphCPU(N, H, lr=0.01, tied=tied, params=[p.copy() for p in ref.params()])
    for step in range(3):
        x_pos = np.stack([rng.integers(0, B, 40), rng.integers(0, N, 40)], 1); x_val = np.ones(40, np.float32)
        y_pos = np.stack([rng.integers(0, B, 70), rng.integers(0, N, 70)], 1); y_val = np.ones(70, np.float32)
        c_ref, g, f = ref.loss_and_grads(x_pos, x_val, y_pos, y_val, B, 0.8, 0.7, seed=9, step=step)
        ref.apply_grads(g)
        keep_dense = np.zeros((B, N), bool); keep_dense[O.csr_rows(f["row_ptr"]), f["col"]] = f["keep_in"]
        c_port = port.train_step(x_pos, x_val, y_pos, y_val, B, 0.8, 0.7,
                                 keep_in_dense=torch.tensor(keep_dense), keep_h=torch.tensor(f["keep_h"]))
        assert abs(c_ref - c_port) < 1e-4 * abs(c_ref)
    np.testing.assert_allclose(port.vars["W_enc"].numpy(), ref.W_enc, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(port.vars["b_dec"].numpy(), ref.b_dec, rtol=1e-4, atol=2e-5)
    x_pos = np.stack([rng.integers(0, B, 30), rng.integers(0, N, 30)], 1)
    np.testing.assert_allclose(port.predict(x_pos, np.ones(30), B), ref.predict(x_pos, np.ones(30), B),
                               rtol=1e-4, atol=1e-6)


def test_bf16_round_matches_torch():
    x = np.random.default_rng(0).normal(0, 1, 10000).astype(np.float32)
    x[:4] = [0.0, 1.0, -1.0000001, 3.3895314e38]
    ours = O.bf16_round(x)
    theirs = torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(ours, theirs)


def test_b200_mode_close_to_fp32():
    rng = np.random.default_rng(4)
    B, N, H = 8, 200, 64
    a = O.DAEOracle(N, H, lr=0.01, seed=1, mode="fp32")
    b = O.DAEOracle(N, H, lr=0.01, seed=1, mode="b200")
    x_pos = np.stack([rng.integers(0, B, 50), rng.integers(0, N, 50)], 1)
    pa, pb = a.predict(x_pos, np.ones(50), B), b.predict(x_pos, np.ones(50), B)
    np.testing.assert_allclose(pa, pb, rtol=1e-2)


# ---------------------------------------------------------------- ranking + metrics vs the reference's own code
def test_ranking_matches_reference_cand_generate():
    g = np.load(os.path.join(GOLDEN, "ranking_golden.npz"), allow_pickle=False)
    for s, c, sd in zip(g["scores"], g["cands"], g["seeds"]):
        ours = ranking.topk_excluding_seeds(s, json.loads(str(sd)), 500)
        assert np.array_equal(ours, c)


def test_ranking_ties_canonical():
    s = np.array([0.5, 1.0, 1.0, 0.5, 1.0], np.float32)
    assert ranking.topk_excluding_seeds(s, [2], 3).tolist() == [1, 4, 0]


def test_metrics_match_reference():
    with open(os.path.join(GOLDEN, "metrics_golden.json")) as f:
        cases = json.load(f)
    for c in cases:
        assert ranking.r_precision(c["answer"], c["cand"]) == pytest.approx(c["r_precision"], abs=0)
        assert ranking.ndcg(c["answer"], c["cand"]) == pytest.approx(c["ndcg"], rel=1e-12)
        assert ranking.rsc(c["answer"], c["cand"]) == c["rsc"]


def test_merge_sharded_topk_equals_unsharded():
    rng = np.random.default_rng(5)
    T, K = 4000, 100
    s = rng.permutation(T).astype(np.float32)
    s[rng.integers(0, T, 500)] = 7.0      # ties
    full = ranking.topk_excluding_seeds(s, [3, 5], K)
    idxs, scs = [], []
    for g in range(8):
        lo, hi = g * T // 8, (g + 1) * T // 8
        loc = ranking.topk_excluding_seeds(s[lo:hi], [x - lo for x in (3, 5) if lo <= x < hi], K)
        idxs.append(loc + lo); scs.append(s[loc + lo])
    m_idx, _ = ranking.merge_sharded_topk(idxs, scs, K)
    assert np.array_equal(m_idx, full)


# ---------------------------------------------------------------- title branch (Char_CNN.py, DAEs.py:153-201)
def _title_setup(B=6, N=40, H=8, seed=0):
    from oracle import title_oracle as TO
    dae = O.DAEOracle(N, H, 0.01, tied=False, seed=seed)
    dae.b_dec[:] = np.random.default_rng(1).normal(0, 0.3, N)
    cnn = TO.CharCNNOracle(charsize=11, strmaxlen=12, char_emb=5, filter_num=4, filter_size=[3, 5], n_output=N, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    titles = rng.integers(0, 11, (B, 12))
    titles[0, 7:] = -1; titles[1, :] = -1; titles[2, 3:] = -1          # padded, empty and short titles
    n = B * 5
    x = np.stack([np.sort(rng.integers(0, B, n)), rng.integers(0, N, n)], 1)
    y = np.stack([np.sort(rng.integers(0, B, 2 * n)), rng.integers(0, N, 2 * n)], 1)
    use = (rng.random(B) < 0.7).astype(np.float32)
    return TO, dae, cnn, titles, x, np.ones(n, np.float32), y, np.ones(2 * n, np.float32), use


def test_title_backward_matches_autograd():
    """Closed-form backward of the title branch vs torch autograd in fp64 on the same masks."""
    import torch
    TO, dae, cnn, titles, x, xv, y, yv, use = _title_setup()
    B = titles.shape[0]
    m = TO.DAETitleOracle(dae, cnn, 0.01)
    cost, grads, f = m.loss_and_grads(x, xv, y, yv, titles, use, B, 0.8, 0.7, 0.6, seed=3, step=2)
    # torch restatement, differentiable w.r.t. every title variable
    P = [torch.tensor(p.astype(np.float64), requires_grad=True) for p in cnn.params()]
    emb = P[0]
    tt = torch.tensor(titles)
    ok = (tt >= 0).double().unsqueeze(-1)
    xe = emb[tt.clamp(min=0)] * ok                                       # [B, L, E]
    feats = []
    for i, w in enumerate(cnn.fs):
        W, b = P[1 + 2 * i], P[2 + 2 * i]
        win = xe.unfold(1, w, 1).permute(0, 1, 3, 2)                       # [B, P, w, E]
        conv = torch.relu(torch.einsum("bpke,kef->bpf", win, W) + b)
        feats.append(conv.max(1).values)
    feat = torch.cat(feats, 1)
    keep = torch.tensor(f["title"]["keep"].astype(np.float64))
    feat_d = feat / 0.6 * keep
    t = torch.sigmoid(feat_d @ P[-2] + P[-1])
    p = torch.tensor(f["dae"]["p"].astype(np.float64))
    q = t * torch.tensor(f["w_t"].astype(np.float64)) + p * torch.tensor(f["w_p"].astype(np.float64))
    yt = torch.tensor(f["y"].astype(np.float64))
    L = -(yt * torch.log(q + 1e-10) + 0.55 * (1 - yt) * torch.log(1 - q + 1e-10)).sum(1)
    c = L.mean()
    c.backward()
    assert abs(float(c) - cost) < 1e-5 * abs(cost)
    for g, pt in zip(grads, P):
        ref = pt.grad.numpy()
        assert np.abs(g - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-6), (g.shape, np.abs(g - ref).max())


def test_title_mix_weights_and_padding():
    TO, dae, cnn, titles, x, xv, y, yv, use = _title_setup()
    B = titles.shape[0]
    m = TO.DAETitleOracle(dae, cnn, 0.01)
    f = m.forward(x, xv, titles, use, B)
    # titles_use == 0 -> pure DAE score; an all-pad title still produces a finite score
    for r in range(B):
        if use[r] == 0 and f["dae"]["s"][r] > 0:
            np.testing.assert_allclose(f["q"][r], f["dae"]["p"][r], rtol=1e-6)
    assert np.isfinite(f["q"]).all()
    xe, ok = cnn.embed(titles)
    assert np.all(xe[1] == 0) and not ok[1].any()                       # pad id -1 embeds to zeros
    np.testing.assert_allclose(f["w_t"] + f["w_p"], np.where((use[:, None] + f["dae"]["s"][:, None]) > 0, 1.0, 0.0), atol=1e-6)


def test_title_training_lowers_cost():
    TO, dae, cnn, titles, x, xv, y, yv, use = _title_setup()
    B = titles.shape[0]
    m = TO.DAETitleOracle(dae, cnn, 0.02)
    costs = [m.train_step(x, xv, y, yv, titles, np.ones(B, np.float32), B, 1.0, 1.0, 1.0) for _ in range(15)]
    assert costs[-1] < costs[0]


def test_decode_epilogue_fast_path_algebra():
    """The G1 TRAIN epilogue (csrc/gemm_sm100.cu: train_chunk_y0) does not evaluate the reference's formulas literally:
    for y = 0 cells it takes dz = rcp((1 + e) / c_neg) with e = 2^(-z log2 e) (the reciprocal IS the gradient
    0.55 p / B), 1 - p = 1 - dz / c_neg, and the loss term log(1 - p) once per EIGHT cells on their product, skipping
    the 1e-10 inside the log unless some 1 - p < 2e-3 (then the exact path redoes the chunk).  This restates that
    arithmetic in fp32 NumPy and checks it against the oracle's literal formulas: gradients within one bf16 ulp (they
    are stored as bf16), the loss of every 8-cell group within 2e-5 relative (the worst case sits next to the 2e-3
    saturation cut, where 1 - p itself carries 3e-5 of fp32 rounding in any evaluation order; the contract on the cost is
    1e-3) -- for logits over the whole range the fast path accepts."""
    rng = np.random.default_rng(0)
    B = 256
    inv_b = np.float32(1.0 / B)
    z = rng.uniform(-14.0, 6.0, size=(4096, 8)).astype(np.float32)          # 1 - sigmoid(6) = 2.5e-3 > 2e-3
    y = np.zeros_like(z)
    p = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    want_dz = O.bce_dz(p, y, inv_b)
    # fast path, fp32 throughout
    c1 = np.float32(-1.4426950408889634)
    c_neg = np.float32(0.55) * inv_b
    ic = np.float32(1.0) / c_neg
    e = np.exp2((z * c1).astype(np.float32)).astype(np.float32)
    dz = (np.float32(1.0) / (e * ic + ic).astype(np.float32)).astype(np.float32)
    omp = (np.float32(1.0) - dz * ic).astype(np.float32)
    assert omp.min() >= 2e-3                                                # the fast path's own precondition
    prod = np.prod(omp, axis=1, dtype=np.float32)                           # 8 cells -> one log2
    assert prod.min() > 1e-30                                               # no underflow: (2e-3)^8 = 2.6e-22
    loss_fast = np.float32(-0.6931471805599453 * 0.55) * np.log2(prod).astype(np.float32)
    loss_exact = -(0.55 * np.log(1.0 - p.astype(np.float64) + 1e-10)).sum(axis=1)
    np.testing.assert_allclose(loss_fast, loss_exact, rtol=2e-5, atol=1e-7)
    assert abs(loss_fast.sum(dtype=np.float64) - loss_exact.sum()) <= 2e-6 * loss_exact.sum()      # and it averages out
    ulp = np.abs(O.bf16_round(want_dz)) * 2.0 ** -7 + 1e-30
    assert (np.abs(O.bf16_round(dz) - O.bf16_round(want_dz)) <= ulp).all()
    np.testing.assert_allclose(dz, want_dz, rtol=3e-6, atol=1e-12)


@pytest.mark.parametrize("ties", [False, True])
def test_threshold_pass_topk_is_exact(ties):
    """The fused decode + top-K never loses a member of the top-k: thresholds taken from a prefix are lower bounds.
    Random logits (optionally quantised: heavy ties, the case where '>=' rather than '>' matters), random seeds, random
    prefix lengths, popularity-skewed and adversarial (best items last) orders -> identical to ranking the whole range."""
    rng = np.random.default_rng(7 + ties)
    for trial in range(60):
        T = int(rng.integers(300, 4000))
        k = int(rng.integers(5, 120))
        z = rng.normal(0, 2, T).astype(np.float32)
        if trial % 3 == 1:
            z = np.sort(z)                                        # adversarial: the best items come last
        if trial % 3 == 2:
            z = np.sort(z)[::-1].copy()                           # popularity-ranked ids: the prefix bound is tight
        if ties:
            z = np.round(z * 4) / 4
        seeds = rng.choice(T, int(rng.integers(0, 40)), replace=False).tolist()
        m1 = int(rng.integers(1, T // 2))
        m2 = int(rng.integers(m1, T))
        got, counts, _ = ranking.topk_by_threshold_passes(z, seeds, k, [m1, m2, T])
        p = (1.0 / (1.0 + np.exp(-z, dtype=np.float32))).astype(np.float32)
        want = ranking.topk_excluding_seeds(p, seeds, k)
        assert np.array_equal(got, want), (trial, T, k, m1, m2)
        assert counts[-1] >= min(T, k + len(seeds)) or counts[-1] == T
