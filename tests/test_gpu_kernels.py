"""GPU parity tests, kernel level, through the C ABI (libdae_b200.so) against the CPU oracle.
Integer / index work is bit-exact; floating point within the tolerance stated in each test."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import dae_oracle as O
from oracle import ranking
from tests.gpu_util import P, check, lib, stream_ptr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
dev = "cuda"


# ------------------------------------------------------------------ COO -> CSR, last wins: bit-exact
@pytest.mark.parametrize("B,N,nnz", [(8, 50, 0), (8, 50, 100), (256, 290000, 34000), (250, 6000, 60000),
                                      (4, 10, 7)])
def test_coo_to_csr_bit_exact(B, N, nnz):
    rng = np.random.default_rng(B * 1000 + nnz)
    rows = rng.integers(0, B, nnz)
    if B > 4:
        rows[rows == 3] = 2                     # row 3 is empty
    cols = np.minimum((np.exp(rng.random(nnz) * np.log(N + 1.0)) - 1).astype(np.int64), N - 1)
    pos = np.stack([rows, cols], 1).astype(np.int64)
    # two-block structure (tracks of all rows, then artists of all rows): not globally row-sorted
    order = np.concatenate([np.argsort(rows[: nnz // 2], kind="stable"),
                            nnz // 2 + np.argsort(rows[nnz // 2:], kind="stable")]) if nnz else np.zeros(0, np.int64)
    pos = pos[order]
    val = rng.integers(0, 3, nnz).astype(np.float32)
    rp, col, v = O.coo_to_csr_last_wins(pos, val, B, N)
    d_pos = torch.tensor(pos.reshape(-1), device=dev) if nnz else torch.zeros(2, dtype=torch.int64, device=dev)
    d_val = torch.tensor(val, device=dev) if nnz else torch.zeros(1, device=dev)
    d_rp = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    d_len = torch.zeros(B, dtype=torch.int32, device=dev)
    d_col = torch.zeros(max(nnz, 1), dtype=torch.int32, device=dev)
    d_v = torch.zeros(max(nnz, 1), dtype=torch.float32, device=dev)
    check(lib().dae_coo_to_csr_device(P(d_pos), P(d_val), nnz, B, N, P(d_rp), P(d_len), P(d_col), P(d_v),
                                      stream_ptr()))
    torch.cuda.synchronize()
    g_rp, g_len, g_col, g_v = d_rp.cpu().numpy(), d_len.cpu().numpy(), d_col.cpu().numpy(), d_v.cpu().numpy()
    assert np.array_equal(g_len, np.diff(rp))
    for r in range(B):
        a, n = g_rp[r], g_len[r]
        assert np.array_equal(g_col[a:a + n], col[rp[r]:rp[r + 1]])
        assert np.array_equal(g_v[a:a + n], v[rp[r]:rp[r + 1]])


def test_coo_to_csr_rejects_bad_index():
    pos = torch.tensor([0, 5, 1, 99], dtype=torch.int64, device=dev)
    val = torch.ones(2, device=dev)
    out = [torch.zeros(8, dtype=torch.int32, device=dev) for _ in range(3)] + [torch.zeros(8, device=dev)]
    rc = lib().dae_coo_to_csr_device(P(pos), P(val), 2, 2, 50, P(out[0]), P(out[1]), P(out[2]), P(out[3]),
                                     stream_ptr())
    assert rc != 0


# ------------------------------------------------------------------ TF1 Adam: bit-exact
@pytest.mark.parametrize("n,lam", [(1000003, 0.0), (4096, 0.01), (7, 0.0)])
def test_adam_bit_exact(n, lam):
    rng = np.random.default_rng(n)
    w = rng.normal(0, 0.05, n).astype(np.float32)
    ora = O.AdamTF1(0.005)
    w_o = w.copy()
    d_w = torch.tensor(w, device=dev); d_m = torch.zeros(n, device=dev); d_v = torch.zeros(n, device=dev)
    d_b = torch.zeros(n, dtype=torch.int16, device=dev)
    for step in range(3):
        g = (rng.normal(0, 1e-3, n) * (rng.random(n) < 0.7)).astype(np.float32)     # many exact zeros
        ora.apply("w", w_o, g + np.float32(lam) * w_o if lam else g)
        check(lib().dae_adam_device(P(d_w), P(d_m), P(d_v), P(torch.tensor(g, device=dev)), P(d_b), n,
                                    C.c_float(0.005), C.c_float(float(ora.b1_pow)), C.c_float(float(ora.b2_pow)),
                                    C.c_float(lam), stream_ptr()))
        ora.finish_step()
        torch.cuda.synchronize()
        assert np.array_equal(d_w.cpu().numpy(), w_o), "step %d" % step
        m_o, v_o = ora.state["w"]
        assert np.array_equal(d_m.cpu().numpy(), m_o) and np.array_equal(d_v.cpu().numpy(), v_o)
    assert np.array_equal(d_b.cpu().numpy().view(np.uint16), O.bf16_bits(w_o))


# ------------------------------------------------------------------ top-K: bit-exact index lists
def _run_topk(scores, seeds, K, T=None, idx_base=0):
    B, ld = scores.shape
    T = T or ld
    sp = np.zeros(B + 1, np.int32); sp[1:] = np.cumsum([len(s) for s in seeds])
    si = np.array([x for s in seeds for x in s] or [0], np.int32)
    d_s = torch.tensor(scores, device=dev)
    d_sp, d_si = torch.tensor(sp, device=dev), torch.tensor(si, device=dev)
    d_idx = torch.zeros(B, K, dtype=torch.int32, device=dev); d_sc = torch.zeros(B, K, device=dev)
    check(lib().dae_topk_device(P(d_s), ld, B, T, K, P(d_sp), P(d_si), idx_base, P(d_idx), P(d_sc), stream_ptr()))
    torch.cuda.synchronize()
    return d_idx.cpu().numpy(), d_sc.cpu().numpy()


def test_topk_reference_golden():
    g = np.load(os.path.join(GOLDEN, "ranking_golden.npz"))
    seeds = [[s for s in json.loads(str(x))] for x in g["seeds"]]
    idx, sc = _run_topk(np.ascontiguousarray(g["scores"]), [[s for s in sd if -2**31 < s < 2**31] for sd in seeds], 500)
    assert np.array_equal(idx, g["cands"])           # the reference's own cand_generate output


@pytest.mark.parametrize("T,K,ties", [(250000, 500, False), (250000, 500, True), (5000, 500, True), (700, 500, True),
                                      (300, 500, False), (2000000, 500, False)])
def test_topk_bit_exact(T, K, ties):
    rng = np.random.default_rng(T + ties)
    B = 6
    s = rng.random((B, T)).astype(np.float32)
    if ties:
        s = np.round(s * 50).astype(np.float32) / 50       # ~51 distinct values -> massive ties
        s[1, :] = 1.0                                      # fully saturated row (sigma(z>17) == 1.0f)
        s[2, : T // 2] = 0.0
    seeds = [list(rng.choice(T, n, replace=False)) for n in (0, 1, 5, 25, 100, 250)]
    seeds[3] = seeds[3] + seeds[3][:3] + [T + 10, -1]      # duplicate / absent seeds
    idx, sc = _run_topk(s, seeds, K)
    for r in range(B):
        want = ranking.topk_excluding_seeds(s[r], seeds[r], K)
        n = len(want)
        assert np.array_equal(idx[r, :n], want), "row %d" % r
        assert np.all(idx[r, n:] == -1)
        assert np.array_equal(sc[r, :n], s[r][want])


def test_topk_sharded_merge_equals_unsharded():
    rng = np.random.default_rng(9)
    T, K, B, G = 64000, 500, 4, 8
    s = np.round(rng.random((B, T)) * 1000).astype(np.float32)
    seeds = [list(rng.choice(T, 30, replace=False)) for _ in range(B)]
    full, _ = _run_topk(s, seeds, K)
    parts = [_run_topk(np.ascontiguousarray(s[:, g * T // G:(g + 1) * T // G]), seeds, K, idx_base=g * T // G)
             for g in range(G)]
    for r in range(B):
        m_idx, _ = ranking.merge_sharded_topk([p[0][r] for p in parts], [p[1][r] for p in parts], K)
        assert np.array_equal(m_idx, full[r])


# ------------------------------------------------------------------ tensor-core contractions vs torch (fp32 on the same bf16 inputs)
def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("N,H,B", [(1000, 64, 64), (4097, 256, 256), (290000, 256, 256), (2500, 128, 250),
                                   (3000, 256, 600)])
def test_decode_gemm(N, H, B):
    """op 0: sigmoid(W h_d^T + b).  Tolerance: 1e-3 relative on the scores (north_star)."""
    torch.manual_seed(N)
    bpad = 256 if B > 256 else (B + 63) // 64 * 64
    nbt = (B + bpad - 1) // bpad
    W = _bf16(torch.randn(N, H, device=dev) * 0.3)
    h = torch.zeros(bpad * nbt, H, device=dev); h[:B] = torch.rand(B, H, device=dev)
    h = _bf16(h)
    bias = torch.randn(N, device=dev) * 0.1
    out = torch.full((B, N), -1.0, device=dev)
    check(lib().dae_gemm_test_device(0, P(W), P(h), P(bias), P(out), N, H, B, bpad, 0, 0, None, stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.sigmoid(h[:B].float() @ W.float().T + bias)
    err = ((out - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("N,H,bpad", [(1000, 64, 64), (4097, 256, 256), (290000, 256, 256), (2500, 128, 192)])
def test_dw_gemm(N, H, bpad):
    """op 1: dW[item,:] = sum_b dzT[item,b] h_dT[:,b]."""
    torch.manual_seed(N + 1)
    dzT = _bf16(torch.randn(N, bpad, device=dev) * 1e-3)
    hT = _bf16(torch.rand(H, bpad, device=dev))
    out = torch.full((N, H), 7.0, device=dev)
    check(lib().dae_gemm_test_device(1, P(dzT), P(hT), None, P(out), N, H, bpad, bpad, 0, 0, None, stream_ptr()))
    torch.cuda.synchronize()
    ref = dzT.float() @ hT.float().T
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() < 1e-3 * scale


@pytest.mark.parametrize("N,H,bpad", [(1000, 64, 64), (4097, 256, 256), (290000, 256, 256), (2500, 128, 192),
                                      (100, 256, 128)])
def test_dh_gemm(N, H, bpad):
    """op 2: dh[b,:] = sum_item dzT[item,b] W[item,:], split-K partials, MN-major operands."""
    torch.manual_seed(N + 2)
    dzT = _bf16(torch.randn(N, bpad, device=dev) * 1e-2)
    W = _bf16(torch.randn(N, H, device=dev) * 0.3)
    ns = lib().dae_dh_nsplit(N)
    out = torch.full((ns, bpad, H), 3.0, device=dev)
    check(lib().dae_gemm_test_device(2, P(dzT), P(W), None, P(out), N, H, bpad, bpad, 0, 0, None, stream_ptr()))
    torch.cuda.synchronize()
    got = out.sum(0)
    ref = dzT.float().T @ W.float()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2e-3 * scale
