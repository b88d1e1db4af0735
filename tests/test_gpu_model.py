"""GPU parity tests, model level: one `sess.run([optimizer, cost])` / `sess.run(y_pred)` of the
reference (restated by oracle/dae_oracle.py) against the CUDA path behind the C ABI, on the same
seeded inputs, stage by stage; plus size-independent properties at BASELINE.json's full size."""
import numpy as np
import pytest
import torch

from oracle import dae_oracle as O
from oracle import ranking
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE, DAE_tied
from tests.gpu_util import Conf, model_buf, random_batch

pytestmark = pytest.mark.gpu


def _mk(tied, N, T, H, B, lr=0.005, lam=0.0, seed=11):
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=lr, reg_lambda=lam, seed=seed)
    ora = O.DAEOracle(N, H, lr, reg_lambda=lam, tied=tied, seed=5, mode="b200")
    ora.b_enc[:] = np.random.default_rng(1).normal(0, 0.1, H)
    ora.b_dec[:] = np.random.default_rng(2).normal(0, 0.1, N)
    m = (DAE_tied if tied else DAE)(conf).fit()
    m.set_params(ora.params())
    return conf, ora, m


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.mark.parametrize("tied,N,T,H,B", [(False, 1500, 1200, 64, 64), (True, 1500, 1200, 64, 64),
                                          (False, 6007, 5000, 256, 250), (True, 3001, 2500, 128, 150),
                                          (False, 20000, 17000, 256, 256)])
def test_train_step_stage_by_stage(tied, N, T, H, B):
    conf, ora, m = _mk(tied, N, T, H, B)
    rng = np.random.default_rng(N + B)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=25, empty_rows=(1,))
    xv = np.ones(len(trk), np.float32); xv[::5] = 0.0          # firstN-style zeros, incl. "last value 0 wins"
    yv = np.ones(len(y), np.float32)
    kp, kp_in = 0.8, 0.75
    m.set_debug(3)                                                         # also materialise the raw dW_dec and dW_enc
    m.stage_batch(0, trk, xv, y, yv)
    m.backward_staged(0, kp, kp_in)
    cost = m.sync_cost()
    c_ora, g, f = ora.loss_and_grads(trk, xv, y, yv, B, kp, kp_in, seed=conf.seed, step=0)
    bpad = (B + 63) // 64 * 64

    # ---- sparse side: bit-exact structure ----------------------------------------------
    rp = model_buf(m, "x_row_ptr", torch.int32).cpu().numpy()
    rl = model_buf(m, "x_row_len", torch.int32).cpu().numpy()[:B]
    col = model_buf(m, "x_col", torch.int32).cpu().numpy()
    xn = model_buf(m, "x_val", torch.float32).cpu().numpy()
    assert np.array_equal(rl, np.diff(f["row_ptr"]))
    g_col = np.concatenate([col[rp[r]:rp[r] + rl[r]] for r in range(B)])
    g_xn = np.concatenate([xn[rp[r]:rp[r] + rl[r]] for r in range(B)])
    assert np.array_equal(g_col, f["col"])
    assert np.array_equal(g_xn != 0, f["x_n"] != 0)                        # identical Philox keep mask
    np.testing.assert_allclose(g_xn, f["x_n"], rtol=2e-6)
    np.testing.assert_allclose(model_buf(m, "x_rowsum", torch.float32).cpu().numpy()[:B], f["s"], rtol=2e-6)

    # ---- encode -----------------------------------------------------------------------
    h = model_buf(m, "h", torch.float32).cpu().numpy()[:B * H].reshape(B, H)
    np.testing.assert_allclose(h, f["h"], rtol=1e-5, atol=1e-6)
    h_d = model_buf(m, "h_d", torch.bfloat16).float().cpu().numpy()[:bpad * H].reshape(bpad, H)
    assert np.array_equal(h_d[:B] != 0, f["keep_h"])                       # identical hidden dropout mask
    assert np.all(h_d[B:] == 0)
    np.testing.assert_allclose(h_d[:B], f["h_dq"], rtol=8e-3)              # <= 1 bf16 ulp

    # ---- decode + loss + dz -----------------------------------------------------------
    assert abs(cost - c_ora) <= 1e-3 * abs(c_ora), (cost, c_ora)           # north_star: 1e-3 relative
    dzT = model_buf(m, "dzT", torch.bfloat16).float().cpu().numpy()[:N * bpad].reshape(N, bpad)
    assert np.all(dzT[:, B:] == 0)
    dz_o = f["dzq"].T
    assert np.array_equal(np.sign(dzT[:, :B]), np.sign(dz_o))              # y pattern (negative exactly where y == 1)
    err = _rel(dzT[:, :B], dz_o, 1e-3 * np.abs(dz_o).max())
    assert err.max() < 2e-2 and err.mean() < 2e-3                          # bf16 storage of dz: 2^-8 relative
    np.testing.assert_allclose(model_buf(m, "g_b_dec", torch.float32).cpu().numpy(), g["b_dec"], rtol=5e-3,
                               atol=1e-3 * np.abs(g["b_dec"]).max())

    # ---- contractions: against torch on the device's own operands (tight), and the oracle (bf16-loose)
    dz_dev = model_buf(m, "dzT", torch.bfloat16)[:N * bpad].view(N, bpad).float()
    hd_dev = model_buf(m, "h_d", torch.bfloat16)[:bpad * H].view(bpad, H).float()
    W16 = model_buf(m, "W_dec_bf16", torch.bfloat16)[:N * H].view(N, H).float()
    dh_ref = dz_dev.T @ W16
    dh = model_buf(m, "dh_sum", torch.float32)[:bpad * H].view(bpad, H)        # fixed-order sum of the split-K partials
    assert (dh - dh_ref).abs().max().item() < 2e-3 * dh_ref.abs().max().item()
    np.testing.assert_allclose(dh[:B].cpu().numpy(), f["dh_d"], rtol=0, atol=3e-2 * np.abs(f["dh_d"]).max())

    da = model_buf(m, "da", torch.float32).cpu().numpy()[:B * H].reshape(B, H)
    np.testing.assert_allclose(da, f["da"], rtol=0, atol=3e-2 * np.abs(f["da"]).max())
    np.testing.assert_allclose(model_buf(m, "g_b_enc", torch.float32).cpu().numpy(), g["b_enc"], rtol=0,
                               atol=3e-2 * np.abs(g["b_enc"]).max())

    Np = (N + 127) // 128 * 128                                            # catalogue rows are held in whole 128-item tiles
    g_dec = model_buf(m, "g_dec", torch.float32).view(Np, H)[:N]
    gw_ref = dz_dev @ hd_dev                                               # dW_dec from the device operands
    assert (g_dec - gw_ref).abs().max().item() < 1e-3 * gw_ref.abs().max().item()
    np.testing.assert_allclose(g_dec.cpu().numpy(), g["W_dec"], rtol=0, atol=2e-2 * np.abs(g["W_dec"]).max())
    g_enc = model_buf(m, "g_enc", torch.float32).view(Np, H)[:N].cpu().numpy()
    # "touched" = every catalogue row the batch lists in x (known before the step: the encoder's Adam of all OTHER rows
    # runs in the background from the start of the step); the gradient is non-zero only where x_n != 0
    touched = model_buf(m, "touched", torch.uint8).cpu().numpy().astype(bool)[:N]
    rows_o = np.zeros(N, bool); rows_o[f["col"]] = True
    assert np.array_equal(touched, rows_o)                                 # exactly the rows present in x
    live_o = np.zeros(N, bool); live_o[f["col"][f["x_n"] != 0]] = True
    assert np.array_equal(np.abs(g_enc).sum(1) != 0, live_o & (np.abs(g["W_enc"]).sum(1) != 0))
    assert np.all(g_enc[~touched] == 0)
    np.testing.assert_allclose(g_enc, g["W_enc"], rtol=0, atol=3e-2 * np.abs(g["W_enc"]).max())

    # ---- Adam ------------------------------------------------------------------------------
    ora.apply_grads(g)
    m.apply_adam()
    m.sync_cost()
    got = m.get_params()
    for a, b, name in zip(got, ora.params(), ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        # the first Adam step is lr*sign(g): elements whose tiny gradient flips sign under bf16 noise move 2*lr
        assert (d > 1e-3 * conf.lr).mean() < 5e-3 and d.max() <= 2.001 * conf.lr, name
    assert model_buf(m, "touched", torch.uint8).sum().item() == 0          # gradient buffers are clean again
    assert model_buf(m, "g_enc", torch.float32).abs().sum().item() == 0
    # re-staging the slot clears the previous batch's target bits before setting the new ones
    trk2, art2, y2 = random_batch(np.random.default_rng(99), B, T, N - T, mean_len=10)
    m.stage_batch(0, trk2, np.ones(len(trk2), np.float32), y2, np.ones(len(y2), np.float32))
    m.backward_staged(0, kp, kp_in)
    m.sync_cost()
    bits = model_buf(m, "ybits", torch.int32).cpu().numpy().view(np.uint32)
    assert int(np.unpackbits(bits.view(np.uint8)).sum()) == len(np.unique(y2, axis=0))
    shadow = model_buf(m, "W_dec_bf16", torch.bfloat16)[:N * H].float().cpu().numpy().reshape(N, H)
    assert np.array_equal(shadow, O.bf16_round(got[1]))                    # shadow == bf16(master), bit-exact
    m.close()


@pytest.mark.parametrize("tied", [False, True])
def test_training_trajectory_matches_oracle(tied):
    """20 steps on the same batch stream: cost trajectories within 1e-3 relative early, 1e-2 late."""
    N, T, H, B = 3000, 2500, 64, 128
    conf, ora, m = _mk(tied, N, T, H, B, lr=0.01)
    rng = np.random.default_rng(3)
    for step in range(20):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=20)
        x, xv = (trk, np.ones(len(trk), np.float32)) if step % 2 == 0 else (art, np.ones(len(art), np.float32))
        yv = np.ones(len(y), np.float32)
        c_gpu = m.train_step(x, xv, y, yv, 0.8, 0.75)
        c_ora = ora.train_step(x, xv, y, yv, B, 0.8, 0.75, seed=conf.seed)
        tol = 1e-3 if step < 5 else 1e-2
        assert abs(c_gpu - c_ora) <= tol * abs(c_ora), (step, c_gpu, c_ora)
    m.close()


def test_pipelined_step_equals_synchronous_step():
    """train_step_async (the runner's call) returns the same costs as train_step, one call late, and leaves the
    same parameters."""
    N, T, H, B = 3000, 2500, 64, 128
    rng = np.random.default_rng(3)
    batches = []
    for step in range(6):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=20)
        batches.append((trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32)))
    conf, ora, m1 = _mk(False, N, T, H, B, lr=0.01)
    sync_costs = [m1.train_step(*b, 0.8, 0.75) for b in batches]
    p1 = m1.get_params(); m1.close()
    conf, ora, m2 = _mk(False, N, T, H, B, lr=0.01)
    got = [m2.train_step_async(*b, 0.8, 0.75) for b in batches]
    assert got[0] is None
    got = got[1:] + [m2.flush()]
    assert m2.flush() is None
    p2 = m2.get_params(); m2.close()
    np.testing.assert_allclose(got, sync_costs, rtol=1e-6)
    for a, b in zip(p1, p2):
        assert np.array_equal(a, b)                             # the single-GPU step is deterministic (gather-form dW_enc)


@pytest.mark.parametrize("tied,N,T,H,B,lam", [(False, 1500, 1200, 64, 64, 0.0), (True, 3001, 2500, 128, 150, 0.0),
                                              (False, 6007, 5000, 256, 250, 1e-3), (True, 6007, 5000, 256, 256, 0.0),
                                              (False, 40000, 33000, 256, 256, 0.0)])
def test_fused_dw_adam_equals_two_kernel_path(tied, N, T, H, B, lam):
    """The default step applies Adam to the dW tile in tensor memory (k_dw_adam_fused: bulk-copy staging of w/m/v);
    debug bit 2 forms dW in HBM and runs k_adam_rows_vec4.  Same MMA order, same rounded Adam ops -> the decoder
    master, both moments and the bf16 operand copy must agree BIT FOR BIT after several steps."""
    rng = np.random.default_rng(N)
    batches = []
    for step in range(3):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=20, empty_rows=(2,))
        x = trk if step % 2 == 0 else art
        batches.append((x, np.ones(len(x), np.float32), y, np.ones(len(y), np.float32)))
    names = (("W_dec", torch.float32), ("mW_dec", torch.float32), ("vW_dec", torch.float32), ("W_dec_bf16", torch.int16))

    def snap(m):
        return {k: model_buf(m, k, torch.bfloat16 if k == "W_dec_bf16" else dt).view(dt).clone() for k, dt in names}
    out = []
    for flags in (0, 4):
        conf, ora, m = _mk(tied, N, T, H, B, lr=0.01, lam=lam)
        m.set_debug(flags)
        costs = [m.train_step(*batches[0], 0.8, 0.75)]
        first = snap(m)
        costs += [m.train_step(*b, 0.8, 0.75) for b in batches[1:]]
        out.append((costs, first, snap(m)))
        m.close()
    assert out[0][0] == out[1][0]
    for k, _ in names:
        # no atomics anywhere in the single-GPU step (the sparse-row dW_enc is gathered in a fixed order): every step
        # of both paths agrees bit for bit, tied or not
        assert torch.equal(out[0][1][k], out[1][1][k]), k
        assert torch.equal(out[0][2][k], out[1][2][k]), k


def _snap_state(m):
    names = ("W_enc", "mW_enc", "vW_enc", "W_dec", "mW_dec", "vW_dec", "b_enc", "b_dec")
    return {k: model_buf(m, k, torch.float32).clone() for k in names}


@pytest.mark.parametrize("N,T,H,B,lam", [(6007, 5000, 256, 250, 0.0), (40000, 33000, 128, 256, 1e-3),
                                         (290000, 250000, 256, 256, 0.0)])
def test_background_encoder_adam_is_bit_identical(N, T, H, B, lam):
    """With debug bit 10 the whole-step call streams the encoder's Adam of the rows no playlist lists (g == 0) in the
    background, under the encode / decode / dh kernels (k_adam_bg: chunks claimed from the top of the row range until the
    decoder update starts), and the dense pass does the rest.  Every element must receive exactly ONE dense TF1 Adam
    update per step whatever the split point: parameters and both moments equal the default run (streamer off), bit for
    bit, after several steps (DAEs.py:102: dense Adam on every row, every step)."""
    rng = np.random.default_rng(N + 1)
    batches = []
    for step in range(4):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=40, empty_rows=(3,))
        x = trk if step % 2 == 0 else art
        batches.append((x, np.ones(len(x), np.float32), y, np.ones(len(y), np.float32)))
    out = []
    claimed = []
    for flags in (1024, 0):
        conf, ora, m = _mk(False, N, T, H, B, lr=0.01, lam=lam)
        m.set_debug(flags)
        costs = [m.train_step(*b, 0.8, 0.75) for b in batches]
        claimed.append(int(model_buf(m, "bg_ctl", torch.int32)[0].item()) & 0x3fffffff)
        out.append((costs, _snap_state(m)))
        m.close()
    assert out[0][0] == out[1][0]
    for k in out[0][1]:
        assert torch.equal(out[0][1][k], out[1][1][k]), k
    assert claimed[1] == 0
    if N >= 100000 and lam == 0.0:
        assert claimed[0] > 0, "the background streamer never claimed a chunk at full size"


def test_step_is_deterministic_and_matches_atomic_scatter():
    """SURVEY section 5 (determinism test for the scatter-add): two runs of the same 3 steps are bit-identical with the
    default gather-form dW_enc; the fp32 red.add form (debug bit 11, the multi-GPU default) agrees to summation order."""
    N, T, H, B = 20000, 17000, 256, 256
    rng = np.random.default_rng(5)
    batches = []
    for step in range(3):
        trk, art, y = random_batch(rng, B, T, N - T, mean_len=60)
        batches.append((trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32)))
    runs = []
    for flags in (0, 0, 2048):
        conf, ora, m = _mk(False, N, T, H, B, lr=0.01)
        m.set_debug(flags)
        costs = [m.train_step(*b, 0.8, 0.75) for b in batches]
        runs.append((costs, _snap_state(m)))
        m.close()
    assert runs[0][0] == runs[1][0]
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
    np.testing.assert_allclose(runs[0][0], runs[2][0], rtol=1e-5)
    for k in ("W_enc", "W_dec"):
        d = (runs[0][1][k] - runs[2][1][k]).abs()
        assert (d > 1e-6).float().mean().item() < 2e-3, k


def test_row_offset_keys_forward_and_backward_masks_alike():
    """One rank emulating rows [r0, r0 + B) of a larger global batch (dae_model_backward_staged global_batch / row_offset):
    the hidden-dropout mask of the backward must be the one the forward used (keyed by the GLOBAL row)."""
    N, T, H, B = 3000, 2500, 64, 64
    conf, ora, m = _mk(False, N, T, H, B)
    rng = np.random.default_rng(12)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=20)
    xv = np.ones(len(trk), np.float32); yv = np.ones(len(y), np.float32)
    kp, kp_in, G, r0 = 0.6, 0.75, 256, 128
    m.set_debug(3)
    m.stage_batch(0, trk, xv, y, yv)
    m.backward_staged(0, kp, kp_in, global_batch=G, row_offset=r0)
    cost = m.sync_cost()
    c_ora, g, f = ora.loss_and_grads(trk, xv, y, yv, B, kp, kp_in, seed=conf.seed, step=0, row_offset=r0, global_batch=G)
    assert abs(cost - c_ora) <= 1e-3 * abs(c_ora)
    h_d = model_buf(m, "h_d", torch.bfloat16).float().cpu().numpy()[:B * H].reshape(B, H)
    assert np.array_equal(h_d != 0, f["keep_h"])
    da = model_buf(m, "da", torch.float32).cpu().numpy()[:B * H].reshape(B, H)
    assert np.array_equal(da != 0, (f["da"] != 0))                        # same mask in the backward
    np.testing.assert_allclose(da, f["da"], rtol=0, atol=3e-2 * np.abs(f["da"]).max())
    np.testing.assert_allclose(model_buf(m, "g_b_enc", torch.float32).cpu().numpy(), g["b_enc"], rtol=0,
                               atol=3e-2 * np.abs(g["b_enc"]).max())
    m.close()


def test_cfg2_size_step_against_oracle():
    """BASELINE.json cfg2 (B=256, N=290 000, H=256): cost, dz and the parameters after one whole (default-path) train step
    against DAEOracle -- the dense NumPy restatement of models/DAEs.py at the headline size."""
    N, T, H, B = 290000, 250000, 256, 256
    conf, ora, m = _mk(False, N, T, H, B, lr=0.005)
    rng = np.random.default_rng(21)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=66)
    xv = np.ones(len(trk), np.float32); yv = np.ones(len(y), np.float32)
    kp, kp_in = 0.8, 0.75
    cost = m.train_step(trk, xv, y, yv, kp, kp_in)
    c_ora, g, f = ora.loss_and_grads(trk, xv, y, yv, B, kp, kp_in, seed=conf.seed, step=0)
    assert abs(cost - c_ora) <= 1e-3 * abs(c_ora), (cost, c_ora)
    dzT = model_buf(m, "dzT", torch.bfloat16)[:N * B].view(N, B).float().cpu().numpy()
    dz_o = f["dzq"].T
    assert np.array_equal(np.sign(dzT), np.sign(dz_o))
    err = _rel(dzT, dz_o, 1e-3 * np.abs(dz_o).max())
    assert err.max() < 2e-2 and err.mean() < 2e-3
    del dzT, dz_o, err
    ora.apply_grads(g)
    got = m.get_params()
    for a, b, name in zip(got, ora.params(), ("W_enc", "W_dec", "b_enc", "b_dec")):
        d = np.abs(a - b)
        assert (d > 1e-3 * conf.lr).mean() < 5e-3 and d.max() <= 2.001 * conf.lr, (name, (d > 1e-3 * conf.lr).mean(), d.max())
    m.close()


def test_reg_lambda_cost_and_update():
    N, T, H, B = 1500, 1200, 64, 64
    conf, ora, m = _mk(False, N, T, H, B, lam=1e-3)
    rng = np.random.default_rng(8)
    trk, art, y = random_batch(rng, B, T, N - T)
    c_gpu = m.train_step(trk, np.ones(len(trk)), y, np.ones(len(y)), 0.8, 0.75)
    c_ora = ora.train_step(trk, np.ones(len(trk)), y, np.ones(len(y)), B, 0.8, 0.75, seed=conf.seed)
    assert abs(c_gpu - c_ora) <= 1e-3 * abs(c_ora)
    m.close()


@pytest.mark.parametrize("N,T,H,B", [(1500, 1200, 64, 64), (6007, 5000, 256, 250), (9000, 8000, 256, 700)])
def test_predict_and_recommend(N, T, H, B):
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.01, DAEval=None)
    ora = O.DAEOracle(N, H, 0.01, tied=False, seed=5, mode="b200")
    ora.b_dec[:] = np.random.default_rng(2).normal(0, 0.5, N)
    m = DAE(conf)
    m.trainable = False
    m.fit()
    m.set_params(ora.params())
    rng = np.random.default_rng(N)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=30, empty_rows=(2,))
    xv = np.concatenate([np.ones(len(trk)), 0.5 * np.ones(len(art))]).astype(np.float32)
    p_gpu = m.predict(y, xv)
    p_ora = ora.predict(y, xv, B)
    assert _rel(p_gpu, p_ora, 1e-6).max() < 1e-3                           # north_star: 1e-3 relative on scores
    pt = m.predict(y, xv, tracks_only=True)
    assert np.array_equal(pt, p_gpu[:, :T])
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    idx, sc = m.recommend(y, xv, seeds, k=500, return_scores=True)
    for r in range(0, B, max(B // 16, 1)):
        want = ranking.topk_excluding_seeds(p_gpu[r, :T], seeds[r], 500)   # exact on the device's own scores
        assert np.array_equal(idx[r], want), r
        want_o = ranking.topk_excluding_seeds(p_ora[r, :T], seeds[r], 500)
        # vs the oracle's scores: identical sets except items within 1e-3 of the K-th score
        diff = set(idx[r].tolist()) ^ set(want_o.tolist())
        kth = p_ora[r, want_o[-1]]
        assert all(abs(p_ora[r, i] - kth) <= 1e-3 * kth for i in diff), (r, len(diff))
    m.close()


@pytest.mark.parametrize("N,T,H,B,k", [(40000, 36000, 128, 300, 500), (9000, 8000, 64, 64, 100), (300000, 270000, 256, 700, 500)])
def test_fused_decode_topk_equals_dense_ranking(N, T, H, B, k):
    """Large catalogues rank through the fused decode + top-K (threshold-filtered candidate lists, no [B, T] scores;
    debug bit 4 forces it here).  It must return exactly what the dense path (scores + exact radix select, bit 5)
    returns -- which test_predict_and_recommend pins against the oracle's ranking."""
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.01, DAEval=None)
    ora = O.DAEOracle(N, H, 0.01, tied=False, seed=5, mode="b200")
    ora.b_dec[:] = np.random.default_rng(2).normal(0, 0.5, N)
    m = DAE(conf)
    m.trainable = False
    m.fit()
    m.set_params(ora.params())
    rng = np.random.default_rng(N)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=30, empty_rows=(2,))
    xv = np.ones(len(trk), np.float32)
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    m.set_debug(32)
    idx_d, sc_d = m.recommend(trk, xv, seeds, k=k, return_scores=True)
    m.set_debug(16)
    idx_f, sc_f = m.recommend(trk, xv, seeds, k=k, return_scores=True)
    # the final ranking of the fused path is on p = sigmoid(z) (ties by id), like the dense path and the reference
    assert np.array_equal(sc_f, sc_d)
    assert np.array_equal(idx_f, idx_d)
    for r in range(B):
        assert not (set(idx_f[r].tolist()) & set(seeds[r]))
    # bit 15: 128-row batch tiles decoded in multicast pairs (clusters of two batch tiles share every W chunk) instead of
    # 256-row tiles: the same scores, bit for bit, on both paths
    for flags in (32 | 32768, 16 | 32768):
        m.set_debug(flags)
        idx_u, sc_u = m.recommend(trk, xv, seeds, k=k, return_scores=True)
        assert np.array_equal(sc_u, sc_d) and np.array_equal(idx_u, idx_d), flags
    m.set_debug(16)
    # item-sharded (dp.ShardedRecommender's data path): per-range lists merged on the device by (score desc, id asc)
    # == the host statement of the rule == the whole-catalogue list
    import ctypes as C
    from spotify_recsys_challenge_2018_b200 import _lib
    from spotify_recsys_challenge_2018_b200.dp import item_shard, merge_topk_lists
    world = 3
    ranges = [item_shard(T, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == T and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    parts = [m.recommend(trk, xv, seeds, k=k, return_scores=True, item_range=rg) for rg in ranges]
    for (lo, hi), (pi, ps) in zip(ranges, parts):
        assert ((pi == -1) | ((pi >= lo) & (pi < hi))).all()
    want_i, want_s = merge_topk_lists([p[0] for p in parts], [p[1] for p in parts], k)
    cat_i = torch.from_numpy(np.concatenate([p[0] for p in parts], 1)).cuda().contiguous()
    cat_s = torch.from_numpy(np.concatenate([p[1] for p in parts], 1)).cuda().contiguous()
    out_i = torch.empty((B, k), dtype=torch.int32, device="cuda"); out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().dae_topk_merge_device(C.c_void_p(cat_s.data_ptr()), C.c_void_p(cat_i.data_ptr()), world * k, B, k,
                                                 C.c_void_p(out_i.data_ptr()), C.c_void_p(out_s.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(out_i.cpu().numpy(), want_i)
    assert np.array_equal(out_s.cpu().numpy(), want_s)
    # merged lists == the unsharded call
    assert np.array_equal(want_s, sc_f) and np.array_equal(want_i, idx_f)
    m.close()


@pytest.mark.parametrize("bias_hi", [9.0, 13.0, 30.0])
def test_fused_decode_topk_with_saturated_scores(bias_hi):
    """Adversarial for the fused decode + top-K: the filter compares logits, the final order is on fp32
    sigmoid(z) with ties to the lower id.  A block of high-id items gets a large bias so that the top of every ranking
    sits where distinct logits collapse onto one p (bias 13: 1 - p ~ 2e-6; bias 30: p == 1.0f exactly for thousands of
    items, more than a candidate list holds -> the dense fallback).  The fused path must still equal the dense one."""
    N, T, H, B, k = 200000, 180000, 64, 128, 500
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.01, DAEval=None)
    ora = O.DAEOracle(N, H, 0.01, tied=False, seed=5, mode="b200")
    rng = np.random.default_rng(int(bias_hi))
    ora.b_dec[:] = rng.normal(0, 0.5, N)
    hot = rng.permutation(T)[:20000 if bias_hi >= 30 else 3000]
    ora.b_dec[hot] = bias_hi + rng.normal(0, 0.3, len(hot))
    m = DAE(conf)
    m.trainable = False
    m.fit()
    m.set_params(ora.params())
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=30, empty_rows=(2,))
    xv = np.ones(len(trk), np.float32)
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    m.set_debug(32)
    idx_d, sc_d = m.recommend(trk, xv, seeds, k=k, return_scores=True)
    m.set_debug(16)
    idx_f, sc_f = m.recommend(trk, xv, seeds, k=k, return_scores=True)
    assert np.array_equal(sc_f, sc_d) and np.array_equal(idx_f, idx_d)
    ties = (np.diff(sc_d, axis=1) == 0).mean()
    assert ties > (0.5 if bias_hi >= 30 else 0.0)             # the case really exercises ties in p
    m.close()


def test_errors_are_loud():
    from spotify_recsys_challenge_2018_b200._lib import DaeError
    conf = Conf(batch=8, n_input=100, n_tracks=80, hidden=64, lr=0.01)
    m = DAE(conf).fit()
    with pytest.raises(DaeError):
        m.train_step(np.array([[0, 100]]), [1.0], np.array([[0, 1]]), [1.0], 0.8, 0.8)     # item id out of range
    with pytest.raises(DaeError):
        m.train_step(np.array([[9, 1]]), [1.0], np.array([[0, 1]]), [1.0], 0.8, 0.8)       # row outside the batch
    with pytest.raises(DaeError):
        m.train_step(np.array([[0, 1]]), [1.0], np.array([[0, 1]]), [0.5], 0.8, 0.8)       # non-binary target
    c = m.train_step(np.zeros((0, 2)), [], np.array([[0, 1]]), [1.0], 0.8, 0.8)            # empty input is legal
    assert np.isfinite(c)
    m.close()
    with pytest.raises(DaeError):
        DAE(Conf(batch=8, n_input=100, n_tracks=80, hidden=48, lr=0.01)).fit()             # unsupported width


def test_full_size_properties():
    """BASELINE cfg2 size (B=256, N=290000, H=256): size-independent properties of one train step."""
    N, T, H, B = 290000, 250000, 256, 256
    conf = Conf(batch=B, n_input=N, n_tracks=T, hidden=H, lr=0.005, seed=1)
    m = DAE(conf).fit()
    m.set_debug(1)
    rng = np.random.default_rng(0)
    trk, art, y = random_batch(rng, B, T, N - T, mean_len=66)
    m.stage_batch(0, trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32))
    m.backward_staged(0, 0.8, 0.75)
    cost = m.sync_cost()
    # at Xavier init p ~= 0.5 everywhere: cost ~= N*0.55*ln2 + nnz_y/B*(1-0.55)*ln2
    nnz_y = len(np.unique(y, axis=0))
    expect = N * 0.55 * np.log(2) + nnz_y / B * 0.45 * np.log(2)
    assert abs(cost - expect) < 0.02 * expect, (cost, expect)
    dz = model_buf(m, "dzT", torch.bfloat16)[:N * 256].view(N, 256).float()
    hd = model_buf(m, "h_d", torch.bfloat16).view(256, H).float()
    W16 = model_buf(m, "W_dec_bf16", torch.bfloat16)[:N * H].view(N, H).float()
    # db_dec == row sums of dz (computed from the unrounded values): loose bf16 tolerance
    db = model_buf(m, "g_b_dec", torch.float32)[:N]
    assert ((db - dz.sum(1)).abs().max() / db.abs().max()).item() < 1e-2
    # dW_dec, dh: exact contractions of the stored operands
    g_dec = model_buf(m, "g_dec", torch.float32).view(-1, H)[:N]
    ref = dz @ hd
    assert ((g_dec - ref).abs().max() / ref.abs().max()).item() < 1e-3
    dh = model_buf(m, "dh_sum", torch.float32)[:256 * H].view(256, H)
    ref = dz.T @ W16
    assert ((dh - ref).abs().max() / ref.abs().max()).item() < 2e-3
    # y bitmask consistency: dz is negative exactly on the (row, item) pairs of y
    neg = (dz < 0).sum().item()
    assert neg == nnz_y, (neg, nnz_y)
    m.set_debug(0)
    m.apply_adam()
    c2 = m.train_step(trk, np.ones(len(trk), np.float32), y, np.ones(len(y), np.float32), 0.8, 0.75)
    assert np.isfinite(c2) and c2 < cost                                    # one Adam step lowers the loss on the same batch
    # ranking at full width: sorted, seeds excluded, idempotent
    seeds = [trk[trk[:, 0] == r, 1].tolist() for r in range(B)]
    idx, sc = m.recommend(trk, np.ones(len(trk), np.float32), seeds, k=500, return_scores=True)
    assert np.all(np.diff(sc, axis=1) <= 0)
    for r in (0, 100, 255):
        assert not (set(idx[r].tolist()) & set(seeds[r])) and len(set(idx[r].tolist())) == 500
    idx2 = m.recommend(trk, np.ones(len(trk), np.float32), seeds, k=500)
    assert np.array_equal(idx, idx2)
    m.close()
