"""NumPy restatement of the reference's title branch (test infrastructure): the character CNN of
/root/reference/models/title_models/Char_CNN.py:16-75 (configured by models/title_get.py:14-22) and
the score mixing / loss of DAE_title, /root/reference/models/DAEs.py:153-201.

Every function cites the lines it restates.  [TF1] marks TensorFlow-1.5 op semantics (third party,
not installable here -- parity unpinned upstream, see oracle/__init__.py); the closed-form backward
below is pinned against torch autograd in tests/test_oracle.py.

The DAE inside DAE_title is frozen (`tf.constant`, DAEs.py:164-171): only the title variables
(char embedding, conv filters / biases, output layer) receive gradients.

``mode="b200"`` mirrors the rounding points of the CUDA path: the dropped-out features, the output
weights and dz_t enter their tensor-core contractions as bf16 (fp32 accumulate); the CNN itself and
Adam are fp32 in both modes.
"""
from __future__ import annotations

import numpy as np

from . import dae_oracle as O
from . import philox

F32 = np.float32


def trunc_normal_xavier(rng, shape, fan_in, fan_out):
    """tf.contrib.layers.xavier_initializer(uniform=False) [TF1] = variance_scaling(factor=1, FAN_AVG,
    truncated normal): stddev = sqrt(1.3 * 2 / (fan_in + fan_out)), samples beyond 2 stddev redrawn.
    Char_CNN.py:19, :45-47, :71-73 (the biases use it too)."""
    std = np.sqrt(1.3 * 2.0 / (fan_in + fan_out))
    x = rng.normal(0.0, std, size=shape)
    bad = np.abs(x) > 2 * std
    while bad.any():
        x[bad] = rng.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(x) > 2 * std
    return x.astype(np.float32)


class CharCNNOracle:
    """Char_CNN (Char_CNN.py:5-75) with conv_layers = [[filter_num, fs, -1] for fs in filter_size] (title_get.py:14-22)."""

    def __init__(self, charsize, strmaxlen, char_emb, filter_num, filter_size, n_output, seed=0, mode="fp32"):
        assert char_emb > 0, "the one-hot variant (char_emb == 0, Char_CNN.py:26-28) is not used by any shipped config"
        self.C, self.L, self.E, self.F = charsize, strmaxlen, char_emb, filter_num
        self.fs = list(filter_size)
        self.D = filter_num * len(self.fs)
        self.N = n_output
        self.mode = mode
        rng = np.random.default_rng(seed)
        self.emb = trunc_normal_xavier(rng, (charsize, char_emb), charsize, char_emb)                # Char_CNN.py:19-20
        self.conv_W = [trunc_normal_xavier(rng, (w, char_emb, filter_num), w * char_emb, w * char_emb * filter_num)
                       for w in self.fs]                                                             # [fs, E, 1, F] squeezed, :43-46
        self.conv_b = [trunc_normal_xavier(rng, (filter_num,), filter_num, filter_num) for _ in self.fs]   # :47
        self.out_W = trunc_normal_xavier(rng, (self.D, n_output), self.D, n_output)                  # :72
        self.out_b = trunc_normal_xavier(rng, (n_output,), n_output, n_output)                       # :73

    def params(self):
        """[emb, conv_W0, conv_b0, ..., out_W, out_b]"""
        out = [self.emb]
        for W, b in zip(self.conv_W, self.conv_b):
            out += [W, b]
        return out + [self.out_W, self.out_b]

    def set_params(self, params):
        params = [np.array(p, dtype=np.float32) for p in params]
        self.emb = params[0]
        n = len(self.fs)
        self.conv_W = [params[1 + 2 * i] for i in range(n)]
        self.conv_b = [params[2 + 2 * i] for i in range(n)]
        self.out_W, self.out_b = params[1 + 2 * n], params[2 + 2 * n]

    def _rd(self, x):
        return O.bf16_round(x) if self.mode == "b200" else np.asarray(x, dtype=np.float32)

    def embed(self, titles):
        """tf.nn.embedding_lookup (Char_CNN.py:30); the pad id -1 must give a zero vector (SURVEY a9:
        TF's GPU gather returns zeros for out-of-range ids [TF1])."""
        titles = np.asarray(titles, dtype=np.int64).reshape(-1, self.L)
        ok = (titles >= 0) & (titles < self.C)
        x = self.emb[np.where(ok, titles, 0)] * ok[..., None]
        return x.astype(np.float32), ok

    def features(self, titles):
        """conv2d VALID + bias + ReLU + max over time, concatenated over the filter widths.  Char_CNN.py:37-64
        Returns feat [B, D] and the arg-max position of every feature."""
        x, _ = self.embed(titles)
        B = x.shape[0]
        feat = np.zeros((B, self.D), np.float32)
        arg = np.zeros((B, self.D), np.int64)
        for i, w in enumerate(self.fs):
            P = self.L - w + 1
            win = np.stack([x[:, p:p + w, :] for p in range(P)], 1)                   # [B, P, w, E]
            conv = np.einsum("bpke,kef->bpf", win.astype(np.float64), self.conv_W[i].astype(np.float64))
            conv = np.maximum(conv + self.conv_b[i].astype(np.float64), 0.0).astype(np.float32)     # bias_add, relu :50-52
            feat[:, i * self.F:(i + 1) * self.F] = conv.max(1)                         # reduce_max :58
            arg[:, i * self.F:(i + 1) * self.F] = conv.argmax(1)
        return feat, arg

    def forward(self, titles, kp_t=1.0, keep=None, seed=0, step=0, row_offset=0):
        """-> dict(feat, arg, keep, feat_d, feat_dq, z_t, t).  dropout :67, xw_plus_b + sigmoid :75."""
        feat, arg = self.features(titles)
        B = feat.shape[0]
        if keep is None:
            rr, kk = np.meshgrid(np.arange(B, dtype=np.uint32) + np.uint32(row_offset),
                                 np.arange(self.D, dtype=np.uint32), indexing="ij")
            keep = philox.keep_mask(seed, philox.STREAM_TITLE, step, rr, kk, kp_t)
        feat_d = (feat / F32(kp_t)).astype(np.float32) * keep.astype(np.float32)
        feat_dq = self._rd(feat_d)
        z_t = (feat_dq @ self._rd(self.out_W) + self.out_b).astype(np.float32)
        return dict(feat=feat, arg=arg, keep=keep, feat_d=feat_d, feat_dq=feat_dq, z_t=z_t, t=O.sigmoid32(z_t))

    def backward(self, titles, f, dz_t, kp_t):
        """Gradients of every title variable given d cost / d z_t [B, N] (autodiff of Char_CNN.py:30-75)."""
        x, ok = self.embed(titles)
        titles = np.asarray(titles, dtype=np.int64).reshape(-1, self.L)
        dzq = self._rd(dz_t)
        g_out_W = (f["feat_dq"].T @ dzq).astype(np.float32)
        g_out_b = dz_t.astype(np.float64).sum(0).astype(np.float32)
        dfeat_d = (dzq @ self._rd(self.out_W).T).astype(np.float32)
        d = (dfeat_d * (f["keep"].astype(np.float32) / F32(kp_t)) * (f["feat"] > 0)).astype(np.float32)
        g_emb = np.zeros_like(self.emb, dtype=np.float64)
        g_W, g_b = [], []
        B = x.shape[0]
        for i, w in enumerate(self.fs):
            di = d[:, i * self.F:(i + 1) * self.F].astype(np.float64)                  # [B, F]
            ai = f["arg"][:, i * self.F:(i + 1) * self.F]                              # [B, F]
            gW = np.zeros((w, self.E, self.F), np.float64)
            for b in range(B):
                for k in range(w):
                    pos = ai[b] + k                                                    # [F] input positions
                    gW[k] += x[b, pos, :].astype(np.float64).T * di[b][None, :]
                    contrib = self.conv_W[i][k].astype(np.float64) * di[b][None, :]    # [E, F]
                    ids = titles[b, pos]
                    valid = ok[b, pos]
                    np.add.at(g_emb, ids[valid], contrib.T[valid])
            g_W.append(gW.astype(np.float32))
            g_b.append(di.sum(0).astype(np.float32))
        grads = [g_emb.astype(np.float32)]
        for gw, gb in zip(g_W, g_b):
            grads += [gw, gb]
        return grads + [g_out_W, g_out_b], dict(dfeat_d=dfeat_d, d=d)


class DAETitleOracle:
    """DAE_title (DAEs.py:153-201) around a frozen DAEOracle and a CharCNNOracle."""

    def __init__(self, dae: O.DAEOracle, cnn: CharCNNOracle, lr):
        self.dae, self.cnn = dae, cnn
        self.adam = O.AdamTF1(lr)
        self.step = 0

    def forward(self, x_pos, x_val, titles, titles_use, B, kp=1.0, kp_in=1.0, kp_t=1.0, seed=0, step=0):
        fd = self.dae.forward(x_pos, x_val, B, kp, kp_in, seed, step)
        ft = self.cnn.forward(titles, kp_t, seed=seed, step=step)
        q, w_t, w_p = O.title_mix(fd["p"], ft["t"], fd["s"], kp_in, titles_use)        # DAEs.py:159-162, :180
        return dict(dae=fd, title=ft, q=q, w_t=w_t, w_p=w_p)

    def predict(self, x_pos, x_val, titles, titles_use, B):
        """sess.run(y_pred, keep probabilities 1)                     main_challenge.py:80-85"""
        return self.forward(x_pos, x_val, titles, titles_use, B)["q"]

    def loss_and_grads(self, x_pos, x_val, y_pos, y_val, titles, titles_use, B, kp, kp_in, kp_t, seed=0, step=0):
        f = self.forward(x_pos, x_val, titles, titles_use, B, kp, kp_in, kp_t, seed, step)
        y = O.densify_last_wins(y_pos, y_val, B, self.dae.N)
        cost = float(O.bce_rows(f["q"], y).sum() / B)                                  # DAEs.py:194-196 (no l2 term)
        dq = O.title_dq(f["q"], y, 1.0 / B)
        t = f["title"]["t"]
        dz_t = (dq * f["w_t"] * (t * (F32(1.0) - t))).astype(np.float32)
        grads, aux = self.cnn.backward(titles, f["title"], dz_t, kp_t)
        f.update(y=y, dq=dq, dz_t=dz_t, **aux)
        return cost, grads, f

    def train_step(self, x_pos, x_val, y_pos, y_val, titles, titles_use, B, kp, kp_in, kp_t, seed=0):
        cost, grads, _ = self.loss_and_grads(x_pos, x_val, y_pos, y_val, titles, titles_use, B, kp, kp_in, kp_t,
                                             seed=seed, step=self.step)
        for i, (p, g) in enumerate(zip(self.cnn.params(), grads)):
            self.adam.apply("t%d" % i, p, g)
        self.adam.finish_step()
        self.step += 1
        return cost
