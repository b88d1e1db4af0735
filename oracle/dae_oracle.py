"""NumPy restatement of the reference's DAE train / predict step (test infrastructure).

Follows /root/reference/models/DAEs.py line by line; every function cites the
lines it restates.  TensorFlow-1 op semantics (marked [TF1]) are third-party
(TF 1.5.0, readme.md:33, not vendored, not installable here) and are restated
from their published definitions -- see oracle/__init__.py for what is and is
not pinned.

Two numeric modes:

* ``mode="fp32"``  -- reference-faithful: everything fp32, as TF computes it.
* ``mode="b200"``  -- the rounding points of the B200 kernels are mirrored:
  the decoder weight operand is the bf16 shadow of the fp32 master, ``h_d``
  and ``dz`` are rounded to bf16 before they enter a tensor-core contraction,
  accumulation stays fp32.  Everything else (encode gather, loss, Adam) is
  fp32 in both modes.  GPU parity tests compare against this mode; the
  end-to-end r-precision test compares against ``fp32``.
"""
from __future__ import annotations

import numpy as np

from . import philox

F32 = np.float32
EPS_LOG = F32(1e-10)      # DAEs.py:42, :98-99, :160
NEG_WEIGHT = F32(0.55)    # DAEs.py:99
ADAM_B1 = F32(0.9)        # [TF1] tf.train.AdamOptimizer defaults (DAEs.py:102 passes lr only)
ADAM_B2 = F32(0.999)
ADAM_EPS = F32(1e-8)


# ----------------------------------------------------------------------------
# bf16 helpers
# ----------------------------------------------------------------------------
def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what __float2bfloat16_rn does)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    out = rounded.astype(np.uint32).view(np.float32).reshape(x.shape)
    return np.where(np.isnan(x), x, out)


def bf16_bits(x):
    """fp32 -> uint16 bf16 bit pattern (RNE)."""
    return (bf16_round(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


# ----------------------------------------------------------------------------
# a2: sparse -> dense, duplicates: last occurrence wins           DAEs.py:32-38
# ----------------------------------------------------------------------------
def densify_last_wins(positions, vals, n_rows, n_cols):
    """tf.sparse_tensor_to_dense(validate_indices=False) [TF1]: out[idx] = val in
    input order, default 0 -> duplicates do not accumulate, the last one wins."""
    out = np.zeros((n_rows, n_cols), dtype=np.float32)
    positions = np.asarray(positions).reshape(-1, 2).astype(np.int64)
    vals = np.asarray(vals, dtype=np.float32).reshape(-1)
    for (r, c), v in zip(positions, vals):       # in order: later entries overwrite
        out[r, c] = v
    return out


def coo_to_csr_last_wins(positions, vals, n_rows, n_cols=None):
    """Same semantics, kept sparse: per row the unique columns in ascending order,
    each with the value of its LAST occurrence in the COO list.  Returns
    (row_ptr int32[n_rows+1], col int32[nnz_u], val f32[nnz_u]).  Entries whose
    last value is 0 stay in the structure (they contribute nothing)."""
    positions = np.asarray(positions).reshape(-1, 2).astype(np.int64)
    vals = np.asarray(vals, dtype=np.float32).reshape(-1)
    assert positions.shape[0] == vals.shape[0]
    if positions.shape[0]:
        assert positions[:, 0].min() >= 0 and positions[:, 0].max() < n_rows
        if n_cols is not None:
            assert positions[:, 1].min() >= 0 and positions[:, 1].max() < n_cols
    e = np.arange(positions.shape[0], dtype=np.int64)
    order = np.lexsort((e, positions[:, 1], positions[:, 0]))   # row, col, then entry index
    r = positions[order, 0]; c = positions[order, 1]
    last = np.ones(order.shape[0], dtype=bool)
    if order.shape[0] > 1:
        last[:-1] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    r = r[last]; c = c[last]; v = vals[order][last]
    row_ptr = np.zeros(n_rows + 1, dtype=np.int32)
    np.add.at(row_ptr, r + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return row_ptr, c.astype(np.int32), v.astype(np.float32)


def csr_rows(row_ptr):
    return np.repeat(np.arange(row_ptr.shape[0] - 1, dtype=np.int32), np.diff(row_ptr))


# ----------------------------------------------------------------------------
# elementary functions in fp32
# ----------------------------------------------------------------------------
def sigmoid32(a):
    """tf.nn.sigmoid in fp32 [TF1]: 1/(1+exp(-a)).  exp overflow -> inf -> 0 is the
    intended saturation; sigma(17) == 1.0f."""
    a = np.asarray(a, dtype=np.float32)
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-a, dtype=np.float32))).astype(np.float32)


# ----------------------------------------------------------------------------
# a3 + a4: input dropout, row-normalise, encode                  DAEs.py:40-42, 64-70
# ----------------------------------------------------------------------------
def normalise_input(row_ptr, val, keep_in, kp_in):
    """x_d = x/kp_in * keep ; s = sum_j x_d ; x_n = x_d/(s+1e-10).   DAEs.py:40-42
    Returns (x_n per nnz, s per row)."""
    rows = csr_rows(row_ptr)
    x_d = (val / F32(kp_in)).astype(np.float32) * keep_in.astype(np.float32)
    s = np.zeros(row_ptr.shape[0] - 1, dtype=np.float64)
    np.add.at(s, rows, x_d.astype(np.float64))
    s = s.astype(np.float32)
    x_n = (x_d / (s[rows] + EPS_LOG)).astype(np.float32)
    return x_n, s


def encode(W_enc, b_enc, row_ptr, col, x_n, keep_h, kp):
    """a = x_n . W_enc + b_enc ; h = sigmoid(a) ; h_d = h/kp * keep_h.  DAEs.py:66-68"""
    B = row_ptr.shape[0] - 1
    rows = csr_rows(row_ptr)
    a = np.zeros((B, W_enc.shape[1]), dtype=np.float64)
    np.add.at(a, rows, x_n[:, None].astype(np.float64) * W_enc[col].astype(np.float64))
    a = (a + b_enc.astype(np.float64)).astype(np.float32)
    h = sigmoid32(a)
    h_d = (h / F32(kp)).astype(np.float32) * keep_h.astype(np.float32)
    return a, h, h_d


# ----------------------------------------------------------------------------
# a5 + a6: decode and weighted BCE                                DAEs.py:73-77, 98-100
# ----------------------------------------------------------------------------
def decode(h_d, W_dec, b_dec):
    """z = h_d . W_dec^T + b_dec ; p = sigmoid(z).                  DAEs.py:75-76 / 143-144"""
    z = (h_d.astype(np.float32) @ W_dec.astype(np.float32).T + b_dec.astype(np.float32)).astype(np.float32)
    return z, sigmoid32(z)


def bce_rows(p, y):
    """L_i = -sum_j [ y log(p+eps) + 0.55 (1-y) log(1-p+eps) ]       DAEs.py:98-99"""
    p = p.astype(np.float32); y = y.astype(np.float32)
    one = F32(1.0)
    t = y * np.log(p + EPS_LOG, dtype=np.float32) + NEG_WEIGHT * (one - y) * np.log(one - p + EPS_LOG, dtype=np.float32)
    return -t.astype(np.float64).sum(axis=1)


def bce_dz(p, y, inv_batch):
    """d cost / d z for cost = mean_i L_i  (closed form of TF autodiff):
    dz = [ -y p(1-p)/(p+eps) + 0.55 (1-y) p(1-p)/(1-p+eps) ] / B"""
    p = p.astype(np.float32); y = y.astype(np.float32)
    one = F32(1.0)
    omp = one - p
    pq = p * omp
    pos = pq / (p + EPS_LOG)
    neg = NEG_WEIGHT * (pq / (omp + EPS_LOG))
    return ((-(y * pos) + (one - y) * neg) * F32(inv_batch)).astype(np.float32)


# ----------------------------------------------------------------------------
# a8: TF1 Adam                                                     DAEs.py:102
# ----------------------------------------------------------------------------
class AdamTF1:
    """[TF1] tf.train.AdamOptimizer / ApplyAdam functor (training_ops.cc), all in fp32:
        alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)         (b^t kept as fp32 running products)
        m += (g - m) * (1 - b1)
        v += (g*g - v) * (1 - b2)
        var -= (m * alpha) / (sqrt(v) + eps)             (eps outside the bias correction)
    Dense: every element of every trainable variable is updated every step."""

    def __init__(self, lr):
        self.lr = F32(lr)
        self.b1_pow = F32(ADAM_B1)
        self.b2_pow = F32(ADAM_B2)
        self.state = {}

    def alpha(self):
        one = F32(1.0)
        return F32(self.lr * np.sqrt(one - self.b2_pow, dtype=np.float32) / (one - self.b1_pow))

    def apply(self, name, var, g):
        m, v = self.state.setdefault(name, (np.zeros_like(var), np.zeros_like(var)))
        one = F32(1.0)
        a = self.alpha()
        g = g.astype(np.float32)
        m += (g - m) * (one - ADAM_B1)
        v += (g * g - v) * (one - ADAM_B2)
        var -= (m * a) / (np.sqrt(v, dtype=np.float32) + ADAM_EPS)

    def finish_step(self):
        self.b1_pow = F32(self.b1_pow * ADAM_B1)
        self.b2_pow = F32(self.b2_pow * ADAM_B2)


# ----------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------
def xavier_uniform(rng, fan_in, fan_out):
    """tf.contrib.layers.xavier_initializer() [TF1]: U(+-sqrt(6/(fan_in+fan_out))).  DAEs.py:54-55"""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


class DAEOracle:
    """DAE_tied (DAEs.py:13-111) when tied=True, DAE (DAEs.py:114-150) otherwise."""

    def __init__(self, n_input, n_hidden, lr, reg_lambda=0.0, tied=False, seed=0, mode="fp32", params=None):
        assert mode in ("fp32", "b200")
        self.N, self.H, self.tied, self.mode = n_input, n_hidden, tied, mode
        self.reg_lambda = F32(reg_lambda)
        rng = np.random.default_rng(seed)
        if params is None:
            W_enc = xavier_uniform(rng, n_input, n_hidden)
            W_dec = W_enc if tied else xavier_uniform(rng, n_input, n_hidden)
            params = [W_enc, W_dec, np.zeros(n_hidden, np.float32), np.zeros(n_input, np.float32)]
        self.W_enc = np.array(params[0], dtype=np.float32)
        self.W_dec = self.W_enc if tied else np.array(params[1], dtype=np.float32)
        self.b_enc = np.array(params[2], dtype=np.float32)
        self.b_dec = np.array(params[3], dtype=np.float32)
        self.adam = AdamTF1(lr)
        self.step = 0

    # d_params order of save_model                                  DAEs.py:60-61, 107-111, 137-138
    def params(self):
        return [self.W_enc, self.W_dec, self.b_enc, self.b_dec]

    def _rd(self, x):
        return bf16_round(x) if self.mode == "b200" else np.asarray(x, dtype=np.float32)

    def forward(self, x_pos, x_val, B, kp=1.0, kp_in=1.0, seed=0, step=0, row_offset=0,
                keep_in=None, keep_h=None):
        row_ptr, col, val = coo_to_csr_last_wins(x_pos, x_val, B, self.N)
        rows = csr_rows(row_ptr)
        if keep_in is None:
            keep_in = philox.keep_mask(seed, philox.STREAM_INPUT, step, rows + row_offset, col, kp_in)
        if keep_h is None:
            rr, kk = np.meshgrid(np.arange(B, dtype=np.uint32) + np.uint32(row_offset),
                                 np.arange(self.H, dtype=np.uint32), indexing="ij")
            keep_h = philox.keep_mask(seed, philox.STREAM_HIDDEN, step, rr, kk, kp)
        x_n, s = normalise_input(row_ptr, val, keep_in, kp_in)
        a, h, h_d = encode(self.W_enc, self.b_enc, row_ptr, col, x_n, keep_h, kp)
        h_dq = self._rd(h_d)
        z, p = decode(h_dq, self._rd(self.W_dec), self.b_dec)
        return dict(row_ptr=row_ptr, col=col, val=val, x_n=x_n, s=s, a=a, h=h, h_d=h_d, h_dq=h_dq,
                    keep_in=keep_in, keep_h=keep_h, z=z, p=p)

    def predict(self, x_pos, x_val, B):
        """sess.run(y_pred, keep_prob=1, input_keep_prob=1)           main_train.py:66-68"""
        return self.forward(x_pos, x_val, B)["p"]

    def loss_and_grads(self, x_pos, x_val, y_pos, y_val, B, kp, kp_in, seed=0, step=0, row_offset=0,
                       global_batch=None, keep_in=None, keep_h=None):
        f = self.forward(x_pos, x_val, B, kp, kp_in, seed, step, row_offset, keep_in, keep_h)
        y = densify_last_wins(y_pos, y_val, B, self.N)
        inv_b = 1.0 / float(global_batch or B)
        L = bce_rows(f["p"], y)
        l2 = F32(0.0)
        if self.reg_lambda != 0:                                   # DAEs.py:79-82 / 147-150; l2_loss = sum(t^2)/2 [TF1]
            l2 = sum(float((t.astype(np.float64) ** 2).sum()) * 0.5
                     for t in ([self.W_enc, self.b_dec, self.b_enc] + ([] if self.tied else [self.W_dec])))
        cost = float(L.sum() * inv_b + float(self.reg_lambda) * float(l2))      # DAEs.py:100
        dz = bce_dz(f["p"], y, inv_b)
        db_dec = dz.astype(np.float64).sum(axis=0).astype(np.float32)
        dzq = self._rd(dz)
        dW_dec = (dzq.T @ f["h_dq"]).astype(np.float32)
        dh_d = (dzq @ self._rd(self.W_dec)).astype(np.float32)
        da = (dh_d * (f["keep_h"].astype(np.float32) / F32(kp)) * (f["h"] * (F32(1.0) - f["h"]))).astype(np.float32)
        db_enc = da.astype(np.float64).sum(axis=0).astype(np.float32)
        dW_enc = np.zeros_like(self.W_enc)
        rows = csr_rows(f["row_ptr"])
        np.add.at(dW_enc, f["col"], f["x_n"][:, None] * da[rows])
        f.update(y=y, dz=dz, dzq=dzq, dh_d=dh_d, da=da)
        grads = dict(W_enc=dW_enc, W_dec=dW_dec, b_enc=db_enc, b_dec=db_dec)
        return cost, grads, f

    def apply_grads(self, grads):
        lam = self.reg_lambda
        if self.tied:
            g = grads["W_enc"] + grads["W_dec"] + lam * self.W_enc
            self.adam.apply("W_enc", self.W_enc, g)
        else:
            self.adam.apply("W_enc", self.W_enc, grads["W_enc"] + lam * self.W_enc)
            self.adam.apply("W_dec", self.W_dec, grads["W_dec"] + lam * self.W_dec)
        self.adam.apply("b_enc", self.b_enc, grads["b_enc"] + lam * self.b_enc)
        self.adam.apply("b_dec", self.b_dec, grads["b_dec"] + lam * self.b_dec)
        self.adam.finish_step()
        self.step += 1

    def train_step(self, x_pos, x_val, y_pos, y_val, B, kp, kp_in, seed=0, **kw):
        """sess.run([optimizer, cost], feed_dict)                      main_train.py:204-213"""
        cost, grads, _ = self.loss_and_grads(x_pos, x_val, y_pos, y_val, B, kp, kp_in, seed=seed,
                                             step=self.step, **kw)
        self.apply_grads(grads)
        return cost


# ----------------------------------------------------------------------------
# a9: title mixing (DAE_title)                                   DAEs.py:153-181, 194-196
# ----------------------------------------------------------------------------
def title_mix(p, title_score, s, kp_in, titles_use):
    """x_count = s*kp_in ; w_t = u/(u+x_count+eps) ; w_p = x_count/(u+x_count+eps) ;
    y_pred = title_score*w_t + p*w_p.                              DAEs.py:159-162, 180"""
    x_count = (s.astype(np.float32) * F32(kp_in)).reshape(-1, 1)
    u = np.asarray(titles_use, dtype=np.float32).reshape(-1, 1)
    deno = u + x_count + EPS_LOG
    w_t = u / deno
    w_p = x_count / deno
    return (title_score.astype(np.float32) * w_t + p.astype(np.float32) * w_p).astype(np.float32), w_t, w_p


def title_dq(q, y, inv_batch):
    """d cost/d y_pred for the title-mode loss (no sigmoid chain here).  DAEs.py:194-196"""
    q = q.astype(np.float32); y = y.astype(np.float32)
    one = F32(1.0)
    return ((-(y / (q + EPS_LOG)) + NEG_WEIGHT * (one - y) / (one - q + EPS_LOG)) * F32(inv_batch)).astype(np.float32)
