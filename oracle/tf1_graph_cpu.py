"""The reference's TF1 graph restated op-for-op on torch-CPU (test infrastructure).

This is the reported CPU comparator ("port"): TensorFlow 1.5 is not installable
here (Python 3.12, no network), so the op sequence TF executes for one
``sess.run([optimizer, cost])`` of /root/reference/models/DAEs.py is restated
with dense fp32 torch-CPU ops on all host cores:

  sparse->dense x,y (DAEs.py:33-38) -> dense dropout (:40) -> reduce_sum / divide (:41-42)
  -> dense x.W_enc + b, sigmoid, dropout (:66-68) -> dense h.W_dec^T + b, sigmoid (:75-76/:143-144)
  -> weighted BCE, reduce_mean, + lambda*l2 (:98-100) -> the dense matmuls TF autodiff emits
  -> dense ApplyAdam on every variable (:102).

It is validated against oracle/dae_oracle.py (mode="fp32") in tests/test_oracle.py.
Only bench.py (cpu_baseline / --impl reference) and tests may import it.
"""
from __future__ import annotations

import numpy as np
import torch


class TF1GraphCPU:
    def __init__(self, n_input, n_hidden, lr, reg_lambda=0.0, tied=False, params=None, seed=0):
        g = torch.Generator().manual_seed(seed)
        lim = (6.0 / (n_input + n_hidden)) ** 0.5
        def xav():
            return (torch.rand(n_input, n_hidden, generator=g) * 2 - 1) * lim
        if params is None:
            W_enc = xav(); W_dec = W_enc if tied else xav()
            b_enc = torch.zeros(n_hidden); b_dec = torch.zeros(n_input)
        else:
            W_enc = torch.tensor(np.array(params[0]), dtype=torch.float32)
            W_dec = W_enc if tied else torch.tensor(np.array(params[1]), dtype=torch.float32)
            b_enc = torch.tensor(np.array(params[2]), dtype=torch.float32)
            b_dec = torch.tensor(np.array(params[3]), dtype=torch.float32)
        self.tied = tied
        self.N, self.H = n_input, n_hidden
        self.lr, self.lam = lr, reg_lambda
        self.vars = {"W_enc": W_enc, "b_enc": b_enc, "b_dec": b_dec}
        if not tied:
            self.vars["W_dec"] = W_dec
        self.m = {k: torch.zeros_like(v) for k, v in self.vars.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.vars.items()}
        self.b1p = np.float32(0.9); self.b2p = np.float32(0.999)

    @staticmethod
    def _densify(pos, val, B, N):
        x = torch.zeros(B, N)
        pos = np.asarray(pos).reshape(-1, 2).astype(np.int64)
        if pos.shape[0]:
            # assignment scatter; duplicates: last wins (index_put on CPU is in-order for
            # non-accumulate mode; tests pin this against the numpy oracle)
            flat = torch.from_numpy(pos[:, 0] * N + pos[:, 1])
            v = torch.tensor(np.asarray(val, dtype=np.float32))
            # enforce last-wins deterministically: keep only the last occurrence of each index
            rev = torch.flip(flat, [0])
            uniq, first_in_rev = np.unique(rev.numpy(), return_index=True)
            keep = (flat.shape[0] - 1 - first_in_rev)
            x.view(-1)[torch.from_numpy(uniq)] = v[torch.from_numpy(keep)]
        return x

    def train_step(self, x_pos, x_val, y_pos, y_val, B, kp, kp_in, keep_in_dense=None, keep_h=None):
        N = self.N
        W_enc = self.vars["W_enc"]; W_dec = W_enc if self.tied else self.vars["W_dec"]
        b_enc = self.vars["b_enc"]; b_dec = self.vars["b_dec"]
        x = self._densify(x_pos, x_val, B, N)
        y = self._densify(y_pos, y_val, B, N)
        if keep_in_dense is None:                       # tf.nn.dropout draws one uniform per CELL
            keep_in_dense = (torch.rand(B, N) < kp_in)
        x_d = x / kp_in * keep_in_dense
        s = x_d.sum(1, keepdim=True)
        x_n = x_d / (s + 1e-10)
        a = x_n @ W_enc + b_enc
        h = torch.sigmoid(a)
        if keep_h is None:
            keep_h = (torch.rand(B, self.H) < kp)
        mh = keep_h / kp
        h_d = h * mh
        z = h_d @ W_dec.t() + b_dec
        p = torch.sigmoid(z)
        L = -(y * torch.log(p + 1e-10) + 0.55 * (1 - y) * torch.log(1 - p + 1e-10)).sum(1)
        cost = L.mean()
        if self.lam != 0:
            l2 = sum((t * t).sum() / 2 for t in self.vars.values())
            cost = cost + self.lam * l2
        # the matmuls TF's autodiff emits (dense, including the dense-x one)
        pq = p * (1 - p)
        dz = (-(y * (pq / (p + 1e-10))) + 0.55 * (1 - y) * (pq / (1 - p + 1e-10))) / B
        dW_dec = dz.t() @ h_d
        db_dec = dz.sum(0)
        dh_d = dz @ W_dec
        da = dh_d * mh * (h * (1 - h))
        dW_enc = x_n.t() @ da
        db_enc = da.sum(0)
        grads = {"W_enc": dW_enc + dW_dec if self.tied else dW_enc, "b_enc": db_enc, "b_dec": db_dec}
        if not self.tied:
            grads["W_dec"] = dW_dec
        one = np.float32(1.0)
        alpha = float(np.float32(self.lr) * np.sqrt(one - self.b2p, dtype=np.float32) / (one - self.b1p))
        for k, var in self.vars.items():
            g = grads[k] + self.lam * var if self.lam != 0 else grads[k]
            m, v = self.m[k], self.v[k]
            m += (g - m) * (1 - 0.9)
            v += (g * g - v) * (1 - 0.999)
            var -= (m * alpha) / (v.sqrt() + 1e-8)
        self.b1p = np.float32(self.b1p * np.float32(0.9)); self.b2p = np.float32(self.b2p * np.float32(0.999))
        return float(cost)

    def predict(self, x_pos, x_val, B):
        W_enc = self.vars["W_enc"]; W_dec = W_enc if self.tied else self.vars["W_dec"]
        x = self._densify(x_pos, x_val, B, self.N)
        x_n = x / (x.sum(1, keepdim=True) + 1e-10)
        h = torch.sigmoid(x_n @ W_enc + self.vars["b_enc"])
        return torch.sigmoid(h @ W_dec.t() + self.vars["b_dec"]).numpy()
