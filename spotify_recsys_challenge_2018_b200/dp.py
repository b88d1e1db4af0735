"""Data-parallel training of the DAE across the GPUs of one box (one process per GPU).

The reference has no data parallelism; the train step shards naturally by playlist:
every rank holds full replicas of the parameters and Adam state, processes B_local rows of the
global batch, scales its gradients by 1/B_global (the loss is a mean over the GLOBAL batch,
models/DAEs.py:100) and the gradients are sum-all-reduced once per step over NCCL / NVLink; then
every rank applies the identical dense Adam update.  Dropout masks are keyed by the global row, so
N ranks x B_local reproduce one rank x (N * B_local) up to fp32 summation order.

`shard_coo` and `allreduce_grads` are backend-agnostic (tested with gloo on CPU, world_size 2).
"""
from __future__ import annotations

import numpy as np


def shard_coo(positions, vals, rank, b_local):
    """Rows [rank*b_local, (rank+1)*b_local) of a reader batch, re-based to local row 0.
    Order inside the shard is preserved (last-wins de-duplication depends on it)."""
    pos = np.asarray(positions).reshape(-1, 2).astype(np.int64)
    val = np.asarray(vals, dtype=np.float32).reshape(-1)
    lo = rank * b_local
    keep = (pos[:, 0] >= lo) & (pos[:, 0] < lo + b_local)
    out = pos[keep].copy()
    out[:, 0] -= lo
    return out, val[keep]


def allreduce_grads(tensors, group=None, flags=None):
    """Sum-all-reduce every gradient tensor (and max-reduce the row flags) in place."""
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if flags is not None:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)


class _DevArr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class DataParallelDAE:
    """Wraps a models.DAEs model created on this rank's GPU and stream."""

    def __init__(self, model, group=None):
        import torch
        import torch.distributed as dist
        self.model, self.group = model, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        view = lambda name, ts: torch.as_tensor(_DevArr(*model.buffer(name)[:2], ts), device="cuda")
        names = ["g_dec", "g_b_enc", "g_b_dec", "cost"] + ([] if model.tied else ["g_enc"])
        self.grads = [view(n, "<f4") for n in names]
        self.flags = None if model.tied else view("touched", "|u1")

    def train_step_staged(self, slot, keep_prob, input_keep_prob):
        b = self.model.n_batch
        self.model.backward_staged(slot, keep_prob, input_keep_prob, global_batch=b * self.world,
                                   row_offset=b * self.rank)
        allreduce_grads(self.grads, self.group, self.flags)
        self.model.apply_adam()

    def stage_global_batch(self, slot, x_positions, x_vals, y_positions, y_vals):
        b = self.model.n_batch
        xp, xv = shard_coo(x_positions, x_vals, self.rank, b)
        yp, yv = shard_coo(y_positions, y_vals, self.rank, b)
        self.model.stage_batch(slot, xp, xv, yp, yv)
