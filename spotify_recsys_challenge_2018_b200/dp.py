"""Data-parallel training of the DAE across the GPUs of one box (one process per GPU).

The reference has no data parallelism.  The train step shards naturally by playlist:
rank r runs encode / decode / loss / dh on rows [r*B, (r+1)*B) of the global batch (the loss is a
mean over the GLOBAL batch, models/DAEs.py:100; dropout is keyed by the global row, so N ranks x B
reproduce one rank x N*B).  The catalogue-sized state -- W_enc, W_dec, their Adam moments -- is
row-sharded tile-cyclically over the ranks, and every exchange is done by the CUDA kernels
themselves with loads / stores into the peers' memory over NVLink / NVSwitch (include/dae_b200.h,
"data parallelism"): there is no collective call in the step.  torch.distributed is only the
out-of-band channel that carries the 64-byte CUDA IPC handles at start-up.

`shard_coo` and `exchange_handles` are backend-agnostic (tested with gloo on CPU, world_size 2).
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


def tile_owner(item, world, tile=128):
    """Rank that owns catalogue row `item` (csrc/kernels.h item_owner): 128-row tiles, cyclic."""
    return (np.asarray(item) // tile) % world


def tile_local_row(item, world, tile=128):
    """Row of `item` inside its owner's shard (csrc/kernels.h item_local)."""
    item = np.asarray(item)
    return (item // tile // world) * tile + item % tile


def exchange_handles(handle, group=None):
    """All-gather one opaque bytes object per rank, in rank order."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(handle), group=group)
    return out


def attach_peers(model, group=None):
    """Map every rank's arena into this rank's model (CUDA IPC), then line the ranks up."""
    import torch.distributed as dist
    model.attach_ipc(exchange_handles(model.ipc_handle(), group))
    dist.barrier(group)


class DataParallelDAE:
    """A models.DAEs model created with conf.world / conf.rank on this rank's GPU, attached to its peers.

    train_step_staged is the single-GPU call: the library enqueues the two cross-GPU flag barriers
    of the step itself.  Every rank must stage batches of the same size and call in the same order."""

    def __init__(self, model, group=None):
        import torch.distributed as dist
        self.model, self.group = model, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if model.world != self.world or model.rank != self.rank:
            raise ValueError("model was created for rank %d/%d, process group is %d/%d"
                             % (model.rank, model.world, self.rank, self.world))
        attach_peers(model, group)

    def train_step_staged(self, slot, keep_prob, input_keep_prob):
        self.model.train_step_staged(slot, keep_prob, input_keep_prob)

    def stage_global_batch(self, slot, x_positions, x_vals, y_positions, y_vals):
        """Every rank is handed the GLOBAL reader batch and keeps its own rows."""
        b = self.model.n_batch
        xp, xv = shard_coo(x_positions, x_vals, self.rank, b)
        yp, yv = shard_coo(y_positions, y_vals, self.rank, b)
        self.model.stage_batch(slot, xp, xv, yp, yv)

    def get_params(self):
        """Gather the four parameter arrays on this rank (all ranks idle: barrier on both sides)."""
        import torch
        import torch.distributed as dist
        self.model.sync_cost()
        dist.barrier(self.group)
        out = self.model.get_params()
        torch.cuda.synchronize()
        dist.barrier(self.group)
        return out
