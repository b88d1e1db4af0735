"""Data-parallel training of the DAE across the GPUs of one box (one process per GPU).

The reference has no data parallelism.  The train step shards naturally by playlist:
rank r runs encode / decode / loss / dh on rows [r*B, (r+1)*B) of the global batch (the loss is a
mean over the GLOBAL batch, models/DAEs.py:100; dropout is keyed by the global row, so N ranks x B
reproduce one rank x N*B).  The catalogue-sized state -- W_enc, W_dec, their Adam moments -- is
row-sharded tile-cyclically over the ranks, and every exchange is done by the CUDA kernels
themselves with loads / stores into the peers' memory over NVLink / NVSwitch (include/dae_b200.h,
"data parallelism"): there is no collective call in the step.  torch.distributed is only the
out-of-band channel that carries the 64-byte CUDA IPC handles at start-up.

Challenge inference shards on the ITEM axis instead (ShardedRecommender): every rank ranks its own slice of the
track catalogue with the fused decode + top-K; the per-shard lists are stored into every peer's merge buffer over
NVLink (no collective call either) and merged on the device.

`shard_coo`, `item_shard`, `merge_topk_lists` and `exchange_handles` are backend-agnostic (tested with gloo on CPU,
world_size 2).
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

    train_step_staged is the single-GPU call: the library enqueues the four cross-GPU flag barriers
    of the step itself (include/dae_b200.h, "data parallelism").  Every rank must stage batches of the same size and call in the same order."""

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


def item_shard(n_tracks, rank, world, tile=128):
    """Contiguous slice [lo, hi) of the track catalogue ranked by `rank`: equal numbers of 128-item tiles (the
    decode kernel's granularity), the last rank takes the ragged end.  The slices partition [0, n_tracks)."""
    tiles = (n_tracks + tile - 1) // tile
    per = (tiles + world - 1) // world
    lo = min(rank * per * tile, n_tracks)
    hi = min((rank + 1) * per * tile, n_tracks)
    return lo, hi


def merge_topk_lists(idx_lists, score_lists, k):
    """Host statement of the merge rule (score desc, id asc; -1 = padding) over per-shard [B, k] lists -> [B, k].
    The device path (dae_topk_merge_device) must return exactly this."""
    idx = np.concatenate(idx_lists, axis=1)
    sc = np.concatenate(score_lists, axis=1).astype(np.float64)
    sc = np.where(idx >= 0, sc, -np.inf)
    order = np.lexsort((idx, -sc), axis=1)[:, :k]
    out_i = np.take_along_axis(idx, order, 1)
    out_s = np.take_along_axis(sc, order, 1)
    return np.where(np.isfinite(out_s), out_i, -1).astype(np.int32), out_s.astype(np.float32)


class ShardedRecommender:
    """Challenge-mode inference sharded on the item axis (main_challenge.py:80-90 on one GPU upstream).

    Every rank holds a replica of the inference model (a plain world = 1 model on its own GPU: 2 M x 256 parameters are
    ~5 GB of 180), encodes the batch redundantly (microseconds) and ranks the tracks of ITS slice with the fused
    decode + top-K.  The per-shard lists never leave the devices and no collective library is involved: each rank's
    final select is followed by plain stores of its B x k (id, score) pairs into every peer's merge buffer over NVLink
    (`dae_exchange_*`, CUDA IPC mapping), one flag barrier, and the (score desc, id asc) merge of world x k candidates
    per playlist on every rank: exactly the unsharded list.  torch.distributed only carries the 64-byte IPC handles at
    construction.  `local_peers`: the exchanges of all ranks living in THIS process (tests on one GPU)."""

    def __init__(self, model, group=None, rank=None, world=None, max_k=500):
        import ctypes as C
        from . import _lib
        self.model, self.group = model, group
        explicit = rank is not None              # ranks of one process (tests): the caller attaches them with attach_local
        if not explicit:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.world = rank, world
        self.range = item_shard(model.n_tracks, self.rank, self.world)
        self._lib = _lib.load()
        self._x = C.c_void_p()
        _lib.check(self._lib.dae_exchange_create(model.device, world, rank, model.n_batch, int(max_k), C.byref(self._x)))
        self.max_k = int(max_k)
        self.share_thresholds = True
        if not explicit and world > 1:
            self.attach_ipc(group)

    def ipc_handle(self):
        import ctypes as C
        buf = C.create_string_buffer(64)
        from . import _lib
        _lib.check(self._lib.dae_exchange_ipc_handle(self._x, buf))
        return buf.raw

    def attach_ipc(self, group=None):
        import torch.distributed as dist
        from . import _lib
        blob = b"".join(exchange_handles(self.ipc_handle(), group))
        _lib.check(self._lib.dae_exchange_attach_ipc(self._x, blob, self.world))
        dist.barrier(group)

    def attach_local(self, recommenders):
        import ctypes as C
        from . import _lib
        arr = (C.c_void_p * len(recommenders))(*[r._x for r in recommenders])
        _lib.check(self._lib.dae_exchange_attach_local(self._x, arr, len(recommenders)))

    def rank_shard(self, x_positions, x_vals, seeds, k=500):
        """This rank's slice: lists stay on the device (model buffers "topk_idx" / "topk_score").  Collective when
        `share_thresholds` (default): the shards agree on the filter thresholds of the fused decode + top-K (the minimum
        over the shards of each one's ceil(kp / world)-th largest logit), which shortens every shard's candidate lists by
        ~world; the lists returned are the same either way."""
        from . import _lib
        share = self.share_thresholds and self.world > 1
        if share:
            _lib.check(self._lib.dae_model_set_threshold_exchange(self.model._h, self._x))
        try:
            self.model.recommend(x_positions, x_vals, seeds, k=k, item_range=self.range, on_device=True)
        finally:
            if share:
                _lib.check(self._lib.dae_model_set_threshold_exchange(self.model._h, None))

    def merge(self, k=500, return_scores=False, reuse_output=False, rows="all"):
        """Collective: store this rank's lists into every peer's merge buffer, barrier, merge locally -> host arrays
        (`reuse_output=True`: page-locked arrays owned by this object, overwritten by the next call).
        rows="own": the merge is sharded by playlist -- this rank receives, merges and returns only ITS rows
        [r0, r1) of the batch (`self.rows` after the call; the returned arrays are those rows): 1 / world of the stores,
        the merge and the read-back, for callers that write their own part of the submission."""
        import ctypes as C
        from . import _lib
        from .models.DAEs import _PinnedPool
        B = self.model.n_batch
        pi, _, _ = self.model.buffer("topk_idx")
        ps, _, _ = self.model.buffer("topk_score")
        if reuse_output:
            if not hasattr(self, "_pinned"):
                self._pinned = _PinnedPool()
            out_i = self._pinned.get((B, k), np.int32)
            out_s = self._pinned.get((B, k), np.float32) if return_scores else None
        else:
            out_i = np.empty((B, k), np.int32)
            out_s = np.empty((B, k), np.float32) if return_scores else None
        st = C.c_void_p(self.model.stream) if self.model.stream else None
        po_i = out_i.ctypes.data_as(C.c_void_p)
        po_s = out_s.ctypes.data_as(C.c_void_p) if out_s is not None else None
        if rows == "own":
            r0, r1 = C.c_int32(), C.c_int32()
            _lib.check(self._lib.dae_exchange_merge_topk_rows(self._x, C.c_void_p(pi), C.c_void_p(ps), B, int(k),
                                                              C.byref(r0), C.byref(r1), po_i, po_s, st))
            self.rows = (r0.value, r1.value)
            out_i = out_i[r0.value:r1.value]
            out_s = out_s[r0.value:r1.value] if out_s is not None else None
        elif rows == "all":
            _lib.check(self._lib.dae_exchange_merge_topk(self._x, C.c_void_p(pi), C.c_void_p(ps), B, int(k), po_i, po_s, st))
            self.rows = (0, B)
        else:
            raise ValueError("rows must be 'all' or 'own'")
        return (out_i, out_s) if return_scores else out_i

    def recommend(self, x_positions, x_vals, seeds, k=500, return_scores=False, reuse_output=False, rows="all"):
        self.rank_shard(x_positions, x_vals, seeds, k)
        return self.merge(k, return_scores, reuse_output, rows)

    def set_profiling(self, on):
        from . import _lib
        _lib.check(self._lib.dae_exchange_set_profiling(self._x, 1 if on else 0))

    def phase_times(self):
        """ms per call: stores, barrier, merge, read-back (device times, mean over the profiled calls)."""
        import ctypes as C
        from . import _lib
        out = (C.c_float * 4)()
        _lib.check(self._lib.dae_exchange_phase_ms(self._x, out))
        return dict(zip(("exchange_store", "exchange_barrier", "exchange_merge", "exchange_d2h"), [float(v) for v in out]))

    def launch_count(self):
        return int(self._lib.dae_exchange_launch_count(self._x))

    def close(self):
        if self._x:
            self._lib.dae_exchange_destroy(self._x)
            self._x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
