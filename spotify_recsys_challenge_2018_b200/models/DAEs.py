"""DAE_tied / DAE / DAE_title over libdae_b200.so -- host mirror of the reference's models/DAEs.py.

Same class names, constructor signatures and `conf` fields as the reference
(models/DAEs.py:13-21, :114-117, :153-156); `fit()` builds the device model instead of a TF graph;
`sess.run([optimizer, cost], feed_dict)` becomes `train_step(...)`, `sess.run(y_pred, feed_dict)`
becomes `predict(...)`, and `recommend(...)` fuses the ranking the runners do on the host
(utils/metrics.py:58-68).  `save_model()` writes the reference's 4-array pickle (DAEs.py:107-111).

No compute happens in Python and there is no CPU path: every method calls the C ABI.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np

from .. import _lib


def _coo(positions, vals):
    """Reader output -> contiguous int64 [nnz,2] + float32 [nnz].  The reference readers emit
    float64 positions whenever a batch row is empty and Python lists for values."""
    pos = np.ascontiguousarray(np.asarray(positions).reshape(-1, 2), dtype=np.int64)
    val = np.ascontiguousarray(np.asarray(vals, dtype=np.float32).reshape(-1))
    if pos.shape[0] != val.shape[0]:
        raise ValueError("positions (%d) and values (%d) differ in length" % (pos.shape[0], val.shape[0]))
    return pos, val


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class _PinnedPool:
    """Page-locked host arrays for results that are consumed before the next call (`reuse_output=True`): a D2H copy into
    pinned memory runs at PCIe speed, one into a fresh pageable NumPy array is staged by the driver at a fraction of it
    (8 MB of ids per 4096-playlist challenge batch: ~1 ms vs ~0.2 ms).  torch is only the allocator here."""

    def __init__(self):
        self._bufs = {}

    def get(self, shape, dtype):
        import torch
        key = (tuple(shape), np.dtype(dtype).str)
        if key not in self._bufs:
            tdt = {"<i4": torch.int32, "<f4": torch.float32, "<f8": torch.float64}[np.dtype(dtype).str]
            self._bufs[key] = torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
        return self._bufs[key].numpy()


def _csr(lists, n_rows, pad_value=None):
    """Per-row id lists -> (ptr int32 [n_rows + 1], flat int32).  Rows past len(lists) (a short last batch) are empty, or hold
    the single id `pad_value` when one is given (the metrics kernel divides by the row length)."""
    lists = [np.asarray(s, dtype=np.int64).reshape(-1) for s in lists]
    if pad_value is not None:
        lists = lists + [np.asarray([pad_value], np.int64)] * (n_rows - len(lists))
    ptr = np.zeros(n_rows + 1, np.int32)
    lens = [len(s) for s in lists]
    ptr[1:len(lens) + 1] = np.cumsum(lens)
    ptr[len(lens) + 1:] = ptr[len(lens)]
    flat = np.concatenate(lists) if lens and sum(lens) else np.zeros(0, np.int64)
    return ptr, np.ascontiguousarray(np.clip(flat, -1, 2 ** 31 - 1).astype(np.int32))


class DAE_tied:
    """Tied-weight DAE used by --pretrain (reference models/DAEs.py:13-111)."""

    tied = True
    trainable = True

    def __init__(self, conf):
        self.save_dir = conf.save                         # DAEs.py:15
        self.n_batch = int(conf.batch)                    # DAEs.py:17
        self.n_input = int(conf.n_input)                  # DAEs.py:18
        self.n_hidden = int(conf.hidden)                  # DAEs.py:19
        self.learning_rate = float(conf.lr)               # DAEs.py:20
        self.reg_lambda = float(conf.reg_lambda)          # DAEs.py:21
        self.n_tracks = int(getattr(conf, "n_tracks", self.n_input))
        self.seed = int(getattr(conf, "seed", 0))
        self.device = int(getattr(conf, "device", 0))
        self.stream = getattr(conf, "stream", None)
        self.world = int(getattr(conf, "world", 1))       # data-parallel ranks (one model per GPU), dp.py
        self.rank = int(getattr(conf, "rank", 0))
        self._h = None
        self._lib = None

    # ---- lifecycle ---------------------------------------------------------------
    def _create(self):
        lib = _lib.load()
        cfg = _lib.DaeConfig(self.n_input, self.n_tracks, self.n_hidden, self.n_batch, int(self.tied),
                             self.learning_rate, self.reg_lambda, self.seed, self.device, int(self.trainable),
                             self.stream, self.world, self.rank)
        h = C.c_void_p()
        _lib.check(lib.dae_model_create(C.byref(cfg), C.byref(h)))
        self._h, self._lib = h, lib

    def init_weight(self):
        """Xavier-uniform W, zero biases (DAEs.py:53-61)."""
        _lib.check(self._lib.dae_model_init_xavier(self._h, self.seed))

    def fit(self):
        """Build the model on the device and initialise it (DAEs.py:84-105 + sess.run(init_op))."""
        self._create()
        self.init_weight()
        return self

    # ---- data-parallel attachment (world > 1; see dp.py) ----------------------------------
    def ipc_handle(self):
        """64-byte CUDA IPC handle of this rank's arena."""
        buf = C.create_string_buffer(64)
        _lib.check(self._lib.dae_model_ipc_handle(self._h, buf))
        return buf.raw

    def attach_ipc(self, handles):
        """Map the arenas of all ranks (handles in rank order, one per process)."""
        blob = b"".join(handles)
        _lib.check(self._lib.dae_model_attach_ipc(self._h, blob, len(handles)))

    def attach_local(self, models):
        """Peers living in this process (all ranks in rank order, including self)."""
        arr = (C.c_void_p * len(models))(*[m._h for m in models])
        _lib.check(self._lib.dae_model_attach_local(self._h, arr, len(models)))

    def set_debug(self, flags):
        _lib.check(self._lib.dae_model_set_debug(self._h, int(flags)))

    def close(self):
        if self._h is not None:
            self._lib.dae_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ----------------------------------------------------------------
    def set_params(self, params):
        W_enc, W_dec, b_enc, b_dec = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
        if W_enc.shape != (self.n_input, self.n_hidden) or b_enc.shape != (self.n_hidden,) \
                or b_dec.shape != (self.n_input,) or W_dec.shape != W_enc.shape:
            raise ValueError("parameter shapes do not match the model")
        _lib.check(self._lib.dae_model_set_params(self._h, _ptr(W_enc), _ptr(W_dec), _ptr(b_enc), _ptr(b_dec)))

    def get_params(self):
        """d_params order: [encoder_h, decoder_h (== encoder_h when tied), encoder_b, decoder_b] (DAEs.py:60-61)."""
        W_enc = np.empty((self.n_input, self.n_hidden), np.float32)
        b_enc = np.empty(self.n_hidden, np.float32)
        b_dec = np.empty(self.n_input, np.float32)
        if self.tied:
            _lib.check(self._lib.dae_model_get_params(self._h, _ptr(W_enc), None, _ptr(b_enc), _ptr(b_dec)))
            W_dec = W_enc
        else:
            W_dec = np.empty_like(W_enc)
            _lib.check(self._lib.dae_model_get_params(self._h, _ptr(W_enc), _ptr(W_dec), _ptr(b_enc), _ptr(b_dec)))
        return [W_enc, W_dec, b_enc, b_dec]

    def save_model(self, sess=None):
        """pickle.dump(sess.run(d_params)) (DAEs.py:107-111); `sess` accepted and ignored."""
        with open(self.save_dir, "wb") as f:
            pickle.dump(self.get_params(), f)

    # ---- the hot path -----------------------------------------------------------------
    def train_step(self, x_positions, x_vals, y_positions, y_vals, keep_prob, input_keep_prob):
        """One `sess.run([optimizer, cost])` (main_train.py:204-213) -> cost (float)."""
        xp, xv = _coo(x_positions, x_vals)
        yp, yv = _coo(y_positions, y_vals)
        cost = C.c_float()
        _lib.check(self._lib.dae_model_train_step(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(yp), _ptr(yv),
                                                  yp.shape[0], self.n_batch, float(keep_prob),
                                                  float(input_keep_prob), C.byref(cost)))
        return float(cost.value)

    def train_step_async(self, x_positions, x_vals, y_positions, y_vals, keep_prob, input_keep_prob):
        """The same step, pipelined: returns the cost of the PREVIOUS step (None on the first call) while this one
        runs; `flush()` returns the last pending cost.  The runner only accumulates the cost (main_train.py:223)."""
        xp, xv = _coo(x_positions, x_vals)
        yp, yv = _coo(y_positions, y_vals)
        cost = C.c_float(); has = C.c_int32()
        _lib.check(self._lib.dae_model_train_step_async(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(yp), _ptr(yv),
                                                        yp.shape[0], self.n_batch, float(keep_prob),
                                                        float(input_keep_prob), C.byref(cost), C.byref(has)))
        return float(cost.value) if has.value else None

    def flush(self):
        cost = C.c_float(); has = C.c_int32()
        _lib.check(self._lib.dae_model_train_flush(self._h, C.byref(cost), C.byref(has)))
        return float(cost.value) if has.value else None

    def predict(self, x_positions, x_vals, tracks_only=False):
        """`sess.run(y_pred, keep_prob=1, input_keep_prob=1)` (main_train.py:66-68) -> [batch, n_input]
        (or [batch, n_tracks], the slice the runner keeps, main_train.py:86)."""
        xp, xv = _coo(x_positions, x_vals)
        n_cols = self.n_tracks if tracks_only else self.n_input
        out = np.empty((self.n_batch, n_cols), np.float32)
        _lib.check(self._lib.dae_model_predict(self._h, _ptr(xp), _ptr(xv), xp.shape[0], self.n_batch, n_cols,
                                               _ptr(out)))
        return out

    def recommend(self, x_positions, x_vals, seeds, k=500, return_scores=False, item_range=None, on_device=False,
                  reuse_output=False):
        """Top-k track ids per playlist with the seeds removed (metrics.py:58-68, main_challenge.py:26-36),
        decode + ranking on the device.  `seeds`: list of per-row seed id lists, or the CSR pair (seed_ptr, seed_idx).
        -> int32 [batch, k].
        `item_range=(lo, hi)` ranks only that slice of the track catalogue (item-sharded inference; ids stay global).
        `on_device=True` skips the copy to the host: the lists stay in the device buffers "topk_idx" / "topk_score"
        (self.buffer(name) -> pointer) and None is returned.  `reuse_output=True` returns page-locked arrays owned by the
        model that the NEXT call overwrites (the runners consume a batch's candidates before asking for the next)."""
        xp, xv = _coo(x_positions, x_vals)
        if isinstance(seeds, tuple):
            # already CSR: (seed_ptr int32 [batch + 1], seed_idx int32 [nnz]) -- large batches skip the Python list walk
            seed_ptr = np.ascontiguousarray(seeds[0], dtype=np.int32)
            flat = np.ascontiguousarray(seeds[1], dtype=np.int32)
            if seed_ptr.shape[0] != self.n_batch + 1:
                raise ValueError("seed_ptr must have batch + 1 = %d entries" % (self.n_batch + 1))
        else:
            seed_ptr = np.zeros(self.n_batch + 1, np.int32)
            lens = [len(s) for s in seeds]
            seed_ptr[1:len(lens) + 1] = np.cumsum(lens)
            seed_ptr[len(lens) + 1:] = seed_ptr[len(lens)]
            flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int64).reshape(-1) for s in seeds])
                                        if lens and sum(lens) else np.zeros(0, np.int64))
            flat = np.clip(flat, -1, 2 ** 31 - 1).astype(np.int32)
        if reuse_output and not on_device:
            if not hasattr(self, "_pinned"):
                self._pinned = _PinnedPool()
            idx = self._pinned.get((self.n_batch, k), np.int32)
            sc = self._pinned.get((self.n_batch, k), np.float32) if return_scores else None
        else:
            idx = np.empty((self.n_batch, k), np.int32) if not on_device else None
            sc = np.empty((self.n_batch, k), np.float32) if (return_scores and not on_device) else None
        lo, hi = item_range if item_range is not None else (0, self.n_tracks)
        _lib.check(self._lib.dae_model_recommend_range(self._h, _ptr(xp), _ptr(xv), xp.shape[0], self.n_batch,
                                                       _ptr(seed_ptr), _ptr(flat), int(k), int(lo), int(hi), _ptr(idx),
                                                       _ptr(sc)))
        if on_device:
            return None
        return (idx, sc) if return_scores else idx

    def evaluate(self, x_positions, x_vals, seeds, answers, k=500):
        """`recommend` + `met.single_eval` of every playlist on the device (main_train.py:62-100): float64 [n, 3] =
        (r-precision, ndcg, recommended-songs clicks) for the n = len(answers) playlists of the batch; rows past n are
        padding of a short last batch.  Only 24 bytes per playlist cross PCIe."""
        xp, xv = _coo(x_positions, x_vals)
        seed_ptr, flat = _csr(seeds, self.n_batch)
        ans_ptr, ans = _csr(answers, self.n_batch, pad_value=-1)
        out = np.empty((self.n_batch, 3), np.float64)
        _lib.check(self._lib.dae_model_evaluate(self._h, _ptr(xp), _ptr(xv), xp.shape[0], self.n_batch, _ptr(seed_ptr),
                                                _ptr(flat), _ptr(ans_ptr), _ptr(ans), int(k), _ptr(out)))
        return out[:len(answers)]

    # ---- staged / asynchronous surface (bench, data-parallel trainer) -------------------
    def stage_batch(self, slot, x_positions, x_vals, y_positions, y_vals):
        xp, xv = _coo(x_positions, x_vals)
        yp, yv = _coo(y_positions, y_vals)
        _lib.check(self._lib.dae_model_stage_batch(self._h, slot, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(yp),
                                                   _ptr(yv), yp.shape[0], self.n_batch))

    def restage(self, slot):
        _lib.check(self._lib.dae_model_restage(self._h, slot))

    def train_step_staged(self, slot, keep_prob, input_keep_prob):
        _lib.check(self._lib.dae_model_train_step_staged(self._h, slot, float(keep_prob), float(input_keep_prob)))

    def backward_staged(self, slot, keep_prob, input_keep_prob, global_batch=0, row_offset=0):
        _lib.check(self._lib.dae_model_backward_staged(self._h, slot, float(keep_prob), float(input_keep_prob),
                                                       int(global_batch), int(row_offset)))

    def apply_adam(self):
        _lib.check(self._lib.dae_model_apply_adam(self._h))

    def sync_cost(self):
        cost = C.c_float()
        _lib.check(self._lib.dae_model_sync_cost(self._h, C.byref(cost)))
        return float(cost.value)

    def set_profiling(self, on):
        _lib.check(self._lib.dae_model_set_profiling(self._h, int(bool(on))))

    def phase_times(self):
        """{phase name: (total ms, launches timed)} accumulated since set_profiling(True)."""
        out = {}
        for k in range(self._lib.dae_model_phase_count()):
            ms = C.c_double(); n = C.c_int64()
            _lib.check(self._lib.dae_model_phase_time(self._h, k, C.byref(ms), C.byref(n)))
            out[self._lib.dae_model_phase_name(k).decode()] = (ms.value, n.value)
        return out

    def launch_count(self):
        return int(self._lib.dae_model_launch_count(self._h))

    def buffer(self, name):
        """(device pointer, n_elem, elem_size) of a named internal buffer (include/dae_b200.h)."""
        p = C.c_void_p(); n = C.c_int64(); s = C.c_int32()
        _lib.check(self._lib.dae_model_buffer(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value


class DAE(DAE_tied):
    """Untied DAE used by --dae, optionally initialised from a pickle (reference models/DAEs.py:114-150)."""

    tied = False

    def __init__(self, conf):
        DAE_tied.__init__(self, conf)
        self.initval_dir = conf.initval                   # DAEs.py:117

    def init_weight(self):
        # DAEs.py:120-135 ('NULL' may arrive joined under --dir, main.py:44)
        if self.initval_dir == "NULL" or str(self.initval_dir).endswith("NULL"):
            _lib.check(self._lib.dae_model_init_xavier(self._h, self.seed))
        else:
            with open(self.initval_dir, "rb") as f:
                emb = pickle.load(f)
            self.set_params(emb)


class DAE_title(DAE):
    """Constant DAE whose scores are mixed with a title model's scores (reference models/DAEs.py:153-201).

    The DAE weights come from conf.DAEval and are constants (DAEs.py:164-171): the device model holds
    no Adam state; it is created "frozen with targets" (trainable = 2) so the title branch can train
    against y.  The mixing (DAEs.py:159-162, :180), the loss on y_pred (:194-196) and the title
    variables live in the title model (models/title_models/Char_CNN.py), which is attached with
    `title_model.fit(self)`; `train_step` / `predict` / `recommend` here forward to it.
    """

    trainable = 2

    def __init__(self, conf, title_score=None):
        DAE_tied.__init__(self, conf)
        self.DAEval_dir = conf.DAEval                     # DAEs.py:156
        self.initval_dir = conf.DAEval
        self.title_score = title_score                    # the Char_CNN object (the reference passes its .output tensor)

    def train_step(self, x_positions, x_vals, y_positions, y_vals, keep_prob, input_keep_prob, titles=None,
                   titles_use=1.0, title_keep_prob=1.0):
        return self.title_score.train_step(self, x_positions, x_vals, titles, keep_prob, title_keep_prob,
                                           input_keep_prob, y_positions, y_vals, titles_use)

    def predict(self, x_positions, x_vals, titles=None, titles_use=1.0, tracks_only=False):
        if self.title_score is None or titles is None:    # no title branch attached: titles_use = 0 -> the DAE's own scores
            return DAE.predict(self, x_positions, x_vals, tracks_only)
        return self.title_score.predict(self, x_positions, x_vals, titles, titles_use, tracks_only)

    def recommend(self, x_positions, x_vals, seeds, titles=None, titles_use=1.0, k=500, return_scores=False, **kw):
        if self.title_score is None or titles is None:
            return DAE.recommend(self, x_positions, x_vals, seeds, k, return_scores, **kw)
        return self.title_score.recommend(self, x_positions, x_vals, titles, seeds, titles_use, k, return_scores)

    def evaluate(self, x_positions, x_vals, seeds, answers, titles=None, titles_use=1.0, k=500):
        if self.title_score is None or titles is None:
            return DAE.evaluate(self, x_positions, x_vals, seeds, answers, k)
        return self.title_score.evaluate(self, x_positions, x_vals, titles, seeds, answers, titles_use, k)
