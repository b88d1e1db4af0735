"""Char_CNN over libdae_b200.so -- host mirror of the reference's models/title_models/Char_CNN.py.

Same constructor (`Char_CNN(config, conv_layers)`, Char_CNN.py:6-12) and configuration fields.  The
reference object is a TF graph fragment whose `.output` is handed to DAE_title; here the object owns
the device-side title branch (`dae_title_*` in include/dae_b200.h) that is attached to the constant
DAE with `fit(dae_model)`.  `sess.run([optimizer, cost], feed)` of title mode (main_train.py:214-221)
becomes `train_step`, `sess.run(y_pred, feed)` becomes `predict` / `recommend`.  No compute in Python.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np

from ... import _lib
from ..DAEs import _coo, _csr, _ptr


class Char_CNN:
    def __init__(self, config, conv_layers):
        self.config = config
        self.embedding = int(config.char_emb)                 # Char_CNN.py:8
        self.input_len = int(config.strmaxlen)                # Char_CNN.py:9
        self.output_dim = int(config.n_output)                # Char_CNN.py:10
        self.char_size = int(config.charsize)                 # Char_CNN.py:11
        self.conv_layers = conv_layers                        # [[filters, width, -1], ...]  Char_CNN.py:12
        if any(layer[-1] != -1 for layer in conv_layers):
            raise NotImplementedError("intermediate max-pooling (Char_CNN.py:53-56) is not used by title_get.get_model")
        if len({layer[0] for layer in conv_layers}) != 1:
            raise ValueError("every conv layer has filter_num filters (title_get.py:20)")
        self.filter_num = int(conv_layers[0][0])
        self.filter_size = [int(layer[1]) for layer in conv_layers]
        self.learning_rate = float(getattr(config, "lr", 0.001))
        self.seed = int(getattr(config, "seed", 0))
        self.trainable = bool(getattr(config, "title_trainable", True))
        self._h = None
        self._lib = None
        self._dae = None

    # ---- lifecycle -------------------------------------------------------------------
    def fit(self, dae_model):
        """Create the device-side branch on top of the constant DAE and initialise it (Char_CNN.py:19-73)."""
        lib = _lib.load()
        cfg = _lib.DaeTitleConfig(self.char_size, self.input_len, self.embedding, self.filter_num, len(self.filter_size),
                                  (C.c_int32 * 8)(*(self.filter_size + [0] * (8 - len(self.filter_size)))),
                                  self.learning_rate, int(self.trainable))
        h = C.c_void_p()
        _lib.check(lib.dae_title_create(dae_model._h, C.byref(cfg), C.byref(h)))
        self._h, self._lib, self._dae = h, lib, dae_model
        _lib.check(lib.dae_title_init(self._h, self.seed))
        return self

    def close(self):
        if self._h is not None:
            self._lib.dae_title_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------
    def _shapes(self):
        E, F = self.embedding, self.filter_num
        shapes = [(self.char_size, E)]
        for w in self.filter_size:
            shapes += [(w, E, F), (F,)]
        D = F * len(self.filter_size)
        return shapes + [(D, self.output_dim), (self.output_dim,)]

    def get_params(self):
        """[char_embedding, Conv_W0, Conv_b0, ..., Output_W, Output_b] in the reference's shapes."""
        arrs = [np.empty(s, np.float32) for s in self._shapes()]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        _lib.check(self._lib.dae_title_get_params(self._h, ptrs))
        return arrs

    def set_params(self, params):
        arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
        for a, s in zip(arrs, self._shapes()):
            if a.shape != s:
                raise ValueError("title parameter has shape %s, expected %s" % (a.shape, s))
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        _lib.check(self._lib.dae_title_set_params(self._h, ptrs))

    def save(self, path):
        """Stands in for tf.train.Saver().save(sess, conf.save) (main_train.py:247): a pickle of get_params()."""
        with open(path, "wb") as f:
            pickle.dump(self.get_params(), f)

    def restore(self, path):
        """saver.restore(sess, conf.save) (main_train.py:184, main_challenge.py:69)."""
        with open(path, "rb") as f:
            self.set_params(pickle.load(f))

    # ---- the hot path -------------------------------------------------------------------
    def _titles(self, titles, titles_use, batch):
        t = np.full((batch, self.input_len), -1, np.int64)     # short last batch: pad rows with -1 (main_challenge.py:74-78)
        tt = np.asarray(titles, dtype=np.int64).reshape(-1, self.input_len)
        t[:tt.shape[0]] = tt
        u = np.zeros(batch, np.float32)
        if np.isscalar(titles_use):
            u[:] = float(titles_use)                           # [[1]] * conf.batch (main_train.py:221)
        else:
            uu = np.asarray(titles_use, dtype=np.float32).reshape(-1)
            u[:uu.shape[0]] = uu
        return t, u

    def train_step(self, dae_model, x_positions, x_vals, titles, keep_prob, title_keep_prob, input_keep_prob,
                   y_positions=None, y_vals=None, titles_use=1.0):
        """One `sess.run([optimizer, cost])` of title mode (main_train.py:214-221): x = y = tracks + artists
        unless a separate target is given."""
        xp, xv = _coo(x_positions, x_vals)
        yp, yv = (xp, xv) if y_positions is None else _coo(y_positions, y_vals)
        B = dae_model.n_batch
        t, u = self._titles(titles, titles_use, B)
        cost = C.c_float()
        _lib.check(self._lib.dae_title_train_step(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(yp), _ptr(yv), yp.shape[0],
                                                  _ptr(t), _ptr(u), B, float(keep_prob), float(input_keep_prob),
                                                  float(title_keep_prob), C.byref(cost)))
        return float(cost.value)

    def train_step_async(self, dae_model, x_positions, x_vals, titles, keep_prob, title_keep_prob, input_keep_prob,
                         y_positions=None, y_vals=None, titles_use=1.0):
        """The same step pipelined: returns the PREVIOUS step's cost (None on the first call) while this one runs;
        `flush()` returns the last pending cost.  The runner only accumulates the cost (main_train.py:223)."""
        xp, xv = _coo(x_positions, x_vals)
        yp, yv = (xp, xv) if y_positions is None else _coo(y_positions, y_vals)
        B = dae_model.n_batch
        t, u = self._titles(titles, titles_use, B)
        cost = C.c_float(); has = C.c_int32()
        _lib.check(self._lib.dae_title_train_step_async(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(yp), _ptr(yv),
                                                        yp.shape[0], _ptr(t), _ptr(u), B, float(keep_prob),
                                                        float(input_keep_prob), float(title_keep_prob), C.byref(cost),
                                                        C.byref(has)))
        return float(cost.value) if has.value else None

    def flush(self):
        cost = C.c_float(); has = C.c_int32()
        _lib.check(self._lib.dae_title_train_flush(self._h, C.byref(cost), C.byref(has)))
        return float(cost.value) if has.value else None

    def predict(self, dae_model, x_positions, x_vals, titles, titles_use=1.0, tracks_only=False):
        """`sess.run(y_pred)` with every keep probability 1 (main_train.py:69-79, main_challenge.py:80-85)."""
        xp, xv = _coo(x_positions, x_vals)
        B = dae_model.n_batch
        t, u = self._titles(titles, titles_use, B)
        n_cols = dae_model.n_tracks if tracks_only else dae_model.n_input
        out = np.empty((B, n_cols), np.float32)
        _lib.check(self._lib.dae_title_predict(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(t), _ptr(u), B, n_cols, _ptr(out)))
        return out

    def recommend(self, dae_model, x_positions, x_vals, titles, seeds, titles_use=1.0, k=500, return_scores=False):
        """y_pred[:, :n_tracks] -> cand_generate (main_challenge.py:26-36, :87-90) on the device."""
        xp, xv = _coo(x_positions, x_vals)
        B = dae_model.n_batch
        t, u = self._titles(titles, titles_use, B)
        seed_ptr = np.zeros(B + 1, np.int32)
        lens = [len(s) for s in seeds]
        seed_ptr[1:len(lens) + 1] = np.cumsum(lens)
        seed_ptr[len(lens) + 1:] = seed_ptr[len(lens)]
        flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int64).reshape(-1) for s in seeds])
                                    if lens and sum(lens) else np.zeros(0, np.int64))
        flat = np.clip(flat, -1, 2 ** 31 - 1).astype(np.int32)
        idx = np.empty((B, k), np.int32)
        sc = np.empty((B, k), np.float32) if return_scores else None
        _lib.check(self._lib.dae_title_recommend(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(t), _ptr(u), B,
                                                 _ptr(seed_ptr), _ptr(flat), int(k), _ptr(idx), _ptr(sc)))
        return (idx, sc) if return_scores else idx

    def evaluate(self, dae_model, x_positions, x_vals, titles, seeds, answers, titles_use=1.0, k=500):
        """`recommend` + met.single_eval per playlist on the device -> float64 [len(answers), 3] (main_train.py:69-100)."""
        xp, xv = _coo(x_positions, x_vals)
        B = dae_model.n_batch
        t, u = self._titles(titles, titles_use, B)
        seed_ptr, flat = _csr(seeds, B)
        ans_ptr, ans = _csr(answers, B, pad_value=-1)
        out = np.empty((B, 3), np.float64)
        _lib.check(self._lib.dae_title_evaluate(self._h, _ptr(xp), _ptr(xv), xp.shape[0], _ptr(t), _ptr(u), B, _ptr(seed_ptr),
                                                _ptr(flat), _ptr(ans_ptr), _ptr(ans), int(k), _ptr(out)))
        return out[:len(answers)]

    def buffer(self, name):
        p = C.c_void_p(); n = C.c_int64(); s = C.c_int32()
        _lib.check(self._lib.dae_title_buffer(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value

    def launch_count(self):
        return int(self._lib.dae_title_launch_count(self._h))

    def set_profiling(self, on):
        """Per-phase device times of the title step are kept by the underlying DAE model object (phases "title_*")."""
        self._dae.set_profiling(on)

    def phase_times(self):
        return self._dae.phase_times()

    def __str__(self):                                          # Char_CNN.py:77-83
        return "\n".join(["Wide CNN", "Embedding Size : " + str(self.embedding),
                          "Number of Filters : " + str(self.filter_num), "Conv Layers : " + str(self.conv_layers)])
