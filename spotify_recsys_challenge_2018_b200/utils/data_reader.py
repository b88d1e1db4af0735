"""Sparse-batch readers: host mirror of the reference's utils/data_reader.py (same class names,
constructor arguments, return tuples and Python-`random` consumption, so a run with the same
`random.seed` yields the same batches as the reference -- pinned by tests/golden/reader_golden.json).

Format (the drop-in contract): `*_positions` = int64 [nnz, 2] with column 0 the row in
the batch and column 1 the item id (tracks 0..T-1, artists T..N-1), playlist order, duplicates
preserved; train `y_positions` / challenge `x_positions` are the track block of all rows followed
by the artist block of all rows.  Positions are always int64 here (the reference silently degrades
to float64 when a row is empty and relies on the feed to cast back).
"""
from __future__ import annotations

import json
import os
import random

import numpy as np


def _load(data_dir, filename):
    with open(os.path.join(data_dir, filename)) as f:
        return json.load(f)


def _ragged(lists, dtype=np.int32):
    """list of id lists -> (flat [total], ptr [n + 1])"""
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    flat = np.fromiter((v for x in lists for v in x), dtype=dtype, count=int(ptr[-1]))
    return flat, ptr


def _gather_rows(flat, ptr, ids):
    """The rows `ids` of a ragged array, concatenated: -> (values int64 [total], lens int64 [len(ids)])."""
    beg = ptr[ids]
    lens = ptr[ids + 1] - beg
    total = int(lens.sum())
    if total == 0:
        return np.empty(0, np.int64), lens
    starts = np.cumsum(lens) - lens
    idx = np.arange(total, dtype=np.int64) + np.repeat(beg - starts, lens)
    return flat[idx].astype(np.int64), lens


def _coo_block(vals, lens):
    """int64 [nnz, 2]: column 0 = row in the batch, column 1 = item id, playlist order."""
    out = np.empty((len(vals), 2), dtype=np.int64)
    out[:, 0] = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    out[:, 1] = vals
    return out


class data_reader:
    """Whole-playlist reader (reference utils/data_reader.py:7-54).

    Same batches, same Python-`random` consumption and the same epoch rule as the reference (a batch may straddle the
    wrap: the list is reshuffled the moment its last playlist has been consumed, data_reader.py:43-46) -- but assembled
    from a flat CSR copy of the playlists with a handful of NumPy calls instead of a per-playlist Python loop
    (at < 1 ms per device step the reference's loop would be the bottleneck).  The value vectors come back
    as float32 arrays (the reference returns Python lists of 1); `titles` is an int64 [batch, max_title_len] array."""

    def __init__(self, data_dir, filename, batch_size):
        d = _load(data_dir, filename)
        self.num_tracks = len(d["track_uri2id"])
        self.num_items = self.num_tracks + len(d["artist_uri2id"])
        self.max_title_len = d["max_title_len"]
        self.num_char = d["num_char"]
        pls = d["playlists"]
        self.class_divpnt = d.get("class_divpnt", [])
        self.batch_size = batch_size
        self.train_idx = 0
        self._n = len(pls)
        self._trk, self._trk_ptr = _ragged([p[0] for p in pls])
        self._art, self._art_ptr = _ragged([p[1] for p in pls])
        L = self.max_title_len
        self._titles = np.full((self._n, L), -1, dtype=np.int64)
        for i, p in enumerate(pls):
            t = p[2][:L]
            self._titles[i, :len(t)] = t
        # current order of the playlist list; random.shuffle on a list of the same length consumes the RNG exactly as the
        # reference's shuffle of the playlists themselves and yields the same permutation
        self._order_list = list(range(self._n))
        self._order = np.arange(self._n, dtype=np.int64)

    @property
    def playlists(self):
        """The playlists in their current order, as the reference keeps them: [[tracks], [artists], [title ids]]."""
        out = []
        for i in self._order_list:
            out.append([self._trk[self._trk_ptr[i]:self._trk_ptr[i + 1]].tolist(),
                        self._art[self._art_ptr[i]:self._art_ptr[i + 1]].tolist(), self._titles[i].tolist()])
        return out

    def __len__(self):
        return self._n

    def _take(self, count):
        """ids of the next `count` playlists, with the reference's wrap + reshuffle rule (data_reader.py:43-46)."""
        parts, need = [], count
        while need > 0:
            k = min(need, self._n - self.train_idx)
            parts.append(self._order[self.train_idx:self.train_idx + k])
            self.train_idx += k
            need -= k
            if self.train_idx == self._n:
                self.train_idx = 0
                random.shuffle(self._order_list)
                self._order = np.asarray(self._order_list, dtype=np.int64)
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def next_batch(self):
        ids = self._take(self.batch_size)
        trk, trk_lens = _gather_rows(self._trk, self._trk_ptr, ids)
        art, art_lens = _gather_rows(self._art, self._art_ptr, ids)
        trk_positions = _coo_block(trk, trk_lens)
        art_positions = _coo_block(art, art_lens)
        y_positions = np.concatenate((trk_positions, art_positions), 0)       # data_reader.py:50
        return (trk_positions, art_positions, y_positions, self._titles[ids],
                np.ones(len(trk_positions), np.float32), np.ones(len(art_positions), np.float32))


class data_reader_firstN(data_reader):
    """Reader that marks only the first `given_num` entries of each modality as input
    (value 1, the rest 0; reference utils/data_reader.py:57-128).  `given_num` is drawn with the reference's own
    `random.randrange` calls, in the reference's order (tracks then artists, per playlist)."""

    def __init__(self, data_dir, filename, batch_size, from_to):
        data_reader.__init__(self, data_dir, filename, batch_size)
        self.from_to = from_to

    def _given(self, n_items):
        f0, f1 = self.from_to[0], self.from_to[1]
        if f0 >= 1:                                            # data_reader.py:85-87
            n = int(f0)
            m = int(min(n_items, f1))
            n = min(n, m)      # the reference raises ValueError for playlists shorter than f0; clamp instead
        else:                                                  # data_reader.py:88-90
            n = int(max(n_items * f0, 1))
            m = int(max(n_items * f1, 1))
        return random.randrange(n, m + 1)                      # data_reader.py:91

    def _bounds(self, n_items):
        """Vector form of _given's range: -> (n, width = m + 1 - n), width 0 where the modality is empty."""
        f0, f1 = self.from_to[0], self.from_to[1]
        n_items = np.asarray(n_items, dtype=np.int64)
        if f0 >= 1:
            m = np.minimum(n_items.astype(np.float64), float(f1)).astype(np.int64)
            n = np.minimum(int(f0), m)
        else:
            n = np.maximum(n_items * float(f0), 1.0).astype(np.int64)
            m = np.maximum(n_items * float(f1), 1.0).astype(np.int64)
        width = np.where(n_items != 0, m + 1 - n, 0)
        return n, width

    def next_batch(self):
        # the draws must interleave with the epoch reshuffle exactly as in the reference (both use `random`): walk the
        # batch in runs that do not cross the wrap
        id_parts, gt_parts, ga_parts = [], [], []
        need = self.batch_size
        while need > 0:
            k = min(need, self._n - self.train_idx)
            ids = self._order[self.train_idx:self.train_idx + k]
            tl = self._trk_ptr[ids + 1] - self._trk_ptr[ids]
            al = self._art_ptr[ids + 1] - self._art_ptr[ids]
            # [n, m] of every draw at once (same float arithmetic as _given), then the draws themselves in the reference's
            # order with random.randrange's own algorithm (CPython _randbelow_with_getrandbits) on random.getrandbits
            lo_t, w_t = self._bounds(tl)
            lo_a, w_a = self._bounds(al)
            lo = np.stack([lo_t, lo_a], 1).reshape(-1).tolist()
            wd = np.stack([w_t, w_a], 1).reshape(-1).tolist()          # 0: empty modality, no draw
            g = [0] * (2 * k)
            getrandbits = random.getrandbits
            for j in range(2 * k):
                w = wd[j]
                if w > 0:
                    nb = w.bit_length()
                    r = getrandbits(nb)
                    while r >= w:
                        r = getrandbits(nb)
                    g[j] = lo[j] + r
            gt, ga = g[0::2], g[1::2]
            id_parts.append(ids); gt_parts.append(gt); ga_parts.append(ga)
            self.train_idx += k
            need -= k
            if self.train_idx == self._n:
                self.train_idx = 0
                random.shuffle(self._order_list)
                self._order = np.asarray(self._order_list, dtype=np.int64)
        ids = id_parts[0] if len(id_parts) == 1 else np.concatenate(id_parts)
        gt = np.asarray([g for part in gt_parts for g in part], dtype=np.int64)
        ga = np.asarray([g for part in ga_parts for g in part], dtype=np.int64)
        trk, trk_lens = _gather_rows(self._trk, self._trk_ptr, ids)
        art, art_lens = _gather_rows(self._art, self._art_ptr, ids)
        trk_positions = _coo_block(trk, trk_lens)
        art_positions = _coo_block(art, art_lens)
        y_positions = np.concatenate((trk_positions, art_positions), 0)

        def first_n(lens, given):          # 1 for the first given[r] entries of row r, 0 for the rest
            within = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
            return (within < np.repeat(given, lens)).astype(np.float32)
        return (trk_positions, art_positions, y_positions, self._titles[ids],
                first_n(trk_lens, gt), first_n(art_lens, ga))


class data_reader_test:
    """Held-out reader (reference utils/data_reader.py:131-254).  Records are the writer's
    [seed_trks, seed_arts, title_ixs, answers] (spotify_reader.py:286); the 5-tuple form the committed
    reader unpacks (seed, seed_art, answer, seed_cls, answer_cls; data_reader.py:158) is accepted too.
    The COO blocks come from a flat CSR copy of the seeds (as in data_reader); seeds / answers / titles are handed out
    as the stored lists, which is what metrics.single_eval and the rankers take."""

    def __init__(self, data_dir, filename, batch_size, test_num):
        d = _load(data_dir, filename)
        self.playlists = d["playlists"][:test_num]
        self.batch_size = batch_size
        self.test_idx = 0
        recs = [self._unpack(r) for r in self.playlists]
        self._seeds = [r[0] for r in recs]
        self._answers = [r[3] for r in recs]
        self._titles = [r[2] for r in recs]
        self._trk, self._trk_ptr = _ragged(self._seeds)
        self._art, self._art_ptr = _ragged([r[1] for r in recs])

    @staticmethod
    def _unpack(rec):
        if len(rec) == 4:
            seed, seed_art, title, answer = rec
        else:
            seed, seed_art, answer = rec[0], rec[1], rec[2]
            title = None
        return seed, seed_art, title, answer

    def next_batch_test(self, with_artists=False):
        """-> (x_positions, seeds, answers, titles, x_vals): the signature the eval loop unpacks
        (main_train.py:64).  Seed tracks weigh 1; with_artists adds the seed artists at 0.5
        (data_reader.py:251-254).  The last batch of a file is short (data_reader.py:186-188)."""
        lo = self.test_idx
        hi = min(lo + self.batch_size, len(self.playlists))
        ids = np.arange(lo, hi, dtype=np.int64)
        self.test_idx = 0 if hi == len(self.playlists) else hi
        trk, trk_lens = _gather_rows(self._trk, self._trk_ptr, ids)
        trk_positions = _coo_block(trk, trk_lens)
        seeds, answers, titles = self._seeds[lo:hi], self._answers[lo:hi], self._titles[lo:hi]
        if not with_artists:
            return trk_positions, seeds, answers, titles, np.ones(len(trk_positions), np.float32)
        art, art_lens = _gather_rows(self._art, self._art_ptr, ids)
        art_positions = _coo_block(art, art_lens)
        x_positions = np.concatenate((trk_positions, art_positions), 0)
        x_vals = np.concatenate((np.ones(len(trk_positions), np.float32), np.full(len(art_positions), 0.5, np.float32)))
        return x_positions, seeds, answers, titles, x_vals


class data_reader_challenge:
    """Challenge-set reader (reference utils/data_reader.py:257-319), COO blocks from a flat CSR copy of the seeds."""

    def __init__(self, data_dir, filename, batch_size):
        d = _load(data_dir, filename)
        self.playlists = d["playlists"]
        self.id2uri = d["id2uri"]
        self.num_tracks = d["num_tracks"]
        self.num_items = d["num_items"]
        self.is_in_order = d["in_order"]
        self.max_title_len = d["max_title_len"]
        self.num_char = d["num_char"]
        self.batch_size = batch_size
        self.ch_idx = 0
        self._trk, self._trk_ptr = _ragged([p[0] for p in self.playlists])
        self._art, self._art_ptr = _ragged([p[1] for p in self.playlists])

    def next_batch(self):
        lo = self.ch_idx
        hi = min(lo + self.batch_size, len(self.playlists))
        ids = np.arange(lo, hi, dtype=np.int64)
        self.ch_idx = 0 if hi == len(self.playlists) else hi                  # data_reader.py:309-311
        trk, trk_lens = _gather_rows(self._trk, self._trk_ptr, ids)
        art, art_lens = _gather_rows(self._art, self._art_ptr, ids)
        trk_positions = _coo_block(trk, trk_lens)
        art_positions = _coo_block(art, art_lens)
        # in-order playlists with more than 50 seeds: the last 15 tracks weigh 1.0, the earlier ones 0.15 (data_reader.py:288-291)
        trk_ones = np.ones(len(trk), np.float64)          # float64: 0.15 exactly as the reference's Python floats
        if self.is_in_order and len(trk):
            within = np.arange(len(trk), dtype=np.int64) - np.repeat(np.cumsum(trk_lens) - trk_lens, trk_lens)
            n_row = np.repeat(trk_lens, trk_lens)
            trk_ones[(n_row > 50) & (within < n_row - 15)] = 0.15
        recs = self.playlists[lo:hi]
        x_positions = np.concatenate((trk_positions, art_positions), 0)
        x_ones = np.concatenate((trk_ones, np.full(len(art_positions), 0.5, np.float64)))   # data_reader.py:317
        return (x_positions, [p[0] for p in recs], [p[2] for p in recs], [p[3] for p in recs], [p[4] for p in recs],
                x_ones)
