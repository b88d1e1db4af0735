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


def _block(rows_ids):
    """[(row, ids), ...] -> int64 [nnz, 2] in order."""
    lens = np.fromiter((len(ids) for _, ids in rows_ids), dtype=np.int64, count=len(rows_ids))
    total = int(lens.sum())
    out = np.empty((total, 2), dtype=np.int64)
    if total:
        out[:, 0] = np.repeat(np.fromiter((r for r, _ in rows_ids), dtype=np.int64, count=len(rows_ids)), lens)
        out[:, 1] = np.concatenate([np.asarray(ids, dtype=np.int64).reshape(-1) for _, ids in rows_ids])
    return out


def _load(data_dir, filename):
    with open(os.path.join(data_dir, filename)) as f:
        return json.load(f)


class data_reader:
    """Whole-playlist reader (reference utils/data_reader.py:7-54)."""

    def __init__(self, data_dir, filename, batch_size):
        d = _load(data_dir, filename)
        self.num_tracks = len(d["track_uri2id"])
        self.num_items = self.num_tracks + len(d["artist_uri2id"])
        self.max_title_len = d["max_title_len"]
        self.num_char = d["num_char"]
        self.playlists = d["playlists"]
        self.class_divpnt = d.get("class_divpnt", [])
        self.batch_size = batch_size
        self.train_idx = 0

    def _advance(self):
        self.train_idx += 1
        if self.train_idx == len(self.playlists):          # data_reader.py:44-46
            self.train_idx = 0
            random.shuffle(self.playlists)

    def next_batch(self):
        trk, art, titles = [], [], []
        for i in range(self.batch_size):
            t, a, title = self.playlists[self.train_idx]
            trk.append((i, t)); art.append((i, a)); titles.append(title)
            self._advance()
        trk_positions = _block(trk)
        art_positions = _block(art)
        y_positions = np.concatenate((trk_positions, art_positions), 0)       # data_reader.py:50
        return (trk_positions, art_positions, y_positions, titles,
                [1] * len(trk_positions), [1] * len(art_positions))


class data_reader_firstN(data_reader):
    """Reader that marks only the first `given_num` entries of each modality as input
    (value 1, the rest 0; reference utils/data_reader.py:57-128)."""

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

    def next_batch(self):
        trk, art, titles = [], [], []
        trk_val, art_val = [], []
        for i in range(self.batch_size):
            t, a, title = self.playlists[self.train_idx]
            if len(t) != 0:
                g = self._given(len(t))
                trk.append((i, t))
                trk_val += [1] * g + [0] * (len(t) - g)
            if len(a) != 0:
                g = self._given(len(a))
                art.append((i, a))
                art_val += [1] * g + [0] * (len(a) - g)
            titles.append(title)
            self._advance()
        trk_positions = _block(trk)
        art_positions = _block(art)
        y_positions = np.concatenate((trk_positions, art_positions), 0)
        return trk_positions, art_positions, y_positions, titles, trk_val, art_val


class data_reader_test:
    """Held-out reader (reference utils/data_reader.py:131-254).  Records are the writer's
    [seed_trks, seed_arts, title_ixs, answers] (spotify_reader.py:286); the 5-tuple form the committed
    reader unpacks (seed, seed_art, answer, seed_cls, answer_cls; data_reader.py:158) is accepted too."""

    def __init__(self, data_dir, filename, batch_size, test_num):
        d = _load(data_dir, filename)
        self.playlists = d["playlists"][:test_num]
        self.batch_size = batch_size
        self.test_idx = 0

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
        (data_reader.py:251-254)."""
        trk, art, seeds, answers, titles = [], [], [], [], []
        for i in range(self.batch_size):
            seed, seed_art, title, answer = self._unpack(self.playlists[self.test_idx])
            trk.append((i, seed)); art.append((i, seed_art))
            seeds.append(seed); answers.append(answer); titles.append(title)
            self.test_idx += 1
            if self.test_idx == len(self.playlists):           # data_reader.py:186-188
                self.test_idx = 0
                break
        trk_positions = _block(trk)
        if not with_artists:
            return trk_positions, seeds, answers, titles, [1] * len(trk_positions)
        art_positions = _block(art)
        x_positions = np.concatenate((trk_positions, art_positions), 0)
        return x_positions, seeds, answers, titles, [1] * len(trk_positions) + [0.5] * len(art_positions)


class data_reader_challenge:
    """Challenge-set reader (reference utils/data_reader.py:257-319)."""

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

    def next_batch(self):
        trk, art = [], []
        trk_ones = []
        ch_seed, ch_titles, ch_titles_exist, ch_pid = [], [], [], []
        for i in range(self.batch_size):
            seed, seed_art, title, title_exist, pid = self.playlists[self.ch_idx]
            n = len(seed)
            if n > 50 and self.is_in_order:                    # data_reader.py:288-291
                trk_ones += [0.15] * (n - 15) + [1.0] * 15
            else:
                trk_ones += [1.0] * n
            trk.append((i, seed)); art.append((i, seed_art))
            ch_seed.append(seed); ch_titles.append(title); ch_titles_exist.append(title_exist); ch_pid.append(pid)
            self.ch_idx += 1
            if self.ch_idx == len(self.playlists):             # data_reader.py:309-311
                self.ch_idx = 0
                break
        trk_positions = _block(trk)
        art_positions = _block(art)
        x_positions = np.concatenate((trk_positions, art_positions), 0)
        x_ones = trk_ones + [0.5] * len(art_positions)         # data_reader.py:317
        return x_positions, ch_seed, ch_titles, ch_titles_exist, ch_pid, x_ones
