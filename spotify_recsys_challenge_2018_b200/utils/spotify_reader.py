"""MPD slices -> the JSON files the readers consume: host mirror of the reference's utils/spotify_reader.py.

Offline, one-shot string / JSON work: it stays on the host.  Same class names and constructor signatures as the
reference (`Spotify_train`, `Spotify_test`, `Spotify_challenge`), same output files and schemas (the fixture format of
utils/data_reader.py), pinned against the reference's own classes executed on a small synthetic MPD
(tests/golden/make_golden.py --mpd -> tests/golden/mpd*, tests/test_host.py).

What is restated (file:line in the reference):
  normalize_name / change_title2ixs                      spotify_reader.py:21-37
  vocabulary: popularity-ranked ids, tracks 0..T-1, artists T..T+A-1, minimum counts     :63-70, :135-144
  playlist filter (empty in both modalities, or > 250 entries in either)                 :82-85
  popularity class cut points                                                              :74-75, :157-174
  test sets: first-n seeds / remaining answers, optional shuffle, length windows          :215-286
  challenge sets: in-order / random split by `pos`, seed-count buckets                     :343-369

Deliberate differences from the snapshot:
  * D11: `Spotify_test` upstream reads `self.class_divpnt` / `self.get_class`, which nothing defines, so it cannot run;
    the per-class lists it would build are never written anyway.  Here the class cut points come from the train file and
    the record is the writer's own 4-tuple `[seed_tracks, seed_artists, title_ixs, answers]`.
  * `create_uri2id` upstream cuts the vocabulary at `count_list.index(min_count - 1)` and raises when no item has
    exactly that count; here the cut is the first count < min_count (the same cut whenever upstream's exists).
  * a missing "various artists" id is not an error.
"""
from __future__ import annotations

import json
import os
import random
import re
from collections import Counter

random.seed(180610)                                           # spotify_reader.py:13 (test-set shuffling)

VARIOUS_ARTISTS_URI = "0LyfQWJT6nXafLPZqxe9Of"                # :15
MAX_TITLE_LEN = 25                                            # :16
CHARS = "abcdefghijklmnopqrstuvwxyz/<>+-1234567890"           # :17
CHAR2IX = {ch: i for i, ch in enumerate(CHARS)}
NUM_CHAR = len(CHAR2IX)
_PUNCT = re.compile(r"[.,#!$%\^\*;:{}=\_`~()@]")
_SPACE = re.compile(r"\s+")


def normalize_name(name):
    """lower-case, punctuation -> space, runs of white space collapsed (:21-25)."""
    return _SPACE.sub(" ", _PUNCT.sub(" ", name.lower())).strip()


def change_title2ixs(title):
    """The first MAX_TITLE_LEN known characters as ids, -1 padded (:28-37)."""
    ixs = [CHAR2IX[ch] for ch in title if ch in CHAR2IX][:MAX_TITLE_LEN]
    return ixs + [-1] * (MAX_TITLE_LEN - len(ixs))


def _uri(spotify_uri):
    return spotify_uri.split(":")[2]


def _slices(fullpaths):
    for path in fullpaths:
        with open(path) as f:
            yield from json.load(f)["playlists"]


def _dump(obj, path):
    with open(path, "w") as f:
        json.dump(obj, f, indent="\t")


def _title_ixs(playlist, normalize):
    name = playlist["name"]
    return change_title2ixs(normalize_name(name) if normalize else name)


def build_vocabulary(histogram, min_count, start_from):
    """Popularity-ranked vocabulary (:135-144): every uri by count descending (ties: first seen first, the order of
    Counter.most_common), the counts of the kept ones, and uri -> id for those seen at least `min_count` times."""
    ranked = sorted(histogram.items(), key=lambda kv: -kv[1])            # stable: insertion order within a count
    uris = [u for u, _ in ranked]
    counts = [c for _, c in ranked]
    keep = len(counts)
    if min_count > 1:
        keep = next((i for i, c in enumerate(counts) if c < min_count), len(counts))
    return uris, counts[:keep], {u: start_from + i for i, u in enumerate(uris[:keep])}


def class_cut_points(counts, points=(0.3, 0.8, 0.9)):
    """Ids at which the cumulative share of track occurrences passes each point (:157-174): for every point, scanning
    on from the previous cut, the id BEFORE the first one whose cumulative share exceeds it.  (A cut of -1 -- the most
    popular track alone exceeds the point -- makes the next scan start at index -1, i.e. at the LAST id, as Python's
    negative indexing does upstream; kept, so that tiny vocabularies give the reference's numbers.)"""
    total = float(sum(counts))
    cdf, run = [], 0
    for c in counts:
        run += c
        cdf.append(run / total)
    cuts, start = [], 0
    for p in points:
        for i in range(start, len(cdf)):
            if cdf[i] > p:
                cuts.append(i - 1)
                start = i - 1
                break
    return cuts


class Spotify_train:
    """MPD slices -> `<save_dir>/train` (:41-131)."""

    def __init__(self, train_fullpaths, trk_min_count, art_min_count, is_title_normalize, save_dir):
        self.is_title_normalize = is_title_normalize
        titles, tracks, artists = [], [], []
        trk_hist, art_hist = Counter(), Counter()
        for pl in _slices(train_fullpaths):
            name = pl["name"]
            titles.append(normalize_name(name) if is_title_normalize else name)
            t = [_uri(x["track_uri"]) for x in pl["tracks"]]
            a = [_uri(x["artist_uri"]) for x in pl["tracks"]]
            trk_hist.update(t)
            art_hist.update(a)
            tracks.append(t)
            artists.append(a)
        art_hist.pop(VARIOUS_ARTISTS_URI, None)                               # :66
        track_total, track_count, track_uri2id = build_vocabulary(trk_hist, trk_min_count, 0)
        _, _, artist_uri2id = build_vocabulary(art_hist, art_min_count, len(track_uri2id))

        playlists = []
        for t, a, title in zip(tracks, artists, titles):
            t_ids = [track_uri2id[u] for u in t if u in track_uri2id]
            a_ids = [artist_uri2id[u] for u in a if u in artist_uri2id]
            if (not t_ids and not a_ids) or len(t_ids) > 250 or len(a_ids) > 250:      # :82-85
                continue
            playlists.append([t_ids, a_ids, change_title2ixs(title)])
        self.num_playlists = len(playlists)
        os.makedirs(save_dir, exist_ok=True)
        _dump({"is_title_normalize": is_title_normalize, "max_title_len": MAX_TITLE_LEN, "num_char": NUM_CHAR,
               "track_total": track_total, "track_count": track_count, "track_uri2id": track_uri2id,
               "artist_uri2id": artist_uri2id, "playlists": playlists,
               "class_divpnt": class_cut_points(track_count)}, os.path.join(save_dir, "train"))


# length windows of the answer part per seed count (:216-228): (min, max) inclusive, None = unbounded
_ANSWER_WINDOW = {0: (10, 50), 1: (9, 77), 5: (5, 95), 10: (30, 90), 25: (76, None), 100: (50, None)}


class Spotify_test:
    """MPD slices + the train vocabulary -> `<save_dir>/test-<n>[r]` (:177-286, with D11 resolved as documented above)."""

    def __init__(self, test_fullpaths, train_json, test_seeds_num, save_dir, is_shuffle):
        with open(train_json) as f:
            train = json.load(f)
        trk2id, art2id = train["track_uri2id"], train["artist_uri2id"]
        seen = set(train["track_total"])
        normalize = bool(train["is_title_normalize"])
        n = test_seeds_num
        self.playlists = []
        for pl in _slices(test_fullpaths):
            # tracks that never occur in the training set are dropped; known ones below the minimum count become -1 (:203-211)
            pairs = [(trk2id.get(_uri(x["track_uri"]), -1), art2id.get(_uri(x["artist_uri"]), -1))
                     for x in pl["tracks"] if _uri(x["track_uri"]) in seen]
            n_ans = len(pairs) - n
            if n_ans <= 0:
                continue
            lo, hi = _ANSWER_WINDOW.get(n, (None, None))
            if (lo is not None and n_ans < lo) or (hi is not None and n_ans > hi):
                continue
            if is_shuffle:
                order = list(range(len(pairs)))
                random.shuffle(order)                                          # :232-233
                pairs = [pairs[i] for i in order]
            seed_trk = [t for t, _ in pairs[:n] if t != -1]
            seed_art = [a for _, a in pairs[:n] if a != -1]
            answers = []
            for t, _ in pairs[n:]:                                             # unique, not a seed; every -1 is kept (:252-258)
                if t not in seed_trk and (t == -1 or t not in answers):
                    answers.append(t)
            self.playlists.append([seed_trk, seed_art, _title_ixs(pl, normalize), answers])
        self.num_playlists = len(self.playlists)
        _dump({"playlists": self.playlists, "class_divpnt": train.get("class_divpnt", [])},
              os.path.join(save_dir, "test-%d%s" % (n, "r" if is_shuffle else "")))


class Spotify_challenge:
    """Challenge-set slices + the train vocabulary -> `<save_dir>/challenge_{inorder|random}_<a>[to<b>]` (:289-369)."""

    def __init__(self, challenge_fullpaths, train_json, save_dir, num_trk_lst, in_order):
        with open(train_json) as f:
            train = json.load(f)
        trk2id, art2id = train["track_uri2id"], train["artist_uri2id"]
        normalize = bool(train["is_title_normalize"])
        self.playlists = []
        for pl in _slices(challenge_fullpaths):
            last_pos = pl["tracks"][-1]["pos"] if pl["tracks"] else -1
            if (last_pos + 1 == pl["num_samples"]) != bool(in_order) or pl["num_samples"] not in num_trk_lst:   # :349-351
                continue
            t_ids = [trk2id[_uri(x["track_uri"])] for x in pl["tracks"] if _uri(x["track_uri"]) in trk2id]
            a_ids = [art2id[_uri(x["artist_uri"])] for x in pl["tracks"] if _uri(x["artist_uri"]) in art2id]
            has_name = "name" in pl
            ixs = _title_ixs(pl, normalize) if has_name else [-1] * MAX_TITLE_LEN
            self.playlists.append([t_ids, a_ids, ixs, [int(has_name)], pl["pid"]])
        self.num_playlists = len(self.playlists)
        os.makedirs(save_dir, exist_ok=True)
        span = "%d" % num_trk_lst[0] if len(num_trk_lst) == 1 else "%dto%d" % (num_trk_lst[0], num_trk_lst[-1])
        _dump({"max_title_len": MAX_TITLE_LEN, "num_char": NUM_CHAR, "in_order": in_order, "num_tracks": len(trk2id),
               "num_items": len(trk2id) + len(art2id), "id2uri": {v: k for k, v in trk2id.items()},
               "playlists": self.playlists},
              os.path.join(save_dir, "challenge_%s_%s" % ("inorder" if in_order else "random", span)))
