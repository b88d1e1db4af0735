"""Synthetic Million-Playlist-Dataset generator in the reference's JSON schema.

Replaces the reference's offline preprocessing (utils/spotify_reader.py +
data_generator.py, out of scope as compute) as the *producer of
fixtures*: the files written here are consumed unchanged by the readers of
utils/data_reader.py (reference) and of this repo.

Schemas (reference file:line)
  train      : spotify_reader.py:93-104  keys is_title_normalize, max_title_len, num_char,
               track_total, track_count, track_uri2id, artist_uri2id,
               playlists=[[track ids],[artist ids],[25 title char ids, -1 padded]], class_divpnt
  test-<n>[r]: spotify_reader.py:203-205,286  playlists=[[seed trks],[seed arts],[title],[answers]]
               (the writer's 4-tuple), class_divpnt
  challenge_*: spotify_reader.py:314-321,369  keys max_title_len, num_char, in_order, num_tracks,
               num_items, id2uri, playlists=[[trks],[arts],[title],[is_name],pid]

Data model: track ids are popularity ranked (id 0 most popular, spotify_reader.py:63-64);
each playlist belongs to one of C latent clusters and draws most of its tracks from the
cluster's own Zipf(1) over the tracks with id = c (mod C), the rest from a global Zipf(1);
duplicates are allowed; every track has one fixed artist (spotify_reader.py:128-129) and a
playlist carries one artist entry per track entry (heavy duplication, :120-127).
Playlist length ~ clipped log-normal on [5,250] with mean ~66 (<=250 cap: :84).
"""
from __future__ import annotations

import json
import os

import numpy as np

MAX_TITLE_LEN = 25     # spotify_reader.py:9
NUM_CHAR = 41          # spotify_reader.py:10 (26 letters + 10 digits + 5 symbols)
GEN_SEED = 180610      # the reference's only fixed seed (spotify_reader.py:13)


class SynthMPD:
    def __init__(self, n_tracks, n_artists, n_clusters=32, seed=GEN_SEED, home_frac=0.75,
                 mean_len=66.0, min_len=5, max_len=250):
        self.T, self.A, self.C = int(n_tracks), int(n_artists), int(n_clusters)
        self.home_frac = home_frac
        self.min_len, self.max_len = min_len, max_len
        # log-normal with sigma 0.8; mu chosen so the clipped mean is ~ mean_len
        self.sigma = 0.8
        self.mu = np.log(mean_len) - 0.5 * self.sigma ** 2 + 0.08
        rng = np.random.default_rng(seed)
        # fixed artist of each track: Zipf(1) over artists, more popular tracks -> more popular artists
        u = rng.random(self.T)
        a = np.floor(np.exp(u * np.log(self.A + 1.0))).astype(np.int64) - 1
        self.track2artist = (self.T + np.clip(a, 0, self.A - 1)).astype(np.int64)
        self.rng = np.random.default_rng(seed + 1)

    # -- core sampler ----------------------------------------------------
    def _zipf(self, rng, n, size):
        """ids in [0,n) with P(k) ~ 1/(k+1)."""
        u = rng.random(size)
        return np.clip(np.floor(np.exp(u * np.log(n + 1.0))).astype(np.int64) - 1, 0, n - 1)

    def sample(self, n_playlists, rng=None):
        """-> (lengths[n], tracks_flat, artists_flat, clusters[n])"""
        rng = rng or self.rng
        L = np.exp(rng.normal(self.mu, self.sigma, n_playlists))
        L = np.clip(np.rint(L), self.min_len, self.max_len).astype(np.int64)
        total = int(L.sum())
        cl = rng.integers(0, self.C, n_playlists)
        cl_flat = np.repeat(cl, L)
        home = rng.random(total) < self.home_frac
        n_home = max(self.T // self.C, 1)
        r_home = self._zipf(rng, n_home, total) * self.C + cl_flat
        r_glob = self._zipf(rng, self.T, total)
        trk = np.where(home, np.minimum(r_home, self.T - 1), r_glob)
        art = self.track2artist[trk]
        return L, trk, art, cl

    def titles(self, clusters, rng=None):
        """25 char ids per playlist, -1 padded; the first chars encode the cluster so that a
        title model has signal (change_title2ixs: spotify_reader.py:28-37)."""
        rng = rng or self.rng
        n = len(clusters)
        out = np.full((n, MAX_TITLE_LEN), -1, dtype=np.int64)
        lens = rng.integers(3, MAX_TITLE_LEN + 1, n)
        body = rng.integers(0, NUM_CHAR, (n, MAX_TITLE_LEN))
        body[:, 0] = clusters % NUM_CHAR
        body[:, 1] = (clusters // NUM_CHAR) % NUM_CHAR
        body[:, 2] = (clusters * 7 + 3) % NUM_CHAR
        m = np.arange(MAX_TITLE_LEN)[None, :] < lens[:, None]
        out[m] = body[m]
        return out

    # -- batches straight in the reader's COO format (bench / large scale) ------------
    def coo_batch(self, batch, rng=None):
        """One train batch in data_reader.next_batch's return format (data_reader.py:48-54),
        built vectorised: trk_positions, art_positions, y_positions, titles, trk_val, art_val."""
        L, trk, art, cl = self.sample(batch, rng)
        rows = np.repeat(np.arange(batch, dtype=np.int64), L)
        trk_pos = np.stack([rows, trk], axis=1)
        art_pos = np.stack([rows, art], axis=1)
        y_pos = np.concatenate([trk_pos, art_pos], axis=0)
        titles = self.titles(cl, rng)
        return trk_pos, art_pos, y_pos, titles, np.ones(len(trk_pos), np.float32), np.ones(len(art_pos), np.float32)

    # -- JSON fixtures ----------------------------------------------------------
    def _ragged(self, L, flat):
        off = np.concatenate([[0], np.cumsum(L)])
        return [flat[off[i]:off[i + 1]].tolist() for i in range(len(L))]

    def train_dict(self, n_playlists):
        L, trk, art, cl = self.sample(n_playlists)
        titles = self.titles(cl)
        tr = self._ragged(L, trk); ar = self._ragged(L, art)
        playlists = [[t, a, ti.tolist()] for t, a, ti in zip(tr, ar, titles)]
        counts = np.bincount(trk, minlength=self.T)
        cdf = np.cumsum(np.maximum(counts, 1)) / np.maximum(counts, 1).sum()
        divpnt = [int(np.searchsorted(cdf, p)) for p in (0.3, 0.8, 0.9)]
        return {
            "is_title_normalize": True, "max_title_len": MAX_TITLE_LEN, "num_char": NUM_CHAR,
            "track_total": ["t%d" % i for i in range(self.T)], "track_count": counts.tolist(),
            "track_uri2id": {"t%d" % i: i for i in range(self.T)},
            "artist_uri2id": {"a%d" % i: self.T + i for i in range(self.A)},
            "playlists": playlists, "class_divpnt": divpnt,
        }

    # answer-count windows of the reference's test split (spotify_reader.py:231-242)
    _WINDOW = {0: (10, 50), 1: (9, 77), 5: (5, 95), 10: (30, 90), 25: (76, 10 ** 9), 100: (50, 10 ** 9)}

    def test_dict(self, n_playlists, n_seeds, shuffle=False, class_divpnt=(), enforce_window=True):
        out = []
        guard = 0
        while len(out) < n_playlists and guard < 200:
            guard += 1
            L, trk, art, cl = self.sample(max(n_playlists, 64))
            titles = self.titles(cl)
            tr = self._ragged(L, trk); ar = self._ragged(L, art)
            for t, a, ti in zip(tr, ar, titles):
                if len(t) <= n_seeds:
                    continue
                n_ans = len(t) - n_seeds
                lo, hi = self._WINDOW.get(n_seeds, (1, 10 ** 9))
                if enforce_window and not (lo <= n_ans <= hi):
                    continue
                if shuffle:
                    perm = self.rng.permutation(len(t))
                    t = [t[i] for i in perm]; a = [a[i] for i in perm]
                seeds_t = t[:n_seeds]; seeds_a = a[:n_seeds]
                answers = []
                for x in t[n_seeds:]:
                    if x not in seeds_t and x not in answers:
                        answers.append(x)
                if not answers:
                    continue
                out.append([seeds_t, seeds_a, ti.tolist(), answers])
                if len(out) == n_playlists:
                    break
        return {"playlists": out, "class_divpnt": list(class_divpnt)}

    def challenge_dict(self, n_playlists, n_seeds, in_order=True, with_title=True, pid0=1000000):
        L, trk, art, cl = self.sample(n_playlists)
        titles = self.titles(cl)
        tr = self._ragged(L, trk); ar = self._ragged(L, art)
        pls = []
        for i, (t, a, ti) in enumerate(zip(tr, ar, titles)):
            k = min(n_seeds, len(t))
            is_name = 1 if with_title else 0
            ixs = ti.tolist() if with_title else [-1] * MAX_TITLE_LEN
            pls.append([t[:k], a[:k], ixs, [is_name], pid0 + i])
        return {
            "max_title_len": MAX_TITLE_LEN, "num_char": NUM_CHAR, "in_order": bool(in_order),
            "num_tracks": self.T, "num_items": self.T + self.A,
            "id2uri": {str(i): "t%d" % i for i in range(self.T)}, "playlists": pls,
        }


def write_dataset(data_dir, n_tracks=5000, n_artists=1000, n_train=1000, n_test=100, n_challenge=64,
                  n_clusters=16, seed=GEN_SEED):
    """Write train, test-{0,1,5,10,25,100}, test-{25,100}r and two challenge files
    (the sets the shipped config.ini files ask for: */config.ini:13,35,47)."""
    os.makedirs(data_dir, exist_ok=True)
    g = SynthMPD(n_tracks, n_artists, n_clusters, seed)
    train = g.train_dict(n_train)
    with open(os.path.join(data_dir, "train"), "w") as f:
        json.dump(train, f)
    for n in (0, 1, 5, 10, 25, 100):
        with open(os.path.join(data_dir, "test-%d" % n), "w") as f:
            json.dump(g.test_dict(n_test, n, False, train["class_divpnt"], enforce_window=(n < 25)), f)
    for n in (25, 100):
        with open(os.path.join(data_dir, "test-%dr" % n), "w") as f:
            json.dump(g.test_dict(n_test, n, True, train["class_divpnt"], enforce_window=False), f)
    with open(os.path.join(data_dir, "challenge_inorder_0to1"), "w") as f:
        json.dump(g.challenge_dict(n_challenge, 1, True), f)
    with open(os.path.join(data_dir, "challenge_inorder_10to100"), "w") as f:
        json.dump(g.challenge_dict(n_challenge, 100, True), f)
    return g


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="synthetic MPD-shaped fixtures in the reference JSON schema")
    ap.add_argument("--out", default="./data")
    ap.add_argument("--tracks", type=int, default=5000)
    ap.add_argument("--artists", type=int, default=1000)
    ap.add_argument("--train", type=int, default=1000)
    ap.add_argument("--test", type=int, default=100)
    ap.add_argument("--challenge", type=int, default=64)
    a = ap.parse_args()
    write_dataset(a.out, a.tracks, a.artists, a.train, a.test, a.challenge)


def write_mpd_slices(out_dir, seed=180610, n_tracks=400, n_artists=150, train=(2, 60), test=(1, 70), n_challenge=48):
    """Raw Million-Playlist-Dataset slices in Spotify's own JSON layout (what utils/spotify_reader.py consumes):
    `<out>/train/mpd.slice.*.json`, `<out>/test/…` = {"playlists": [{"name", "pid", "tracks": [{"pos", "track_uri":
    "spotify:track:<id>", "artist_uri": "spotify:artist:<id>", …}]}]} and `<out>/challenge/challenge_set.json` (playlists
    with "num_samples", optionally without "name", in-order or randomly sampled seeds).  Small and adversarial: Zipf
    popularity (so that minimum counts cut the vocabulary), the "various artists" id, punctuation / upper case / unknown
    characters in titles, playlists longer than 250 tracks, tracks that only occur in the test slices."""
    import json
    rng = np.random.default_rng(seed)
    words = ["chill", "WORKOUT!!", "Road-Trip", "90s (rock)", "study;time", "Party_2018", "sleep", "Jazz & Soul", "país",
             "gym+run", "love <3", "summer.vibes", "#throwback", "Country", "k-pop", "lo/fi"]
    artist_of = rng.integers(0, n_artists, n_tracks + 40)
    def track(i, pos):
        a = artist_of[i]
        art = "0LyfQWJT6nXafLPZqxe9Of" if a == 0 else "A%05d" % a       # artist 0 = "various artists"
        return {"pos": pos, "track_uri": "spotify:track:T%04d" % i, "artist_uri": "spotify:artist:" + art}   # (the fields read)
    p = 1.0 / np.arange(1, n_tracks + 1)
    p /= p.sum()
    def playlist(pid, length, extra_unseen=0):
        ids = rng.choice(n_tracks, size=length, p=p).tolist()
        ids += (n_tracks + rng.integers(0, 40, extra_unseen)).tolist()     # tracks outside the training slices
        rng.shuffle(ids)
        name = " ".join(rng.choice(words, size=int(rng.integers(1, 4))).tolist())
        return {"name": name, "pid": pid, "num_tracks": len(ids), "tracks": [track(i, k) for k, i in enumerate(ids)]}
    pid = 0
    for split, (n_files, per) in (("train", train), ("test", test)):
        os.makedirs(os.path.join(out_dir, split), exist_ok=True)
        for f in range(n_files):
            pls = []
            for _ in range(per):
                length = int(np.clip(rng.lognormal(3.6, 0.8), 3, 300)) if split == "train" else int(rng.integers(8, 140))
                pls.append(playlist(pid, length, extra_unseen=int(rng.integers(0, 3)) if split == "test" else 0))
                pid += 1
            with open(os.path.join(out_dir, split, "mpd.slice.%d-%d.json" % (f * per, f * per + per - 1)), "w") as fh:
                json.dump({"info": {"slice": "%d-%d" % (f * per, f * per + per - 1)}, "playlists": pls}, fh, separators=(",", ":"))
    os.makedirs(os.path.join(out_dir, "challenge"), exist_ok=True)
    pls = []
    for c in range(n_challenge):
        n_seed = int(rng.choice([0, 1, 5, 10, 25, 100]))
        full = playlist(1000000 + c, n_seed + int(rng.integers(5, 60)), extra_unseen=int(rng.integers(0, 2)))
        in_order = bool(c % 3) or n_seed < 25
        keep = list(range(n_seed)) if in_order else sorted(rng.choice(len(full["tracks"]), n_seed, replace=False).tolist())
        pl = {"pid": full["pid"], "num_holdouts": len(full["tracks"]) - n_seed, "num_tracks": len(full["tracks"]),
              "num_samples": n_seed, "tracks": [full["tracks"][k] for k in keep]}
        if c % 5:
            pl["name"] = full["name"]
        pls.append(pl)
    with open(os.path.join(out_dir, "challenge", "challenge_set.json"), "w") as fh:
        json.dump({"playlists": pls}, fh, separators=(",", ":"))
