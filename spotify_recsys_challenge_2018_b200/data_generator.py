"""MPD -> train / test-* / challenge_* files: host mirror of the reference's data_generator.py with the defects of the
snapshot repaired: the `Spotify_train` call passes `is_title_normalize`, every test file the shipped configs
ask for is written (`test-{0,1,5,10,25,100}` in order and `test-{25,100}r` shuffled, */config.ini:13,35), and the
challenge files documented in readme.md:53-56 (`--mpd_ch`, `--divide_ch`) are produced.

    python -m spotify_recsys_challenge_2018_b200.data_generator --datadir ./data --mpd_tr ./mpd_train --mpd_te ./mpd_test \\
        --mpd_ch ./challenge
"""
from __future__ import annotations

import argparse
import os

from .utils.spotify_reader import Spotify_challenge, Spotify_test, Spotify_train


def fullpaths_generator(path):
    return [os.path.join(path, name) for name in sorted(os.listdir(path))]     # sorted: the vocabulary's tie order is reproducible


def parse_divide(spec):
    """'0-1,5,10-25,10-25r' -> [([0, 1], True), ([5], True), ([10, 25], True), ([10, 25], False)]: seed-count buckets of the
    challenge set, a trailing r = the randomly-ordered playlists (readme.md:56; 100-seed playlists join the 10-25 buckets
    as in the shipped configs' 10to100 names when the range end is 100)."""
    out = []
    for tok in spec.split(","):
        tok = tok.strip()
        in_order = not tok.endswith("r")
        lo, _, hi = tok.rstrip("r").partition("-")
        seeds = [int(lo)] if not hi else [n for n in (0, 1, 5, 10, 25, 100) if int(lo) <= n <= int(hi)]
        out.append((seeds, in_order))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="convert the MPD's format into the models' (readme.md:44-60)")
    ap.add_argument("--datadir", default="./data")
    ap.add_argument("--mpd_tr", default="./mpd_train")
    ap.add_argument("--mpd_te", default="./mpd_test")
    ap.add_argument("--mpd_ch", default="NULL")
    ap.add_argument("--mincount_trk", type=int, default=5)
    ap.add_argument("--mincount_art", type=int, default=3)
    ap.add_argument("--divide_ch", default="0-1,5,10-100,25-100r")
    ap.add_argument("--no_title_normalize", action="store_true")
    a = ap.parse_args(argv)
    Spotify_train(fullpaths_generator(a.mpd_tr), a.mincount_trk, a.mincount_art, not a.no_title_normalize, a.datadir)
    train_json = os.path.join(a.datadir, "train")
    if a.mpd_te != "NULL":
        paths = fullpaths_generator(a.mpd_te)
        for n in (0, 1, 5, 10, 25, 100):
            Spotify_test(paths, train_json, n, a.datadir, False)
        for n in (25, 100):
            Spotify_test(paths, train_json, n, a.datadir, True)
    if a.mpd_ch != "NULL":
        paths = fullpaths_generator(a.mpd_ch)
        for seeds, in_order in parse_divide(a.divide_ch):
            Spotify_challenge(paths, train_json, a.datadir, seeds, in_order)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
