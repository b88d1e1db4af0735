"""Result merger: host mirror of the reference's merge_results.py (:6-23).

Concatenates the challenge pickles of a directory (rows [pid, 'spotify:track:<uri>' x 500], written by
main_runner/main_challenge.py) under the submission header row and writes `results.csv` -- through pandas exactly as
the reference does (a ragged frame: short rows are padded with empty fields), so the file is byte-identical
(tests/golden/merge_results_golden.csv, produced by the reference's own script).

    python -m spotify_recsys_challenge_2018_b200.merge_results --dir challenge_results
"""
from __future__ import annotations

import argparse
import os
import pickle

HEADER = ["team_info", "track", "team_name", "email@address.com"]      # merge_results.py:13


def merge(dir, out="results.csv", verbose=True):
    import pandas as pd
    total_cands = [list(HEADER)]
    for result in os.listdir(dir):                                      # merge_results.py:14-18 (directory order)
        with open(os.path.join(dir, result), "rb") as f:
            total_cands += pickle.load(f)
    if verbose:
        print("num_playlist: ", len(total_cands) - 1)                   # merge_results.py:20-21
        print("num_rec: ", len(total_cands[1]) - 1 if len(total_cands) > 1 else 0)
    pd.DataFrame(total_cands).to_csv(out, index=False, header=False)    # merge_results.py:22-23
    return len(total_cands) - 1


def main(argv=None):
    ap = argparse.ArgumentParser(description="args")
    ap.add_argument("--dir", type=str, default="challenge_results")
    args = ap.parse_args(argv)
    merge("./" + args.dir)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
