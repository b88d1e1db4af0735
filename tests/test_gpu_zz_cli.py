"""End to end through the reference's entry points (main.py:97-139): `--pretrain`, `--dae`, `--title`, `--challenge`
on a small synthetic dataset in the reference's JSON schema, with a config.ini in the shipped layout.  Checks what a
user of the reference relies on: the log lines, the saved pickles (shapes of models/DAEs.py:107-111), the title
checkpoint and the challenge result rows [pid, 'spotify:track:<uri>' x 500] (main_challenge.py:89-96)."""
import os
import pickle
import random

import numpy as np
import pytest

from spotify_recsys_challenge_2018_b200 import main as cli
from tools.synth_mpd import write_dataset

pytestmark = pytest.mark.gpu

INI = """[BASE]
verbose = False
data_dir = {data}
result_dir = {res}
testsize = 32

[DAE]
epochs = 1
batch = 64
lr = 0.005
reg_lambda = 0.0
hidden = 64
test_seed = 1,5
update_seed = 1
keep_prob = 0.8
input_kp = 0.5,0.8
firstn_range = 0.0,0.3
initval = w_pretrain
save = w_dae

[PRETRAIN]
epochs = 2
batch = 64
lr = 0.01
reg_lambda = 0.0
save = w_pretrain

[TITLE]
epochs = 1
batch = 64
lr = 0.001
keep_prob = 0.8
title_kp = 0.8
input_kp = 0.01
test_seed = 1,5
update_seed = 1
char_model = Char_CNN
filter_num = 16
filter_size = 3,5
char_emb = 50
daeval = w_dae
save = graph/model.ckpt

[CHALLENGE]
batch = 64
challenge_data = challenge_inorder_0to1
result = result_inorder_0to1
"""


def test_cli_pretrain_dae_title_challenge(tmp_path, monkeypatch):
    data = tmp_path / "data"
    write_dataset(str(data), n_tracks=1500, n_artists=200, n_train=400, n_test=32, n_challenge=70, n_clusters=8)
    run_dir = tmp_path / "run1"
    run_dir.mkdir()
    (run_dir / "config.ini").write_text(INI.format(data=str(data), res=str(tmp_path / "challenge_results")))
    monkeypatch.chdir(tmp_path)
    random.seed(0)
    np.random.seed(0)

    assert cli.main(["--dir", "run1", "--pretrain"]) == 0
    w = pickle.load(open(run_dir / "w_pretrain", "rb"))
    assert [a.shape for a in w] == [(1700, 64), (1700, 64), (64,), (1700,)]
    assert np.array_equal(w[0], w[1])                                         # tied (DAEs.py:107-111)
    log = (run_dir / "log.txt").read_text()
    assert "[pretrain mode]" in log and "epoch 2" in log and "rprecision:" in log and "training loss:" in log

    assert cli.main(["--dir", "run1", "--dae"]) == 0
    w2 = pickle.load(open(run_dir / "w_dae", "rb"))
    assert [a.shape for a in w2] == [(1700, 64), (1700, 64), (64,), (1700,)]
    assert not np.array_equal(w2[0], w2[1])                                   # untied after one epoch from the tied init
    assert all(np.isfinite(a).all() for a in w2)

    assert cli.main(["--dir", "run1", "--title"]) == 0
    assert os.path.exists(run_dir / "graph" / "model.ckpt")

    assert cli.main(["--dir", "run1", "--challenge"]) == 0
    rows = pickle.load(open(tmp_path / "challenge_results" / "result_inorder_0to1", "rb"))
    assert len(rows) == 70
    for r in rows:
        assert isinstance(r[0], int) and len(r) == 501
        assert all(isinstance(u, str) and u.startswith("spotify:track:t") for u in r[1:])
        assert len(set(r[1:])) == 500
    log = (run_dir / "log.txt").read_text()
    assert "[dae mode]" in log and "[title mode]" in log and "[challenge mode]" in log
