"""Generate golden vectors by running the REFERENCE's own Python code (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Everything pure-Python/NumPy on the hot path's host side is executed from
/root/reference and its outputs are frozen here as small fixtures, because
/root/reference does not exist on the GPU box:

  metrics_golden.json   utils/metrics.py get_r_precision (:20-27), get_ndcg (:29-42), get_rsc (:44-49)
  ranking_golden.npz    main_runner/main_challenge.py cand_generate (:26-41)  (tensorflow/pandas stubbed:
                        they are imported at module level but not used by cand_generate)
  reader_golden.json    utils/data_reader.py data_reader / data_reader_firstN / data_reader_challenge
                        next_batch outputs on tests/golden/data (np.int -> int shim for numpy>=1.24)
  conf_golden.json      main.Conf parsed from the four shipped config.ini files (main_runner stubbed)

  merge_results_golden.csv / merge_results_input.pkl   the reference's merge_results.py run (subprocess, cwd = a temp
                        dir) on one small challenge pickle: `python tests/golden/make_golden.py --merge`

TensorFlow-1 arithmetic (models/DAEs.py) cannot be executed here; see oracle/__init__.py.
"""
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def make_merge_golden():
    """results.csv of the reference's own merge_results.py for one small challenge pickle."""
    import pickle
    import shutil
    import subprocess
    import tempfile
    d = tempfile.mkdtemp()
    os.makedirs(os.path.join(d, "challenge_results"))
    rows = [[1000000 + i] + ["spotify:track:t%d" % ((i * 7 + j) % 900) for j in range(500)] for i in range(3)]
    with open(os.path.join(d, "challenge_results", "result_a"), "wb") as f:
        pickle.dump(rows, f)
    subprocess.run([sys.executable, os.path.join(REF, "merge_results.py")], cwd=d, check=True, capture_output=True)
    shutil.copy(os.path.join(d, "results.csv"), os.path.join(HERE, "merge_results_golden.csv"))
    with open(os.path.join(HERE, "merge_results_input.pkl"), "wb") as f:
        pickle.dump(rows, f)


def main():
    assert os.path.isdir(REF), "reference not mounted"
    if "--merge" in sys.argv:
        make_merge_golden()
        return
    if not hasattr(np, "int"):
        np.int = int
    sys.path.insert(0, REF)
    _stub("tensorflow"); _stub("pandas")
    import utils.metrics as met                            # reference
    import utils.data_reader as rdr                        # reference
    import main_runner.main_challenge as mch               # reference (tf/pandas stubbed)

    rng = np.random.default_rng(20181007)

    # ---- metrics -------------------------------------------------------------
    cases = []
    for i in range(40):
        T = 400
        n_ans = int(rng.integers(1, 60))
        answer = rng.choice(T, n_ans, replace=False).tolist()
        if i % 5 == 0:
            answer[int(rng.integers(0, n_ans))] = -1       # below-min-count answers (spotify_reader.py:225)
        cand = rng.permutation(T)[:int(rng.integers(max(n_ans, 5), 200))].tolist()
        if i % 7 == 0:                                     # force some hits
            cand[:3] = [a for a in answer if a != -1][:3] + cand[len([a for a in answer if a != -1][:3]):3]
        cases.append(dict(answer=answer, cand=cand,
                          r_precision=met.get_r_precision(answer, cand, [], []),
                          ndcg=met.get_ndcg(answer, cand), rsc=met.get_rsc(answer, cand)))
    with open(os.path.join(HERE, "metrics_golden.json"), "w") as f:
        json.dump(cases, f)

    # ---- ranking: cand_generate -------------------------------------------------
    T, K = 3000, 500
    id2uri = {str(i): str(i) for i in range(T)}
    scores, seeds, cands = [], [], []
    for i in range(12):
        s = rng.permutation(T).astype(np.float32) / np.float32(T)      # distinct -> unstable sort irrelevant
        n_seed = [0, 1, 5, 10, 25, 100, 250, 3, 7, 50, 2, 0][i]
        sd = rng.choice(T, n_seed, replace=False).tolist() if n_seed else []
        if i == 7:
            sd = sd + sd[:2]                                           # duplicate seeds
        if i == 8:
            sd = sd + [T + 5, -1]                                      # absent seeds (artists / -1)
        out = mch.cand_generate(s, sd, id2uri)
        ids = [int(u.split(":")[2]) for u in out]
        scores.append(s); seeds.append(sd); cands.append(ids)
    np.savez_compressed(os.path.join(HERE, "ranking_golden.npz"), scores=np.stack(scores),
                        cands=np.array(cands, dtype=np.int32),
                        seeds=np.array([json.dumps(s) for s in seeds]))

    # ---- readers -----------------------------------------------------------------
    from tools.synth_mpd import SynthMPD
    data_dir = os.path.join(HERE, "data")
    os.makedirs(data_dir, exist_ok=True)
    g = SynthMPD(60, 12, n_clusters=4, seed=7, mean_len=8.0, min_len=1, max_len=20)
    train = g.train_dict(23)
    # edge rows the reference readers must handle: empty modality, duplicates
    train["playlists"][3][0] = []                                       # no tracks, artists only
    train["playlists"][5][1] = []                                       # no artists
    train["playlists"][7][0] = [0, 2, 0]; train["playlists"][7][1] = [63, 63, 64]
    with open(os.path.join(data_dir, "train"), "w") as f:
        json.dump(train, f)
    ch = g.challenge_dict(9, 100, in_order=True)
    ch["playlists"][2][0] = list(range(55)); ch["playlists"][2][1] = [60 + (i % 12) for i in range(55)]   # >50 seeds
    ch["playlists"][4][3] = [0]; ch["playlists"][4][2] = [-1] * 25
    with open(os.path.join(data_dir, "challenge_inorder_10to100"), "w") as f:
        json.dump(ch, f)

    def pack(out):
        res = []
        for o in out:
            if isinstance(o, np.ndarray):
                res.append(np.asarray(o, dtype=np.float64).reshape(-1, 2).astype(np.int64).tolist())
            else:
                res.append(o)
        return res

    golden = {}
    random.seed(1234)
    r = rdr.data_reader(data_dir, "train", 5)
    golden["data_reader"] = [pack(r.next_batch()) for _ in range(7)]      # crosses the epoch wrap + reshuffle
    for name, ft in (("firstN_frac", [0.0, 0.3]), ("firstN_abs", [1.0, 4.0])):
        random.seed(4321)
        # drop empty-track/artist rows? no: reference handles len==0 by skipping the modality
        r = rdr.data_reader_firstN(data_dir, "train", 5, ft)
        golden[name] = [pack(r.next_batch()) for _ in range(6)]
    r = rdr.data_reader_challenge(data_dir, "challenge_inorder_10to100", 4)
    outs = []
    while True:
        outs.append(pack(r.next_batch()))
        if r.ch_idx == 0:
            break
    golden["challenge"] = outs
    with open(os.path.join(HERE, "reader_golden.json"), "w") as f:
        json.dump(golden, f)

    # ---- Conf ---------------------------------------------------------------------
    _stub("main_runner", main_train=types.SimpleNamespace(), main_challenge=types.SimpleNamespace())
    import configparser
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_main", os.path.join(REF, "main.py"))
    ref_main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_main)
    conf_gold = {}
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for d in ("0to1_inorder", "5_inorder", "10to100_inorder", "25to100_random"):
                ini = configparser.ConfigParser()
                ini.read(os.path.join(REF, d, "config.ini"))
                per_mode = {}
                for mode in ("pretrain", "dae", "title", "challenge"):
                    c = ref_main.Conf(os.path.join(".", d), ini)
                    c.set_dae_conf()
                    if mode == "pretrain":
                        c.set_pretrain_conf()
                    elif mode == "dae":
                        c.set_dae_conf()
                    elif mode == "title":
                        c.set_title_conf()
                    else:
                        c.set_title_conf(); c.set_challenge_oonf()
                    per_mode[mode] = {k: v for k, v in vars(c).items() if k != "ini"}
                conf_gold[d] = per_mode
        finally:
            os.chdir(cwd)
    with open(os.path.join(HERE, "conf_golden.json"), "w") as f:
        json.dump(conf_gold, f, indent=1)
    # keep a copy of the shipped ini VALUES as fixtures? no: they are read from tests/golden/ini/*.ini,
    # written below from configparser (data, not code).
    ini_dir = os.path.join(HERE, "ini")
    os.makedirs(ini_dir, exist_ok=True)
    for d in ("0to1_inorder", "5_inorder", "10to100_inorder", "25to100_random"):
        ini = configparser.ConfigParser()
        ini.read(os.path.join(REF, d, "config.ini"))
        with open(os.path.join(ini_dir, d + ".ini"), "w") as f:
            ini.write(f)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
