"""CLI + Conf: host mirror of the reference's main.py (same flags, same config.ini schema).

    python -m spotify_recsys_challenge_2018_b200.main --dir D {--pretrain|--dae|--title|--challenge} [--testmode]

Conf attribute names and parsing follow reference main.py:12-95 (pinned by
tests/golden/conf_golden.json); booleans and title_kp are parsed properly.
Optional extra keys with defaults, so reference ini files load unchanged: [BASE] seed, device.
"""
from __future__ import annotations

import argparse
import configparser
import os


def _bool(s):
    return str(s).strip().lower() in ("1", "true", "yes", "on")


class Conf:
    def __init__(self, dir, ini):
        self.dir = dir
        self.ini = ini
        self.data_dir = ini.get("BASE", "data_dir")               # main.py:16
        self.result_dir = ini.get("BASE", "result_dir")           # main.py:17 (code wins over readme, D15)
        self.testsize = int(ini.get("BASE", "testsize"))          # main.py:18
        self.verbose = _bool(ini.get("BASE", "verbose"))          # main.py:19 (bool('False') bug fixed, D13)
        self.seed = int(ini.get("BASE", "seed", fallback="0"))
        self.device = int(ini.get("BASE", "device", fallback="0"))

    def set_dae_conf(self):                                       # main.py:21-47
        g = lambda k: self.ini.get("DAE", k)
        self.epochs = int(g("epochs"))
        self.batch = int(g("batch"))
        self.lr = float(g("lr"))
        self.reg_lambda = float(g("reg_lambda"))
        self.test_seed = ["test-" + s for s in g("test_seed").split(",")]
        self.update_seed = ["test-" + s for s in g("update_seed").split(",")]
        self.input_kp = [float(s) for s in g("input_kp").split(",")]
        self.kp = float(g("keep_prob"))
        self.firstN = [float(s) for s in g("firstN_range").split(",")]
        if len(self.firstN) == 1:                                 # main.py:33-43
            assert self.firstN[0] == -1.0
        else:
            assert self.firstN[0] <= self.firstN[1]
            if self.firstN[1] < 1:
                assert self.firstN[0] == 0 or self.firstN[0].is_integer() is False
            else:
                assert self.firstN[0] >= 1
                assert self.firstN[0].is_integer() is True and self.firstN[1].is_integer() is True
        self.initval = os.path.join(self.dir, g("initval"))
        self.save = os.path.join(self.dir, g("save"))
        self.hidden = int(g("hidden"))
        self.mode = "dae"

    def set_pretrain_conf(self):                                  # main.py:49-56
        g = lambda k: self.ini.get("PRETRAIN", k)
        self.epochs = int(g("epochs"))
        self.batch = int(g("batch"))
        self.lr = float(g("lr"))
        self.reg_lambda = float(g("reg_lambda"))
        self.is_pretrain = True
        self.save = os.path.join(self.dir, g("save"))
        self.mode = "pretrain"

    def set_title_conf(self):                                     # main.py:58-86
        g = lambda k: self.ini.get("TITLE", k)
        self.epochs = int(g("epochs"))
        self.batch = int(g("batch"))
        self.lr = float(g("lr"))
        self.input_kp = [float(s) for s in g("input_kp").split(",")]
        self.title_kp = float(g("title_kp"))                      # D14: the reference keeps a str
        self.test_seed = ["test-" + s for s in g("test_seed").split(",")]
        self.update_seed = ["test-" + s for s in g("update_seed").split(",")]
        self.char_emb = int(g("char_emb"))
        self.char_model = g("char_model")
        if self.char_model == "Char_CNN":
            self.filter_num = int(g("filter_num"))
            self.filter_size = [int(s) for s in g("filter_size").split(",")]
        elif self.char_model == "Char_LSTM":
            self.rnn_hidden = int(g("rnn_hidden"))
            self.bi = _bool(g("bi"))
        self.DAEval = os.path.join(self.dir, g("DAEval"))
        self.save = os.path.join(self.dir, g("save"))
        os.makedirs(os.path.dirname(self.save) or ".", exist_ok=True)
        self.mode = "title"

    def set_challenge_oonf(self):                                 # main.py:88-94 (name kept, typo included)
        os.makedirs(self.result_dir, exist_ok=True)
        self.challenge_data = self.ini.get("CHALLENGE", "challenge_data")
        self.result = os.path.join(self.result_dir, self.ini.get("CHALLENGE", "result"))
        self.batch = int(self.ini.get("CHALLENGE", "batch"))

    set_challenge_conf = set_challenge_oonf


def load_conf(dir):
    ini = configparser.ConfigParser()
    ini.read(os.path.join(dir, "config.ini"))
    return Conf(dir, ini)


def main(argv=None):
    ap = argparse.ArgumentParser(description="args")
    ap.add_argument("--dir", type=str, default="qwerty", help="directory name which contains config file")
    ap.add_argument("--pretrain", action="store_true", default=False, help="pretrain mode if Specified")
    ap.add_argument("--dae", action="store_true", default=False, help="DAE mode if Specified")
    ap.add_argument("--title", action="store_true", default=False, help="title mode if Specified")
    ap.add_argument("--challenge", action="store_true", default=False, help="challenge mode if Specified")
    ap.add_argument("--testmode", action="store_true", default=False, help="test mode if Specified(just check the result)")
    args = ap.parse_args(argv)
    dir = os.path.join(".", args.dir)
    if not os.path.isdir(dir):
        print("ERROR: Cannot find " + dir + " ->Create directory and config.ini file first")
        return 0
    if "config.ini" not in os.listdir(dir):
        print("ERROR: Cannot find config.ini in " + dir + " ->Create config.ini file in the directory first")
        return 0
    conf = load_conf(dir)
    conf.set_dae_conf()                                           # always first (main.py:121)
    from .main_runner import main_challenge, main_train
    if args.pretrain:
        conf.set_pretrain_conf()
        main_train.run(conf, args.testmode)
    elif args.dae:
        conf.set_dae_conf()
        main_train.run(conf, args.testmode)
    elif args.title:
        conf.set_title_conf()
        main_train.run(conf, args.testmode)
    elif args.challenge:
        conf.set_title_conf()
        conf.set_challenge_oonf()
        main_challenge.run(conf)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
