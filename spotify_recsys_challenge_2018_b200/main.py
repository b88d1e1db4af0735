"""Command line + configuration object of the drop-in: same flags, same `config.ini` schema and the same `conf`
attribute names as the reference's main.py (:12-139), so its config directories and runners work unchanged.

    python -m spotify_recsys_challenge_2018_b200.main --dir D {--pretrain|--dae|--title|--challenge} [--testmode]

The schema is data here: one table per ini section says which key becomes which attribute through which parser
(`SCHEMA`), and `Conf.apply(section)` walks it.  `set_dae_conf` / `set_pretrain_conf` / `set_title_conf` /
`set_challenge_oonf` keep the reference's method names (typo included) because the runners and users call them.
Attribute values are pinned against the reference's own parser for the four shipped ini files
(tests/golden/conf_golden.json); the two places where the reference's parse is a bug are fixed: booleans
(`bool('False')` is True upstream) and `title_kp` (left a string upstream, D14).
Extra optional keys, so that reference ini files load unchanged: [BASE] seed, device.
"""
from __future__ import annotations

import argparse
import configparser
import os


def _flag(text):
    return str(text).strip().lower() in ("1", "true", "yes", "on")


def _csv(cast):
    return lambda text: [cast(tok) for tok in text.split(",")]


def _test_files(text):
    return ["test-" + tok for tok in text.split(",")]            # 'test-<n>[r]' files of the data directory


class _InDir:
    """marks a value that is a path relative to the run directory"""

    def __call__(self, text):
        return text


IN_DIR = _InDir()

# section -> [(ini key, conf attribute, parser)], in the reference's order (main.py:21-94)
SCHEMA = {
    "BASE": [("data_dir", "data_dir", str), ("result_dir", "result_dir", str), ("testsize", "testsize", int),
             ("verbose", "verbose", _flag)],
    "DAE": [("epochs", "epochs", int), ("batch", "batch", int), ("lr", "lr", float), ("reg_lambda", "reg_lambda", float),
            ("test_seed", "test_seed", _test_files), ("update_seed", "update_seed", _test_files),
            ("input_kp", "input_kp", _csv(float)), ("keep_prob", "kp", float), ("firstN_range", "firstN", _csv(float)),
            ("initval", "initval", IN_DIR), ("save", "save", IN_DIR), ("hidden", "hidden", int)],
    "PRETRAIN": [("epochs", "epochs", int), ("batch", "batch", int), ("lr", "lr", float),
                 ("reg_lambda", "reg_lambda", float), ("save", "save", IN_DIR)],
    "TITLE": [("epochs", "epochs", int), ("batch", "batch", int), ("lr", "lr", float), ("input_kp", "input_kp", _csv(float)),
              ("title_kp", "title_kp", float), ("test_seed", "test_seed", _test_files),
              ("update_seed", "update_seed", _test_files), ("char_emb", "char_emb", int), ("char_model", "char_model", str)],
    "TITLE/Char_CNN": [("filter_num", "filter_num", int), ("filter_size", "filter_size", _csv(int))],
    "TITLE/Char_LSTM": [("rnn_hidden", "rnn_hidden", int), ("bi", "bi", _flag)],
    "TITLE/paths": [("DAEval", "DAEval", IN_DIR), ("save", "save", IN_DIR)],
    "CHALLENGE": [("challenge_data", "challenge_data", str), ("batch", "batch", int)],
}


def check_firstN(rng):
    """The reference's consistency rules for firstN_range (main.py:33-43): -1 = whole playlists; fractions in [0, 1);
    or absolute counts >= 1, both integral."""
    if len(rng) == 1:
        assert rng[0] == -1.0
        return
    lo, hi = rng[0], rng[1]
    assert lo <= hi
    if hi < 1:
        assert lo == 0 or not lo.is_integer()
    else:
        assert lo >= 1 and lo.is_integer() and hi.is_integer()


class Conf:
    def __init__(self, dir, ini):
        self.dir, self.ini = dir, ini
        self.apply("BASE")
        self.seed = int(ini.get("BASE", "seed", fallback="0"))
        self.device = int(ini.get("BASE", "device", fallback="0"))

    def apply(self, table, section=None):
        section = section or table.split("/")[0]
        for key, attr, parse in SCHEMA[table]:
            value = parse(self.ini.get(section, key))
            if parse is IN_DIR:
                value = os.path.join(self.dir, value)
            setattr(self, attr, value)

    # ---- the reference's method names ---------------------------------------------------------------
    def set_dae_conf(self):                                       # main.py:21-47
        self.apply("DAE")
        check_firstN(self.firstN)
        self.mode = "dae"

    def set_pretrain_conf(self):                                  # main.py:49-56
        self.apply("PRETRAIN")
        self.is_pretrain = True
        self.mode = "pretrain"

    def set_title_conf(self):                                     # main.py:58-86
        self.apply("TITLE")
        if "TITLE/" + self.char_model in SCHEMA:
            self.apply("TITLE/" + self.char_model)
        self.apply("TITLE/paths")
        os.makedirs(os.path.dirname(self.save) or ".", exist_ok=True)    # the checkpoint's directory must exist
        self.mode = "title"

    def set_challenge_oonf(self):                                 # main.py:88-94 (name kept, typo included)
        os.makedirs(self.result_dir, exist_ok=True)
        self.apply("CHALLENGE")
        self.result = os.path.join(self.result_dir, self.ini.get("CHALLENGE", "result"))
        # optional key (not in the reference's schema): rank with the DAE scores alone instead of failing when the
        # title checkpoint is missing (main_runner/main_challenge.py)
        self.challenge_dae_only = _flag(self.ini.get("CHALLENGE", "dae_only", fallback="False"))

    set_challenge_conf = set_challenge_oonf


def load_conf(dir):
    ini = configparser.ConfigParser()
    ini.read(os.path.join(dir, "config.ini"))
    return Conf(dir, ini)


MODES = (("pretrain", "pretrain mode if Specified"), ("dae", "DAE mode if Specified"),
         ("title", "title mode if Specified"), ("challenge", "challenge mode if Specified"),
         ("testmode", "test mode if Specified(just check the result)"))


def main(argv=None):
    ap = argparse.ArgumentParser(description="args")
    ap.add_argument("--dir", type=str, default="qwerty", help="directory name which contains config file")
    for flag, text in MODES:
        ap.add_argument("--" + flag, action="store_true", default=False, help=text)
    args = ap.parse_args(argv)
    run_dir = os.path.join(".", args.dir)
    if not os.path.isdir(run_dir):
        print("ERROR: Cannot find " + run_dir + " ->Create directory and config.ini file first")
        return 0
    if "config.ini" not in os.listdir(run_dir):
        print("ERROR: Cannot find config.ini in " + run_dir + " ->Create config.ini file in the directory first")
        return 0
    conf = load_conf(run_dir)
    conf.set_dae_conf()                                           # always first: every mode starts from the [DAE] values
    from .main_runner import main_challenge, main_train
    if args.pretrain:
        conf.set_pretrain_conf()
        main_train.run(conf, args.testmode)
    elif args.dae:
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
