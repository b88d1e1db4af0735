"""Train / eval loop: host mirror of the reference's main_runner/main_train.py, rebuilt to the
intended behaviour where the committed snapshot is broken.

What is kept verbatim: reader selection by firstN (main_train.py:129-133), the hide-and-seek coin
flip between track-only and artist-only input with tracks+artists as target (:199-213), input_kp
drawn uniformly per batch (:199), the epoch boundary rule (:227), per-epoch evaluation over every
test seed (:235-241), "save when the summed r-precision over update_seed improves" (:243-249) and
the log lines (:124-125, :232-239).
What changes: `sess.run` -> model.train_step / model.recommend (decode + top-500 on the device
instead of a [B,N] D2H copy and np.argsort per playlist).
"""
from __future__ import annotations

import datetime
import os
import random
import time

import numpy as np

from ..models.DAEs import DAE, DAE_tied
from ..utils import metrics as met
from ..utils.data_reader import data_reader, data_reader_firstN, data_reader_test


def log_write(conf, log):
    """main_challenge.py:17-23 (the 2-argument form every call site uses, D2)."""
    with open(os.path.join(conf.dir, "log.txt"), "a") as f:
        f.write(log)
        f.write("\n")
    if conf.verbose:
        print(log)


def show_result(rprecision, ndcg, rsc):
    return "rprecision: %f ndcg: %f rsc: %f" % (rprecision, ndcg, rsc)      # main_train.py:124-125


def eval(reader_test, conf, model, model_title=None):
    """Average (r-precision, ndcg, rsc) over one test file (main_train.py:48-121, intended form D3)."""
    total = np.zeros(3)
    test_size = len(reader_test.playlists)
    while True:
        x_positions, test_seed, test_answer, titles, x_ones = reader_test.next_batch_test()
        # y_pred[:, :n_tracks] + met.single_eval of every playlist (main_train.py:66-90), ranked AND scored on the device
        if model_title is not None:
            res = model_title.evaluate(model, x_positions, x_ones, titles, test_seed, test_answer, titles_use=1.0)
        else:
            res = model.evaluate(x_positions, x_ones, test_seed, test_answer, k=500)
        total += res.sum(axis=0)
        if reader_test.test_idx == 0:                                         # main_train.py:97-98
            break
    total /= test_size
    return total[0], total[1], total[2]


def run(conf, only_testmode):
    if -1 in conf.firstN:                                                     # main_train.py:129-133
        reader = data_reader(data_dir=conf.data_dir, filename="train", batch_size=conf.batch)
    else:
        reader = data_reader_firstN(data_dir=conf.data_dir, filename="train", batch_size=conf.batch,
                                    from_to=conf.firstN)
    conf.class_divpnt = getattr(reader, "class_divpnt", [])                   # optional (D5)
    conf.n_tracks = reader.num_tracks
    conf.n_input = reader.num_items
    conf.n_output = reader.num_items
    conf.charsize = reader.num_char
    conf.strmaxlen = reader.max_title_len

    kp_range = conf.input_kp
    readers_test = {}
    for seed in conf.test_seed:                                               # main_train.py:146-149
        readers_test[seed] = data_reader_test(data_dir=conf.data_dir, filename=seed, batch_size=conf.batch,
                                              test_num=conf.testsize)
    model_title = None
    if conf.mode == "pretrain":
        info = "[pretrain mode]"
        model = DAE_tied(conf)
    elif conf.mode == "dae":
        if only_testmode:
            conf.initval = conf.save                                          # main_train.py:158-159
        info = "[dae mode]"
        model = DAE(conf)
    elif conf.mode == "title":
        info = "[title mode]"
        from ..models.title_get import get_model
        from ..models.DAEs import DAE_title
        model_title = get_model(conf)
        model = DAE_title(conf, model_title)
    else:
        raise ValueError("unknown mode %r" % (conf.mode,))
    info += " start at " + str(datetime.datetime.now())
    log_write(conf, "*" * 10)
    log_write(conf, info)

    model.fit()
    if model_title is not None:
        model_title.fit(model)

    if only_testmode:                                                         # main_train.py:181-191
        log_write(conf, "<<only test mode>>")
        if model_title is not None:
            model_title.restore(conf.save)
        for seed_num, reader_test in readers_test.items():
            log_write(conf, "seed num: " + seed_num)
            log_write(conf, show_result(*eval(reader_test, conf, model, model_title)))
        return

    epoch, it, loss, max_eval = 0, 0, 0.0, 0.0                                # max_eval initialised (D4)
    t0 = time.time()
    n_seen = 0
    while True:
        start_idx = reader.train_idx
        trk_positions, art_positions, y_positions, titles, trk_val, art_val = reader.next_batch()
        end_idx = reader.train_idx
        input_kp = random.uniform(kp_range[0], kp_range[-1])                  # main_train.py:199
        if conf.mode in ("pretrain", "dae"):
            # pipelined step: the device works on this batch while the reader builds the next one; the cost that
            # comes back is the previous step's (the loop only accumulates it, main_train.py:223)
            y_ones = np.ones(len(y_positions), np.float32)
            if np.random.randint(2) == 0:                                     # main_train.py:202-213
                l = model.train_step_async(trk_positions, trk_val, y_positions, y_ones, conf.kp, input_kp)
            else:
                l = model.train_step_async(art_positions, art_val, y_positions, y_ones, conf.kp, input_kp)
        else:                                                                  # main_train.py:214-221
            l = model_title.train_step_async(model, y_positions, np.ones(len(y_positions), np.float32), titles,
                                             conf.kp, conf.title_kp, input_kp)            # pipelined like the DAE modes
        if l is not None:
            loss += l
        it += 1
        n_seen += conf.batch
        if start_idx > end_idx or end_idx == 0:                               # main_train.py:227
            loss += (model.flush() if conf.mode in ("pretrain", "dae") else model_title.flush()) or 0.0   # the epoch's last step
            epoch += 1
            loss = loss / it
            log_write(conf, "epoch " + str(epoch))
            log_write(conf, "training loss: " + str(loss))
            log_write(conf, "playlists/s (host loop incl. reader): %.1f" % (n_seen / max(time.time() - t0, 1e-9)))
            cur_eval = 0.0
            for seed_num, reader_test in readers_test.items():
                log_write(conf, "seed num: " + seed_num)
                rprec, ndcg, rsc = eval(reader_test, conf, model, model_title)
                log_write(conf, show_result(rprec, ndcg, rsc))
                if seed_num in conf.update_seed:
                    cur_eval += rprec
            if cur_eval >= max_eval:                                          # main_train.py:243-249
                if conf.mode in ("pretrain", "dae"):
                    model.save_model()
                else:
                    model_title.save(conf.save)
                max_eval = cur_eval
                log_write(conf, "The highest score is updated. Parameters are saved")
            loss, it = 0.0, 0
            if epoch == conf.epochs:
                break
    return max_eval
