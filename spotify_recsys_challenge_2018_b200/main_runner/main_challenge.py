"""Challenge inference: host mirror of the reference's main_runner/main_challenge.py.

Batched inference over a challenge file, top-500 per playlist with the seeds removed, ids mapped
to 'spotify:track:<uri>' and pickled as [pid, uri x 500] rows (main_challenge.py:26-41, :72-96).
Decode and ranking run on the device (model.recommend); only the id -> uri mapping stays here.
"""
from __future__ import annotations

import datetime
import os
import pickle

from ..models.DAEs import DAE_title
from ..utils.data_reader import data_reader_challenge
from .main_train import log_write


def cand_to_uri(cand, id2uri):
    """main_challenge.py:37-41"""
    return ["spotify:track:" + id2uri[str(int(i))] for i in cand if i >= 0]


def run(conf):
    reader = data_reader_challenge(data_dir=conf.data_dir, filename=conf.challenge_data, batch_size=conf.batch)
    conf.n_tracks = reader.num_tracks                                         # main_challenge.py:49-53
    conf.n_input = reader.num_items
    conf.n_output = reader.num_items
    conf.charsize = reader.num_char
    conf.strmaxlen = reader.max_title_len

    info = "[challenge mode]"
    model_title = None
    title_ckpt = conf.save
    # The reference restores the title checkpoint unconditionally and fails when it is missing (saver.restore,
    # main_challenge.py:68-69).  Same here: a mistyped path must not silently produce a lower-quality submission.
    # DAE-only ranking (titles_use = 0 for every playlist) is an explicit opt-in: [CHALLENGE] dae_only = True.
    dae_only = bool(getattr(conf, "challenge_dae_only", False))
    if not dae_only:
        if not os.path.exists(title_ckpt):
            raise FileNotFoundError("title checkpoint %r not found (main_challenge.py:69 restores it unconditionally); set "
                                    "[CHALLENGE] dae_only = True to rank with the DAE scores alone" % title_ckpt)
        from ..models.title_get import get_model
        model_title = get_model(conf)
    model = DAE_title(conf, model_title)
    info += " start at " + str(datetime.datetime.now())
    log_write(conf, "*" * 10)
    log_write(conf, info)
    model.fit()
    if model_title is not None:
        model_title.fit(model)
        model_title.restore(title_ckpt)                                        # saver.restore (main_challenge.py:69)
    else:
        log_write(conf, "[CHALLENGE] dae_only: ranking by the DAE scores alone (titles_use = 0)")

    total_cands = []
    while True:
        x_positions, seed, titles, titles_exist, pid, x_ones = reader.next_batch()
        if model_title is not None:
            cand = model_title.recommend(model, x_positions, x_ones, titles, seed,
                                         titles_use=[t[0] for t in titles_exist])
        else:
            cand = model.recommend(x_positions, x_ones, seed, k=500, reuse_output=True)
        for i in range(len(seed)):
            total_cands.append([pid[i]] + cand_to_uri(cand[i], reader.id2uri))  # main_challenge.py:89-90
        if reader.ch_idx == 0:
            break
    with open(conf.result, "wb") as f:                                        # main_challenge.py:95-96
        pickle.dump(total_cands, f)
    return total_cands
