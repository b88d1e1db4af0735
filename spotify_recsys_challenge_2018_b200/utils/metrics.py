"""Challenge metrics: host mirror of the reference's utils/metrics.py.

The ranking itself (np.argsort over n_tracks + seed removal + [:500], metrics.py:58-68) runs on
the device (`model.recommend`, csrc/topk.cu); what stays here are the integer set operations on
<= 500 candidate ids per playlist, restated from metrics.py:20-49 and pinned by
tests/golden/metrics_golden.json.
"""
from __future__ import annotations

import math


def get_r_precision(answer, cand, answer_cls=None, class_divpnt=None):
    """|set(answer) & set(cand[:len(answer)])| / len(answer)   (metrics.py:25-27).
    `answer` may contain -1 (never matched, still counted).  The two trailing arguments are the
    reference's abandoned per-class statistics and are ignored."""
    answer = list(answer)
    return len(set(answer) & set(list(cand)[:len(answer)])) / len(answer)


def get_ndcg(answer, cand):
    """metrics.py:29-42 (IDCG grows with the number of hits found after rank 0)."""
    answer = set(answer)
    cand = list(cand)
    idcg, idcg_idx = 1.0, 2
    dcg = 1.0 if cand[0] in answer else 0.0
    for i in range(1, len(cand)):
        if cand[i] in answer:
            dcg += 1 / math.log(i + 1, 2)
            idcg += 1 / math.log(idcg_idx, 2)
            idcg_idx += 1
    return dcg / idcg


def get_rsc(answer, cand):
    """Recommended-songs clicks: first hit index // 10, 51 when there is none (metrics.py:44-49)."""
    answer = set(answer)
    for i, c in enumerate(cand):
        if c in answer:
            return i // 10
    return 51


def get_metrics(answer, cand):
    """(r_precision, ndcg, rsc) -- the triple the runner logs (main_train.py:124-125, 237)."""
    cand = [int(c) for c in cand if c >= 0]
    return get_r_precision(answer, cand), get_ndcg(answer, cand) if cand else 0.0, get_rsc(answer, cand)


def single_eval(cand, answer):
    """Metrics of one playlist from its device-ranked candidates (metrics.py:58-70 minus the sort)."""
    return get_metrics(answer, cand)
