"""CPU oracle for the DAE hot path of hojinYang/spotify_recSys_challenge_2018.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the reported CPU comparator.  The product
(``spotify_recsys_challenge_2018_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity status
-------------
* The arithmetic of the path lives in TensorFlow 1.5.0 (third party, not
  vendored under /root/reference, not installable here: Python 3.12, no
  network).  The reference ships no tests, fixtures or golden vectors for it.
  The DAE arithmetic restated in ``dae_oracle.py`` / ``tf1_graph_cpu.py`` is
  therefore **parity unpinned** upstream; its pins are created here
  (closed-form backward vs torch autograd in fp64, hand-computed known-answer
  cases, TF1-form Adam restated from the published ApplyAdam functor).
* The pure-Python/NumPy parts of the path that *do* import from
  /root/reference in the build container (``utils/metrics.py`` r-precision /
  ndcg / rsc, ``main_runner/main_challenge.cand_generate`` ranking,
  ``utils/data_reader.py`` sparse-batch producers, ``main.Conf``) are pinned
  by golden vectors generated from the reference itself:
  ``tests/golden/make_golden.py`` -> ``tests/golden/*.json|npz``.
"""
