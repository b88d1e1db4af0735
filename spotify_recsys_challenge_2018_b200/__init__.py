"""B200-native denoising-autoencoder hot path of hojinYang/spotify_recSys_challenge_2018.

Host side (Python, mirrors the reference's module tree for this path):
    main.py                      Conf + CLI            (reference main.py)
    main_runner/main_train.py    train / eval loop     (reference main_runner/main_train.py)
    main_runner/main_challenge.py challenge inference  (reference main_runner/main_challenge.py)
    models/DAEs.py               DAE_tied / DAE / DAE_title over the C ABI (reference models/DAEs.py)
    utils/data_reader.py         sparse-batch readers  (reference utils/data_reader.py)
    utils/metrics.py             r-precision / ndcg / clicks (reference utils/metrics.py)
Device side: csrc/ -> libdae_b200.so (hand-written sm_100a CUDA behind include/dae_b200.h).
There is no CPU fallback: importing the models without the built library raises.
"""
__version__ = "0.1.0"
