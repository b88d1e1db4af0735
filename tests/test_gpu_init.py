"""Initialisers on the device: tf.contrib.layers.xavier_initializer() [TF1] = U(+-sqrt(6 / (fan_in + fan_out)))
for the DAE matrices with zero biases (models/DAEs.py:53-61, :119-128), and xavier_initializer(uniform=False) [TF1] =
truncated normal, stddev sqrt(2.6 / (fan_in + fan_out)), resampled beyond two standard deviations, for EVERY title
variable, biases included (models/title_models/Char_CNN.py:19, :45-47, :71-73).  TF's generator is not reproducible
here (unseeded upstream); what is pinned is the distribution: support, mean, variance, truncation."""
import numpy as np
import pytest

from spotify_recsys_challenge_2018_b200.models.DAEs import DAE, DAE_tied, DAE_title
from spotify_recsys_challenge_2018_b200.models.title_get import get_model
from tests.gpu_util import Conf

pytestmark = pytest.mark.gpu

TRUNC_STD = 0.87962566103423978       # std of a standard normal truncated to [-2, 2]


def _check_uniform(w, lim):
    n = w.size
    assert np.abs(w).max() <= lim * (1 + 1e-6) and np.abs(w).max() > 0.999 * lim     # (the device rounds the limit to fp32)
    assert abs(w.mean()) < 5 * (lim / np.sqrt(3)) / np.sqrt(n)
    assert abs(w.var() / (lim * lim / 3) - 1) < 0.01
    hist, _ = np.histogram(w, bins=20, range=(-lim, lim))
    assert np.abs(hist / (n / 20) - 1).max() < 0.02              # flat


@pytest.mark.parametrize("tied", [False, True])
def test_xavier_uniform_init(tied):
    N, T, H = 40000, 33000, 256
    m = (DAE_tied if tied else DAE)(Conf(batch=64, n_input=N, n_tracks=T, hidden=H, lr=0.01, seed=3)).fit()
    W_enc, W_dec, b_enc, b_dec = m.get_params()
    lim = np.sqrt(6.0 / (N + H))
    _check_uniform(W_enc, lim)
    assert not b_enc.any() and not b_dec.any()                   # zeros_initializer (DAEs.py:56-59)
    if tied:
        assert W_dec is W_enc
    else:
        _check_uniform(W_dec, lim)
        assert abs(np.corrcoef(W_enc.ravel()[:200000], W_dec.ravel()[:200000])[0, 1]) < 0.01    # independent draws
    # the values are keyed by the element's GLOBAL index: the same matrix whatever the row sharding
    ms = [(DAE_tied if tied else DAE)(Conf(batch=64, n_input=N, n_tracks=T, hidden=H, lr=0.01, seed=3, world=2, rank=r)).fit()
          for r in range(2)]
    for x in ms:
        x.attach_local(ms)
    got = ms[0].get_params()
    assert np.array_equal(got[0], W_enc) and np.array_equal(got[1], W_dec)
    # and by the seed
    m2 = (DAE_tied if tied else DAE)(Conf(batch=64, n_input=N, n_tracks=T, hidden=H, lr=0.01, seed=4)).fit()
    assert not np.array_equal(m2.get_params()[0], W_enc)
    for x in ms + [m, m2]:
        x.close()


def test_title_truncated_normal_init():
    N, T, H, B = 20000, 17000, 64, 64
    conf = Conf(batch=B, n_input=N, n_tracks=T, n_output=N, hidden=H, lr=0.005, seed=11, DAEval=None, charsize=41,
                strmaxlen=25, char_emb=50, char_model="Char_CNN", filter_num=100, filter_size=[3, 5, 7, 9])
    tm = get_model(conf)
    m = DAE_title(conf, tm)
    m._create()
    tm.fit(m)
    params = tm.get_params()
    E, F, D = 50, 100, 400
    fans = [(41, E)]
    for w in (3, 5, 7, 9):
        fans += [(w * E, w * E * F), (F, F)]                     # conv kernel [w, E, 1, F]: receptive field x channels; bias [F]
    fans += [(D, N), (N, N)]
    assert len(fans) == len(params)
    for p, (fi, fo) in zip(params, fans):
        sd = np.sqrt(2.6 / (fi + fo))
        assert np.abs(p).max() <= 2 * sd * (1 + 1e-6)           # truncation at two standard deviations
        assert p.any()
        if p.size >= 2000:
            assert np.abs(p).max() > 1.9 * sd
            assert abs(p.std() / (TRUNC_STD * sd) - 1) < 0.03
            assert abs(p.mean()) < 5 * TRUNC_STD * sd / np.sqrt(p.size)
    # Output_W [D, N]: every feature row and every item column is populated (no dead padding leaks into the host layout)
    assert (np.abs(params[-2]).sum(1) > 0).all() and (np.abs(params[-2]).sum(0) > 0).all()
    tm.close(); m.close()
