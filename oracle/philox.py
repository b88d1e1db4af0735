"""Philox4x32-10 counter-based RNG, NumPy restatement (test infrastructure).

The reference draws its dropout masks from TF1's unseeded stateful RNG
(models/DAEs.py:40, :68 -> tf.nn.dropout), so no mask is reproducible
upstream.  The B200 build keys every Bernoulli draw by
(seed, stream, step, row, col) with Philox4x32-10 (Salmon et al., SC'11) so the
same mask can be re-created here and is independent of the batch sharding.
Device twin: spotify_recsys_challenge_2018_b200/csrc/philox.cuh.

Known-answer vectors (Random123 kat_vectors) are checked in
tests/test_oracle.py.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)

STREAM_INPUT = 0   # input dropout (DAEs.py:40), counter = (item, row, step, 0)
STREAM_HIDDEN = 1  # hidden dropout (DAEs.py:68), counter = (unit, row, step, 1)
STREAM_TITLE = 2   # title-feature dropout (Char_CNN.py:67)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
    c0 = c0.astype(np.uint64); c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64); c3 = c3.astype(np.uint64)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0
            p1 = PHILOX_M1 * c2
            hi0 = p0 >> np.uint64(32); lo0 = p0 & _MASK32
            hi1 = p1 >> np.uint64(32); lo1 = p1 & _MASK32
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n1 = lo1
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def uniform24(seed, stream, step, row, col):
    """u in [0,1) with 24 random bits: (word0 >> 8) * 2^-24, as float32."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    step = int(step) & 0xFFFFFFFFFFFFFFFF
    c2 = np.uint32(step & 0xFFFFFFFF)
    c3 = np.uint32((((step >> 32) & 0xFFFFFF) << 8) | (int(stream) & 0xFF))
    w0 = philox4x32_10(col, row, c2, c3, seed & 0xFFFFFFFF, seed >> 32)[0]
    return (w0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def keep_mask(seed, stream, step, row, col, keep_prob):
    """Bernoulli(keep_prob) keep decision; keep_prob >= 1 keeps everything.

    Same law as TF1's dropout (floor(keep_prob + U) == 1  <=>  U >= 1-keep_prob)
    [TF1 nn_ops.dropout], expressed as u < keep_prob on the 24-bit uniform.
    """
    u = uniform24(seed, stream, step, row, col)
    return u < np.float32(keep_prob)
