"""Helpers for the GPU parity tests: device views of library-owned buffers, ctypes plumbing."""
import ctypes as C

import numpy as np
import torch

from spotify_recsys_challenge_2018_b200 import _lib


class _CudaArr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def dev_view(ptr, n, dtype):
    """torch view (no copy) of n elements of device memory owned by the library."""
    if dtype == torch.bfloat16:
        return torch.as_tensor(_CudaArr(ptr, n, "<i2"), device="cuda").view(torch.bfloat16)
    ts = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_CudaArr(ptr, n, ts), device="cuda")


def model_buf(model, name, dtype):
    ptr, n, es = model.buffer(name)
    return dev_view(ptr, n, dtype)


def P(t):
    """data pointer of a torch tensor (or None) as c_void_p"""
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def lib():
    return _lib.load()


def check(rc):
    _lib.check(rc)


class Conf:
    def __init__(self, **kw):
        self.save = "/tmp/dae_test_w"
        self.initval = "NULL"
        self.reg_lambda = 0.0
        self.seed = 11
        self.__dict__.update(kw)


def random_batch(rng, B, T, A, mean_len=20, dup=True, empty_rows=()):
    """A reader-shaped batch: (trk_pos, art_pos, y_pos) with duplicates and the block structure."""
    lens = np.clip(rng.poisson(mean_len, B), 1, 250)
    for r in empty_rows:
        lens[r] = 0
    rows = np.repeat(np.arange(B), lens)
    # Zipf-ish ids so that columns repeat across rows
    trk = np.minimum((np.exp(rng.random(rows.size) * np.log(T + 1.0)) - 1).astype(np.int64), T - 1)
    if dup and rows.size > 4:
        trk[1::7] = trk[0::7][:len(trk[1::7])]          # in-row duplicates
    art = T + (trk * 7919 % A)
    trk_pos = np.stack([rows, trk], 1)
    art_pos = np.stack([rows, art], 1)
    return trk_pos, art_pos, np.concatenate([trk_pos, art_pos], 0)
