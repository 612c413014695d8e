"""Drop-in for the reference's pybind11 `sampling` module (code/sources/sampling.cpp:95-106), backed by liblgcn_b200.so:

    seed(seed)                                                     sampling.cpp:88-91
    randint(end)                                                   sampling.cpp:22-25
    sample_negative(user_num, item_num, train_num, allPos, neg_num) -> int32[user_num*(train_num//user_num), 2+neg_num]   :27-56
    sample_negative_ByUser(users, item_num, allPos, neg_num)        -> int32[len(users), 2+neg_num]                       :58-86

Same glibc rand() stream and draw order (tests/golden/sampler.npz was recorded from the reference's file compiled as it lies),
so `utils.sampling = lgcn_b200.sampling` leaves the reference's UniformSample_original unchanged (code/utils.py:68-81).
allPos is the reference's list of per-user item arrays; a (indptr, items) CSR pair is accepted as well.
"""
import ctypes

import numpy as np

from . import _lib


def _csr(allPos):
    if isinstance(allPos, tuple) and len(allPos) == 2:
        return np.ascontiguousarray(allPos[0], dtype=np.int64), np.ascontiguousarray(allPos[1], dtype=np.int32)
    lens = np.fromiter((len(a) for a in allPos), dtype=np.int64, count=len(allPos))
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = np.concatenate([np.asarray(a, dtype=np.int32) for a in allPos]) if len(allPos) else np.zeros(0, np.int32)
    return np.ascontiguousarray(indptr), np.ascontiguousarray(items, dtype=np.int32)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _cancel_prefetch():
    from . import utils
    utils._prefetch.cancel()            # an epoch sample drawn ahead of time is rewound: this call sees the stream position it expects


def seed(s):
    _cancel_prefetch()
    _lib.load().lgcn_sampler_seed(ctypes.c_uint32(int(s) & 0xffffffff))


def randint(end):
    _cancel_prefetch()
    r = _lib.load().lgcn_randint(int(end))
    if r < 0:
        raise RuntimeError(_lib.load().lgcn_last_error().decode())
    return int(r)


def sample_negative(user_num, item_num, train_num, allPos, neg_num):
    _cancel_prefetch()
    indptr, items = _csr(allPos)
    out = np.empty((int(user_num) * (int(train_num) // int(user_num)), 2 + int(neg_num)), dtype=np.int32)
    rows = _lib.load().lgcn_sample_negative(int(user_num), int(item_num), int(train_num), _vp(indptr), _vp(items), int(neg_num), _vp(out))
    if rows < 0:
        raise RuntimeError(_lib.load().lgcn_last_error().decode())
    return out


def sample_negative_ByUser(users, item_num, allPos, neg_num):
    _cancel_prefetch()
    indptr, items = _csr(allPos)
    users = np.ascontiguousarray(users, dtype=np.int32)
    out = np.empty((users.size, 2 + int(neg_num)), dtype=np.int32)
    rows = _lib.load().lgcn_sample_negative_by_user(_vp(users), users.size, indptr.size - 1, int(item_num), _vp(indptr), _vp(items), int(neg_num), _vp(out))
    if rows < 0:
        raise RuntimeError(_lib.load().lgcn_last_error().decode())
    return out
