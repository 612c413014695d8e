"""TEST INFRASTRUCTURE — numpy restatement of the reference's LightGCN hot path.

Every function cites the reference lines (LightGCN_work/code/...) it restates.  fp64 by default
(the "true" value the fp32 paths are compared to), fp32 where the reference's rounding matters.
See oracle/__init__.py for how this oracle is pinned.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def clib():
    """C part of the oracle (oracle/c/score_topk_ref.c), built by oracle/Makefile."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _CLIB = ctypes.CDLL(path)
    return _CLIB


# ------------------------------------------------------------------------------ graph build
def build_norm_adj(train_user, train_item, n_users, m_items):
    """dataloader.py:133-136 (csr_matrix sums duplicate pairs), :223-227 (A = [[0,R],[R^T,0]]),
    :230-234 (rowsum, d^-1/2, zero-degree -> 0, D.A.D with row scaling first, all float32).

    Returns indptr int32[N+1], indices int32[nnz] (sorted per row), vals float32[nnz],
    deg float32[N] (weighted row sums), dinv float32[N]."""
    tu = np.asarray(train_user, dtype=np.int64)
    ti = np.asarray(train_item, dtype=np.int64)
    N = n_users + m_items
    rows = np.concatenate([tu, ti + n_users])
    cols = np.concatenate([ti + n_users, tu])
    key, mult = np.unique(rows * N + cols, return_counts=True)       # sorted by (row, col); duplicates summed
    r, c = key // N, key % N
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    indptr = np.cumsum(indptr)
    deg = np.zeros(N, dtype=np.float64)
    np.add.at(deg, r, mult.astype(np.float64))
    deg32 = deg.astype(np.float32)
    with np.errstate(divide='ignore'):
        dinv = np.where(deg32 > 0, (1.0 / np.sqrt(deg.clip(min=1e-300))), 0.0).astype(np.float32)   # correctly rounded d^-1/2
    w = mult.astype(np.float32)
    vals = ((dinv[r] * w).astype(np.float32) * dinv[c]).astype(np.float32)   # fl32(fl32(d_r * a) * d_c)
    return indptr.astype(np.int32), c.astype(np.int32), vals, deg32, dinv


# ------------------------------------------------------------------------------ propagation
def spmm(indptr, indices, vals, X):
    """Y = A @ X for a CSR A (model.py:217 torch.sparse.mm), in X's dtype, pure numpy."""
    n_rows = indptr.size - 1
    Y = np.zeros((n_rows, X.shape[1]), dtype=X.dtype)
    nnz = indices.size
    if nnz == 0:
        return Y
    chunk = max(1, (1 << 24) // max(X.shape[1], 1))
    row_of = np.repeat(np.arange(n_rows), np.diff(indptr).astype(np.int64))
    for lo in range(0, nnz, chunk):
        hi = min(nnz, lo + chunk)
        contrib = vals[lo:hi, None].astype(X.dtype) * X[indices[lo:hi]]
        np.add.at(Y, row_of[lo:hi], contrib)
    return Y


def spmm_scipy(indptr, indices, vals, X):
    """Same product through scipy (fast path for the big cases: gowalla KAT, CPU baseline)."""
    import scipy.sparse as sp
    A = sp.csr_matrix((vals.astype(X.dtype), indices, indptr), shape=(indptr.size - 1, X.shape[0]))
    return np.asarray(A @ X)


def propagate(indptr, indices, vals, E0, n_layers, dtype=np.float64, fast=False):
    """model.py:201-225: embs=[E0]; x = A x (L times); out = mean(stack(embs), dim=1)."""
    mm = spmm_scipy if fast else spmm
    x = E0.astype(dtype)
    v = vals.astype(dtype)
    acc = x.copy()
    for _ in range(n_layers):
        x = mm(indptr, indices, v, x)
        acc += x
    return acc / (n_layers + 1)


def propagate_backward(indptr, indices, vals, G, n_layers, dtype=np.float64, fast=False):
    """Adjoint of propagate (A is symmetric): g_L = G/(L+1); g_k = G/(L+1) + A g_{k+1}; returns g_0
    (what SparseAddmmBackward + mean/stack backward compute, utils.py:61)."""
    mm = spmm_scipy if fast else spmm
    s = 1.0 / (n_layers + 1)
    G = G.astype(dtype)
    v = vals.astype(dtype)
    g = s * G
    for _ in range(n_layers):
        g = s * G + mm(indptr, indices, v, g)
    return g


# ------------------------------------------------------------------------------ BPR loss
def bpr_loss(out, users, pos, neg, n_users, dtype=np.float64):
    """model.py:162-173 on PROPAGATED embeddings: bpr = -mean(logsigmoid(pos-neg)),
    reg = 0.5(|u|^2+|p|^2+|n|^2)/B.  Also returns dbpr/dout and dreg/dout (closed form, §3.2)."""
    out = out.astype(dtype)
    U, P, Nn = out[users], out[n_users + pos], out[n_users + neg]
    B = float(len(users))
    z = (U * P).sum(1) - (U * Nn).sum(1)
    bpr = np.mean(np.maximum(-z, 0) + np.log1p(np.exp(-np.abs(z))))          # softplus(-z)
    reg = 0.5 * ((U * U).sum() + (P * P).sum() + (Nn * Nn).sum()) / B
    s = np.where(z >= 0, np.exp(-np.abs(z)) / (1 + np.exp(-np.abs(z))), 1 / (1 + np.exp(-np.abs(z))))   # sigmoid(-z)
    Gb = np.zeros_like(out)
    Gr = np.zeros_like(out)
    np.add.at(Gb, users, (s[:, None] * (Nn - P)) / B)
    np.add.at(Gb, n_users + pos, (-s[:, None] * U) / B)
    np.add.at(Gb, n_users + neg, (s[:, None] * U) / B)
    np.add.at(Gr, users, U / B)
    np.add.at(Gr, n_users + pos, P / B)
    np.add.at(Gr, n_users + neg, Nn / B)
    return bpr, reg, Gb, Gr


def adam_step(p, m, v, g, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (utils.py:51,62; defaults, step count t >= 1)."""
    m = m + (1 - b1) * (g - m)
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    p = p - (lr / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + eps))
    return p, m, v


def train_step(indptr, indices, vals, E0, m, v, t, users, pos, neg, n_users, n_layers, decay, lr, dtype=np.float64, fast=False):
    """utils.py:53-64 (stageOne): loss = bpr + decay*reg -> backward -> Adam.  Returns (loss, E0', m', v')."""
    out = propagate(indptr, indices, vals, E0, n_layers, dtype, fast)
    bpr, reg, Gb, Gr = bpr_loss(out, users, pos, neg, n_users, dtype)
    g0 = propagate_backward(indptr, indices, vals, Gb + decay * Gr, n_layers, dtype, fast)
    p2, m2, v2 = adam_step(E0.astype(dtype), m.astype(dtype), v.astype(dtype), g0, t, lr)
    return bpr + decay * reg, p2, m2, v2, (bpr, reg, g0)


# ------------------------------------------------------------------------------ scoring / top-k
def topk_stable(scores, k):
    """Indices of the k largest per row, ties -> lowest index first (the fixed tie-break rule)."""
    order = np.argsort(-scores, axis=1, kind='stable')
    return order[:, :k]


def masked_scores(out_users, out_items, users, indptr, indices, n_users, dtype=np.float64):
    """model.py:122 + Procedure.py:177-181: scores with the users' train items set to -(1<<10)."""
    S = out_users[users].astype(dtype) @ out_items.astype(dtype).T
    for b, u in enumerate(users):
        S[b, indices[indptr[u]:indptr[u + 1]] - n_users] = -(1 << 10)
    return S


def score_topk_exact(out_users, out_items, users, k, indptr=None, indices=None, mask_col_offset=0):
    """fp32 FMA-chain scores + mask + top-k through the C restatement (bit-exact contract of K3)."""
    lib = clib()
    U = np.ascontiguousarray(out_users, dtype=np.float32)
    V = np.ascontiguousarray(out_items, dtype=np.float32)
    users = None if users is None else np.ascontiguousarray(users, dtype=np.int64)
    Bt = U.shape[0] if users is None else users.size
    idx = np.zeros((Bt, k), dtype=np.int64)
    val = np.zeros((Bt, k), dtype=np.float32)
    vp = ctypes.c_void_p
    ip = None if indptr is None else np.ascontiguousarray(indptr, dtype=np.int32)
    ii = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    lib.oracle_score_topk(U.ctypes.data_as(vp), V.ctypes.data_as(vp), None if users is None else users.ctypes.data_as(vp),
                          ctypes.c_int(Bt), ctypes.c_int(V.shape[0]), ctypes.c_int(V.shape[1]),
                          None if ip is None else ip.ctypes.data_as(vp), None if ii is None else ii.ctypes.data_as(vp),
                          ctypes.c_int(mask_col_offset), ctypes.c_int(k), idx.ctypes.data_as(vp), val.ctypes.data_as(vp))
    return idx, val


def score_dense_exact(out_users, out_items, users):
    lib = clib()
    U = np.ascontiguousarray(out_users, dtype=np.float32)
    V = np.ascontiguousarray(out_items, dtype=np.float32)
    users = None if users is None else np.ascontiguousarray(users, dtype=np.int64)
    Bt = U.shape[0] if users is None else users.size
    out = np.zeros((Bt, V.shape[0]), dtype=np.float32)
    vp = ctypes.c_void_p
    lib.oracle_score_dense(U.ctypes.data_as(vp), V.ctypes.data_as(vp), None if users is None else users.ctypes.data_as(vp),
                           ctypes.c_int(Bt), ctypes.c_int(V.shape[0]), ctypes.c_int(V.shape[1]), out.ctypes.data_as(vp))
    return out


# ------------------------------------------------------------------------------ metrics
def metrics_at_k(topk_items, ground_truth, ks):
    """Procedure.py:89-121 + utils.py:173-200,212-217, per-user then mean over users.
    topk_items int[n,kmax]; ground_truth list of item lists.  Returns dict of float64 arrays [len(ks)]."""
    n = len(ground_truth)
    res = {m: np.zeros(len(ks)) for m in ('precision', 'recall', 'ndcg')}
    for row, gt in zip(topk_items, ground_truth):
        gts = set(int(x) for x in gt)
        r = np.array([1.0 if int(x) in gts else 0.0 for x in row])
        for j, k in enumerate(ks):
            hits = r[:k].sum()
            res['precision'][j] += hits / k
            res['recall'][j] += hits / len(gt)
            disc = 1.0 / np.log2(np.arange(2, k + 2))
            idcg = disc[:min(k, len(gt))].sum()
            res['ndcg'][j] += (r[:k] * disc).sum() / (idcg if idcg != 0 else 1.0)
    return {m: v / max(n, 1) for m, v in res.items()}


# ------------------------------------------------------------------------------ sampler (oracle/c/sampler_ref.c)
def sample_negative_ref(seed, user_num, item_num, train_num, indptr, items, neg_num=1):
    """The reference's C++ sampler (code/sources/sampling.cpp:27-56,88-91) restated in C: srand(seed), then per user
    train_num // user_num triples.  indptr int64[user_num+1], items int32 -> int32[rows, 2+neg_num]."""
    lib = clib()
    lib.oracle_sample_negative.restype = ctypes.c_int64
    lib.oracle_sample_negative.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    indptr = np.ascontiguousarray(indptr, dtype=np.int64); items = np.ascontiguousarray(items, dtype=np.int32)
    out = np.empty((user_num * (train_num // user_num), 2 + neg_num), dtype=np.int32)
    if seed is not None:
        lib.oracle_sampler_seed(ctypes.c_uint(int(seed) & 0xffffffff))
    rows = lib.oracle_sample_negative(user_num, item_num, train_num, indptr.ctypes.data, items.ctypes.data, neg_num, out.ctypes.data)
    if rows < 0:
        raise RuntimeError("oracle sampler: a user has no positive item")
    return out


def sample_negative_by_user_ref(users, item_num, indptr, items, neg_num=1):
    """code/sources/sampling.cpp:58-86, continuing the current rand() stream."""
    lib = clib()
    lib.oracle_sample_negative_by_user.restype = ctypes.c_int64
    lib.oracle_sample_negative_by_user.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    users = np.ascontiguousarray(users, dtype=np.int32)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64); items = np.ascontiguousarray(items, dtype=np.int32)
    out = np.empty((users.size, 2 + neg_num), dtype=np.int32)
    if lib.oracle_sample_negative_by_user(users.ctypes.data, users.size, item_num, indptr.ctypes.data, items.ctypes.data, neg_num, out.ctypes.data) < 0:
        raise RuntimeError("oracle sampler: a user has no positive item")
    return out


def randint_ref(end):
    """code/sources/sampling.cpp:22-25."""
    lib = clib()
    lib.oracle_randint.restype = ctypes.c_int
    return int(lib.oracle_randint(ctypes.c_int(int(end))))
