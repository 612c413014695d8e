"""Loss wrapper, sampler glue, batching and metric helpers with the reference's names and
semantics (code/utils.py), minus its three breakages: `timer` is complete and
`minibatch` has the single-tensor case that Procedure.Test relies on.
"""
import ctypes
import os
from time import time

import numpy as np
import torch
from torch import optim

from . import _lib, world


# ---------------------------------------------------------------------------- BPR loss / optimiser
class _AdamStateView(optim.Adam):
    """torch.optim.Adam whose state tensors ALIAS the engine's fused-Adam buffers, so that
    optimizer.state_dict()/load_state_dict() keep the reference checkpoint layout
    (code/main.py:56-87: 'optimizer_state') while the update itself runs inside the kernels."""

    def __init__(self, model, lr):
        super().__init__(model.parameters(), lr=lr)
        self._model = model
        self._bind()

    def _bind(self, full=None):
        m = self._model
        eng, nu = m._engine, m.n_users
        # single GPU / replicas: the state tensors ALIAS the engine's buffers.  Row partition: a rank holds only its rows of
        # M and V, so the aliases are replaced by gathered full tables whenever a state_dict is asked for.
        sharded = eng.dist_mode in ('rowpart', 'featpart') and eng.world > 1
        M, V = full if full is not None else ((None, None) if sharded else (eng.M, eng.V))
        for p, sl in ((m.embedding_user.weight, slice(0, nu)), (m.embedding_item.weight, slice(nu, None))):
            st = self.state[p]
            st['step'] = torch.tensor(float(eng._host_step))
            if M is not None:
                st['exp_avg'] = M[sl]
                st['exp_avg_sq'] = V[sl]
        if eng.pg is not None:          # pop-gate MLPs: the 8 tensors are consecutive slices of the engine's flat blocks
            off = 0
            for t in m.popgate_tensors():
                st = self.state[t]
                st['step'] = torch.tensor(float(eng._host_step))
                st['exp_avg'] = eng.pg['M'][off:off + t.numel()].view(t.shape)
                st['exp_avg_sq'] = eng.pg['V'][off:off + t.numel()].view(t.shape)
                off += t.numel()

    def state_dict(self):
        eng = self._model._engine
        if eng.dist_mode in ('rowpart', 'featpart') and eng.world > 1:
            self._bind(full=eng.adam_state_full())          # collective: every rank must call state_dict()
        for st in self.state.values():
            st['step'] = torch.tensor(float(eng._host_step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        m = self._model
        eng, nu = m._engine, m.n_users
        step = 0
        parts = []
        for p, sl in ((m.embedding_user.weight, slice(0, nu)), (m.embedding_item.weight, slice(nu, None))):
            st = self.state.get(p, {})
            if 'exp_avg' in st:
                parts.append((st['exp_avg'], st['exp_avg_sq']))
                step = int(float(st['step']))
        if len(parts) == 2:
            dev = eng.M.device
            eng.load_adam_state(torch.cat([parts[0][0].to(dev), parts[1][0].to(dev)]), torch.cat([parts[0][1].to(dev), parts[1][1].to(dev)]))
        if eng.pg is not None:
            off = 0
            for t in m.popgate_tensors():
                st = self.state.get(t, {})
                if 'exp_avg' in st:
                    eng.pg['M'][off:off + t.numel()].copy_(st['exp_avg'].reshape(-1)); eng.pg['V'][off:off + t.numel()].copy_(st['exp_avg_sq'].reshape(-1))
                off += t.numel()
        eng.set_lr(self.param_groups[0]['lr'])
        eng.set_adam_step(step)
        self._bind()


class BPRLoss:
    """Same constructor and `stageOne(users, pos, neg) -> float` as the reference (code/utils.py:38-64).
    With this package's LightGCN the step is the fused kernel sequence (no autograd graph, no
    materialised gradients); with any other nn.Module it is the reference's generic recipe."""

    def __init__(self, recmodel, config):
        self.model = recmodel
        self.weight_decay = config['decay']
        self.lr = config['lr']
        self.fused = hasattr(recmodel, 'fused_train_step') and getattr(recmodel, 'plain', True)
        if self.fused:
            recmodel._engine.decay = float(self.weight_decay)
            recmodel._engine.set_lr(self.lr)
            self.opt = _AdamStateView(recmodel, self.lr)
        else:
            self.opt = optim.Adam(recmodel.parameters(), lr=self.lr)

    def stageOne_async(self, users, pos, neg):
        """Enqueue one step; returns the device tensor {bpr, reg, total, running} without synchronising."""
        if not self.fused:
            raise RuntimeError("stageOne_async needs lgcn_b200.LightGCN")
        eng = self.model.fused_train_step(users, pos, neg, lr=self.opt.param_groups[0]['lr'])
        return eng.loss_out

    def stageOne(self, users, pos, neg):
        if self.fused:
            eng = self.model.fused_train_step(users, pos, neg, lr=self.opt.param_groups[0]['lr'])
            return float(eng.loss_to_host()[2])          # loss.cpu().item() (code/utils.py:64)
        loss, reg_loss = self.model.bpr_loss(users, pos, neg)
        loss = loss + reg_loss * self.weight_decay
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.cpu().item()


# ---------------------------------------------------------------------------- sampling
sample_ext = True     # the C sampler ships inside liblgcn_b200.so


class _EpochPrefetch:
    """The NEXT epoch's sample, drawn by a background thread while the GPU runs the current epoch (the ctypes call releases
    the GIL; the C sampler is ~30 ms per gowalla epoch, a quarter of a BPR_train_original call).  Exactness: the sampler's
    generator state is snapshotted before the prefetch; if the next sampler call is not the epoch sample it ran ahead for
    (another dataset, sample_negative_ByUser, randint, a reseed ...), the state is rewound and the prefetch dropped — every
    interleaving of calls sees the stream the reference's module would have produced."""

    def __init__(self):
        self.thread, self.key, self.result, self.snapshot, self.error = None, None, None, None, None

    def cancel(self):
        if self.thread is None:
            return
        self.thread.join()
        _lib.load().lgcn_sampler_set_state(self.snapshot.ctypes.data_as(ctypes.c_void_p))       # rewind: as if it never ran
        self.thread, self.key, self.result, self.snapshot, self.error = None, None, None, None, None

    def take(self, key):
        if self.thread is None:
            return None
        if key != self.key:
            self.cancel()
            return None
        self.thread.join()
        res, err = self.result, self.error
        self.thread, self.key, self.result, self.snapshot, self.error = None, None, None, None, None
        if err is not None:
            raise err
        return res

    def start(self, key, fn):
        import threading
        self.snapshot = np.empty(33, dtype=np.int32)
        _lib.load().lgcn_sampler_get_state(self.snapshot.ctypes.data_as(ctypes.c_void_p))
        self.key, self.result, self.error = key, None, None

        def run():
            try:
                self.result = fn()
            except Exception as e:                  # noqa: BLE001 — re-raised by take()
                self.error = e
        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()


_prefetch = _EpochPrefetch()


def sampler_seed(seed):
    _prefetch.cancel()
    _lib.load().lgcn_sampler_seed(ctypes.c_uint32(int(seed) & 0xffffffff))


def _epoch_sample(indptr, items, n_users, m_items, train_size, neg_ratio):
    per_user = train_size // n_users
    out = np.empty((n_users * per_user, 2 + neg_ratio), dtype=np.int32)
    rows = _lib.load().lgcn_sample_negative(n_users, m_items, train_size,
                                           indptr.ctypes.data_as(ctypes.c_void_p), items.ctypes.data_as(ctypes.c_void_p),
                                           neg_ratio, out.ctypes.data_as(ctypes.c_void_p))
    if rows < 0:
        raise RuntimeError(_lib.load().lgcn_last_error().decode())
    return out


def UniformSample_original(dataset, neg_ratio=1, prefetch_next=False):
    """One epoch of (user, pos, neg) triples, int32 [n_users * (trainDataSize // n_users), 2+neg_ratio],
    with the draw order of the reference's C++ sampler (code/sources/sampling.cpp:27-56).
    prefetch_next (used by Procedure.BPR_train_original): start drawing the following epoch's sample in the background."""
    if hasattr(dataset, 'allPos_csr'):
        indptr, items = dataset.allPos_csr()
    else:
        ap = dataset.allPos
        lens = np.fromiter((len(a) for a in ap), dtype=np.int64, count=len(ap))
        indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        items = np.concatenate([np.asarray(a, dtype=np.int32) for a in ap]) if len(ap) else np.zeros(0, np.int32)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    items = np.ascontiguousarray(items, dtype=np.int32)
    args = (indptr, items, int(dataset.n_users), int(dataset.m_items), int(dataset.trainDataSize), int(neg_ratio))
    key = (id(dataset), args[2], args[3], args[4], args[5], int(items.size))
    out = _prefetch.take(key)
    if out is None:
        out = _epoch_sample(*args)
    if prefetch_next:
        _prefetch.start(key, lambda: _epoch_sample(*args))
    return out


def UniformSample_original_python(dataset):
    """numpy fallback with the reference's semantics (code/utils.py:84-110): trainDataSize iid users from the global
    numpy RNG, one uniform positive each, negatives redrawn until they miss the user's positives; users without
    positives are skipped.  Kept for comparisons — the C sampler above is what the procedures use."""
    picked = np.random.randint(0, dataset.n_users, dataset.trainDataSize)
    positives = dataset.allPos
    triples = []
    for u in picked:
        mine = positives[u]
        if len(mine) == 0:
            continue
        p = np.random.choice(mine)
        n = np.random.randint(0, dataset.m_items)
        while n in mine:
            n = np.random.randint(0, dataset.m_items)
        triples.append((u, p, n))
    return np.asarray(triples)


# ---------------------------------------------------------------------------- helpers
def set_seed(seed):
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.manual_seed(seed)


def getFileName():
    if world.model_name == 'mf':
        file = f"mf-{world.dataset}-{world.config['latent_dim_rec']}.pth.tar"
    else:
        file = f"lgn-{world.dataset}-{world.config['lightGCN_n_layers']}-{world.config['latent_dim_rec']}.pth.tar"
    return os.path.join(world.PATH, file)


def minibatch(*tensors, **kwargs):
    batch_size = kwargs.get('batch_size', world.config['bpr_batch_size'])
    if len(tensors) == 1:
        tensor = tensors[0]
        for i in range(0, len(tensor), batch_size):
            yield tensor[i:i + batch_size]
    else:
        for i in range(0, len(tensors[0]), batch_size):
            yield tuple(x[i:i + batch_size] for x in tensors)


def shuffle(*arrays, **kwargs):
    """Same permutation stream as the reference: np.random.shuffle of arange(n) (code/utils.py:142-151)."""
    require_indices = kwargs.get('indices', False)
    if len(set(len(x) for x in arrays)) != 1:
        raise ValueError("All inputs to shuffle must have the same length.")
    shuffle_indices = np.arange(len(arrays[0]))
    np.random.shuffle(shuffle_indices)
    if len(arrays) == 1:
        result = arrays[0][shuffle_indices]
    else:
        result = tuple(x[shuffle_indices] for x in arrays)
    return (result, shuffle_indices) if require_indices else result


class timer:
    """Named accumulating timers: `with timer(name="Sample"): ...`; timer.dict() -> "|Sample:0.21|"."""
    TAPE = [-1]
    NAMED_TAPE = {}

    @staticmethod
    def get():
        return timer.TAPE.pop() if len(timer.TAPE) > 1 else -1

    @staticmethod
    def dict(select_keys=None):
        keys = timer.NAMED_TAPE.keys() if select_keys is None else select_keys
        return "|" + "|".join(f"{k}:{timer.NAMED_TAPE[k]:.2f}" for k in keys) + "|" if len(timer.NAMED_TAPE) else "|"

    @staticmethod
    def zero(select_keys=None):
        keys = list(timer.NAMED_TAPE.keys()) if select_keys is None else select_keys
        for k in keys:
            timer.NAMED_TAPE[k] = 0

    def __init__(self, tape=None, **kwargs):
        self.named = kwargs.get('name')
        if self.named:
            timer.NAMED_TAPE.setdefault(self.named, 0.)
            self.tape = None
        else:
            self.tape = tape or timer.TAPE

    def __enter__(self):
        self.start = time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.named:
            timer.NAMED_TAPE[self.named] += time() - self.start
        else:
            self.tape.append(time() - self.start)


# ---------------------------------------------------------------------------- metrics (host, numpy)
# Same definitions as the reference (code/utils.py:173-200,212-217); the procedures use the device kernel
# (ops.rank_metrics), these stay for API parity and as its cross-check.
def _discounts(k):
    return 1.0 / np.log2(np.arange(k) + 2.0)


def RecallPrecision_ATk(test_data, r, k):
    """r: {0,1} hit matrix [n_users, >=k] in rank order -> sums over users of recall@k and precision@k."""
    hits = np.asarray(r)[:, :k].sum(axis=1)
    n_relevant = np.fromiter((len(t) for t in test_data), dtype=np.float64, count=len(test_data))
    return {'recall': float(np.sum(hits / n_relevant)), 'precision': float(np.sum(hits) / k)}


def NDCGatK_r(test_data, r, k):
    """Sum over users of DCG@k / IDCG@k with IDCG over min(k, |ground truth|) ones (0 -> 1)."""
    r = np.asarray(r)
    if len(r) != len(test_data):
        raise AssertionError("one hit row per user expected")
    disc = _discounts(k)
    dcg = (r[:, :k] * disc).sum(axis=1)
    ideal = np.array([disc[:min(k, len(t))].sum() for t in test_data])
    ideal[ideal == 0.0] = 1.0
    return float(np.sum(dcg / ideal))


def getLabel(groundTruth, predictTopK):
    """Hit vector (float32) of a ranked list against a ground-truth collection (or a single item)."""
    truth = groundTruth if isinstance(groundTruth, (list, set, tuple, np.ndarray)) else [groundTruth]
    return np.isin(np.asarray(predictTopK).astype(np.int64), np.fromiter((int(x) for x in truth), dtype=np.int64)).astype(np.float32)
