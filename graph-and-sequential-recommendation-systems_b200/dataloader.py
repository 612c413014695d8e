"""Datasets with the reference's `BasicDataset` surface (code/dataloader.py:26-48) whose
`getSparseGraph()` runs the device CSR builder (K4) instead of the SciPy dok/lil path
(code/dataloader.py:203-246, 83.8 s on gowalla).

`Loader(config, path)` reads `<path>/train.txt` and `<path>/test.txt` in the reference's format
("uid item item ..." per line, lines without items skipped, ids sized by max over train and test:
code/dataloader.py:82-119).  `InteractionDataset` takes the same information as arrays (synthetic
graphs, fixtures).  Both expose: n_users, m_items, trainDataSize, testDict, allPos, trainUser,
trainItem, users_D, items_D, UserItemNet, getUserPosItems, getUserItemFeedback, getSparseGraph —
plus getCSRGraph() (the int32 CSR handle the kernels use).
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops, world


class BasicDataset(Dataset):
    """Abstract surface, same members as the reference's BasicDataset."""

    @property
    def n_users(self): raise NotImplementedError

    @property
    def m_items(self): raise NotImplementedError

    @property
    def trainDataSize(self): raise NotImplementedError

    @property
    def testDict(self): raise NotImplementedError

    @property
    def allPos(self): raise NotImplementedError

    def getUserItemFeedback(self, users, items): raise NotImplementedError

    def getUserPosItems(self, users): raise NotImplementedError

    def getSparseGraph(self): raise NotImplementedError


class InteractionDataset(BasicDataset):
    def __init__(self, n_users, m_items, train_user, train_item, test_user, test_item, config=None, name='synthetic'):
        config = world.config if config is None else config
        self.name = name
        self.n_user, self.m_item = int(n_users), int(m_items)
        self.trainUser = np.ascontiguousarray(train_user, dtype=np.int64)
        self.trainItem = np.ascontiguousarray(train_item, dtype=np.int64)
        self.testUser = np.ascontiguousarray(test_user, dtype=np.int64)
        self.testItem = np.ascontiguousarray(test_item, dtype=np.int64)
        if self.trainUser.shape != self.trainItem.shape or self.testUser.shape != self.testItem.shape:
            raise ValueError("user/item arrays differ in length")
        for arr, hi, what in ((self.trainUser, self.n_user, 'train user'), (self.trainItem, self.m_item, 'train item'),
                              (self.testUser, self.n_user, 'test user'), (self.testItem, self.m_item, 'test item')):
            if arr.size and (arr.min() < 0 or arr.max() >= hi):
                raise ValueError(f"{what} id outside [0,{hi})")
        self.traindataSize = int(self.trainUser.size)
        self.testDataSize = int(self.testUser.size)
        self.trainUniqueUsers = np.unique(self.trainUser)
        self.testUniqueUsers = np.unique(self.testUser)
        self.seg_len = int(config.get('spmm_seg_len', ops.DEFAULT_SEG_LEN))
        self.split = config.get('A_split', False)
        if self.split:
            raise NotImplementedError("A_split folding is dead code in the reference (SURVEY.md §2 row 2); "
                                      "use dist_mode='rowpart' to partition the adjacency across GPUs")
        self.Graph = None
        self._csr = None
        self._allPos = None
        self._testDict = None
        self._test_csr = None
        self._degrees = None
        self._uinet = None

    # ---- basic properties -----------------------------------------------------------------
    @property
    def n_users(self): return self.n_user

    @property
    def m_items(self): return self.m_item

    @property
    def trainDataSize(self): return self.traindataSize

    @property
    def testDict(self):
        """{user: [test items in file order]} — insertion order = first appearance (code/dataloader.py:165-171)."""
        if self._testDict is None:
            d = {}
            for u, i in zip(self.testUser.tolist(), self.testItem.tolist()):
                d.setdefault(u, []).append(i)
            self._testDict = d
        return self._testDict

    @property
    def allPos(self):
        """Per-user sorted unique train items (what UserItemNet[u].indices yields, code/dataloader.py:178-180)."""
        if self._allPos is None:
            key = np.unique(self.trainUser * self.m_item + self.trainItem)
            users, items = key // self.m_item, (key % self.m_item).astype(np.int32)
            counts = np.bincount(users, minlength=self.n_user)
            self._allPos_indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            self._allPos_items = np.ascontiguousarray(items)
            self._allPos = np.split(items, self._allPos_indptr[1:-1])
        return self._allPos

    def allPos_csr(self):
        """(indptr int64[n_users+1], items int32[nnz]) host arrays behind allPos (for the C sampler)."""
        self.allPos
        return self._allPos_indptr, self._allPos_items

    def _deg(self):
        if self._degrees is None:
            du = np.bincount(self.trainUser, minlength=self.n_user).astype(np.float32)
            di = np.bincount(self.trainItem, minlength=self.m_item).astype(np.float32)
            self._degrees = (du, di)
        return self._degrees

    @property
    def users_D(self):
        d = self._deg()[0].copy(); d[d == 0.] = 1.; return d      # code/dataloader.py:139-140

    @property
    def items_D(self):
        d = self._deg()[1].copy(); d[d == 0.] = 1.; return d      # code/dataloader.py:141-142

    @property
    def UserItemNet(self):
        """SciPy CSR of the interaction matrix, for API compatibility only (not used by the kernels)."""
        if self._uinet is None:
            import scipy.sparse as sp
            self._uinet = sp.csr_matrix((np.ones(self.trainUser.size, dtype=np.float32), (self.trainUser, self.trainItem)),
                                        shape=(self.n_user, self.m_item))
        return self._uinet

    def getUserItemFeedback(self, users, items):
        return np.array(self.UserItemNet[users, items]).astype('uint8').reshape((-1,))

    def getUserPosItems(self, users):
        ap = self.allPos
        return [ap[u] for u in users]

    # ---- adjacency --------------------------------------------------------------------------
    def getCSRGraph(self):
        """Normalised symmetric bipartite adjacency as an int32 CSR on the device (K4)."""
        if self._csr is None:
            dev = world.device
            if dev.type != 'cuda':
                raise RuntimeError("building the adjacency needs a CUDA device (K4 has no CPU path)")
            tu = torch.from_numpy(self.trainUser).to(dev)
            ti = torch.from_numpy(self.trainItem).to(dev)
            self._csr = ops.csr_build(tu, ti, self.n_user, self.m_item, seg_len=self.seg_len)
        return self._csr

    def getRowBlockBuilder(self):
        """Graph source of the memory-partitioned row partition (dist_mode='rowpart'): an ops.RowBlockBuilder over the
        train edges; each rank assembles only the rows it owns (K4 on the owned keys)."""
        if getattr(self, '_builder', None) is None:
            dev = world.device
            if dev.type != 'cuda':
                raise RuntimeError("building the adjacency needs a CUDA device (K4 has no CPU path)")

            def chunks():
                yield torch.from_numpy(self.trainUser).to(dev), torch.from_numpy(self.trainItem).to(dev)
            self._builder = ops.RowBlockBuilder(self.n_user, self.m_item, chunks, seg_len=self.seg_len, device=dev)
        return self._builder

    def train_mask_csr(self):
        """(indptr int32[n_users+1], items int32[nnz]) on the device: every user's sorted train items (= allPos), the
        mask of the full-ranking evaluation when no rank holds the whole adjacency."""
        if getattr(self, '_mask_csr', None) is None:
            indptr, items = self.allPos_csr()
            dev = world.device
            self._mask_csr = (torch.from_numpy(indptr.astype(np.int32)).to(dev), torch.from_numpy(np.ascontiguousarray(items, dtype=np.int32)).to(dev))
        return self._mask_csr

    def getSparseGraph(self):
        """D^-1/2 [[0,R],[R^T,0]] D^-1/2 as a torch sparse tensor on world.device (accepted by
        torch.sparse.mm like the reference's coalesced COO); shares storage with getCSRGraph()."""
        if self.Graph is None:
            self.Graph = self.getCSRGraph().to_torch_sparse_csr()
        return self.Graph

    def test_csr(self):
        """Test interactions as a device CSR over testDict's users (key order), items sorted: the
        ground truth the on-device metric kernel searches.  Returns (users int64, indptr, indices)."""
        if self._test_csr is None:
            # users in testDict key order = order of first appearance in the test file; items sorted per user
            uniq, first = np.unique(self.testUser, return_index=True)
            users = uniq[np.argsort(first, kind='stable')]
            rank_of = np.empty(self.n_user, dtype=np.int64); rank_of[users] = np.arange(users.size)
            r = rank_of[self.testUser]
            order = np.lexsort((self.testItem, r))
            items = self.testItem[order].astype(np.int32)
            indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=users.size))]).astype(np.int32)
            dev = world.device
            self._test_csr = (torch.from_numpy(users.astype(np.int64)).to(dev), torch.from_numpy(indptr).to(dev),
                              torch.from_numpy(items).to(dev))
        return self._test_csr

    def __getitem__(self, idx):
        return self.trainUniqueUsers[idx]

    def __len__(self):
        return len(self.trainUniqueUsers)


def _parse_interactions(path):
    """'uid i1 i2 ...' lines -> (users int64[E], items int64[E]) in file order; lines without items are skipped.
    Native single-pass parser (lgcn_parse_interactions) instead of the reference's per-line Python loop."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    raw = os.fsencode(path)
    size = os.path.getsize(path) if os.path.exists(path) else 0
    if size <= (256 << 20):
        cap = size // 2 + 1                     # an item token and its separator take at least two bytes: one pass
    else:
        cap = lib.lgcn_parse_interactions(raw, None, None, 0, None, None)     # size the arrays first
        if cap < 0:
            raise RuntimeError(lib.lgcn_last_error().decode())
    users = np.empty(cap, dtype=np.int64); items = np.empty(cap, dtype=np.int64)
    n = lib.lgcn_parse_interactions(raw, users.ctypes.data_as(ctypes.c_void_p), items.ctypes.data_as(ctypes.c_void_p), cap, None, None)
    if n < 0:
        raise RuntimeError(lib.lgcn_last_error().decode())
    if n > cap:
        raise RuntimeError(f"{path} changed while it was being read")
    users, items = users[:n].copy(), items[:n].copy()
    return users, items


class Loader(InteractionDataset):
    """File-backed dataset: Loader(config, path) with path = <data>/<dataset> (code/dataloader.py:62-66)."""

    def __init__(self, config=None, path=None):
        config = world.config if config is None else config
        if path is None:
            path = os.path.join(world.DATA_PATH, world.dataset)
        self.path = path
        world.cprint(f'loading [{self.path}]')
        tu, ti = _parse_interactions(os.path.join(path, 'train.txt'))
        su, si = _parse_interactions(os.path.join(path, 'test.txt'))
        n_users = int(max(tu.max(initial=-1), su.max(initial=-1))) + 1
        m_items = int(max(ti.max(initial=-1), si.max(initial=-1))) + 1
        super().__init__(n_users, m_items, tu, ti, su, si, config=config, name=os.path.basename(path.rstrip('/')))
        print(f"{self.trainDataSize} interactions for training")
        print(f"{self.testDataSize} interactions for testing")
        print(f"{self.name} Sparsity : {(self.trainDataSize + self.testDataSize) / self.n_users / self.m_items:.12f}")
        print(f"{self.name} is ready to go")
