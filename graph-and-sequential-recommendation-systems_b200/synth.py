"""Synthetic interaction graphs with the shapes BASELINE.json names.

User degrees come from a discretised log-normal calibrated to each dataset's train-degree
median/mean, items from a rank-popularity law, pairs are de-duplicated, every user keeps at least one
train item, and ~20 % of each user's interactions are held out as test.  Deterministic in `seed`.
"""
import numpy as np

# name: (n_users, m_items, train_edges, median_train_deg, mean_train_deg, max_train_deg, item_slope)
SHAPES = {
    'gowalla': (29858, 40981, 810128, 16, 27.1, 811, 0.55),
    'yelp2018': (31668, 38048, 1237259, 27, 39.1, 1800, 0.45),
    'amazon-book': (52643, 91599, 2380730, 28, 45.2, 10400, 0.49),
    'tiny': (300, 500, 6000, 14, 20.0, 120, 0.5),
}


def make_graph(name='yelp2018', seed=2020, scale=1.0, test_frac=0.2):
    """Returns dict(n_users, m_items, train_user, train_item, test_user, test_item) (int64 arrays)."""
    nu, ni, e_train, med, mean, dmax, slope = SHAPES[name]
    nu, ni = max(8, int(nu * scale)), max(16, int(ni * scale))
    e_train = int(e_train * scale)
    rng = np.random.default_rng(seed)
    # total (train+test) degree per user: log-normal with mean/median = exp(sigma^2/2)
    sigma = np.sqrt(2.0 * np.log(max(mean / med, 1.0001)))
    mu = np.log(med / (1.0 - test_frac))
    deg = np.exp(rng.normal(mu, sigma, nu))
    deg = np.clip(np.rint(deg), 2, min(dmax / (1.0 - test_frac), 0.5 * ni)).astype(np.int64)
    if name == 'amazon-book':                       # the real data has one hub user (~10.4 k train items)
        deg[rng.integers(nu)] = int(min(dmax / (1.0 - test_frac), 0.5 * ni))
    target_total = e_train / (1.0 - test_frac)
    deg = np.maximum(2, np.rint(deg * (target_total / deg.sum()))).astype(np.int64)
    # item popularity: p_r ~ (r + r0)^-a  (head slope ~ -0.45..-0.6 on log-log rank plots)
    ranks = np.arange(ni, dtype=np.float64)
    w = (ranks + 0.002 * ni + 1.0) ** (-(slope + 0.35))
    w[rng.random(ni) < 0.02] *= 1e-4                # a few % of items are (almost) never seen in train
    cdf = np.cumsum(w / w.sum())
    perm = rng.permutation(ni)                      # popularity is not aligned with the item id
    users = np.repeat(np.arange(nu, dtype=np.int64), (deg * 1.15 + 2).astype(np.int64))
    items = perm[np.minimum(np.searchsorted(cdf, rng.random(users.size)), ni - 1)]
    key = np.unique(users * ni + items)             # de-duplicate (u,i); sorted by user then item
    users, items = key // ni, key % ni
    # trim every user back to its target degree (random subset), then split train/test
    order = np.lexsort((rng.random(users.size), users))
    users, items = users[order], items[order]
    start = np.concatenate([[0], np.cumsum(np.bincount(users, minlength=nu))])
    rank_in_user = np.arange(users.size) - start[users]
    keep = rank_in_user < deg[users]
    users, items, rank_in_user = users[keep], items[keep], rank_in_user[keep]
    have = np.bincount(users, minlength=nu)
    n_test = np.minimum(np.floor(have * test_frac).astype(np.int64), np.maximum(have - 1, 0))
    is_test = rank_in_user < n_test[users]
    tr_u, tr_i, te_u, te_i = users[~is_test], items[~is_test], users[is_test], items[is_test]
    # users that ended up with no interaction at all get one random train item (sampler needs >= 1)
    missing = np.setdiff1d(np.arange(nu), tr_u, assume_unique=False)
    if missing.size:
        tr_u = np.concatenate([tr_u, missing])
        tr_i = np.concatenate([tr_i, rng.integers(0, ni, missing.size)])
    shuf = rng.permutation(tr_u.size)               # file order is not sorted order
    return dict(name=name, n_users=nu, m_items=ni, train_user=tr_u[shuf].astype(np.int64), train_item=tr_i[shuf].astype(np.int64),
                test_user=te_u.astype(np.int64), test_item=te_i.astype(np.int64))


def make_dataset(name='yelp2018', seed=2020, scale=1.0, config=None):
    from .dataloader import InteractionDataset
    g = make_graph(name, seed, scale)
    return InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'],
                              config=config, name=f"{name}-shape")


def write_txt(graph, path):
    """Write train.txt / test.txt in the reference's format (one 'uid items...' line per user)."""
    import os
    os.makedirs(path, exist_ok=True)
    for split in ('train', 'test'):
        u, i = graph[f'{split}_user'], graph[f'{split}_item']
        order = np.lexsort((i, u))
        u, i = u[order], i[order]
        with open(os.path.join(path, f'{split}.txt'), 'w') as f:
            if u.size:
                cuts = np.flatnonzero(np.diff(u)) + 1
                for uu, its in zip(u[np.concatenate([[0], cuts])], np.split(i, cuts)):
                    f.write(f"{uu} {' '.join(map(str, its.tolist()))}\n")


def powerlaw_chunks(n_users, m_items, n_edges, seed=2020, device=None, chunk=1 << 26):
    """BASELINE config 5 (scaled power-law graph) generated ON THE DEVICE, one chunk of edges at a time: yields
    (train_user, train_item) int64 tensors.  u = floor(n_users * r^2), i = floor(m_items * r^2.5): hub user ~
    n_edges/sqrt(n_users), hub item ~ n_edges/m_items^0.4.  Duplicate pairs are left in (K4 sums them like scipy's
    csr_matrix).  Deterministic in (seed, chunk): every call — and every rank — walks the same edge stream, which is what
    lets each rank of the row partition filter out its own rows without any rank ever holding the whole edge list."""
    import torch
    device = device or torch.device('cuda')
    gen = torch.Generator(device=device).manual_seed(seed)
    for lo in range(0, n_edges, chunk):
        hi = min(n_edges, lo + chunk)
        r1 = torch.rand(hi - lo, device=device, generator=gen, dtype=torch.float64)
        r2 = torch.rand(hi - lo, device=device, generator=gen, dtype=torch.float64)
        tu = (r1 * r1 * n_users).long().clamp_(0, n_users - 1)
        ti = (r2.pow(2.5) * m_items).long().clamp_(0, m_items - 1)
        del r1, r2
        yield tu, ti


def make_powerlaw_device(n_users, m_items, n_edges, seed=2020, device=None, chunk=1 << 26):
    """The whole edge list of powerlaw_chunks as two int64 device tensors (single-GPU path)."""
    import torch
    device = device or torch.device('cuda')
    tu = torch.empty(n_edges, dtype=torch.int64, device=device)
    ti = torch.empty(n_edges, dtype=torch.int64, device=device)
    lo = 0
    for cu, ci in powerlaw_chunks(n_users, m_items, n_edges, seed, device, chunk):
        tu[lo:lo + cu.numel()] = cu; ti[lo:lo + ci.numel()] = ci
        lo += cu.numel()
    return tu, ti


class DeviceGraphDataset:
    """Minimal dataset over device-resident edge arrays (no host copies): what LightGCN needs to build its graph.
    Training uses the device sampler (world.config['device_sampler']).  With `chunks` (a callable returning an iterable of
    edge chunks, e.g. functools.partial(powerlaw_chunks, ...)) instead of arrays, it is the graph source of the
    memory-partitioned row partition: getRowBlockBuilder() — no rank materialises the edge list or the whole CSR."""

    def __init__(self, n_users, m_items, train_user_dev=None, train_item_dev=None, seg_len=128, chunks=None, n_edges=None):
        self.n_users, self.m_items = int(n_users), int(m_items)
        self.trainDataSize = int(train_user_dev.numel()) if train_user_dev is not None else int(n_edges or 0)
        self._tu, self._ti, self._seg_len = train_user_dev, train_item_dev, seg_len
        self._chunks = chunks
        self._csr, self.Graph, self._builder = None, None, None
        self.testDict = {}

    def getCSRGraph(self):
        if self._csr is None:
            from . import ops
            if self._tu is None:
                raise RuntimeError("this dataset streams its edges in chunks: use getRowBlockBuilder() (dist_mode='rowpart')")
            self._csr = ops.csr_build(self._tu, self._ti, self.n_users, self.m_items, seg_len=self._seg_len)
            self._tu = self._ti = None            # the edge list is no longer needed
        return self._csr

    def getRowBlockBuilder(self):
        if self._builder is None:
            from . import ops
            chunks = self._chunks
            if chunks is None:
                tu, ti = self._tu, self._ti
                chunks = lambda: iter([(tu, ti)])         # noqa: E731
            self._builder = ops.RowBlockBuilder(self.n_users, self.m_items, chunks, seg_len=self._seg_len)
            self.trainDataSize = self._builder.n_edges
        return self._builder

    def getSparseGraph(self):
        if self.Graph is None:
            self.Graph = self.getCSRGraph().to_torch_sparse_csr()
        return self.Graph
