"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (row partition bounds,
uneven in-place all-gather, batch sharding).  The SpMM itself is played by the numpy oracle — each
output row is produced by exactly one rank, so the partitioned result must equal the unpartitioned one
bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import lgcn_b200 as lg
        from oracle import lightgcn_oracle as orc
        g = lg.synth.make_graph('tiny', seed=1)
        nu, ni = g['n_users'], g['m_items']
        indptr, indices, vals, _, _ = orc.build_norm_adj(g['train_user'], g['train_item'], nu, ni)
        N, d, L = nu + ni, 16, 3
        rng = np.random.default_rng(0)
        E0 = rng.normal(0, 0.1, (N, d)).astype(np.float32)
        bounds = lg.engine.balanced_row_bounds(torch.from_numpy(indptr), world)
        assert bounds[0] == 0 and bounds[-1] == N and all(b0 <= b1 for b0, b1 in zip(bounds, bounds[1:]))
        nnz_parts = [int(indptr[b1] - indptr[b0]) for b0, b1 in zip(bounds, bounds[1:])]
        assert max(nnz_parts) - min(nnz_parts) <= 2 * int(np.diff(indptr).max())
        r0, r1 = bounds[rank], bounds[rank + 1]
        lp = indptr[r0:r1 + 1] - indptr[r0]
        li, lv = indices[indptr[r0]:indptr[r1]], vals[indptr[r0]:indptr[r1]]
        # partitioned propagation: local rows, then the uneven all-gather
        x = torch.from_numpy(E0.copy())
        acc = x.clone()
        for _ in range(L):
            y = torch.zeros_like(x)
            y[r0:r1] = torch.from_numpy(orc.spmm(lp, li, lv, x.numpy()))
            lg.engine.allgather_rows(y, bounds)
            acc += y
            x = y
        full = E0.copy(); acc_full = full.copy()
        for _ in range(L):
            full = orc.spmm(indptr, indices, vals, full); acc_full += full
        ok_prop = np.array_equal(acc.numpy(), acc_full)
        # batch sharding covers the batch exactly once
        lo, hi = lg.engine.shard_batch(2049, rank, world)
        cover = torch.zeros(2049); cover[lo:hi] = 1
        dist.all_reduce(cover)
        ok_shard = bool((cover == 1).all())
        # data-parallel gradient: sum of per-shard closed-form gradients == full-batch gradient
        users = rng.integers(0, nu, 64); pos = rng.integers(0, ni, 64); neg = rng.integers(0, ni, 64)
        out = (acc_full / (L + 1)).astype(np.float64)
        _, _, Gb, Gr = orc.bpr_loss(out, users, pos, neg, nu)
        lo, hi = lg.engine.shard_batch(64, rank, world)
        _, _, Gb_l, Gr_l = orc.bpr_loss(out, users[lo:hi], pos[lo:hi], neg[lo:hi], nu)
        Gl = torch.from_numpy((Gb_l + 1e-4 * Gr_l) * (hi - lo) / 64.0)     # local mean -> global mean
        dist.all_reduce(Gl)
        ok_dp = np.allclose(Gl.numpy(), Gb + 1e-4 * Gr, rtol=1e-12, atol=1e-15)
        q.put((rank, ok_prop, ok_shard, ok_dp))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_rowpart_and_dp_host_logic_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True, True, True), (1, True, True, True)]


def test_rebalance_by_time_equalises_predicted_time():
    """engine.rebalance_by_time: blocks whose measured time per unit of cost differs (item rows gather from the larger
    table) are re-cut so that every block is predicted to take the same time."""
    import lgcn_b200 as lg
    rng = np.random.default_rng(0)
    deg = rng.integers(1, 60, 4000)
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int64))
    speed = np.where(np.arange(4000) < 3000, 1.0, 3.0)                    # last quarter of the rows is 3x slower per non-zero
    true_prefix = np.concatenate([[0], np.cumsum(deg * speed)])
    bounds = lg.engine.balanced_row_bounds(indptr, 4)
    for _ in range(4):
        times = [true_prefix[b1] - true_prefix[b0] for b0, b1 in zip(bounds, bounds[1:])]
        bounds = lg.engine.rebalance_by_time(indptr, bounds, times)
        assert bounds[0] == 0 and bounds[-1] == 4000 and all(a <= b for a, b in zip(bounds, bounds[1:]))
    times = [true_prefix[b1] - true_prefix[b0] for b0, b1 in zip(bounds, bounds[1:])]
    assert max(times) < 1.1 * min(times), times


class _StubEngine:
    def __init__(self, rank, world):
        self.dist_mode, self.rank, self.world, self.group = 'rowpart', rank, world, None


class _StubModel:
    """Stands in for LightGCN in the host-side sharding logic of Procedure.rank_all: the 'ranking' of user u is a fixed
    function of u, so the multi-rank assembly can be compared with the single-process result."""
    def __init__(self, rank, world):
        self._engine = _StubEngine(rank, world)
        self.calls = 0

    def computer(self):
        self.calls += 1

    def rank_topk(self, users, k):
        idx = users.view(-1, 1) * 1000 + torch.arange(k, dtype=torch.int64).view(1, -1)
        return idx, idx.to(torch.float32)


class _StubDataset:
    def __init__(self, n):
        self._users = torch.arange(7, 7 + n, dtype=torch.int64)

    def test_csr(self):
        return self._users, None, None


def _procedure_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import lgcn_b200 as lg
        from lgcn_b200 import Procedure
        ok = True
        for n in (1, 5, 64, 101):                                   # fewer users than ranks, ragged shards, several tiles
            ds = _StubDataset(n)
            got = Procedure.rank_all(ds, _StubModel(rank, world), 4, user_tile=16)
            want, _ = _StubModel(0, 1).rank_topk(ds._users, 4)
            ok = ok and got.shape == (n, 4) and bool(torch.equal(got, want))
        # every rank must feed the same triples: identical passes, different raises on every rank
        same = torch.arange(30, dtype=torch.int64).view(3, 10)
        Procedure._assert_same_on_every_rank(same, None, "triples")
        diff = same.clone(); diff[1, 3] += rank
        raised = False
        try:
            Procedure._assert_same_on_every_rank(diff, None, "triples")
        except RuntimeError:
            raised = True
        q.put((rank, bool(ok), raised))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_procedure_sharding_host_logic_world2():
    """Procedure.rank_all under dist_mode='rowpart' (users sharded in contiguous blocks, lists all-gathered, trimmed) equals
    the single-process ranking, and the same-triples guard fires — over gloo, no GPU."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_procedure_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True, True), (1, True, True)]


class _FeatStub:
    """The attributes Engine.gather_columns reads (dist_mode='featpart' on `world` ranks, slices of d columns)."""
    def __init__(self, rank, world, d_full):
        self.dist_mode, self.rank, self.world, self.group = 'featpart', rank, world, None
        self.d_full, self.d, self.c0 = d_full, d_full // world, rank * (d_full // world)


def _feat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import lgcn_b200 as lg
        from oracle import lightgcn_oracle as orc
        g = lg.synth.make_graph('tiny', seed=1)
        nu, ni = g['n_users'], g['m_items']
        indptr, indices, vals, _, _ = orc.build_norm_adj(g['train_user'], g['train_item'], nu, ni)
        N, d, L = nu + ni, 16, 3
        rng = np.random.default_rng(0)
        E0 = rng.normal(0, 0.1, (N, d))
        st = _FeatStub(rank, world, d)
        c0, c1 = st.c0, st.c0 + st.d
        # (1) the propagation of a column slice needs nothing from the other ranks and equals the slice of the full propagation
        x = E0[:, c0:c1].copy(); acc = x.copy()
        for _ in range(L):
            x = orc.spmm(indptr, indices, vals, x); acc += x
        out_slice = acc / (L + 1)
        full = E0.copy(); acc_full = full.copy()
        for _ in range(L):
            full = orc.spmm(indptr, indices, vals, full); acc_full += full
        out_full = acc_full / (L + 1)
        ok_prop = np.array_equal(out_slice, out_full[:, c0:c1])
        # (2) Engine.gather_columns: every rank ends with the full table, columns in rank order
        got = lg.engine.Engine.gather_columns(st, torch.from_numpy(out_slice))
        ok_gather = got.shape == (N, d) and np.array_equal(got.numpy(), out_full)
        # (3) K2 split: the dot products are sums of the ranks' partial dot products; given the sums, the gradient rows of a
        #     slice are the slice of the full gradient
        users = rng.integers(0, nu, 64); pos = rng.integers(0, ni, 64); neg = rng.integers(0, ni, 64)
        U, P, Nn = out_slice[users], out_slice[nu + pos], out_slice[nu + neg]
        part = torch.from_numpy(np.stack([(U * P).sum(1), (U * Nn).sum(1), (U * U).sum(1), (P * P).sum(1), (Nn * Nn).sum(1)]))
        dist.all_reduce(part)
        pu, nuu, uu, pp, nn = part.numpy()
        z = pu - nuu
        bpr = float(np.mean(np.logaddexp(0.0, -z))); reg = float(0.5 * (uu + pp + nn).sum() / 64)
        bpr_ref, reg_ref, Gb, Gr = orc.bpr_loss(out_full, users, pos, neg, nu)
        ok_loss = abs(bpr - bpr_ref) < 1e-12 and abs(reg - reg_ref) < 1e-12
        sg = 1.0 / (1.0 + np.exp(z))                                  # sigmoid(-z)
        G = np.zeros((N, st.d))
        np.add.at(G, users, (sg[:, None] * (Nn - P) + 1e-4 * U) / 64)
        np.add.at(G, nu + pos, (-sg[:, None] * U + 1e-4 * P) / 64)
        np.add.at(G, nu + neg, (sg[:, None] * U + 1e-4 * Nn) / 64)
        ok_grad = np.allclose(G, (Gb + 1e-4 * Gr)[:, c0:c1], rtol=1e-10, atol=1e-15)
        q.put((rank, bool(ok_prop), bool(ok_gather), bool(ok_loss), bool(ok_grad)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_feature_partition_host_logic_world2():
    """dist_mode='featpart' over gloo, no GPU: a column slice propagates on its own, Engine.gather_columns reassembles the table,
    and the BPR loss/gradient follow from the all-reduced partial dot products (the algebra csrc/bpr.cu's split K2 relies on)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_feat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True, True, True, True), (1, True, True, True, True)], res
