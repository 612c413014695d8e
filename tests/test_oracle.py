"""CPU: pins the oracle against outputs of the real reference (tests/golden/*.npz, made by
oracle/gen_golden.py) and against the one known answer in the reference's run artefacts."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import lightgcn_oracle as orc


def _adj(g):
    return orc.build_norm_adj(g['train_user'], g['train_item'], int(g['n_users']), int(g['m_items']))


def test_graph_structure_bit_exact(golden):
    indptr, indices, vals, deg, dinv = _adj(golden)
    N = int(golden['n_users']) + int(golden['m_items'])
    ref_indptr = np.concatenate([[0], np.cumsum(np.bincount(golden['adj_row'], minlength=N))])
    assert np.array_equal(indptr, ref_indptr)
    assert np.array_equal(indices, golden['adj_col'])
    # values: numpy's float32 power(-0.5) is not correctly rounded -> <= 2 ulp
    assert np.max(np.abs(vals - golden['adj_val']) / golden['adj_val']) < 3e-7
    nu = int(golden['n_users'])
    ud, idg = deg[:nu].copy(), deg[nu:].copy()
    ud[ud == 0] = 1; idg[idg == 0] = 1
    assert np.array_equal(ud, golden['users_D']) and np.array_equal(idg, golden['items_D'])


def test_spmm_paths_agree(golden):
    indptr, indices, vals, _, _ = _adj(golden)
    X = golden['E0'].astype(np.float64)
    a = orc.spmm(indptr, indices, vals.astype(np.float64), X)
    b = orc.spmm_scipy(indptr, indices, vals.astype(np.float64), X)
    assert rel_err(a, b) < 1e-13


def test_propagation(golden):
    indptr, indices, vals, _, _ = _adj(golden)
    out = orc.propagate(indptr, indices, vals, golden['E0'], int(golden['L']))
    assert rel_err(out, golden['out']) < 2e-6          # fp64 oracle vs the reference's fp32


def test_bpr_loss_and_gradient(golden):
    indptr, indices, vals, _, _ = _adj(golden)
    nu, L, decay = int(golden['n_users']), int(golden['L']), float(golden['decay'])
    out = orc.propagate(indptr, indices, vals, golden['E0'], L)
    bpr, reg, Gb, Gr = orc.bpr_loss(out, golden['users'], golden['pos'], golden['neg'], nu)
    assert abs(bpr - float(golden['loss'])) < 2e-6 * abs(float(golden['loss']))
    assert abs(reg - float(golden['reg'])) < 2e-6 * abs(float(golden['reg']))
    g0 = orc.propagate_backward(indptr, indices, vals, Gb + decay * Gr, L)
    assert rel_err(g0, golden['grad']) < 2e-5          # reference autograd in fp32


def test_three_adam_steps(golden):
    indptr, indices, vals, _, _ = _adj(golden)
    nu, L = int(golden['n_users']), int(golden['L'])
    decay, lr = float(golden['decay']), float(golden['lr'])
    p = golden['E0'].astype(np.float64); m = np.zeros_like(p); v = np.zeros_like(p)
    B = len(golden['users'])
    for s in range(3):
        sh = (s * 17) % B
        u, pp, nn = (np.roll(golden[k], sh) for k in ('users', 'pos', 'neg'))
        loss, p, m, v, _ = orc.train_step(indptr, indices, vals, p, m, v, s + 1, u, pp, nn, nu, L, decay, lr)
        assert abs(loss - golden['step_losses'][s]) < 5e-6 * abs(golden['step_losses'][s])
        assert rel_err(p, golden['params_after'][s]) < 5e-5
    assert rel_err(m, golden['exp_avg']) < 1e-4
    assert rel_err(v, golden['exp_avg_sq']) < 1e-4


def test_scores_topk_metrics(golden):
    indptr, indices, vals, _, _ = _adj(golden)
    nu, ni, L = int(golden['n_users']), int(golden['m_items']), int(golden['L'])
    out = golden['out_after']                         # the reference's own embeddings at Test time
    users = golden['test_users']
    k = int(max(golden['topks']))
    # fp32 FMA-chain scores (C oracle) vs the reference's matmul
    dense = orc.score_dense_exact(out[:nu], out[nu:], users)
    S64 = out[:nu][users].astype(np.float64) @ out[nu:].astype(np.float64).T
    assert rel_err(dense, S64) < 1e-6
    idx, val = orc.score_topk_exact(out[:nu], out[nu:], users, k, indptr, indices, nu)
    # selection exactness: the C top-k is the stable descending sort of its own masked scores
    masked = dense.copy()
    for b, u in enumerate(users):
        masked[b, indices[indptr[u]:indptr[u + 1]] - nu] = -1024.0
    assert np.array_equal(idx, orc.topk_stable(masked, k))
    assert np.array_equal(val, np.take_along_axis(masked, idx, axis=1))
    # against the reference's torch.topk: may differ only at (near-)ties
    ref = golden['topk']
    diff = idx != ref
    if diff.any():
        rows, cols = np.nonzero(diff)
        a = np.take_along_axis(masked, idx, 1)[rows, cols]; b = masked[rows, ref[rows, cols]]
        assert np.all(np.abs(a - b) <= 1e-6 * (np.abs(a) + 1e-3)), "top-k differs from the reference away from ties"
    gt = {}
    for u, i in zip(golden['test_user'].tolist(), golden['test_item'].tolist()):
        gt.setdefault(u, []).append(i)
    m = orc.metrics_at_k(idx, [gt[int(u)] for u in users], [int(x) for x in golden['topks']])
    for name in ('precision', 'recall', 'ndcg'):
        assert np.allclose(m[name], golden[name], rtol=0, atol=1e-6), (name, m[name], golden[name])


def test_edge_case_k_exceeds_unmasked_items():
    """edge.npz user 0 has 45 train items of 60: top-20 must contain masked (-1024) items, lowest id first."""
    from conftest import load_golden
    g = load_golden('edge')
    indptr, indices, vals, _, _ = _adj(g)
    nu = int(g['n_users'])
    out = g['out_after']
    idx, val = orc.score_topk_exact(out[:nu], out[nu:], np.array([0]), 20, indptr, indices, nu)
    n_masked = int((val[0] == -1024.0).sum())
    assert n_masked == 20 - (60 - 45)
    tail = idx[0][val[0] == -1024.0]
    train0 = np.sort(indices[indptr[0]:indptr[1]] - nu)
    assert np.array_equal(tail, train0[:n_masked])


def test_gowalla_step0_kat(gowalla):
    """The reference's only known answer (author's tfevents): seed 2020 -> construct ->
    Test at step 0 on gowalla gives P/R/NDCG@20 = 0.0001875544 / 0.0005374941 / 0.00040836."""
    import torch
    nu, ni = int(gowalla['n_users']), int(gowalla['m_items'])
    tu = np.repeat(np.arange(nu), np.diff(gowalla['train_indptr'])).astype(np.int64)
    ti = gowalla['train_items'].astype(np.int64)
    indptr, indices, vals, _, _ = orc.build_norm_adj(tu, ti, nu, ni)
    assert indices.size == 1620256
    torch.manual_seed(2020)                            # utils.set_seed -> model.py:57-60 RNG order
    eu = torch.nn.Embedding(nu, 64); ei = torch.nn.Embedding(ni, 64)
    torch.nn.init.normal_(eu.weight, std=0.1); torch.nn.init.normal_(ei.weight, std=0.1)
    E0 = torch.cat([eu.weight, ei.weight]).detach().numpy()
    out = orc.propagate(indptr, indices, vals, E0, 3, dtype=np.float32, fast=True)
    tip = gowalla['test_indptr']; titems = gowalla['test_items'].astype(np.int64)
    users = gowalla['test_users'].astype(np.int64)
    hits_p = hits_r = ndcg = 0.0
    disc = 1.0 / np.log2(np.arange(2, 22))
    V = out[nu:]
    for lo in range(0, users.size, 2000):
        ub = users[lo:lo + 2000]
        S = out[:nu][ub] @ V.T
        for b, u in enumerate(ub):
            S[b, indices[indptr[u]:indptr[u + 1]] - nu] = -1024.0
        part = np.argpartition(-S, 20, axis=1)[:, :20]
        ps = np.take_along_axis(S, part, 1)
        order = np.lexsort((part, -ps), axis=1)
        top = np.take_along_axis(part, order, 1)
        for b in range(ub.size):
            gt = titems[tip[lo + b]:tip[lo + b + 1]]
            r = np.isin(top[b], gt).astype(np.float64)
            hits_p += r.sum() / 20; hits_r += r.sum() / gt.size
            ndcg += (r * disc).sum() / disc[:min(20, gt.size)].sum()
    n = users.size
    assert abs(hits_p / n - float(gowalla['kat_precision'])) < 2e-9
    assert abs(hits_r / n - float(gowalla['kat_recall'])) < 2e-9
    assert abs(ndcg / n - float(gowalla['kat_ndcg'])) < 2e-7      # two author runs differ in the 4th digit


def _replay_epochs_with_port(g, train_user, train_item, test_user, test_item, nu, ni, epochs, B):
    """Replays the reference's epoch loop (code/Procedure.py:28-83) with oracle/ref_port on the CPU: the library's C
    sampler (bit-identical to sources/sampling.cpp, tests/test_host.py) seeded 2020, numpy shuffle, minibatch."""
    import torch
    import lgcn_b200 as lg
    from oracle import ref_port
    ds = lg.InteractionDataset(nu, ni, train_user, train_item, test_user, test_item, config=dict(lg.world.config))
    graph, _, _ = ref_port.build_graph(ds.trainUser, ds.trainItem, nu, ni)
    lg.utils.set_seed(2020)
    lg.utils.sampler_seed(2020)
    m = ref_port.RefLightGCN(nu, ni, int(g['d']), int(g['L']), graph)
    bpr = ref_port.RefBPRLoss(m, float(g['decay']), float(g['lr']))
    losses, first_S = [], None
    for _ in range(epochs):
        S = lg.utils.UniformSample_original(ds)
        if first_S is None:
            first_S = S.copy()
        u, p, n = (torch.from_numpy(S[:, j].astype(np.int64)) for j in range(3))
        u, p, n = lg.utils.shuffle(u, p, n)
        tot = 0.0
        for bu, bp, bn in lg.utils.minibatch(u, p, n, batch_size=B):
            tot += bpr.stageOne(bu, bp, bn)
        losses.append(tot / (len(u) // B + 1))
    return ds, m, np.array(losses), first_S


def test_fixed_epochs_port_reproduces_the_reference_procedure():
    """north_star: Recall@20 / NDCG@20 within 1e-4 after a fixed number of epochs.  tests/golden/tiny_epochs.npz holds
    what the REAL reference's BPR_train_original x 20 + Test produced (oracle/gen_epochs_golden.py); the CPU port with
    the library's sampler/shuffle/minibatch glue must land on the same triples, losses, parameters and metrics."""
    import torch
    from conftest import load_golden
    from oracle import ref_port
    torch.set_num_threads(1)
    g, t = load_golden('tiny_epochs'), load_golden('tiny')
    E, B = int(g['epochs']), int(g['batch'])
    ds, m, losses, first_S = _replay_epochs_with_port(g, t['train_user'], t['train_item'], t['test_user'], t['test_item'],
                                                      int(t['n_users']), int(t['m_items']), E, B)
    assert np.array_equal(first_S[:64], g['first_epoch_triples_head'])
    assert np.array_equal(first_S.astype(np.int64).sum(axis=0), g['first_epoch_triples_sum'])
    assert np.allclose(np.round(losses, 3), g['epoch_loss_3dp'], atol=1.01e-3)
    P = torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().numpy()
    assert rel_err(P, g['params']) < 1e-6
    res, _ = ref_port.ref_test(m, ds.testDict, ds.allPos, [20])
    for k in ('precision', 'recall', 'ndcg'):
        assert abs(float(res[k][0]) - float(g[k][0])) < 1e-9, (k, res[k], g[k])
