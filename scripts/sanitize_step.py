"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
K4 build (+ row-block build), K1 (plain, segmented hub rows, column-blocked, Adam epilogue), K2 (atomic + deterministic),
pop-gate step, K5 sampler, K3 exact + tensor-core ranking, metrics.  Eager launches (no CUDA graph) so that every kernel is
seen.  Prints 'sanitize ok' at the end; the sanitizer's own summary is what matters.
    compute-sanitizer --tool memcheck python scripts/sanitize_step.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
lg.world.configure(device='cuda:0', checkpoint_dir='/tmp/lgcn_b200_sanitize', bpr_batch_size=256, topks=[20])
which = set((sys.argv[1] if len(sys.argv) > 1 else 'train,popgate,blocked,rowblock,tc').split(','))
g = lg.synth.make_graph('tiny', seed=3)
g['train_user'] = np.concatenate([g['train_user'], np.zeros(200, np.int64)])        # a hub user: segmented row
g['train_item'] = np.concatenate([g['train_item'], np.arange(200, dtype=np.int64)])
if 'train' in which:
    for det in (False, True):
        cfg = dict(lg.world.config); cfg.update(cuda_graph=False, deterministic=det, spmm_seg_len=32)
        ds = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'], config=cfg)
        lg.utils.set_seed(1); lg.utils.sampler_seed(1)
        m = lg.LightGCN(cfg, ds); bpr = lg.utils.BPRLoss(m, cfg)
        lg.Procedure.BPR_train_original(ds, m, bpr, 1)
        S = lg.ops.sample_bpr(ds.getCSRGraph(), ds.n_users, ds.m_items, ds.trainDataSize, 1, 0)
        bpr.stageOne(S[0, :256], S[1, :256], S[2, :256])
        lg.Procedure.Test(ds, m, 1)
        m.getUsersRating(torch.arange(10))
if 'popgate' in which:
    cfg = dict(lg.world.config); cfg.update(cuda_graph=False, use_pop_gate=True)
    ds = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'], config=cfg)
    m = lg.LightGCN(cfg, ds); bpr = lg.utils.BPRLoss(m, cfg)
    assert bpr.fused
    lg.Procedure.BPR_train_original(ds, m, bpr, 1)
    lg.Procedure.Test(ds, m, 1)
if 'blocked' in which:
    ds = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'])
    csr = ds.getCSRGraph()
    assert csr.block_plans(64, slab_bytes=100 * 64 * 4) is not None
    X = torch.randn(csr.n_rows, 64, device='cuda'); Y = torch.empty_like(X)
    lg.ops.spmm(csr, X, Y, 0.5, 0.5, [X])
if 'rowblock' in which:
    tu, ti = torch.from_numpy(g['train_user']).cuda(), torch.from_numpy(g['train_item']).cuda()
    b = lg.ops.RowBlockBuilder(g['n_users'], g['m_items'], lambda: iter([(tu, ti)]))
    blk = b.build(100, 600)
    X = torch.randn(b.N, 64, device='cuda'); Y = torch.empty(500, 64, device='cuda')
    lg.ops.spmm(blk, X, Y)
if 'tc' in which:
    nu, ni = 300, 16384 + 128
    U = torch.randn(nu, 64, device='cuda') * 0.1; V = torch.randn(ni, 64, device='cuda') * 0.1
    rng = np.random.default_rng(0)
    tu = np.repeat(np.arange(nu), 12).astype(np.int64); ti = rng.integers(0, ni, tu.size).astype(np.int64)
    csr = lg.ops.csr_build(torch.from_numpy(tu).cuda(), torch.from_numpy(ti).cuda(), nu, ni)
    users = torch.arange(nu, device='cuda')
    idx, val, redone = lg.ops.score_topk_tc(U, V, users, 20, csr.indptr, csr.indices, nu)
    idx2, val2 = lg.ops.score_topk(U, V, users, 20, csr.indptr, csr.indices, nu)
    assert torch.equal(idx, idx2) and torch.equal(val, val2)
torch.cuda.synchronize()
print('sanitize ok')
