"""GPU, N ranks over NCCL (skipped with fewer than 2 devices; LGCN_TEST_RANKS picks N, default 2): the data-parallel
and row-partitioned engines must reproduce the single-GPU / reference result, the row partition must really partition
memory, and the reference-facing procedures must run under it."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

NRANKS = int(os.environ.get('LGCN_TEST_RANKS', '2'))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _spawn(worker, world, *args, timeout=500):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
        assert p.exitcode == 0
    return sorted(q.get(timeout=10) for _ in range(world))


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import lgcn_b200 as lg
    lg.world.configure(device=f'cuda:{rank}')
    return dist, lg


def _tiny_model(lg, g, **over):
    cfg = dict(lg.world.config)
    cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), bpr_batch_size=len(g['users']),
               decay=float(g['decay']), lr=float(g['lr']), deterministic=True)
    cfg.update(over)
    ds = lg.InteractionDataset(int(g['n_users']), int(g['m_items']), g['train_user'], g['train_item'],
                               g['test_user'], g['test_item'], config=cfg)
    m = lg.LightGCN(cfg, ds)
    nu = int(g['n_users'])
    with torch.no_grad():
        m.embedding_user.weight.copy_(torch.from_numpy(g['E0'][:nu])); m.embedding_item.weight.copy_(torch.from_numpy(g['E0'][nu:]))
    return cfg, ds, m


def _worker(rank, world, port, q, mode):
    variant = mode
    p2p = mode not in ('rowpart_nccl',)
    part_mem = mode != 'rowpart_shared'
    multicast = mode != 'rowpart_unicast'
    mode = 'rowpart' if mode.startswith('rowpart') else mode
    dist, lg = _init(rank, world, port)
    try:
        g = load_golden('tiny')
        cfg, ds, m = _tiny_model(lg, g, dist_mode=mode, rowpart_p2p=p2p, rowpart_partition_memory=part_mem, rowpart_multicast=multicast)
        eng = m._engine
        assert eng.p2p == (mode == 'rowpart' and p2p)
        if mode == 'rowpart':
            N, d = eng.N, eng.d
            nnz_local = torch.tensor([eng.local.nnz], device='cuda')
            dist.all_reduce(nnz_local)
            full_nnz = 2 * np.unique(g['train_user'] * int(g['m_items']) + g['train_item']).size
            assert int(nnz_local.item()) == full_nnz                                     # the blocks tile the matrix
            if part_mem:                                                                 # ... and nobody holds more than its block
                assert eng.csr is None and m._csr is None and m.Graph is None
                assert eng.M.shape == (eng.r1 - eng.r0, d) and eng.V.shape == (eng.r1 - eng.r0, d)
                assert eng.local.indptr.numel() == eng.r1 - eng.r0 + 1
            assert eng.use_graph == bool(eng.p2p)
        B = len(g['users'])
        losses = []
        for s in range(3):
            sh = (s * 17) % B
            u, p, n = (torch.from_numpy(np.roll(g[k], sh)).long().cuda() for k in ('users', 'pos', 'neg'))
            if mode in ('dp', 'dp_idx'):
                lo, hi = lg.engine.shard_batch(B, rank, world)
                eng.step(u[lo:hi], p[lo:hi], n[lo:hi], B_global=B)
            else:
                eng.step(u, p, n)
            losses.append(float(eng.loss_to_host()[2]))
        eng.sync_params_for_read()
        params = eng.E0.cpu().numpy()
        with torch.no_grad():
            out = torch.cat(m.computer()).cpu().numpy()
        ok = (rel_err(params, g['params_after'][2]) < 1e-4 and rel_err(out, g['out_after']) < 1e-4
              and all(abs(a - b) < 1e-5 * abs(b) for a, b in zip(losses, g['step_losses'])))
        # every rank must hold the same replica
        chk = torch.from_numpy(params).cuda().double().sum()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same = all(float(x) == float(lst[0]) for x in lst)
        bitwise = True
        if mode == 'rowpart':
            # the partition changes who computes a row, not how: bit-identical to the single-GPU engine (deterministic mode)
            _, _, m1 = _tiny_model(lg, g, dist_mode=None)
            for s in range(3):
                sh = (s * 17) % B
                m1._engine.step(*(torch.from_numpy(np.roll(g[k], sh)).long().cuda() for k in ('users', 'pos', 'neg')))
            bitwise = bool(np.array_equal(m1._engine.E0.cpu().numpy(), params))
            Mf, Vf = eng.adam_state_full()
            bitwise = bitwise and torch.equal(Mf, m1._engine.M) and torch.equal(Vf, m1._engine.V)
            if eng._barrier is not None:
                eng._barrier.check()
        q.put((rank, bool(ok), bool(same), bool(bitwise), variant))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["dp", "dp_idx", "rowpart", "rowpart_unicast", "rowpart_shared", "rowpart_nccl"])
@pytest.mark.timeout(600)
def test_n_rank_training_matches_reference(mode):
    res = _spawn(_worker, NRANKS, mode)
    assert all(r[1] and r[2] and r[3] for r in res), res


def _procedure_worker(rank, world, port, q, tmp):
    dist, lg = _init(rank, world, port)
    try:
        g, t = load_golden('tiny_epochs'), load_golden('tiny')
        lg.world.configure(checkpoint_dir=os.path.join(tmp, f'r{rank}'), bpr_batch_size=int(g['batch']), topks=[20])

        def run(dist_mode):
            cfg = dict(lg.world.config)
            cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), decay=float(g['decay']), lr=float(g['lr']),
                       deterministic=True, dist_mode=dist_mode)
            ds = lg.InteractionDataset(int(t['n_users']), int(t['m_items']), t['train_user'], t['train_item'],
                                       t['test_user'], t['test_item'], config=cfg)
            lg.utils.set_seed(2020); lg.utils.sampler_seed(2020)
            m = lg.LightGCN(cfg, ds)
            bpr = lg.utils.BPRLoss(m, cfg)
            infos = [lg.Procedure.BPR_train_original(ds, m, bpr, e) for e in range(1, int(g['epochs']) + 1)]
            res = lg.Procedure.Test(ds, m, int(g['epochs']))
            sd = bpr.opt.state_dict()
            return m, infos, res, sd
        m, infos, res, sd = run('rowpart')
        m1, infos1, res1, sd1 = run(None)                       # the same procedures on this GPU alone
        P, P1 = m._engine.E0.cpu().numpy(), m1._engine.E0.cpu().numpy()
        same_as_single = (np.array_equal(P, P1) and all(np.array_equal(res[k], res1[k]) for k in res)
                          and [s.split('-')[0] for s in infos] == [s.split('-')[0] for s in infos1]
                          and all(torch.equal(sd['state'][i][k].cpu(), sd1['state'][i][k].cpu()) for i in (0, 1) for k in ('exp_avg', 'exp_avg_sq')))
        vs_ref = (abs(float(res['recall'][0]) - float(g['recall'][0])) <= 1e-4 and abs(float(res['ndcg'][0]) - float(g['ndcg'][0])) <= 1e-4
                  and rel_err(P, g['params']) < 1e-3)
        if m._engine._barrier is not None:
            m._engine._barrier.check()
        q.put((rank, bool(same_as_single), bool(vs_ref), float(res['recall'][0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_procedures_under_the_row_partition(tmp_path):
    """BPR_train_original x 20 epochs + Test under dist_mode='rowpart' (users sharded for the ranking, ranked lists gathered):
    metrics, losses, parameters and optimizer state equal the single-GPU run BIT FOR BIT, and the reference's fixed-epoch
    golden to 1e-4 (code/Procedure.py:28-83,127-206)."""
    res = _spawn(_procedure_worker, NRANKS, str(tmp_path))
    assert all(r[1] and r[2] for r in res), res


def _feat_worker(rank, world, port, q, mode):
    """dist_mode='featpart' through the reference-facing API: stageOne (host or device batches, captured graph or eager),
    computer(), state_dict() and the optimizer state, against the real reference's golden."""
    dist, lg = _init(rank, world, port)
    try:
        g = load_golden('tiny')
        cfg, ds, m = _tiny_model(lg, g, dist_mode='featpart', deterministic=(mode == 'deterministic'), cuda_graph=(mode != 'eager'))
        eng = m._engine
        d = int(g['d'])
        assert eng.d == d // world and eng.E0.shape == (eng.N, d // world) and eng.M.shape == (eng.N, d // world)
        assert eng.use_graph == (mode != 'eager') and eng.csr.nnz == 2 * np.unique(g['train_user'] * int(g['m_items']) + g['train_item']).size
        bpr = lg.utils.BPRLoss(m, cfg)
        B = len(g['users'])
        ok = True
        losses = []
        for s in range(3):
            sh = (s * 17) % B
            u, p, n = (torch.from_numpy(np.roll(g[k], sh)).long() for k in ('users', 'pos', 'neg'))
            if mode != 'host_batch':
                u, p, n = u.cuda(), p.cuda(), n.cuda()
            losses.append(bpr.stageOne(u, p, n))
            sd = m.state_dict()                                  # collective: gathers the trained column slices
            P = torch.cat([sd['embedding_user.weight'], sd['embedding_item.weight']]).cpu().numpy()
            ok = ok and rel_err(P, g['params_after'][s]) < 1e-4
        ok = ok and all(abs(a - b) < 1e-5 * abs(b) for a, b in zip(losses, g['step_losses']))
        with torch.no_grad():
            out = torch.cat(m.computer()).cpu().numpy()
        ok = ok and rel_err(out, g['out_after']) < 1e-4
        osd = bpr.opt.state_dict()
        m_cat = np.concatenate([osd['state'][0]['exp_avg'].cpu().numpy(), osd['state'][1]['exp_avg'].cpu().numpy()])
        v_cat = np.concatenate([osd['state'][0]['exp_avg_sq'].cpu().numpy(), osd['state'][1]['exp_avg_sq'].cpu().numpy()])
        ok = ok and rel_err(m_cat, g['exp_avg']) < 1e-4 and rel_err(v_cat, g['exp_avg_sq']) < 1e-4 and float(osd['state'][0]['step']) == 3.0
        # parameters written from outside reach the slices: load the step-1 state into a fresh model, one more step -> step 2
        cfg2, _, m2 = _tiny_model(lg, g, dist_mode='featpart', deterministic=(mode == 'deterministic'), cuda_graph=False)
        bpr2 = lg.utils.BPRLoss(m2, cfg2)
        u, p, n = (torch.from_numpy(g[k]).long().cuda() for k in ('users', 'pos', 'neg'))
        bpr2.stageOne(u, p, n)
        msd, osd2 = m2.state_dict(), bpr2.opt.state_dict()
        cfg3, _, m3 = _tiny_model(lg, g, dist_mode='featpart', deterministic=(mode == 'deterministic'), cuda_graph=False)
        bpr3 = lg.utils.BPRLoss(m3, cfg3)
        m3.load_state_dict(msd); bpr3.opt.load_state_dict(osd2)
        sh = 17 % B
        l2 = bpr3.stageOne(*(torch.from_numpy(np.roll(g[k], sh)).long().cuda() for k in ('users', 'pos', 'neg')))
        sd3 = m3.state_dict()
        P3 = torch.cat([sd3['embedding_user.weight'], sd3['embedding_item.weight']]).cpu().numpy()
        resumed = rel_err(P3, g['params_after'][1]) < 1e-4 and abs(l2 - g['step_losses'][1]) < 1e-5 * abs(g['step_losses'][1])
        # every rank computes the same loss bits (same records, same order)
        chk = torch.tensor(losses, device='cuda', dtype=torch.float64)
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same = all(torch.equal(x, lst[0]) for x in lst)
        eng._barrier.check()
        q.put((rank, bool(ok), bool(same), bool(resumed), mode))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["graph", "eager", "deterministic", "host_batch"])
@pytest.mark.timeout(600)
def test_feature_partition_matches_reference(mode):
    res = _spawn(_feat_worker, NRANKS, mode)
    assert all(r[1] and r[2] and r[3] for r in res), res


def _feat_procedure_worker(rank, world, port, q, tmp):
    dist, lg = _init(rank, world, port)
    try:
        g, t = load_golden('tiny_epochs'), load_golden('tiny')
        lg.world.configure(checkpoint_dir=os.path.join(tmp, f'r{rank}'), bpr_batch_size=int(g['batch']), topks=[20])

        def run(dist_mode):
            cfg = dict(lg.world.config)
            cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), decay=float(g['decay']), lr=float(g['lr']),
                       deterministic=True, dist_mode=dist_mode)
            ds = lg.InteractionDataset(int(t['n_users']), int(t['m_items']), t['train_user'], t['train_item'],
                                       t['test_user'], t['test_item'], config=cfg)
            lg.utils.set_seed(2020); lg.utils.sampler_seed(2020)
            m = lg.LightGCN(cfg, ds)
            bpr = lg.utils.BPRLoss(m, cfg)
            infos = [lg.Procedure.BPR_train_original(ds, m, bpr, e) for e in range(1, int(g['epochs']) + 1)]
            res = lg.Procedure.Test(ds, m, int(g['epochs']))
            sd = m.state_dict()
            return torch.cat([sd['embedding_user.weight'], sd['embedding_item.weight']]).cpu().numpy(), infos, res
        P, infos, res = run('featpart')
        P1, infos1, res1 = run(None)                            # the same procedures on this GPU alone
        close_to_single = (rel_err(P, P1) < 1e-4 and all(abs(float(res[k][0]) - float(res1[k][0])) <= 1e-4 for k in res)
                           and [s.split('-')[0] for s in infos] == [s.split('-')[0] for s in infos1])
        vs_ref = (abs(float(res['recall'][0]) - float(g['recall'][0])) <= 1e-4 and abs(float(res['ndcg'][0]) - float(g['ndcg'][0])) <= 1e-4
                  and rel_err(P, g['params']) < 1e-3)
        q.put((rank, bool(close_to_single), bool(vs_ref), float(res['recall'][0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_procedures_under_the_feature_partition(tmp_path):
    """BPR_train_original x 20 epochs + Test under dist_mode='featpart': the reference's fixed-epoch golden to 1e-4, the
    single-GPU run to rounding (the dot products are summed in a different order, nothing else changes)."""
    res = _spawn(_feat_procedure_worker, NRANKS, str(tmp_path))
    assert all(r[1] and r[2] for r in res) and len({r[3] for r in res}) == 1, res


def _barrier_worker(rank, world, port, q):
    dist, lg = _init(rank, world, port)
    try:
        from lgcn_b200.engine import map_peer_buffers
        lib = lg._lib.load()
        for p_dev in range(torch.cuda.device_count()):
            lib.lgcn_enable_peer_access(p_dev)
        flags = torch.zeros(64, dtype=torch.int32, device='cuda')
        box = torch.zeros(world, 1 << 16, dtype=torch.float32, device='cuda')      # box[src] is written by rank src
        torch.cuda.synchronize()
        bar = lg.ops.RankBarrier(flags, map_peer_buffers(flags), rank, world, timeout_ms=5000)
        peers = map_peer_buffers(box)
        dist.barrier()
        ok = True
        for it in range(1, 201):
            for p in range(world):                       # put: my slice of every rank's box (peer stores over NVLink)
                peers[p][rank].fill_(float(it * 10 + rank))
            bar()                                        # device-side: all puts of round `it` are visible after this
            got = box[:, ::4097].clone()
            bar()                                        # nobody overwrites before everybody has read
            want = torch.tensor([float(it * 10 + s) for s in range(world)], device='cuda').view(world, 1).expand_as(got)
            ok = ok and bool(torch.equal(got, want))
        bar.check()
        # the same inside a CUDA graph, replayed: the epoch counter lives on the device
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(6):
                bar()
        for _ in range(50):
            g.replay()
        torch.cuda.synchronize()
        bar.check()
        q.put((rank, ok, int(bar.epoch.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_device_side_rank_barrier():
    res = _spawn(_barrier_worker, NRANKS, timeout=200)
    assert all(r[1] for r in res) and len({r[2] for r in res}) == 1 and res[0][2] == 400 + 300, res
