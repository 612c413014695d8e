"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): the data-parallel and row-partitioned
engines must reproduce the single-GPU / reference result."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, q):
    p2p = mode != 'rowpart_nccl'
    mode = 'rowpart' if mode.startswith('rowpart') else mode
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import lgcn_b200 as lg
        lg.world.configure(device=f'cuda:{rank}')
        g = load_golden('tiny')
        cfg = dict(lg.world.config)
        cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), bpr_batch_size=len(g['users']),
                   decay=float(g['decay']), lr=float(g['lr']), dist_mode=mode, deterministic=True, rowpart_p2p=p2p)
        ds = lg.InteractionDataset(int(g['n_users']), int(g['m_items']), g['train_user'], g['train_item'],
                                   g['test_user'], g['test_item'], config=cfg)
        m = lg.LightGCN(cfg, ds)
        nu = int(g['n_users'])
        with torch.no_grad():
            m.embedding_user.weight.copy_(torch.from_numpy(g['E0'][:nu])); m.embedding_item.weight.copy_(torch.from_numpy(g['E0'][nu:]))
        eng = m._engine
        assert eng.p2p == (mode == 'rowpart' and p2p)
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
        if mode == 'rowpart' and not eng.p2p:
            eng._allgather_rows(eng.E0)
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
        q.put((rank, bool(ok), bool(same), [float(x) for x in losses]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["dp", "dp_idx", "rowpart", "rowpart_nccl"])
@pytest.mark.timeout(600)
def test_two_rank_training_matches_reference(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert all(r[1] and r[2] for r in res), res
