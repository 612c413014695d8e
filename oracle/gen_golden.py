"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference.

Runs only in the build container (needs /root/reference).  For each case it writes a small dataset in
the reference's train.txt/test.txt format to a temp dir, imports the reference's own dataloader /
model / utils / Procedure (with sys.argv preset and the two procedure shims of SURVEY.md §8c), runs
graph build -> computer() -> bpr_loss -> 3 x stageOne -> Test, and stores inputs and outputs.  It also
asserts that oracle/ref_port.py reproduces the reference exactly, which is what lets the port stand in
for the reference on the GPU box.

usage:  python oracle/gen_golden.py            (all cases, one subprocess each: `world` is per-process)
        python oracle/gen_golden.py --case tiny
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = '/root/reference/LightGCN_work/code'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

CASES = {
    #        users items d  L  topks      batch
    'tiny': (300, 500, 64, 3, '[20]', 256),
    'edge': (40, 60, 32, 2, '[5,20]', 64),
    # model variants of SURVEY.md §8f #4 (same data as 'tiny'): popularity gate / item-item smoothing
    'popgate': (300, 500, 64, 3, '[20]', 256),
    'i2i': (300, 500, 64, 3, '[20]', 256),
}


def make_case(name):
    rng = np.random.default_rng(11 if name == 'edge' else 7)
    nu, ni = CASES[name][:2]
    train, test = {}, {}
    if name != 'edge':
        for u in range(nu):
            deg = int(np.clip(np.rint(np.exp(rng.normal(2.6, 0.8))), 3, 120))
            items = rng.choice(ni - 20, size=min(deg, ni - 20), replace=False)     # last 20 items: never in train
            n_test = max(1, len(items) // 5)
            test[u] = [int(x) for x in items[:n_test]]
            train[u] = [int(x) for x in items[n_test:]]
        test[3].append(ni - 1)            # an item that only appears in test (zero-degree item row)
    else:
        for u in range(nu):
            deg = 50 if u == 0 else int(rng.integers(2, 9))          # user 0: k=20 > #unmasked items (60-45)
            items = rng.choice(ni, size=deg, replace=False)
            n_test = 5 if u == 0 else 1
            test[u] = [int(x) for x in items[:n_test]]
            train[u] = [int(x) for x in items[n_test:]]
        train[5] = train[5] + [train[5][0]]                          # duplicate pair -> weight 2 (A3)
        del test[7]                                                  # a user without test items (A2)
    return nu, ni, train, test


def write_case(path, train, test):
    os.makedirs(path, exist_ok=True)
    for fname, d in (('train.txt', train), ('test.txt', test)):
        with open(os.path.join(path, fname), 'w') as f:
            for u, items in d.items():
                f.write(f"{u} {' '.join(map(str, items))}\n")


def run_case(name):
    import torch
    nu, ni, d, L, topks, B = CASES[name]
    _, _, train, test = make_case(name)
    tmp = tempfile.mkdtemp(prefix=f'golden_{name}_')
    data_dir = os.path.join(tmp, 'data', name)
    write_case(data_dir, train, test)
    sys.argv = ['x', '--dataset', name, '--tensorboard', '0', '--checkpoint_dir', os.path.join(tmp, 'ckpt'),
                '--recdim', str(d), '--layer', str(L), '--topks', topks, '--bpr_batch', str(B)]
    extra = {}
    if name == 'popgate':
        sys.argv += ['--use_pop_gate']
    if name == 'i2i':
        import scipy.sparse as sp
        r2 = np.random.default_rng(5)
        dense = (r2.random((ni, ni)) < 0.02) * r2.random((ni, ni))
        np.fill_diagonal(dense, 0.0)
        i2i = sp.csr_matrix(dense.astype(np.float32))
        i2i_path = os.path.join(tmp, 'i2i.npz')
        sp.save_npz(i2i_path, i2i)
        sys.argv += ['--use_item_item', '--i2i_path', i2i_path, '--i2i_alpha', '0.3']
        extra = dict(i2i_indptr=i2i.indptr.astype(np.int32), i2i_indices=i2i.indices.astype(np.int32), i2i_data=i2i.data.astype(np.float32), i2i_alpha=0.3)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import world                                         # noqa: E402  (the reference's)
    world.device = torch.device('cpu')
    import dataloader, model, utils, Procedure           # noqa: E402

    # shims for the two breakages on the procedure path
    def minibatch(*tensors, **kwargs):
        bs = kwargs.get('batch_size', world.config['bpr_batch_size'])
        if len(tensors) == 1:
            for i in range(0, len(tensors[0]), bs):
                yield tensors[0][i:i + bs]
        else:
            for i in range(0, len(tensors[0]), bs):
                yield tuple(x[i:i + bs] for x in tensors)
    utils.minibatch = minibatch

    torch.set_num_threads(1)
    ds = dataloader.Loader(world.config, path=data_dir)
    utils.set_seed(2020)
    m = model.LightGCN(world.config, ds)
    variant = name in ('popgate', 'i2i')
    if name == 'popgate':
        extra = {'sd_' + k.replace('.', '__'): v.detach().numpy().copy() for k, v in m.state_dict().items() if not k.startswith('embedding')}
    E0 = torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().numpy().copy()
    g = m.Graph.coalesce()
    out_u, out_i = m.computer()
    out = torch.cat([out_u, out_i]).detach().numpy().copy()

    rng = np.random.default_rng(3)
    users = rng.integers(0, ds.n_users, B)
    users[:8] = users[8:16]                               # repeated users inside the batch (atomics)
    pos = np.array([rng.choice(ds.allPos[u]) for u in users])
    neg = rng.integers(0, ds.m_items, B)
    tu, tp, tn = (torch.from_numpy(x).long() for x in (users, pos, neg))

    loss, reg = m.bpr_loss(tu, tp, tn)
    total = loss + reg * world.config['decay']
    m.zero_grad()
    total.backward()
    grad = torch.cat([m.embedding_user.weight.grad, m.embedding_item.weight.grad]).numpy().copy()
    if name == 'popgate':
        extra.update({'grad_' + k.replace('.', '__'): p.grad.numpy().copy() for k, p in m.named_parameters() if not k.startswith('embedding')})
    m.zero_grad()

    bpr = utils.BPRLoss(m, world.config)
    step_losses, params_after = [], []
    for s in range(3):
        shift = (s * 17) % B
        step_losses.append(bpr.stageOne(torch.roll(tu, shift), torch.roll(tp, shift), torch.roll(tn, shift)))
        params_after.append(torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().numpy().copy())
    st = bpr.opt.state_dict()['state']
    exp_avg = np.concatenate([st[0]['exp_avg'].numpy(), st[1]['exp_avg'].numpy()])
    exp_avg_sq = np.concatenate([st[0]['exp_avg_sq'].numpy(), st[1]['exp_avg_sq'].numpy()])

    with torch.no_grad():
        test_users = list(ds.testDict.keys())
        rating = m.getUsersRating(torch.tensor(test_users).long()).numpy().copy()
    res = Procedure.Test(ds, m, 0)
    out_after = torch.cat(m.computer()).detach().numpy().copy()

    if variant:
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez_compressed(
            os.path.join(GOLDEN, f'{name}.npz'),
            n_users=ds.n_users, m_items=ds.m_items, d=d, L=L, topks=np.array(world.topks), decay=world.config['decay'],
            lr=world.config['lr'], train_user=ds.trainUser, train_item=ds.trainItem, test_user=ds.testUser, test_item=ds.testItem,
            E0=E0, out=out, users=users, pos=pos, neg=neg, loss=loss.item(), reg=reg.item(), grad=grad,
            step_losses=np.array(step_losses), params_after=params_after[-1], test_users=np.array(test_users), rating=rating,
            precision=res['precision'], recall=res['recall'], ndcg=res['ndcg'], **extra)
        print(f"[golden] {name}: loss={loss.item():.6f} reg={reg.item():.6f} recall={res['recall']} ndcg={res['ndcg']}")
        return
    # ---- the port must reproduce the reference bit for bit -----------------------------------
    sys.path.insert(0, ROOT)
    from oracle import ref_port
    graph_p, norm_adj_p, _ = ref_port.build_graph(ds.trainUser, ds.trainItem, ds.n_users, ds.m_items)
    assert torch.equal(graph_p.indices(), g.indices()) and torch.equal(graph_p.values(), g.values()), "port graph != reference"
    utils.set_seed(2020)
    pm = ref_port.RefLightGCN(ds.n_users, ds.m_items, d, L, graph_p)
    assert np.array_equal(torch.cat([pm.embedding_user.weight, pm.embedding_item.weight]).detach().numpy(), E0)
    assert np.array_equal(torch.cat(pm.computer()).detach().numpy(), out), "port computer() != reference"
    pb = ref_port.RefBPRLoss(pm, world.config['decay'], world.config['lr'])
    for s in range(3):
        shift = (s * 17) % B
        l = pb.stageOne(torch.roll(tu, shift), torch.roll(tp, shift), torch.roll(tn, shift))
        assert l == step_losses[s], "port stageOne != reference"
    assert np.array_equal(torch.cat([pm.embedding_user.weight, pm.embedding_item.weight]).detach().numpy(), params_after[-1])
    pres, ptopk = ref_port.ref_test(pm, ds.testDict, ds.allPos, world.topks, world.config['test_u_batch_size'])
    for k_ in res:
        assert np.allclose(pres[k_], res[k_], rtol=0, atol=1e-12), (k_, pres[k_], res[k_])

    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(
        os.path.join(GOLDEN, f'{name}.npz'),
        n_users=ds.n_users, m_items=ds.m_items, d=d, L=L, topks=np.array(world.topks), decay=world.config['decay'],
        lr=world.config['lr'], train_user=ds.trainUser, train_item=ds.trainItem, test_user=ds.testUser, test_item=ds.testItem,
        adj_row=g.indices()[0].numpy(), adj_col=g.indices()[1].numpy(), adj_val=g.values().numpy(),
        users_D=ds.users_D, items_D=ds.items_D, E0=E0, out=out, users=users, pos=pos, neg=neg,
        loss=loss.item(), reg=reg.item(), grad=grad, step_losses=np.array(step_losses),
        params_after=np.stack(params_after), exp_avg=exp_avg, exp_avg_sq=exp_avg_sq,
        test_users=np.array(test_users), rating=rating, topk=ptopk, out_after=out_after,
        precision=res['precision'], recall=res['recall'], ndcg=res['ndcg'])
    print(f"[golden] {name}: loss={loss.item():.6f} reg={reg.item():.6f} recall={res['recall']} ndcg={res['ndcg']}")


def sampler_golden():
    """Compile the reference's sampling.cpp as it lies and record its output for a fixed seed."""
    import sysconfig, importlib
    tmp = tempfile.mkdtemp(prefix='golden_sampler_')
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    inc = subprocess.check_output([sys.executable, '-m', 'pybind11', '--includes']).decode().split()
    subprocess.check_call(['g++', '-O2', '-std=c++11', '-shared', '-fPIC', *inc, os.path.join(REF, 'sources', 'sampling.cpp'),
                           '-o', os.path.join(tmp, 'sampling' + ext)])
    sys.path.insert(0, tmp)
    sampling = importlib.import_module('sampling')
    nu, ni, train, _ = make_case('tiny')
    all_pos = [sorted(set(train[u])) for u in range(nu)]
    train_num = sum(len(v) for v in train.values())
    sampling.seed(2020)
    S = sampling.sample_negative(nu, ni, train_num, all_pos, 1)
    # the rest of the module's ABI, continuing the same rand() stream: randint, then sample_negative_ByUser (2 negatives)
    randints = np.array([sampling.randint(1000) for _ in range(16)], dtype=np.int32)
    by_users = np.array([5, 0, 299, 5, 17, 123, 42, 42], dtype=np.int32)
    S_by_user = sampling.sample_negative_ByUser(by_users.tolist(), ni, all_pos, 2)
    np.savez_compressed(os.path.join(GOLDEN, 'sampler.npz'), S=S, train_num=train_num, n_users=nu, m_items=ni,
                        randints=randints, by_users=by_users, S_by_user=S_by_user,
                        indptr=np.concatenate([[0], np.cumsum([len(a) for a in all_pos])]).astype(np.int64),
                        items=np.concatenate(all_pos).astype(np.int32))
    print(f"[golden] sampler: {S.shape} first rows {S[:3].tolist()}")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default=None)
    a = ap.parse_args()
    if a.case is None:
        for c in list(CASES) + ['sampler']:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), '--case', c])
    elif a.case == 'sampler':
        sampler_golden()
    else:
        run_case(a.case)
