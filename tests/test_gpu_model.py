"""GPU: the reference-facing Python API (LightGCN / BPRLoss / Procedure) against outputs of the real
reference (tests/golden/*.npz) and the gowalla known answer."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope='module')
def lg():
    import lgcn_b200
    return lgcn_b200


def make_model(lg, g, **over):
    cfg = dict(lg.world.config)
    cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), bpr_batch_size=len(g['users']),
               decay=float(g['decay']), lr=float(g['lr']))
    cfg.update(over)
    ds = lg.InteractionDataset(int(g['n_users']), int(g['m_items']), g['train_user'], g['train_item'],
                               g['test_user'], g['test_item'], config=cfg)
    m = lg.LightGCN(cfg, ds)
    nu = int(g['n_users'])
    with torch.no_grad():
        m.embedding_user.weight.copy_(torch.from_numpy(g['E0'][:nu]))
        m.embedding_item.weight.copy_(torch.from_numpy(g['E0'][nu:]))
    return cfg, ds, m


def triples(g, shift=0):
    return tuple(torch.from_numpy(np.roll(g[k], shift)).long() for k in ('users', 'pos', 'neg'))


def params(m):
    return torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().cpu().numpy()


def test_seeded_init_is_the_references(lg, golden):
    """set_seed(2020) -> construct consumes the CPU RNG exactly like code/model.py:57-60."""
    cfg = dict(lg.world.config); cfg.update(latent_dim_rec=int(golden['d']), lightGCN_n_layers=int(golden['L']))
    ds = lg.InteractionDataset(int(golden['n_users']), int(golden['m_items']), golden['train_user'], golden['train_item'],
                               golden['test_user'], golden['test_item'], config=cfg)
    lg.utils.set_seed(2020)
    m = lg.LightGCN(cfg, ds)
    assert np.array_equal(params(m), golden['E0'])
    assert set(m.state_dict().keys()) == {'embedding_user.weight', 'embedding_item.weight'}


def test_computer_matches_reference(lg, golden):
    _, ds, m = make_model(lg, golden)
    with torch.no_grad():
        u, i = m.computer()
    out = torch.cat([u, i]).cpu().numpy()
    assert rel_err(out, golden['out']) < TOL
    # the dataset's graph is a torch sparse tensor usable exactly like the reference's
    ref = torch.sparse.mm(ds.getSparseGraph(), torch.from_numpy(golden['E0']).cuda())
    mine = torch.empty_like(ref); lg.ops.spmm(ds.getCSRGraph(), torch.from_numpy(golden['E0']).cuda(), mine)
    assert rel_err(mine.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_bpr_loss_autograd_matches_reference(lg, golden):
    """Generic path: bpr_loss -> (loss + decay*reg).backward(), as the reference's utils.BPRLoss drives it."""
    _, _, m = make_model(lg, golden)
    u, p, n = (t.cuda() for t in triples(golden))
    loss, reg = m.bpr_loss(u, p, n)
    assert abs(loss.item() - float(golden['loss'])) < TOL * abs(float(golden['loss']))
    assert abs(reg.item() - float(golden['reg'])) < TOL * abs(float(golden['reg']))
    (loss + reg * float(golden['decay'])).backward()
    grad = torch.cat([m.embedding_user.weight.grad, m.embedding_item.weight.grad]).cpu().numpy()
    assert rel_err(grad, golden['grad']) < 2e-5


@pytest.mark.parametrize("mode", ["graph", "eager", "deterministic", "host_batch"])
def test_fused_stageOne_matches_reference(lg, golden, mode):
    over = dict(cuda_graph=(mode != "eager"), deterministic=(mode == "deterministic"))
    cfg, _, m = make_model(lg, golden, **over)
    bpr = lg.utils.BPRLoss(m, cfg)
    B = len(golden['users'])
    for s in range(3):
        u, p, n = triples(golden, (s * 17) % B)
        if mode != "host_batch":
            u, p, n = u.cuda(), p.cuda(), n.cuda()
        loss = bpr.stageOne(u, p, n)
        assert abs(loss - golden['step_losses'][s]) < TOL * abs(golden['step_losses'][s])
        assert rel_err(params(m), golden['params_after'][s]) < 1e-4
    sd = bpr.opt.state_dict()
    assert float(sd['state'][0]['step']) == 3.0
    m_cat = np.concatenate([sd['state'][0]['exp_avg'].cpu().numpy(), sd['state'][1]['exp_avg'].cpu().numpy()])
    v_cat = np.concatenate([sd['state'][0]['exp_avg_sq'].cpu().numpy(), sd['state'][1]['exp_avg_sq'].cpu().numpy()])
    assert rel_err(m_cat, golden['exp_avg']) < 1e-4 and rel_err(v_cat, golden['exp_avg_sq']) < 1e-4
    with torch.no_grad():
        out = torch.cat(m.computer()).cpu().numpy()
    assert rel_err(out, golden['out_after']) < 1e-4


def test_reference_recipe_on_this_model(lg, golden):
    """The reference's own BPRLoss recipe (zero_grad/backward/torch Adam) drives this model unchanged."""
    cfg, _, m = make_model(lg, golden)
    opt = torch.optim.Adam(m.parameters(), lr=cfg['lr'])
    B = len(golden['users'])
    for s in range(3):
        u, p, n = (t.cuda() for t in triples(golden, (s * 17) % B))
        loss, reg = m.bpr_loss(u, p, n)
        total = loss + reg * cfg['decay']
        opt.zero_grad(); total.backward(); opt.step()
        assert abs(total.item() - golden['step_losses'][s]) < TOL * abs(golden['step_losses'][s])
    assert rel_err(params(m), golden['params_after'][2]) < 1e-4


def test_graph_replay_equals_eager_bitwise(lg, golden_tiny):
    res = []
    for use_graph in (True, False):
        cfg, _, m = make_model(lg, golden_tiny, cuda_graph=use_graph, deterministic=True)
        bpr = lg.utils.BPRLoss(m, cfg)
        for s in range(4):
            bpr.stageOne(*(t.cuda() for t in triples(golden_tiny, s * 5)))
        res.append(params(m))
    assert np.array_equal(res[0], res[1])


def test_dead_row_pruning_is_bitwise_neutral(lg, golden):
    """Skipping the rows nobody reads (engine.prune) must not change a single bit of the parameters."""
    res = []
    for prune in (True, False):
        cfg, _, m = make_model(lg, golden, deterministic=True, prune_dead_rows=prune)
        assert m._engine.prune == prune
        bpr = lg.utils.BPRLoss(m, cfg)
        losses = [bpr.stageOne(*(t.cuda() for t in triples(golden, s * 3))) for s in range(4)]
        res.append((params(m), losses))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


def test_getUsersRating_and_Test_match_reference(lg, golden):
    cfg, ds, m = make_model(lg, golden)
    nu = int(golden['n_users'])
    with torch.no_grad():
        m.embedding_user.weight.copy_(torch.from_numpy(golden['params_after'][2][:nu]))
        m.embedding_item.weight.copy_(torch.from_numpy(golden['params_after'][2][nu:]))
    users = torch.from_numpy(golden['test_users']).long()
    with torch.no_grad():
        rating = m.getUsersRating(users.cuda()).cpu().numpy()
    assert rating.shape == golden['rating'].shape
    assert rel_err(rating, golden['rating']) < 2e-5
    lg.world.configure(topks=[int(k) for k in golden['topks']], checkpoint_dir='/tmp/lgcn_b200_test_ckpt')
    try:
        res = lg.Procedure.Test(ds, m, 0)
    finally:
        lg.world.configure(topks=[20])
    for name in ('precision', 'recall', 'ndcg'):
        assert np.allclose(res[name], golden[name], rtol=0, atol=1e-4), (name, res[name], golden[name])
    # top-k ids agree with the reference's torch.topk except at near-ties
    idx, _ = m.rank_topk(users, int(max(golden['topks'])))
    idx = idx.cpu().numpy()
    frac = (idx != golden['topk']).mean()
    assert frac < 0.02, frac


def test_state_dict_and_optimizer_round_trip(lg, golden_tiny):
    cfg, _, m = make_model(lg, golden_tiny)
    bpr = lg.utils.BPRLoss(m, cfg)
    for s in range(2):
        bpr.stageOne(*(t.cuda() for t in triples(golden_tiny, s)))
    sd, osd = {k: v.clone() for k, v in m.state_dict().items()}, bpr.opt.state_dict()
    cfg2, _, m2 = make_model(lg, golden_tiny)
    m2.load_state_dict(sd, strict=True)
    bpr2 = lg.utils.BPRLoss(m2, cfg2)
    bpr2.opt.load_state_dict(osd)
    a = bpr.stageOne(*(t.cuda() for t in triples(golden_tiny, 9)))
    b = bpr2.stageOne(*(t.cuda() for t in triples(golden_tiny, 9)))
    assert abs(a - b) < 1e-6 * abs(a)
    assert rel_err(params(m2), params(m)) < 1e-6
    m3 = m2.to(lg.world.device)                       # the reference does Recmodel.to(world.device)
    assert m3 is m2 and m2._params_packed()


def test_epoch_training_and_eval_on_synthetic(lg, tmp_path):
    lg.world.configure(checkpoint_dir=str(tmp_path), bpr_batch_size=256)
    try:
        cfg = dict(lg.world.config)
        ds = lg.synth.make_dataset('tiny', config=cfg)
        lg.utils.set_seed(2020); lg.utils.sampler_seed(2020)
        m = lg.LightGCN(cfg, ds)
        bpr = lg.utils.BPRLoss(m, cfg)
        r0 = lg.Procedure.Test(ds, m, 0)
        infos = [lg.Procedure.BPR_train_original(ds, m, bpr, e) for e in range(30)]
        r1 = lg.Procedure.Test(ds, m, 30)
        l0, l1 = float(infos[0].split('-')[0][4:]), float(infos[-1].split('-')[0][4:])
        assert l1 < l0 and infos[0].endswith('|') and 'Sample' in infos[0]
        assert r1['recall'][0] > r0['recall'][0]
        assert (tmp_path / 'train_epoch_metrics.csv').read_text().count('\n') == 31
    finally:
        lg.world.configure(checkpoint_dir='./checkpoints', bpr_batch_size=2048)


def test_training_with_device_sampler(lg, tmp_path):
    lg.world.configure(checkpoint_dir=str(tmp_path), bpr_batch_size=256, device_sampler=True)
    try:
        cfg = dict(lg.world.config)
        ds = lg.synth.make_dataset('tiny', config=cfg)
        lg.utils.set_seed(2020)
        m = lg.LightGCN(cfg, ds)
        bpr = lg.utils.BPRLoss(m, cfg)
        r0 = lg.Procedure.Test(ds, m, 0)
        infos = [lg.Procedure.BPR_train_original(ds, m, bpr, e) for e in range(30)]
        r1 = lg.Procedure.Test(ds, m, 30)
        assert float(infos[-1].split('-')[0][4:]) < float(infos[0].split('-')[0][4:])
        assert r1['recall'][0] > r0['recall'][0]
    finally:
        lg.world.configure(checkpoint_dir='./checkpoints', bpr_batch_size=2048, device_sampler=False)


def test_train_driver_checkpoint_resume(lg, tmp_path):
    """The driver's checkpoint uses the reference's schema/keys and resumes to the same state."""
    from lgcn_b200 import train
    args = ['--synthetic', 'tiny', '--epochs', '3', '--bpr_batch', '256', '--checkpoint_dir', str(tmp_path), '--save_every', '1']
    try:
        m = train.main(args)
        ck = torch.load(str(tmp_path / 'last.pth.tar'), map_location='cpu', weights_only=False)
        assert set(ck['model_state'].keys()) == {'embedding_user.weight', 'embedding_item.weight'}
        assert ck['epoch'] == 3 and float(ck['optimizer_state']['state'][0]['step']) > 0
        assert torch.equal(ck['model_state']['embedding_user.weight'], m.embedding_user.weight.detach().cpu())
        m2 = train.main(args[:3] + ['4'] + args[4:] + ['--resume', str(tmp_path / 'last.pth.tar')])
        assert m2._engine._host_step > m._engine._host_step * 0 + int(float(ck['optimizer_state']['state'][0]['step']))
    finally:
        lg.world.configure(checkpoint_dir='./checkpoints', bpr_batch_size=2048, epochs=1000, device_sampler=False)


def test_epoch_mode_equals_step_mode(lg, golden_tiny):
    """Device-resident epoch (window advance + graph replay) == explicit stageOne calls on the same batches."""
    g = golden_tiny
    S = np.stack([np.tile(g[k], 3)[:700] for k in ('users', 'pos', 'neg')])          # 700 triples, B=256 -> 3 steps
    cfg, _, ma = make_model(lg, g, deterministic=True)
    ea = ma._engine
    steps = ea.begin_epoch(torch.from_numpy(S).cuda())
    assert steps == 3
    for _ in range(steps):
        ea.epoch_step()
    run_sum = float(ea.loss_to_host()[3])
    cfg, _, mb = make_model(lg, g, deterministic=True)
    bpr = lg.utils.BPRLoss(mb, cfg)
    losses = [bpr.stageOne(*(torch.from_numpy(S[j, lo:lo + 256]).cuda() for j in range(3))) for lo in (0, 256, 512)]
    assert np.array_equal(params(ma), params(mb))
    assert abs(run_sum - sum(losses)) < 1e-5


def test_gowalla_step0_known_answer_end_to_end(lg, gowalla, tmp_path):
    """Reference run artefacts: seed 2020 -> LightGCN(L=3,d=64) -> Test on gowalla gives
    Precision@20 0.0001875544, Recall@20 0.0005374941, NDCG@20 0.00040836."""
    nu, ni = int(gowalla['n_users']), int(gowalla['m_items'])
    tu = np.repeat(np.arange(nu), np.diff(gowalla['train_indptr'])).astype(np.int64)
    ti = gowalla['train_items'].astype(np.int64)
    su = np.repeat(gowalla['test_users'].astype(np.int64), np.diff(gowalla['test_indptr']))
    si = gowalla['test_items'].astype(np.int64)
    lg.world.configure(checkpoint_dir=str(tmp_path), topks=[20])
    cfg = dict(lg.world.config)
    ds = lg.InteractionDataset(nu, ni, tu, ti, su, si, config=cfg, name='gowalla')
    g = ds.getCSRGraph()
    assert g.nnz == 1620256
    lg.utils.set_seed(2020)
    m = lg.LightGCN(cfg, ds)
    res = lg.Procedure.Test(ds, m, 0)
    assert abs(res['precision'][0] - float(gowalla['kat_precision'])) < 2e-9
    assert abs(res['recall'][0] - float(gowalla['kat_recall'])) < 2e-9
    assert abs(res['ndcg'][0] - float(gowalla['kat_ndcg'])) < 2e-7


# ------------------------------------------------------------------------------------ variants
def _variant_model(lg, g, tmp_path, kind):
    over = {}
    if kind.startswith('popgate'):
        over = dict(use_pop_gate=True, popgate_kernel=(kind == 'popgate'))
    else:
        over_extra = dict(i2i_kernel_step=(kind == 'i2i'))
        import scipy.sparse as sp
        ni = int(g['m_items'])
        m = sp.csr_matrix((g['i2i_data'], g['i2i_indices'], g['i2i_indptr']), shape=(ni, ni))
        path = str(tmp_path / 'i2i.npz'); sp.save_npz(path, m)
        over = dict(use_item_item=True, i2i_path=path, i2i_alpha=float(g['i2i_alpha']), **over_extra)
    cfg, ds, m = make_model(lg, g, **over)
    if kind.startswith('popgate'):
        sd = {k[3:].replace('__', '.'): torch.from_numpy(v) for k, v in g.items() if k.startswith('sd_')}
        missing = m.load_state_dict(sd, strict=False)
        assert set(missing.missing_keys) == {'embedding_user.weight', 'embedding_item.weight'} and not missing.unexpected_keys
    return cfg, ds, m


@pytest.mark.parametrize("kind", ["popgate", "popgate_autograd", "i2i", "i2i_autograd"])
def test_model_variants_match_reference(lg, tmp_path, kind):
    """use_pop_gate / use_item_item: loss, gradients, three optimiser steps and scores against the real reference.
    'popgate' trains through the FUSED step (csrc/popgate.cu: fusion + BPR + closed-form backward + fused Adam on the MLP
    block), 'popgate_autograd' / 'i2i' through bpr_loss().backward() + torch Adam on top of the kernel-backed propagation."""
    g = load_golden('popgate' if kind.startswith('popgate') else 'i2i')
    cfg, ds, m = _variant_model(lg, g, tmp_path, kind)
    assert m.plain == (kind in ('popgate', 'i2i'))
    nu = int(g['n_users'])
    with torch.no_grad():
        out = torch.cat(m.computer()).cpu().numpy()
    assert rel_err(out, g['out']) < TOL
    u, p, n = (t.cuda() for t in triples(g))
    loss, reg = m.bpr_loss(u, p, n)
    assert abs(loss.item() - float(g['loss'])) < 2e-5 * abs(float(g['loss']))
    assert abs(reg.item() - float(g['reg'])) < 2e-5 * abs(float(g['reg']))
    (loss + reg * float(g['decay'])).backward()
    grad = torch.cat([m.embedding_user.weight.grad, m.embedding_item.weight.grad]).cpu().numpy()
    assert rel_err(grad, g['grad']) < 5e-5
    if kind.startswith('popgate'):
        for name, prm in m.named_parameters():
            if not name.startswith('embedding'):
                assert rel_err(prm.grad.cpu().numpy(), g['grad_' + name.replace('.', '__')]) < 1e-4, name
    if kind == 'popgate':
        # the kernel's closed-form backward against the same reference gradients (embeddings AND the 8 MLP tensors)
        eng = m._engine
        eng.forward(); eng.G.zero_(); eng.pg['grad'].zero_()
        eng._stage_batch(u, p, n)
        lg.ops.popgate_bpr_fwd_bwd(eng.out, eng.bu, eng.bp, eng.bn, eng.B_cap, eng.ctl, eng.nu, eng.ni, eng.pg['pop'], eng.pg['params'],
                                   eng.pg['H1'], eng.pg['H2'], eng.pg['temp'], eng.pg['coeff'], float(g['decay']), eng.loss_out, eng.G, eng.pg['grad'], eng.pg['ws'])
        lo = eng.loss_out.cpu().numpy()
        assert abs(lo[0] - float(g['loss'])) < 2e-5 * abs(float(g['loss'])) and abs(lo[1] - float(g['reg'])) < 2e-5 * abs(float(g['reg']))
        gE = eng.backward_to(eng.G.clone(), eng.grad_buffer()).cpu().numpy()
        assert rel_err(gE, g['grad']) < 5e-5
        off = 0
        for name, t in zip(['pop_mlp.0.weight', 'pop_mlp.0.bias', 'pop_mlp.2.weight', 'pop_mlp.2.bias',
                            'gate_mlp.0.weight', 'gate_mlp.0.bias', 'gate_mlp.2.weight', 'gate_mlp.2.bias'], m.popgate_tensors()):
            got = eng.pg['grad'][off:off + t.numel()].view(t.shape).cpu().numpy(); off += t.numel()
            assert rel_err(got, g['grad_' + name.replace('.', '__')]) < 1e-4, name
        eng.G.zero_(); eng.pg['grad'].zero_()
    m.zero_grad()
    bpr = lg.utils.BPRLoss(m, cfg)
    assert bpr.fused == (kind in ('popgate', 'i2i'))
    B = len(g['users'])
    for s in range(3):
        l = bpr.stageOne(*(t.cuda() for t in triples(g, (s * 17) % B)))
        assert abs(l - g['step_losses'][s]) < 5e-5 * abs(g['step_losses'][s])
    assert rel_err(params(m), g['params_after']) < 2e-4
    users = torch.from_numpy(g['test_users']).long()
    with torch.no_grad():
        rating = m.getUsersRating(users.cuda()).cpu().numpy()
    assert rel_err(rating, g['rating']) < 1e-4
    lg.world.configure(checkpoint_dir=str(tmp_path))
    res = lg.Procedure.Test(ds, m, 0)
    for name in ('precision', 'recall', 'ndcg'):
        assert np.allclose(res[name], g[name], rtol=0, atol=1e-4)


def test_host_batches_zero_copy_equal_device_batches(lg, golden_tiny):
    """stageOne with pinned-host / pageable-host batches (pulled in by a kernel at the head of the captured step, loss pushed
    out by a kernel at its tail) == the same steps fed with device tensors, bit for bit — full, short and changing batch sizes."""
    g = golden_tiny
    B = len(g['users'])
    runs = []
    for where in ('device', 'host', 'pinned'):
        cfg, _, m = make_model(lg, g, deterministic=True)
        bpr = lg.utils.BPRLoss(m, cfg)
        losses = []
        for s, n in enumerate((B, B, 100, B, 37, 37, B)):
            t = [torch.from_numpy(np.roll(g[k], s * 11)[:n].copy()).long() for k in ('users', 'pos', 'neg')]
            if where == 'device':
                t = [x.cuda() for x in t]
            elif where == 'pinned':
                t = [x.pin_memory() for x in t]
            losses.append(bpr.stageOne(*t))
        runs.append((losses, params(m)))
    assert runs[0][0] == runs[1][0] == runs[2][0]
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][1], runs[2][1])
    assert m._engine.zero_copy and any(isinstance(k, tuple) and k[0] == 'host' for k in m._engine._graphs)


def _train_epochs(lg, ds, cfg, epochs, tmp_path):
    lg.utils.set_seed(2020)
    lg.utils.sampler_seed(2020)
    m = lg.LightGCN(cfg, ds)
    bpr = lg.utils.BPRLoss(m, cfg)
    infos = [lg.Procedure.BPR_train_original(ds, m, bpr, e) for e in range(1, epochs + 1)]
    res = lg.Procedure.Test(ds, m, epochs)
    return m, np.array([float(s[4:s.index('-')]) for s in infos]), res


@pytest.mark.parametrize('deterministic', [True, False])
def test_fixed_epochs_match_the_reference_procedure(lg, tmp_path, deterministic):
    """north_star: "Recall@20 and NDCG@20 must match to within 1e-4 after a fixed number of epochs".  Golden =
    the REAL reference's set_seed(2020) -> LightGCN -> 20 x Procedure.BPR_train_original (its own C++ sampler seeded
    2020, its numpy shuffle) -> Procedure.Test on the tiny dataset (oracle/gen_epochs_golden.py;
    code/Procedure.py:28-83,127-206, code/utils.py:38-64).  Here: the same calls on this package."""
    g, t = load_golden('tiny_epochs'), load_golden('tiny')
    lg.world.configure(checkpoint_dir=str(tmp_path), bpr_batch_size=int(g['batch']), topks=[20])
    try:
        cfg = dict(lg.world.config)
        cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), decay=float(g['decay']), lr=float(g['lr']),
                   deterministic=deterministic)
        ds = lg.InteractionDataset(int(t['n_users']), int(t['m_items']), t['train_user'], t['train_item'],
                                   t['test_user'], t['test_item'], config=cfg)
        m, losses, res = _train_epochs(lg, ds, cfg, int(g['epochs']), tmp_path)
        assert np.allclose(losses, g['epoch_loss_3dp'], atol=1.01e-3)
        assert rel_err(params(m), g['params']) < 1e-3
        assert abs(float(res['recall'][0]) - float(g['recall'][0])) <= 1e-4
        assert abs(float(res['ndcg'][0]) - float(g['ndcg'][0])) <= 1e-4
        assert abs(float(res['precision'][0]) - float(g['precision'][0])) <= 1e-4
    finally:
        lg.world.configure(checkpoint_dir='./checkpoints', bpr_batch_size=2048)


def test_gowalla_one_epoch_matches_the_reference_procedure(lg, gowalla, tmp_path):
    """The same on the reference's bundled gowalla data: ONE full epoch (395 steps of 2048 triples from the reference's
    sampler stream) + full-ranking Test, against what the real reference produced on the CPU
    (tests/golden/gowalla_epoch.npz, oracle/gen_epochs_golden.py --case gowalla_epoch)."""
    g = load_golden('gowalla_epoch')
    nu, ni = int(gowalla['n_users']), int(gowalla['m_items'])
    tu = np.repeat(np.arange(nu), np.diff(gowalla['train_indptr'])).astype(np.int64)
    ti = gowalla['train_items'].astype(np.int64)
    su = np.repeat(gowalla['test_users'].astype(np.int64), np.diff(gowalla['test_indptr']))
    si = gowalla['test_items'].astype(np.int64)
    lg.world.configure(checkpoint_dir=str(tmp_path), bpr_batch_size=int(g['batch']), topks=[20])
    try:
        cfg = dict(lg.world.config)
        cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), decay=float(g['decay']), lr=float(g['lr']), deterministic=True)
        ds = lg.InteractionDataset(nu, ni, tu, ti, su, si, config=cfg, name='gowalla')
        m, losses, res = _train_epochs(lg, ds, cfg, int(g['epochs']), tmp_path)
        assert np.allclose(losses, g['epoch_loss_3dp'], atol=1.01e-3)
        P = params(m)
        assert rel_err(P[:256], g['params_head']) < 1e-3
        assert rel_err(np.linalg.norm(P[::97], axis=1), g['params_row_norms_sample']) < 1e-4
        for k in ('recall', 'ndcg', 'precision'):
            assert abs(float(res[k][0]) - float(g[k][0])) <= 1e-4, (k, res[k], g[k])
    finally:
        lg.world.configure(checkpoint_dir='./checkpoints', bpr_batch_size=2048)


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_feature_partition_emulated_ranks_match_reference(lg, golden, P, deterministic):
    """dist_mode='featpart' with P ranks emulated on ONE GPU (an engine and a CUDA stream per rank; the record exchange, the
    device barrier and the double-buffered records are the multi-process ones): column slices of width 32 / 16 / 8 trained
    for 3 steps reproduce the reference's losses and parameters (code/utils.py:53-64), every rank sees the same loss bits."""
    from lgcn_b200.engine import Engine, link_feat_engines
    g = golden
    nu, ni, d, L = int(g['n_users']), int(g['m_items']), int(g['d']), int(g['L'])
    if d // P < 8:
        pytest.skip("slices narrower than 8 columns are not supported")
    csr = lg.ops.csr_build(torch.from_numpy(g['train_user']).cuda(), torch.from_numpy(g['train_item']).cuda(), nu, ni)
    E0 = torch.from_numpy(g['E0']).cuda()
    B = len(g['users'])
    batches = [tuple(t.cuda() for t in triples(g, (s * 17) % B)) for s in range(3)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    engines = []
    for p in range(P):
        with torch.cuda.stream(streams[p]):
            e = Engine(csr.rows(0, csr.n_rows), nu, ni, d, L, torch.device('cuda'), lr=float(g['lr']), decay=float(g['decay']), B_cap=B,
                       deterministic=deterministic, use_graph=False, dist_mode='featpart', feat=(p, P))
            assert e.d == d // P and e.E0.shape == (nu + ni, d // P)
            e.E0.copy_(E0[:, e.c0:e.c0 + e.d])
        engines.append(e)
    torch.cuda.synchronize()
    link_feat_engines(engines, timeout_ms=3000)
    for s in range(3):
        for p in range(P):                       # nothing in step() waits on the host, so the P ranks' steps overlap on the device
            with torch.cuda.stream(streams[p]):
                engines[p].step(*batches[s])
        torch.cuda.synchronize()
        try:
            for e in engines:
                e.xchg.barrier.check()
        except RuntimeError as err:
            # the emulation needs the P streams to make progress concurrently (conftest.py asks for 32 hardware queues); where the
            # device serialises them a barrier gives up after 3 s — an environment limit, not a result (tests/test_gpu_dist.py runs
            # the same protocol with one process per GPU)
            pytest.skip(f"streams of this device did not run concurrently: {err}")
        losses = [e.loss_out.cpu().numpy() for e in engines]
        assert all(np.array_equal(x, losses[0]) for x in losses)
        assert abs(float(losses[0][2]) - g['step_losses'][s]) < TOL * abs(g['step_losses'][s])
        P_all = torch.cat([e.E0 for e in engines], dim=1).cpu().numpy()
        assert rel_err(P_all, g['params_after'][s]) < 1e-4
    M_all = torch.cat([e.M for e in engines], dim=1).cpu().numpy()
    V_all = torch.cat([e.V for e in engines], dim=1).cpu().numpy()
    assert rel_err(M_all, g['exp_avg']) < 1e-4 and rel_err(V_all, g['exp_avg_sq']) < 1e-4
    for p in range(P):
        with torch.cuda.stream(streams[p]):
            engines[p].forward()
    torch.cuda.synchronize()
    assert rel_err(torch.cat([e.out for e in engines], dim=1).cpu().numpy(), g['out_after']) < 1e-4
