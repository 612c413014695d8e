"""Training and evaluation procedures with the reference's signatures (code/Procedure.py).

BPR_train_original(dataset, recommend_model, loss_class, epoch, neg_k=1, w=None) -> str
Test(dataset, Recmodel, epoch, w=None, multicore=0) -> {'precision','recall','ndcg': np.ndarray}

What changed underneath:
  * training: the epoch's shuffled triples are moved to the device once (as in the reference,
    code/Procedure.py:52-55) and every step is a device-side window move + one CUDA-graph replay;
    the per-step loss is accumulated on the device and read once per epoch (the reference syncs
    with loss.cpu().item() every step, code/utils.py:64).  The returned string and the CSV row are
    the reference's: sum of step losses / (len//batch + 1).
  * evaluation: propagation runs ONCE (the reference recomputes it for each of the 299 user
    batches), scoring + train-item mask + top-k is one fused kernel, and precision/recall/NDCG are
    reduced on the device.  Users are evaluated in testDict order; tiling is free because results
    are per-user.
"""
import csv
import os

import numpy as np
import torch

from . import ops, utils, world
from .utils import timer


def _save_dir():
    return world.config.get('path', world.config.get('checkpoint_dir', './checkpoints'))


def _append_csv(name, header, row):
    save_path = _save_dir()
    os.makedirs(save_path, exist_ok=True)
    path = os.path.join(save_path, name)
    if not os.path.exists(path):
        with open(path, 'w', newline='') as f:
            csv.writer(f).writerow(header)
    with open(path, 'a', newline='') as f:
        csv.writer(f).writerow(row)


def _dist_info(Recmodel):
    """(mode, rank, world, group) of the model's engine; (None, 0, 1, None) for a single GPU or a foreign model."""
    eng = getattr(Recmodel, '_engine', None)
    if eng is None or eng.dist_mode is None:
        return None, 0, 1, None
    return eng.dist_mode, eng.rank, eng.world, eng.group


def _assert_same_on_every_rank(t, group, what):
    """The row partition replicates the batch: every rank must feed the same triples (same sampler and shuffle seeds)."""
    import torch.distributed as dist
    h = (t.to(torch.int64) * torch.arange(1, t.numel() + 1, dtype=torch.int64, device=t.device).view(t.shape)).sum().view(1)
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group); dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if int(lo.item()) != int(hi.item()):
        raise RuntimeError(f"{what} differ between ranks: seed the sampler and numpy identically on every rank (dist_mode='rowpart')")


def BPR_train_original(dataset, recommend_model, loss_class, epoch, neg_k=1, w=None):
    Recmodel = recommend_model
    Recmodel.train()
    bpr = loss_class
    bs = world.config['bpr_batch_size']
    mode, rank, nranks, group = _dist_info(Recmodel)
    if mode in ('dp', 'dp_idx'):
        raise NotImplementedError("BPR_train_original drives one GPU, the row partition ('rowpart') or the feature partition ('featpart'); the "
                                  "replicated modes 'dp'/'dp_idx' are driven per rank through Engine.step")
    if world.config.get('device_sampler', False) and getattr(bpr, 'fused', False):
        # K5: sample + shuffle on the device, straight into the layout the step reads (no host phase, no H2D)
        eng = Recmodel._engine
        if Recmodel._csr is None:
            raise NotImplementedError("the device sampler reads the user rows of the whole adjacency; the memory-partitioned "
                                      "row partition keeps no such copy — use the (bit-exact) host sampler")
        with timer(name="Sample"):
            S_dev = ops.sample_bpr(Recmodel._csr, dataset.n_users, dataset.m_items, dataset.trainDataSize,
                                   world.seed, epoch)
        if eng.B_cap != bs:
            eng._alloc_batch(bs)
        eng.set_lr(bpr.opt.param_groups[0]['lr'])
        eng.decay = float(bpr.weight_decay)
        if mode in ('rowpart', 'featpart') and nranks > 1:
            _assert_same_on_every_rank(S_dev, group, "the epoch's device-sampled triples")      # same (seed, epoch) key on every rank
        steps = eng.begin_epoch(S_dev)
        for _ in range(steps):
            eng.epoch_step()
        total_batch = S_dev.shape[1] // bs + 1
        aver_loss = float(eng.loss_to_host()[3]) / total_batch
        if rank == 0:
            _append_csv('train_epoch_metrics.csv', ['epoch', 'loss'], [epoch, aver_loss])
        time_info = timer.dict()
        timer.zero()
        return f"loss{aver_loss:.3f}-{time_info}"
    with timer(name="Sample"):
        # the following epoch's sample is drawn in the background while the GPU runs this one (rewound if the next sampler
        # call turns out to be something else: utils._EpochPrefetch)
        S = utils.UniformSample_original(dataset, prefetch_next=bool(world.config.get('sampler_prefetch', True)))
    # int64 on the device directly (the reference's torch.Tensor(...).long() float32 round trip,
    # code/Procedure.py:52-54, is exact only below 2^24 — SURVEY.md A16)
    # Like the reference (code/Procedure.py:52-55) the triples go to the device first and are permuted there; only the
    # permutation itself is drawn on the host, with the same numpy call as utils.shuffle (code/utils.py:148).
    dev = world.device if world.device.type == 'cuda' else torch.device('cpu')
    S_raw = torch.from_numpy(np.ascontiguousarray(S[:, :3])).to(dev, non_blocking=True)      # int32 [n, 3]
    perm = np.arange(S.shape[0])
    np.random.shuffle(perm)
    S_t = S_raw[torch.from_numpy(perm).to(dev)].t().to(torch.int64).contiguous()              # int64 [3, n], shuffled
    n = S_t.shape[1]
    total_batch = n // bs + 1
    if getattr(bpr, 'fused', False):
        eng = Recmodel._engine
        if eng.B_cap != bs:
            eng._alloc_batch(bs)
        eng.set_lr(bpr.opt.param_groups[0]['lr'])
        eng.decay = float(bpr.weight_decay)
        S_dev = S_t
        if mode in ('rowpart', 'featpart') and nranks > 1:
            # every rank replays the SAME epoch (the batch is replicated, the rows of A are what is split); the loss and
            # the returned string are identical on every rank
            _assert_same_on_every_rank(S_dev, group, "the epoch's sampled triples")
        steps = eng.begin_epoch(S_dev)
        for batch_i in range(steps):
            eng.epoch_step()
            if world.tensorboard and w is not None:
                w.add_scalar('BPRLoss/BPR', float(eng.loss_to_host()[2]), epoch * total_batch + batch_i)
        aver_loss = float(eng.loss_to_host()[3])
    else:
        users, posItems, negItems = (S_t[i].to(world.device) for i in range(3))
        aver_loss = 0.
        for batch_i, (u, p, nn_) in enumerate(utils.minibatch(users, posItems, negItems, batch_size=bs)):
            cri = bpr.stageOne(u, p, nn_)
            aver_loss += cri
            if world.tensorboard and w is not None:
                w.add_scalar('BPRLoss/BPR', cri, epoch * total_batch + batch_i)
    aver_loss = aver_loss / total_batch          # code/Procedure.py:57,68 (keeps the +1 quirk)
    if rank == 0:
        _append_csv('train_epoch_metrics.csv', ['epoch', 'loss'], [epoch, aver_loss])
    time_info = timer.dict()
    timer.zero()
    return f"loss{aver_loss:.3f}-{time_info}"


def test_one_batch(X):
    """Host metric arithmetic for one user (code/Procedure.py:89-121); kept for API parity and as the
    cross-check of the device metric kernel."""
    sorted_items = X[0].cpu().numpy() if isinstance(X[0], torch.Tensor) else np.asarray(X[0])
    groundTrue = X[1]
    if not isinstance(groundTrue, (list, set, tuple, np.ndarray)):
        groundTrue = [groundTrue]
    test_data = [groundTrue]
    r = np.expand_dims(utils.getLabel(groundTrue, sorted_items), axis=0)
    pre, recall, ndcg = [], [], []
    for k in world.topks:
        ret = utils.RecallPrecision_ATk(test_data, r, k)
        pre.append(ret['precision'])
        recall.append(ret['recall'])
        ndcg.append(utils.NDCGatK_r(test_data, r, k))
    return {'precision': np.array(pre), 'recall': np.array(recall), 'ndcg': np.array(ndcg)}


def rank_all(dataset, Recmodel, k, user_tile=8192):
    """Top-k item ids for every user of testDict (key order) -> int64 device tensor [n_test_users, k].
    Multi-GPU: the users are sharded over the ranks in contiguous blocks, the item table is replicated (it
    is the exchanged `out`), every rank ranks its block and the blocks are all-gathered — the result is the same tensor
    on every rank and the same bits as the single-GPU call (each row is ranked by the same kernel on the same data)."""
    users_dev, _, _ = dataset.test_csr()
    mode, rank, nranks, group = _dist_info(Recmodel)
    n = users_dev.numel()
    if nranks > 1:
        import torch.distributed as dist
        from .engine import shard_batch
        Recmodel.computer()                              # collective (exchanged layers): every rank, before the shards diverge
        per = (n + nranks - 1) // nranks
        lo_r, hi_r = shard_batch(n, rank, nranks)
        mine = torch.zeros((per, k), dtype=torch.int64, device=users_dev.device)
        for lo in range(lo_r, hi_r, user_tile):
            hi = min(hi_r, lo + user_tile)
            idx, _ = Recmodel.rank_topk(users_dev[lo:hi], k)
            mine[lo - lo_r:hi - lo_r] = idx
        allb = torch.empty((nranks * per, k), dtype=torch.int64, device=users_dev.device)
        dist.all_gather_into_tensor(allb, mine, group=group)
        return allb[:n].contiguous()
    out = []
    for lo in range(0, n, user_tile):
        idx, _ = Recmodel.rank_topk(users_dev[lo:lo + user_tile], k)
        out.append(idx)
    return torch.cat(out, dim=0) if len(out) > 1 else out[0]


def Test(dataset, Recmodel, epoch, w=None, multicore=0):
    Recmodel = Recmodel.eval()
    max_K = max(world.topks)
    results = {m: np.zeros(len(world.topks)) for m in ['precision', 'recall', 'ndcg']}
    with torch.no_grad():
        users_dev, t_indptr, t_indices = dataset.test_csr()
        n_users_eval = users_dev.numel()
        if n_users_eval:
            if hasattr(Recmodel, 'rank_topk'):
                # one call ranks a whole tile of users; the tensor-core path takes everybody at once (its workspace is a few
                # hundred MB for 64 k users), the exact kernel works in tiles of 8192
                tc = bool(getattr(Recmodel, 'config', world.config).get('score_tensor_core', True)) and dataset.m_items >= ops.TC_MIN_ITEMS
                topk = rank_all(dataset, Recmodel, max_K, user_tile=max(int(world.config.get('test_u_batch_size', 100)), 65536 if tc else 8192))
            else:   # a foreign model: the reference's unfused recipe, one user tile at a time
                parts = []
                u_bs = world.config['test_u_batch_size']
                for lo in range(0, n_users_eval, u_bs):
                    bu = users_dev[lo:lo + u_bs]
                    rating = Recmodel.getUsersRating(bu)
                    allPos = dataset.getUserPosItems(bu.cpu().tolist())
                    ex_i = np.concatenate([np.full(len(it), i) for i, it in enumerate(allPos)])
                    ex_j = np.concatenate([np.asarray(it) for it in allPos])
                    rating[torch.from_numpy(ex_i).to(rating.device), torch.from_numpy(ex_j).long().to(rating.device)] = -(1 << 10)
                    parts.append(torch.topk(rating, k=max_K)[1])
                topk = torch.cat(parts, dim=0)
            sums = ops.rank_metrics(topk.contiguous(), t_indptr, t_indices, world.topks).cpu().numpy()
            results['precision'] = sums[:, 0] / n_users_eval
            results['recall'] = sums[:, 1] / n_users_eval
            results['ndcg'] = sums[:, 2] / n_users_eval
    prec, rec, nd = float(results['precision'][0]), float(results['recall'][0]), float(results['ndcg'][0])
    rank = _dist_info(Recmodel)[1]
    if rank != 0:
        return results
    _append_csv('valid_epoch_metrics.csv', ['epoch', 'precision', 'recall', 'ndcg'], [epoch, prec, rec, nd])
    if world.tensorboard and w is not None:
        w.add_scalars(f'Test/Recall@{world.topks}', {str(world.topks[i]): results['recall'][i] for i in range(len(world.topks))}, epoch)
        w.add_scalars(f'Test/Precision@{world.topks}', {str(world.topks[i]): results['precision'][i] for i in range(len(world.topks))}, epoch)
        w.add_scalars(f'Test/NDCG@{world.topks}', {str(world.topks[i]): results['ndcg'][i] for i in range(len(world.topks))}, epoch)
    print(results)
    return results
