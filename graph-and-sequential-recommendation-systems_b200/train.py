"""Working driver with the shape of the reference's (unparseable) main.py epoch loop, code/main.py:185-242:
every 10th epoch Test + best-NDCG checkpoint, then one BPR epoch, CSV rows, atomic `last.pth.tar`.

    python -m lgcn_b200.train --dataset gowalla --data_path data/gowalla --epochs 50
    python -m lgcn_b200.train --synthetic yelp2018 --epochs 20 --device_sampler
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m lgcn_b200.train --synthetic amazon-book --epochs 20

Under torchrun (WORLD_SIZE > 1) the adjacency is row-partitioned over the ranks (dist_mode='rowpart'): every rank runs the
same loop on the same sampled triples (same seeds), holds only its block of the CSR and of the Adam moments, evaluates its
shard of the test users, and rank 0 alone prints and writes checkpoints (the optimizer state is gathered for that).
`--dist_mode featpart` splits the embedding COLUMNS instead (CSR replicated, no exchange in the propagation): the mode for
graphs that fit one GPU, where a row-partitioned layer is bound by the NVLink ingest of the exchanged table.

Checkpoint schema = the reference's (code/main.py:56-67): {'epoch','model_state','optimizer_state','best_metric'} with
parameter keys embedding_user.weight / embedding_item.weight, so files are interchangeable.
"""
import argparse
import os
import sys
import time

import torch

from . import Procedure, synth, utils, world
from .dataloader import Loader
from .model import LightGCN


def make_scheduler(bpr, cfg):
    """The reference's optional MultiStepLR on bpr.opt (code/main.py:38-44); the fused step re-reads
    param_groups[0]['lr'] every step, so stepping the scheduler is all that is needed."""
    if not cfg.get('use_scheduler', False):
        return None
    milestones = cfg.get('sched_milestones', [120, 240, 360, 480])
    if isinstance(milestones, str):
        import ast
        milestones = ast.literal_eval(milestones)
    return torch.optim.lr_scheduler.MultiStepLR(bpr.opt, milestones=[int(m) for m in milestones],
                                                gamma=float(cfg.get('sched_gamma', 0.5)))


def save_checkpoint(path, epoch, model, bpr, best, scheduler=None, rank=0):
    if hasattr(model, '_engine'):
        model._engine.sync_params_for_read()            # row partition without the fused exchange: owners publish their rows
    opt_state = bpr.opt.state_dict()                    # collective under the row partition (gathers the moments): every rank
    if rank != 0:
        return
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    tmp = path + '.tmp'
    torch.save({'epoch': epoch, 'model_state': model.state_dict(), 'optimizer_state': opt_state,
                'scheduler_state': scheduler.state_dict() if scheduler is not None else None,
                'best_metric': float(best) if best is not None and best >= 0 else None}, tmp)   # code/main.py:58-64
    os.replace(tmp, path)                               # atomic, like code/main.py:65-67


def load_checkpoint(path, model, bpr, scheduler=None):
    ck = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(ck, dict) and 'model_state' in ck:    # new schema
        model.load_state_dict(ck['model_state'], strict=True)
        if ck.get('optimizer_state') is not None:
            bpr.opt.load_state_dict(ck['optimizer_state'])
        if scheduler is not None and ck.get('scheduler_state') is not None:
            scheduler.load_state_dict(ck['scheduler_state'])
        best = ck.get('best_metric')                    # the reference stores None until a best exists (code/main.py:63)
        return int(ck.get('epoch', 0)), (-1.0 if best is None else float(best))
    model.load_state_dict(ck, strict=True)              # legacy raw state_dict (code/main.py:80-86)
    return 0, -1.0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--synthetic', default=None, help='gowalla|yelp2018|amazon-book|tiny synthetic shape instead of files')
    ap.add_argument('--device_sampler', action='store_true')
    ap.add_argument('--resume', default=None)
    ap.add_argument('--eval_every', type=int, default=10)
    ap.add_argument('--dist_mode', default='rowpart', choices=['rowpart', 'featpart'],
                    help="under torchrun: 'rowpart' = rows of the adjacency over the ranks (graphs of any size), 'featpart' = embedding "
                         "columns over the ranks, CSR replicated, no exchange in the propagation (graphs that fit one GPU)")
    known, rest = ap.parse_known_args(argv)
    a = world.from_args(rest)
    world.configure(device_sampler=known.device_sampler)
    nranks, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    if nranks > 1:
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        world.configure(device=f'cuda:{local}', dist_mode=known.dist_mode)
    cfg = world.config
    if known.synthetic:
        ds = synth.make_dataset(known.synthetic, seed=world.seed, config=cfg)
    else:
        ds = Loader(cfg, path=a.data_path or os.path.join(world.DATA_PATH, world.dataset))
    utils.set_seed(world.seed)
    utils.sampler_seed(world.seed)
    model = LightGCN(cfg, ds)
    bpr = utils.BPRLoss(model, cfg)
    scheduler = make_scheduler(bpr, cfg)
    start, best = 0, -1.0
    last = os.path.join(world.PATH, 'last.pth.tar')
    if known.resume:
        start, best = load_checkpoint(known.resume, model, bpr, scheduler)
    for epoch in range(start + 1, world.TRAIN_epochs + 1):
        t0 = time.time()
        if (epoch - 1) % known.eval_every == 0:
            res = Procedure.Test(ds, model, epoch)
            if float(res['ndcg'][0]) > best:
                best = float(res['ndcg'][0])
                save_checkpoint(os.path.join(world.PATH, f'best-epoch{epoch}.pth.tar'), epoch, model, bpr, best, scheduler, rank)
        info = Procedure.BPR_train_original(ds, model, bpr, epoch)
        if scheduler is not None:
            scheduler.step()                            # per epoch, code/main.py:222-223
        torch.cuda.synchronize()
        if rank == 0:
            print(f'EPOCH[{epoch}/{world.TRAIN_epochs}] {info} | {time.time() - t0:.3f}s')
        if epoch % a.save_every == 0 or epoch == world.TRAIN_epochs:
            save_checkpoint(last, epoch, model, bpr, best, scheduler, rank)
    return model


if __name__ == '__main__':
    main(sys.argv[1:])
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as _dist
        if _dist.is_initialized():
            torch.cuda.synchronize()
            _dist.barrier()
            _dist.destroy_process_group()
