"""Working driver with the shape of the reference's (unparseable) main.py epoch loop, code/main.py:185-242:
every 10th epoch Test + best-NDCG checkpoint, then one BPR epoch, CSV rows, atomic `last.pth.tar`.

    python -m lgcn_b200.train --dataset gowalla --data_path data/gowalla --epochs 50
    python -m lgcn_b200.train --synthetic yelp2018 --epochs 20 --device_sampler

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


def save_checkpoint(path, epoch, model, bpr, best):
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    tmp = path + '.tmp'
    torch.save({'epoch': epoch, 'model_state': model.state_dict(), 'optimizer_state': bpr.opt.state_dict(),
                'scheduler_state': None, 'best_metric': best}, tmp)
    os.replace(tmp, path)                               # atomic, like code/main.py:65-67


def load_checkpoint(path, model, bpr):
    ck = torch.load(path, map_location='cpu', weights_only=False)
    if 'model_state' in ck:                             # new schema
        model.load_state_dict(ck['model_state'], strict=True)
        if ck.get('optimizer_state') is not None:
            bpr.opt.load_state_dict(ck['optimizer_state'])
        return int(ck.get('epoch', 0)), ck.get('best_metric', -1.0)
    model.load_state_dict(ck, strict=True)              # legacy raw state_dict (code/main.py:80-86)
    return 0, -1.0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--synthetic', default=None, help='gowalla|yelp2018|amazon-book|tiny synthetic shape instead of files')
    ap.add_argument('--device_sampler', action='store_true')
    ap.add_argument('--resume', default=None)
    ap.add_argument('--eval_every', type=int, default=10)
    known, rest = ap.parse_known_args(argv)
    a = world.from_args(rest)
    world.configure(device_sampler=known.device_sampler)
    cfg = world.config
    if known.synthetic:
        ds = synth.make_dataset(known.synthetic, seed=world.seed, config=cfg)
    else:
        ds = Loader(cfg, path=a.data_path or os.path.join(world.DATA_PATH, world.dataset))
    utils.set_seed(world.seed)
    utils.sampler_seed(world.seed)
    model = LightGCN(cfg, ds)
    bpr = utils.BPRLoss(model, cfg)
    start, best = 0, -1.0
    last = os.path.join(world.PATH, 'last.pth.tar')
    if known.resume:
        start, best = load_checkpoint(known.resume, model, bpr)
    for epoch in range(start + 1, world.TRAIN_epochs + 1):
        t0 = time.time()
        if (epoch - 1) % known.eval_every == 0:
            res = Procedure.Test(ds, model, epoch)
            if float(res['ndcg'][0]) > best:
                best = float(res['ndcg'][0])
                save_checkpoint(os.path.join(world.PATH, f'best-epoch{epoch}.pth.tar'), epoch, model, bpr, best)
        info = Procedure.BPR_train_original(ds, model, bpr, epoch)
        torch.cuda.synchronize()
        print(f'EPOCH[{epoch}/{world.TRAIN_epochs}] {info} | {time.time() - t0:.3f}s')
        if epoch % a.save_every == 0 or epoch == world.TRAIN_epochs:
            save_checkpoint(last, epoch, model, bpr, best)
    return model


if __name__ == '__main__':
    main(sys.argv[1:])
