"""TEST INFRASTRUCTURE — fixed-epoch goldens made by running the REAL reference's training procedure.

north_star: "Recall@20 and NDCG@20 must match to within 1e-4 after a fixed number of epochs".  This script runs,
in the build container (needs /root/reference), the reference's own

    utils.set_seed(2020) -> model.LightGCN -> utils.BPRLoss -> E x Procedure.BPR_train_original -> Procedure.Test

(code/Procedure.py:28-83,127-206, code/utils.py:38-64) with the reference's own C++ sampler (sources/sampling.cpp
compiled as it lies, seeded 2020) and its numpy shuffle, and stores what came out: per-epoch returned loss, final
parameters (tiny) or a parameter digest (gowalla), and the metrics.  Shims: the two breakages of SURVEY.md §0
(`utils.timer` is incomplete, `utils.minibatch` lacks the single-tensor case) — nothing on the arithmetic path.

The same epochs are then replayed by (i) oracle/ref_port.py on the CPU (tests/test_oracle.py) and (ii) the CUDA path
through lgcn_b200.Procedure.BPR_train_original (tests/test_gpu_model.py); both must land within 1e-4 of these metrics.

usage:  python oracle/gen_epochs_golden.py --case tiny_epochs      (~10 s)
        python oracle/gen_epochs_golden.py --case gowalla_epoch    (~10 min: one real gowalla epoch + Test on 8 cores)
"""
import argparse
import importlib
import os
import subprocess
import sys
import sysconfig
import tempfile
from time import time

import numpy as np

REF = '/root/reference/LightGCN_work/code'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

EPOCHS = {'tiny_epochs': 20, 'gowalla_epoch': 1}


class Timer:
    """Working stand-in for the reference's truncated utils.timer (API from its call sites)."""
    TAPE = [-1]
    NAMED_TAPE = {}

    @staticmethod
    def get():
        return Timer.TAPE.pop() if len(Timer.TAPE) > 1 else -1

    @staticmethod
    def dict(select_keys=None):
        return "|" + "|".join(f"{k}:{v:.2f}" for k, v in Timer.NAMED_TAPE.items()) + "|"

    @staticmethod
    def zero(select_keys=None):
        for k in Timer.NAMED_TAPE:
            Timer.NAMED_TAPE[k] = 0

    def __init__(self, tape=None, **kw):
        self.named = kw.get('name')
        if self.named:
            Timer.NAMED_TAPE.setdefault(self.named, 0.)

    def __enter__(self):
        self.start = time()
        return self

    def __exit__(self, *a):
        if self.named:
            Timer.NAMED_TAPE[self.named] += time() - self.start


def compile_reference_sampler(tmp):
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    inc = subprocess.check_output([sys.executable, '-m', 'pybind11', '--includes']).decode().split()
    subprocess.check_call(['g++', '-O2', '-std=c++11', '-shared', '-fPIC', *inc, os.path.join(REF, 'sources', 'sampling.cpp'),
                           '-o', os.path.join(tmp, 'sampling' + ext)])
    sys.path.insert(0, tmp)
    return importlib.import_module('sampling')


def write_txt(path, tu, ti, su, si):
    os.makedirs(path, exist_ok=True)
    for fname, u, i in (('train.txt', tu, ti), ('test.txt', su, si)):
        with open(os.path.join(path, fname), 'w') as f:
            if len(u) == 0:
                continue
            cuts = np.flatnonzero(np.diff(u)) + 1
            for uu, its in zip(u[np.concatenate([[0], cuts])], np.split(i, cuts)):
                f.write(f"{int(uu)} {' '.join(map(str, its.tolist()))}\n")


def run(case):
    import torch
    tmp = tempfile.mkdtemp(prefix=f'golden_{case}_')
    E = EPOCHS[case]
    if case == 'tiny_epochs':
        from oracle.gen_golden import make_case, write_case
        _, _, train, test = make_case('tiny')
        data_dir = os.path.join(tmp, 'data', 'tiny')
        write_case(data_dir, train, test)
        d, L, B, name = 64, 3, 256, 'tiny'
    else:
        z = np.load(os.path.join(GOLDEN, 'gowalla.npz'))
        nu = int(z['n_users'])
        tu = np.repeat(np.arange(nu), np.diff(z['train_indptr'])).astype(np.int64)
        su = np.repeat(z['test_users'].astype(np.int64), np.diff(z['test_indptr']))
        data_dir = os.path.join(tmp, 'data', 'gowalla')
        write_txt(data_dir, tu, z['train_items'].astype(np.int64), su, z['test_items'].astype(np.int64))
        # graph cache: the reference loads s_pre_adj_mat.npz blindly if present (code/dataloader.py:210-216); the bmat
        # path builds the bit-identical matrix in 0.1 s instead of 84 s
        import scipy.sparse as sp
        from oracle import ref_port
        _, norm_adj, _ = ref_port.build_graph(tu, z['train_items'].astype(np.int64), nu, int(z['m_items']))
        sp.save_npz(os.path.join(data_dir, 's_pre_adj_mat.npz'), norm_adj)
        d, L, B, name = 64, 3, 2048, 'gowalla'
    sys.argv = ['x', '--dataset', name, '--tensorboard', '0', '--checkpoint_dir', os.path.join(tmp, 'ckpt'),
                '--recdim', str(d), '--layer', str(L), '--topks', '[20]', '--bpr_batch', str(B)]
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import world
    world.device = torch.device('cpu')
    import dataloader, model, utils, Procedure          # noqa: E401

    def minibatch(*tensors, **kwargs):
        bs = kwargs.get('batch_size', world.config['bpr_batch_size'])
        if len(tensors) == 1:
            for i in range(0, len(tensors[0]), bs):
                yield tensors[0][i:i + bs]
        else:
            for i in range(0, len(tensors[0]), bs):
                yield tuple(x[i:i + bs] for x in tensors)
    utils.minibatch = minibatch
    utils.timer = Timer
    Procedure.timer = Timer
    sampling = compile_reference_sampler(tmp)
    utils.sampling, utils.sample_ext = sampling, True

    torch.set_num_threads(1 if case == 'tiny_epochs' else (os.cpu_count() or 1))
    ds = dataloader.Loader(world.config, path=data_dir)
    utils.set_seed(2020)
    sampling.seed(2020)
    m = model.LightGCN(world.config, ds)
    bpr = utils.BPRLoss(m, world.config)
    infos, first_S = [], None
    orig_sample = utils.UniformSample_original

    def recording_sampler(dataset, neg_ratio=1):
        nonlocal first_S
        S = orig_sample(dataset, neg_ratio)
        if first_S is None:
            first_S = np.asarray(S).copy()
        return S
    utils.UniformSample_original = recording_sampler
    t0 = time()
    for epoch in range(1, E + 1):
        infos.append(Procedure.BPR_train_original(ds, m, bpr, epoch))
        print(f"[{case}] epoch {epoch}: {infos[-1]} ({time() - t0:.1f} s)", flush=True)
    res = Procedure.Test(ds, m, E)
    P = torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().numpy()
    losses = np.array([float(s[4:s.index('-')]) for s in infos])
    out = dict(epochs=E, d=d, L=L, batch=B, seed=2020, lr=world.config['lr'], decay=world.config['decay'],
               epoch_loss_3dp=losses, precision=res['precision'], recall=res['recall'], ndcg=res['ndcg'],
               first_epoch_triples_head=first_S[:64].astype(np.int32), first_epoch_triples_sum=first_S.astype(np.int64).sum(axis=0))
    if case == 'tiny_epochs':
        out['params'] = P
    else:           # digest: enough to catch a drift, small enough to commit
        out['params_head'] = P[:256].copy()
        out['params_row_norms_sample'] = np.linalg.norm(P[::97], axis=1)
        out['params_sum'] = P.astype(np.float64).sum()
    np.savez_compressed(os.path.join(GOLDEN, f'{case}.npz'), **out)
    print(f"[golden] {case}: losses {losses.tolist()} recall {res['recall']} ndcg {res['ndcg']} precision {res['precision']}")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', required=True, choices=list(EPOCHS))
    run(ap.parse_args().case)
