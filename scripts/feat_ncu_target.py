"""ncu target: K1 at the slice widths of the feature partition (d = 64 / 32 / 16 / 8) on the amazon-book shape, L2-warm
(two launches per width, the second is the one to read).  python scripts/feat_ncu_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
gr = lg.synth.make_graph("amazon-book", seed=2020)
nu, ni = gr['n_users'], gr['m_items']
csr0 = lg.ops.csr_build(torch.from_numpy(gr['train_user']).cuda(), torch.from_numpy(gr['train_item']).cuda(), nu, ni)
for d in (64, 32, 16, 8):
    csr = csr0.rows(0, csr0.n_rows, seg_len=64 if d == 16 else 128)
    X = (0.1 * torch.randn((csr.n_rows, d), device="cuda")).contiguous(); Y = torch.empty_like(X)
    lg.ops.spmm(csr, X, Y); lg.ops.spmm(csr, Y, X)
    torch.cuda.synchronize()
print("done")
