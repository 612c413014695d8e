"""K1 on the column slices of the feature partition (d = 32 / 16 / 8): the row-per-group kernel (variant 50), the lane-per-non-zero
kernel (spmm_narrow_kernel) and its tuning variants; cold single launches and launches inside a CUDA graph on L2-warm data (what a
layer costs inside the captured step).  python scripts/feat_probe.py [shape ...]"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
lib = lg._lib.load()
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda").view(torch.int64)


def med_us(fn, reps=11, flush=True, per=1):
    ts = []
    for _ in range(reps):
        if flush:
            flush_buf.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(1e3 * a.elapsed_time(b) / per)
    return statistics.median(ts)


VARIANTS = {64: [0], 32: [0, 51, 54], 16: [0], 8: [50, 0]}      # 0 = shipped; 51/54: d = 32 with 8 lanes x 4 in flight / the 4-lane layout; 50: row-per-group kernel at d = 8
for name in sys.argv[1:] or ["gowalla", "amazon-book"]:
    gr = lg.synth.make_graph(name, seed=2020)
    nu, ni = gr['n_users'], gr['m_items']
    csr0 = lg.ops.csr_build(torch.from_numpy(gr['train_user']).cuda(), torch.from_numpy(gr['train_item']).cuda(), nu, ni)
    for d in (64, 32, 16, 8):
        for seg in ((128, 64) if d <= 32 else (128,)):
            csr = csr0.rows(0, csr0.n_rows, seg_len=seg)
            X = (0.1 * torch.randn((csr.n_rows, d), device="cuda")).contiguous(); Y = torch.empty_like(X)
            for v in VARIANTS[d]:
                lib.lgcn_debug_spmm_variant(v)
                lg.ops.spmm(csr, X, Y); lg.ops.spmm(csr, Y, X); torch.cuda.synchronize()
                cold = med_us(lambda: lg.ops.spmm(csr, X, Y))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(3):
                        lg.ops.spmm(csr, X, Y); lg.ops.spmm(csr, Y, X)
                g.replay(); torch.cuda.synchronize()

                def many():
                    for _ in range(5):
                        g.replay()
                warm = med_us(many, flush=False, per=30)
                first_cold = med_us(g.replay, per=6)
                print(json.dumps({"shape": name, "d_slice": d, "seg_len": seg, "variant": v, "cold_us": round(cold, 2),
                                  "in_graph_warm_us": round(warm, 2), "in_graph_after_flush_us": round(first_cold, 2)}), flush=True)
                del g
                X.normal_(0, 0.1)
    lib.lgcn_debug_spmm_variant(0)
