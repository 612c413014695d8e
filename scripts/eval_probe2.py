"""Full-ranking timing (tcgen05 path), random and trained-like embeddings, the three shapes; per-kernel breakdown via CUDA events
around the whole call.  python scripts/eval_probe2.py"""
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
for name in ("gowalla", "yelp2018", "amazon-book"):
    gr = lg.synth.make_graph(name, seed=2020)
    nu, ni = gr['n_users'], gr['m_items']
    g = lg.ops.csr_build(torch.from_numpy(gr['train_user']).cuda(), torch.from_numpy(gr['train_item']).cuda(), nu, ni)
    rng = np.random.default_rng(1)
    for kind in ("random", "trained-like"):
        if kind == "random":
            U = torch.randn(nu, 64, device="cuda") * 0.1; V = torch.randn(ni, 64, device="cuda") * 0.1
        else:
            common = rng.normal(0, 1, 64).astype(np.float32)
            pop = (1.0 / (1.0 + np.arange(ni) / 300.0)).astype(np.float32)[:, None]
            V = torch.from_numpy((0.05 * rng.normal(0, 1, (ni, 64)) + pop * common).astype(np.float32)).cuda()
            U = torch.from_numpy((0.1 * rng.normal(0, 1, (nu, 64)) + 0.3 * common).astype(np.float32)).cuda()
        users = torch.arange(nu, device="cuda")
        idx, val, redone = lg.ops.score_topk_tc(U, V, users, 20, g.indptr, g.indices, nu)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lg.ops.score_topk_tc(U, V, users, 20, g.indptr, g.indices, nu); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        sub = users[:2000].contiguous()
        ei, ev = lg.ops.score_topk(U, V, sub, 20, g.indptr, g.indices, nu)
        ok = bool(torch.equal(idx[:2000], ei) and torch.equal(val[:2000], ev))
        print(json.dumps({"shape": name, "embeddings": kind, "users": nu, "items": ni, "rank_all_ms": ms, "useful_tflops": 2.0 * nu * ni * 64 / (ms * 1e-3) / 1e12,
                          "rows_redone": int(redone), "bit_identical_to_exact_on_2000_rows": ok}), flush=True)
