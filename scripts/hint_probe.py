"""K1 on a graph whose gathered table exceeds L2 (BASELINE config 5 at --scale): one layer without and with L2 eviction hints
(hot columns evict-last, the rest evict-first) for several hot-set sizes.  One JSON line per configuration.
    python scripts/hint_probe.py [--scale 1.0] [--hot 16,32,48,64,96] [--ncu]"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--hot", default="16,32,48,64,96")
ap.add_argument("--ncu", action="store_true")
a = ap.parse_args()
torch.cuda.set_device(0)
nu, ni, ne = int(10_000_000 * a.scale), int(2_000_000 * a.scale), int(500_000_000 * a.scale)
tu, ti = lg.synth.make_powerlaw_device(nu, ni, ne, seed=2020)
g = lg.ops.csr_build(tu, ti, nu, ni)
del tu, ti
torch.cuda.empty_cache()
if g.col_weight is None:
    g.col_weight = g.deg.to(torch.int32)
N, d = nu + ni, 64
X = torch.randn(N, d, device="cuda")
Y = torch.empty(N, d, device="cuda")
alg = 8 * g.nnz + 4 * (N + 1) + 8 * N * d


def layer_ms(reps=5):
    lg.ops.spmm(g, X, Y); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lg.ops.spmm(g, X, Y); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


g.clear_hints(); g._hinted = None
lg.ops.HINTS_DEFAULT = False
g._plan = None
t0 = layer_ms()
Y0 = Y.clone()
print(json.dumps({"config": f"power-law x{a.scale:g}", "nnz": g.nnz, "mode": "no hints", "layer_ms": t0, "alg_gbs": alg / (t0 * 1e-3) / 1e9}), flush=True)
if a.ncu:
    g.hint_indices(d, hot_bytes=48 << 20); g._plan = None
    lg.ops.spmm(g, X, Y); torch.cuda.synchronize()
    sys.exit(0)
for mb in [int(x) for x in a.hot.split(",")]:
    h = g.hint_indices(d, hot_bytes=mb << 20)
    g._plan = None
    t = layer_ms()
    same = bool(torch.equal(Y, Y0))
    print(json.dumps({"mode": "hinted", "hot_mb": mb, "hot_nnz_frac": float((h < 0).float().mean()), "layer_ms": t, "alg_gbs": alg / (t * 1e-3) / 1e9,
                      "speedup": t0 / t, "bit_identical": same}), flush=True)
