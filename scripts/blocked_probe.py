"""K1 on a graph whose gathered table exceeds L2 (BASELINE config 5 at --scale): one layer unblocked vs column-slab blocked
at several slab sizes.  One JSON line per configuration.  With --ncu: exactly one unblocked layer and one blocked layer
(default slab size) are launched after the setup, for `ncu -k regex:spmm_kernel --metrics dram__bytes_read.sum,...`.
    python scripts/blocked_probe.py [--scale 1.0] [--slabs 32,48,64,80,96] [--ncu]"""
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
ap.add_argument("--slabs", default="32,48,64,80,96")
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--rows", default="all", choices=["all", "users", "items"])
a = ap.parse_args()
torch.cuda.set_device(0)
nu, ni, ne = int(10_000_000 * a.scale), int(2_000_000 * a.scale), int(500_000_000 * a.scale)
tu, ti = lg.synth.make_powerlaw_device(nu, ni, ne, seed=2020)
g = lg.ops.csr_build(tu, ti, nu, ni)
del tu, ti
torch.cuda.empty_cache()
if a.rows == "users":
    g = g.rows(0, nu)
elif a.rows == "items":
    g = g.rows(nu, nu + ni)
N, d = nu + ni, 64
X = torch.randn(N, d, device="cuda")
Y = torch.empty(g.n_rows, d, device="cuda")
alg = 8 * g.nnz + 4 * (g.n_rows + 1) + 4 * N * d + 4 * g.n_rows * d


def layer_ms(reps=5):
    lg.ops.spmm(g, X, Y); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lg.ops.spmm(g, X, Y); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


if a.ncu:
    g.use_blocking = False
    lg.ops.spmm(g, X, Y); torch.cuda.synchronize()
    Y0 = Y.clone()
    g.use_blocking = True
    g.block_plans(d, slab_bytes=lg.ops.SLAB_BYTES)
    lg.ops.spmm(g, X, Y); torch.cuda.synchronize()
    print(json.dumps({"ncu": True, "nnz": g.nnz, "alg_bytes": alg, "slabs": g.n_slabs, "block_items": g.n_block_items,
                      "max_abs_diff_vs_unblocked": float((Y - Y0).abs().max())}))
    sys.exit(0)
g.use_blocking = False
t0 = layer_ms()
Y0 = Y.clone()
print(json.dumps({"config": f"power-law x{a.scale:g} rows={a.rows}", "nnz": g.nnz, "rows": g.n_rows, "alg_bytes": alg, "mode": "unblocked",
                  "layer_ms": t0, "alg_gbs": alg / (t0 * 1e-3) / 1e9}), flush=True)
g.use_blocking = True
for mb in [int(x) for x in a.slabs.split(",")]:
    g.clear_blocking()
    torch.cuda.empty_cache()
    plans = g.block_plans(d, slab_bytes=mb << 20)
    for ipg, var in ((1, 101), (2, 103), (4, 100), (8, 102)):
        lg._lib.load().lgcn_debug_spmm_variant(var)
        t = layer_ms()
        err = float((Y - Y0).abs().max() / Y0.abs().max())
        print(json.dumps({"mode": "blocked", "slab_mb": mb, "items_per_group": ipg, "slabs": g.n_slabs, "launches": len(plans), "block_items": g.n_block_items,
                          "pair_traffic_gb": g.n_block_items * 528 / 1e9, "layer_ms": t, "alg_gbs": alg / (t * 1e-3) / 1e9,
                          "speedup_vs_unblocked": t0 / t, "max_rel_diff": err}), flush=True)
    lg._lib.load().lgcn_debug_spmm_variant(100)
