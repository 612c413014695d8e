"""The L2 -> SM gather ceiling of this box (bench.py `roofline_l2`): uniform random 256-byte row gathers from tables of several
sizes (L2-resident up to ~100 MB, HBM beyond), every variant of lgcn_debug_gather_rows, warm L2.  One JSON line per table size.
    python scripts/l2_gather_ceiling.py > gpurun_out/l2_gather_ceiling.jsonl"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
n_idx = 1 << 24
gen = torch.Generator(device="cuda").manual_seed(1)
for rows in (16_384, 70_839, 144_242, 400_000, 1_000_000, 4_000_000, 12_000_000):
    X = torch.randn(rows, 64, device="cuda")
    idx = torch.randint(0, rows, (n_idx,), device="cuda", generator=gen, dtype=torch.int32)
    rec = {"table_rows": rows, "table_mb": rows * 256 / 1e6, "gathers": n_idx}
    for run in (32, 128):
        for variant in range(4):
            out = lg.ops.gather_probe(X, idx, run=run, variant=variant)
            ts = []
            for _ in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); lg.ops.gather_probe(X, idx, run=run, variant=variant, out=out); b.record(); b.synchronize()
                ts.append(a.elapsed_time(b))
            rec[f"run{run}_v{variant}_gbs"] = round(n_idx * 256 / (statistics.median(ts) * 1e-3) / 1e9, 1)
    rec["best_gbs"] = max(v for k, v in rec.items() if k.endswith("_gbs"))
    print(json.dumps(rec), flush=True)
    del X, idx
