"""getUsersRating timing: tensor-core 3xTF32 kernel vs the exact CUDA-core kernel vs torch.matmul (cuBLAS SGEMM), warm.
    python scripts/dense_probe.py"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

torch.cuda.set_device(0)
torch.backends.cuda.matmul.allow_tf32 = False


def med_us(fn, reps=11):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(1e3 * a.elapsed_time(b))
    return statistics.median(ts)


for name, nu, ni in (("gowalla", 29858, 40981), ("yelp2018", 31668, 38048), ("amazon-book", 52643, 91599)):
    U = torch.randn(nu, 64, device="cuda") * 0.1; V = torch.randn(ni, 64, device="cuda") * 0.1
    for Bt in (100, 2048):
        users = torch.randperm(nu, device="cuda")[:Bt]
        rec = {"shape": name, "users": Bt, "items": ni, "out_mb": Bt * ni * 4 / 1e6,
               "tc_3xtf32_us": med_us(lambda: lg.ops.score_dense_tc(U, V, users)),
               "exact_cuda_core_us": med_us(lambda: lg.ops.score_dense(U, V, users)),
               "torch_matmul_fp32_us": med_us(lambda: torch.matmul(U[users], V.t()))}
        rec["hbm_write_bound_us"] = rec["out_mb"] * 1e6 / 6.55e12 * 1e6
        print(json.dumps(rec), flush=True)
