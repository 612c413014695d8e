"""Where a row-partitioned layer's time goes (N ranks, torchrun): local K1, K1 + multicast row stores, the device barrier,
the exchanged layer — each captured in a CUDA graph of REPS back-to-back launches so that host launch latency is out.
    torchrun --nproc-per-node N scripts/rowpart_probe.py [--workload gowalla]"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgcn_b200 as lg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="gowalla")
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--seg_lens", default="128")
ap.add_argument("--multicast", default="1")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
lg.world.configure(device=f"cuda:{lr_}")
g = lg.synth.make_graph(a.workload, seed=2020)


def graph_us(fn, reps=a.reps, outer=9):
    fn(); torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(outer):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); e1.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1) / reps)
    return statistics.median(ts)


for seg_len in [int(x) for x in a.seg_lens.split(",")]:
  for use_mc in [bool(int(x)) for x in a.multicast.split(",")]:
    cfg = dict(lg.world.config); cfg.update(dist_mode='rowpart', spmm_seg_len=seg_len, rowpart_multicast=use_mc)
    ds = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'], config=cfg)
    lg.utils.set_seed(2020)
    m = lg.LightGCN(cfg, ds)
    eng = m._engine
    r0, r1, d = eng.r0, eng.r1, eng.d
    mc = eng._mc.get(eng.X[0].data_ptr(), 0)
    Y = torch.empty((r1 - r0, d), device="cuda")
    res = {
        "k1_local_us": graph_us(lambda: lg.ops.spmm(eng.local, eng.E0, Y)),
        "k1_multicast_us": graph_us(lambda: lg.ops.spmm(eng.local, eng.E0, eng.X[0][r0:r1], mc_y=mc + r0 * d * 4)) if mc else None,
        "barrier_us": graph_us(eng._rank_barrier),
        "layer_us": graph_us(lambda: eng._layer(eng.E0, eng.X[0], 1.0, 0.0, None)),
        "two_layers_pingpong_us": graph_us(lambda: (eng._layer(eng.E0, eng.X[0], 1.0, 0.0, None), eng._layer(eng.X[0], eng.X[1], 1.0, 0.0, None))) / 2,
    }
    eng._barrier.check()
    allr = [None] * world
    dist.all_gather_object(allr, {k: (round(v, 2) if v is not None else None) for k, v in res.items()} | {"rows": r1 - r0, "nnz": eng.local.nnz})
    if rank == 0:
        print(json.dumps({"workload": a.workload, "n_gpus": world, "seg_len": seg_len, "multicast": bool(mc), "per_rank": allr}), flush=True)
    del m, eng, ds
    torch.cuda.empty_cache()
dist.destroy_process_group()
