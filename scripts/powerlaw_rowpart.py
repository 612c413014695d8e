"""BASELINE config 5: scaled power-law graph, adjacency row-partitioned over the ranks, per-layer all-gather.
torchrun --nproc-per-node N scripts/powerlaw_rowpart.py [--users U --items I --edges E --steps K]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lgcn_b200 as lg   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=10_000_000)
ap.add_argument("--items", type=int, default=2_000_000)
ap.add_argument("--edges", type=int, default=500_000_000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--row_cost", type=int, default=-1)
ap.add_argument("--no_p2p", action="store_true")
ap.add_argument("--no_multicast", action="store_true")
ap.add_argument("--rebalance", type=int, default=2)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
lg.world.configure(device=f"cuda:{lr_}")
cfg = dict(lg.world.config)
cfg.update(dist_mode='rowpart' if world > 1 else None, cuda_graph=False, bpr_batch_size=2048,
           rowpart_row_cost=None if a.row_cost < 0 else a.row_cost, rowpart_p2p=not a.no_p2p, rowpart_multicast=not a.no_multicast, rowpart_rebalance=a.rebalance)
t0 = time.perf_counter()
tu, ti = lg.synth.make_powerlaw_device(a.users, a.items, a.edges, seed=2020)
torch.cuda.synchronize(); t_gen = time.perf_counter() - t0
ds = lg.synth.DeviceGraphDataset(a.users, a.items, tu, ti)
del tu, ti
t0 = time.perf_counter()
g = ds.getCSRGraph()
torch.cuda.synchronize(); t_build = time.perf_counter() - t0
torch.cuda.empty_cache()
lg.utils.set_seed(2020)
model = lg.LightGCN(cfg, ds)
eng = model._engine
S = lg.ops.sample_bpr(g, a.users, a.items, min(ds.trainDataSize, a.users * 4), seed=2020, epoch=0)
B = 2048


def step(i):
    lo = (i * B) % (S.shape[1] - B)
    eng.step(S[0, lo:lo + B], S[1, lo:lo + B], S[2, lo:lo + B])


for i in range(a.warmup):
    step(i)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    step(a.warmup + i)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# one local SpMM layer alone (no exchange), for the per-rank roofline
Y = torch.empty_like(eng.out[eng.r0:eng.r1])
for _ in range(2):
    lg.ops.spmm(eng.local, eng.E0, Y)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    lg.ops.spmm(eng.local, eng.E0, Y)
e1.record(); torch.cuda.synchronize()
spmm_ms = e0.elapsed_time(e1) / 5
loss = float(eng.loss_to_host()[2])
# where a step's time goes: the forward propagation alone (L layers with their exchanges), one exchanged layer alone
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(reps):
        fn()
    a1.record(); torch.cuda.synchronize()
    t = torch.tensor([a0.elapsed_time(a1) / reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
r0_, r1_ = eng.r0, eng.r1
def _spmm_only():
    mc = eng._mc.get(eng.X[0].data_ptr(), 0) if getattr(eng, "_mc", None) else 0
    peers = eng._peer.get(eng.X[0].data_ptr()) if eng.p2p else None
    if mc:
        lg.ops.spmm(eng.local, eng.E0, eng.X[0][r0_:r1_], mc_y=mc + r0_ * eng.d * 4)
    elif peers is not None:
        lg.ops.spmm(eng.local, eng.E0, eng.X[0][r0_:r1_], peer_y=[peers[p][r0_:r1_] for p in range(world) if p != rank])
    else:
        lg.ops.spmm(eng.local, eng.E0, eng.X[0][r0_:r1_])
def timed_local(fn, reps=5):          # this rank's own time, no max over ranks
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(reps):
        fn()
    a1.record(); torch.cuda.synchronize()
    return a0.elapsed_time(a1) / reps
mine = torch.tensor([timed_local(_spmm_only), timed_local(lambda: lg.ops.spmm(eng.local, eng.E0, Y)),
                     timed_local(eng._rank_barrier) if world > 1 and eng.p2p else 0.0], device="cuda", dtype=torch.float64)
per_rank = [torch.zeros_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(per_rank, mine)
else:
    per_rank = [mine]
per_rank = [[round(float(x), 3) for x in t.tolist()] for t in per_rank]
fwd_ms = timed(lambda: eng.forward())
layer_ms = timed(lambda: eng._layer(eng.E0, eng.X[0], 1.0, 0.0, None))
if rank == 0:
    nnz_local = eng.local.nnz
    alg = 8 * nnz_local + 4 * (eng.local.n_rows + 1) + 4 * eng.N * 64 + 4 * eng.local.n_rows * 64
    print(json.dumps({"config": f"power-law {a.users} x {a.items}, {a.edges} edges, rowpart over {world} GPUs", "n_gpus": world,
                      "nnz": g.nnz, "n_long": g.n_long, "n_segs": g.n_segs, "gen_s": t_gen, "csr_build_s": t_build,
                      "ms_per_step": float(ms.item()), "samples_per_s": B / (float(ms.item()) * 1e-3), "loss": loss,
                      "per_rank_ms_[spmm+stores, spmm_local, barrier]": per_rank, "forward_ms": fwd_ms, "exchanged_layer_ms": layer_ms, "local_spmm_ms": spmm_ms, "local_spmm_alg_gbs": alg / (spmm_ms * 1e-3) / 1e9,
                      "rows_local": eng.local.n_rows, "nnz_local": nnz_local, "bounds": eng.bounds, "p2p": eng.p2p, "multicast": bool(getattr(eng, "_mc", None)),
                      "mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
if world > 1:
    dist.destroy_process_group()
