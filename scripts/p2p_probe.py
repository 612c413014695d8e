"""Probe: can ranks map each other's torch allocations (CUDA IPC) and store into them from a kernel?"""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = torch.full((1024, 64), float(rank), device=f"cuda:{lr}")
meta = t.untyped_storage()._share_cuda_()
objs = [None] * world
dist.all_gather_object(objs, (lr, meta, t.storage_offset(), tuple(t.shape)))
peers = []
for p, (dev, m, off, shape) in enumerate(objs):
    if p == rank:
        peers.append(t); continue
    st = torch.UntypedStorage._new_shared_cuda(lr, *m[1:])        # open the handle in MY device's context (lazy peer access)
    pt = torch.empty(0, dtype=torch.float32, device=f"cuda:{lr}").set_(st, off, shape)
    peers.append(pt)
print(rank, "peer devices", [str(x.device) for x in peers], "can_access", [torch.cuda.can_device_access_peer(lr, p) for p in range(world) if p != lr], flush=True)
dist.barrier(); torch.cuda.synchronize()
# write my rank id + 10 into row `rank` of every peer's tensor, from MY device, with our own kernel (clear_rows writes zeros; use spmm? simplest: torch copy_ from local)
src = torch.full((64,), 10.0 + rank, device=f"cuda:{lr}")
for p in range(world):
    peers[p][rank].copy_(src)          # P2P store through torch (enables peer access)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
print(rank, "rows now", t[:world, 0].tolist(), flush=True)
# raw-pointer access from our library: spmm with Y = peer buffer
import lgcn_b200 as lg
import numpy as np
lg.world.configure(device=f"cuda:{lr}")
g = lg.ops.csr_build(torch.tensor([0, 1, 2], device=f"cuda:{lr}"), torch.tensor([0, 1, 2], device=f"cuda:{lr}"), 8, 8)
X = torch.ones((16, 64), device=f"cuda:{lr}")
nxt = peers[(rank + 1) % world]
Yview = nxt[100:116]
import ctypes
lib = lg._lib.load()
for p in range(world):
    print(rank, 'enable peer', p, lib.lgcn_enable_peer_access(p), lib.lgcn_last_error(), flush=True)
out = (ctypes.c_int32 * 4)()
st_ = torch.cuda.current_stream().cuda_stream
rc = lib.lgcn_debug_poke(t[200:201].data_ptr(), 5.0, 64, out, st_)
print(rank, "poke local rc", rc, list(out), lib.lgcn_last_error(), flush=True)
rc = lib.lgcn_debug_poke(nxt[201:202].data_ptr(), 7.0 + rank, 64, out, st_)
print(rank, "poke peer rc", rc, list(out), lib.lgcn_last_error(), flush=True)
dist.barrier(); torch.cuda.synchronize()
print(rank, "after poke", t[200:202, 0].tolist(), flush=True)
dist.destroy_process_group()
