"""How fast can SM stores push a row block into a peer GPU?  torch's elementwise copy kernel (vectorised, coalesced)
versus the stores fused into the SpMM epilogue.  torchrun --nproc-per-node 2 scripts/p2p_store_bw.py"""
import os, sys, json, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import lgcn_b200 as lg
from lgcn_b200.engine import map_peer_buffers
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for p in range(torch.cuda.device_count()):
    lg._lib.load().lgcn_enable_peer_access(p)
rows = 700_000
buf = torch.zeros((2 * rows, 64), device=f"cuda:{lr}")
peers = map_peer_buffers(buf)
src = torch.randn((rows, 64), device=f"cuda:{lr}")
dst = peers[(rank + 1) % world][rank * rows:(rank + 1) * rows]
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timed(lambda: dst.copy_(src))
local = torch.empty_like(src)
ms_local = timed(lambda: local.copy_(src))
out = {"rank": rank, "bytes": src.numel() * 4, "peer_copy_ms": ms, "peer_GBs": src.numel() * 4 / ms / 1e6, "local_copy_ms": ms_local}
# the fused path: an identity-like SpMM (diagonal graph) storing into the peer as well
n = rows
idx = torch.arange(n, device=f"cuda:{lr}")
g = lg.ops.coo_to_csr(idx, idx, torch.ones(n, device=f"cuda:{lr}"), n, n)
Y = torch.empty_like(src)
out["spmm_diag_local_ms"] = timed(lambda: lg.ops.spmm(g, src, Y))
out["spmm_diag_fused_peer_ms"] = timed(lambda: lg.ops.spmm(g, src, Y, peer_y=[dst]))
print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
