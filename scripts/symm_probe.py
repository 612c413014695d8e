"""Does this box expose NVSwitch multicast to torch's symmetric memory?  torchrun --nproc-per-node N scripts/symm_probe.py"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm_mem.empty((1024, 64), dtype=torch.float32, device=f"cuda:{lr}")
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "world", h.world_size, flush=True)
t.fill_(float(rank) + 1)
h.barrier()
peer = h.get_buffer((rank + 1) % world, (1024, 64), torch.float32)
print(rank, "peer value", float(peer[0, 0]), flush=True)
h.barrier()
dist.destroy_process_group()
