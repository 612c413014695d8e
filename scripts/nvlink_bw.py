"""Raw GPU-to-GPU copy bandwidth on this box (one process, copy engine and SM-driven P2P), for the rowpart numbers."""
import torch, time, json
n = torch.cuda.device_count()
out = {"gpus": n}
if n >= 2:
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0")      # 1 GiB
    b = torch.empty(1 << 28, dtype=torch.float32, device="cuda:1")
    for name, fn in (("copy_engine_0to1", lambda: b.copy_(a, non_blocking=True)),):
        for _ in range(2): fn()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        out[name + "_GBs"] = 5 * a.numel() * 4 / (time.perf_counter() - t0) / 1e9
    # SM-driven: a kernel on GPU 0 writing into GPU 1's memory (peer access through torch's UVA mapping)
    torch.cuda.set_device(0)
    try:
        bb = b  # torch enables peer access lazily for cross-device elementwise ops
        for _ in range(2): torch.add(a, 1.0, out=a)
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        s.record()
        for _ in range(5): b.copy_(a)          # same-process D2D across devices
        e.record(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        out["copy_0to1_events_GBs"] = 5 * a.numel() * 4 / (s.elapsed_time(e) * 1e-3) / 1e9
    except Exception as ex:
        out["sm_error"] = str(ex)
    # both directions at once
    c = torch.empty_like(a); d = torch.empty_like(b)
    s0 = torch.cuda.Stream(device=0); s1 = torch.cuda.Stream(device=1)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s0): b.copy_(a, non_blocking=True)
        with torch.cuda.stream(s1): c.copy_(d, non_blocking=True)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    out["bidirectional_sum_GBs"] = 10 * a.numel() * 4 / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
