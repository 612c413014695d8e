#!/usr/bin/env python
"""bench.py — BPR train samples/s (+ SpMM propagation GB/s roofline) for LightGCN L=3 d=64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload yelp2018]

A "step" is one BPR training step (B = 2048 sampled triples): 3 forward SpMM layers, fused BPR loss +
gradient, 3 backward SpMM layers, Adam.  N = 1 runs BASELINE.json configs[1] (yelp2018-shape synthetic
graph).  N > 1 (launched by torchrun) runs data-parallel replicas of the same graph, batch sharded
(global batch N x 2048, gradient all-reduce) — weak scaling; `--parallel rowpart` switches to the
row-partitioned adjacency (strong scaling of one 2048-triple step).

value   : samples/s with the epoch's triples already resident in HBM (CUDA events, L2 flushed between
          steps, max over ranks).
e2e     : the same metric through the public API utils.BPRLoss.stageOne(users, pos, neg) with pinned
          HOST tensors: one H2D per step inside the timed region and a D2H read of the loss.
roofline: the dominant kernel (spmm_kernel<64>) — algorithmic bytes B_spmm = 8 nnz + 4 (N+1) + 8 N d per
          launch / mean launch duration measured with CUDA events in an instrumented pass of the same steps.
cpu_baseline: the oracle's torch-CPU port of the reference (oracle/ref_port.py) on this box's host cores.
--impl reference: times that port alone (the reference's own CPU implementation does not travel).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "BPR train samples/s, LightGCN L=3 d=64"
UNIT = "samples/s"
B = 2048


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_graph(name):
    import lgcn_b200 as lg
    return lg.synth.make_graph(name, seed=2020)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread every millisecond (the timed region lasts tens of milliseconds — `nvidia-smi -lms` takes longer than that to
    start); nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.proc, self.index, self.thread, self.nvml = None, index, None, None
        self.sm, self.bits, self.stop_flag = [], 0, False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
            for cand in (uuid, "GPU-" + uuid):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                except Exception:
                    pass
        except Exception:
            pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll(self):
        nv, h = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            self.nvml = self._nvml_handle()
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv, h = self.nvml
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(n for n, bit in self.REASONS if self.bits & bit), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_port_bench(graph, steps, warmup, budget_s=None, seed=2020):
    """Times oracle/ref_port (torch CPU, all host threads) stageOne on the same workload.  The only place
    outside tests/smoke where oracle/ is executed: it is the thing being measured as the CPU baseline."""
    import numpy as np
    import torch
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nu, ni = graph['n_users'], graph['m_items']
    t0 = time.perf_counter()
    g, _, _ = ref_port.build_graph(graph['train_user'], graph['train_item'], nu, ni)
    t_graph = time.perf_counter() - t0
    torch.manual_seed(seed)
    model = ref_port.RefLightGCN(nu, ni, 64, 3, g)
    bpr = ref_port.RefBPRLoss(model, 1e-4, 1e-3)
    rng = np.random.default_rng(seed)
    times = []
    for s in range(warmup + steps):
        u = torch.from_numpy(rng.integers(0, nu, B)); p = torch.from_numpy(rng.integers(0, ni, B)); n = torch.from_numpy(rng.integers(0, ni, B))
        t0 = time.perf_counter()
        bpr.stageOne(u, p, n)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if budget_s is not None and s >= warmup and sum(times) > budget_s:
            break
    total = sum(times)
    return {"value": len(times) * B / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} stageOne steps of B={B} after {warmup} warm-up (torch {torch.__version__} CPU, {cores} threads); graph build {t_graph:.2f} s (bmat path)",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    graph = workload_graph(args.workload)
    steps = min(args.steps, 20)
    r = cpu_port_bench(graph, steps, min(args.warmup, 3))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": min(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shape synthetic graph, LightGCN L=3 d=64, BPR batch {B}, reference CPU path (oracle port)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import lgcn_b200 as lg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lg.world.configure(device=f"cuda:{local_rank}", checkpoint_dir="/tmp/lgcn_b200_bench")
    cfg = dict(lg.world.config)
    mode = None
    if world > 1:
        mode = args.parallel
        cfg.update(dist_mode=mode)
    graph = workload_graph(args.workload)
    ds = lg.InteractionDataset(graph['n_users'], graph['m_items'], graph['train_user'], graph['train_item'],
                               graph['test_user'], graph['test_item'], config=cfg, name=args.workload)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csr = ds.getCSRGraph()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    lg.utils.set_seed(2020)
    model = lg.LightGCN(cfg, ds)
    bpr = lg.utils.BPRLoss(model, cfg)
    eng = model._engine
    lg.utils.sampler_seed(2020 + (rank if mode in ('dp', 'dp_idx') else 0))
    S = lg.utils.UniformSample_original(ds)
    np.random.seed(2020 + (rank if mode in ('dp', 'dp_idx') else 0))
    perm = np.arange(S.shape[0]); np.random.shuffle(perm)
    S_host = torch.from_numpy(np.ascontiguousarray(S[perm, :3].T)).to(torch.int64).pin_memory()
    n_batches = S_host.shape[1] // B
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda").view(torch.int64)

    def flush_l2():
        # read 512 MiB (4x the 126 MB L2): everything the step touched is evicted and the lines left behind are
        # clean, so the timed kernels neither hit stale data nor pay for write-backs of the flush itself
        flush.sum()
    K, W = args.steps, args.warmup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            flush_l2()
            evs[i][0].record()
            step_fn(i)
            evs[i][1].record()
        barrier()
        wall = time.perf_counter() - wall0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # ---- leg 1: inputs resident in HBM ----------------------------------------------------------
    if mode is None:
        n_epoch_steps = eng.begin_epoch(S_host.cuda())
        pos = [0]

        def step1(i):
            if pos[0] >= n_epoch_steps - 1:      # keep to full batches; wrap around to the start of the resident epoch
                eng.rewind_epoch(); pos[0] = 0
            eng.epoch_step(); pos[0] += 1
        for i in range(W):
            step1(i)
    else:
        S_dev = S_host.cuda()
        B_glob = B * world if mode == 'dp' else 0

        def step1(i):
            lo = (i % n_batches) * B
            eng.step(S_dev[0, lo:lo + B], S_dev[1, lo:lo + B], S_dev[2, lo:lo + B], B_global=B_glob)
        for i in range(W):
            step1(i)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_dev, _ = timed_loop(step1)
    clk = clocks.stop() if rank == 0 else None
    samples_per_step = B * (world if mode in ('dp', 'dp_idx') else 1)
    value = samples_per_step * K / (ms_dev * 1e-3)

    # ---- leg 2: end to end through utils.BPRLoss.stageOne with pinned host batches ---------------------
    B_glob = B * world if mode == 'dp' else 0

    def step2(i):
        lo = (i % n_batches) * B
        if mode is None:
            return bpr.stageOne(S_host[0, lo:lo + B], S_host[1, lo:lo + B], S_host[2, lo:lo + B])
        eng.step(S_host[0, lo:lo + B], S_host[1, lo:lo + B], S_host[2, lo:lo + B], B_global=B_glob)
        return float(eng.loss_to_host()[2])
    for i in range(W):
        step2(i)
    ms_e2e, wall_e2e = timed_loop(step2)
    e2e_value = samples_per_step * K / (ms_e2e * 1e-3)
    last_loss = step2(0)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "strong" if mode == 'rowpart' else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shape synthetic graph ({ds.n_users} users x {ds.m_items} items, "
                                   f"{ds.trainDataSize} train edges, nnz {csr.nnz}), LightGCN L=3 d=64, BPR batch {B} per step"
                                   + (f" per rank ({mode})" if mode else ""),
                       "l2": "flushed between timed steps (512 MiB read outside the event brackets)",
                       "parallelism": mode or "single", "cuda_graph": bool(eng.use_graph)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 + 3 * B * 8, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / K, "wall_ms_per_step": 1e3 * wall_e2e / K},
            "gpu_launches": K * (2 * cfg['lightGCN_n_layers'] + 4 + (0 if mode else 0)),
            "clocks": clk, "loss": last_loss, "csr_build_ms": 1e3 * t_build}

    if rank == 0 and world == 1:
        # ---- roofline of the dominant kernel: instrumented eager pass over the same steps ----------------
        peak, peak_src = peaks()
        durs = []
        orig_spmm, orig_adam = lg.ops.spmm, lg.ops.spmm_adam

        def wrap(fn):
            def inner(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(*a, **k); e1.record(); durs.append((e0, e1)); return r
            return inner
        lg.ops.spmm, lg.ops.spmm_adam = wrap(orig_spmm), wrap(orig_adam)
        lg.engine.ops.spmm, lg.engine.ops.spmm_adam = lg.ops.spmm, lg.ops.spmm_adam
        use_graph = eng.use_graph
        eng.use_graph = False
        for i in range(min(K, 20)):
            flush_l2()
            lo = (i % n_batches) * B
            eng.step(S_host[0, lo:lo + B], S_host[1, lo:lo + B], S_host[2, lo:lo + B])
        torch.cuda.synchronize()
        eng.use_graph = use_graph
        lg.ops.spmm, lg.ops.spmm_adam = orig_spmm, orig_adam
        lg.engine.ops.spmm, lg.engine.ops.spmm_adam = orig_spmm, orig_adam
        t_ms = [a.elapsed_time(b) for a, b in durs]
        mean_ms = sum(t_ms) / len(t_ms)
        alg_bytes = csr.algorithmic_bytes(64)
        gather_bytes = 8 * csr.nnz + 4 * (csr.n_rows + 1) + 4 * csr.nnz * 64 + 4 * csr.n_rows * 64
        achieved = alg_bytes / (mean_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get("spmm_dram_bytes_per_launch")
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "kernel": "spmm_kernel<64>", "algorithmic_bytes_per_launch": alg_bytes,
                            "mean_launch_us": 1e3 * mean_ms, "launches_timed": len(t_ms), "peak_source": peak_src,
                            "per_launch_us_by_position": [round(1e3 * sum(t_ms[i::6]) / len(t_ms[i::6]), 1) for i in range(6)] if len(t_ms) % 6 == 0 else None,
                            "spmm_share_of_step": (6 * mean_ms) / (ms_dev / K),
                            "l2_gather_gbs": gather_bytes / (mean_ms * 1e-3) / 1e9}
        # ---- evaluation (K3) timing, reported beside the headline -----------------------------------------
        # (the first Procedure.Test also builds the test CSR and the position-space mask of the tensor-core kernel — once per
        # graph; the steady-state call is the one a training run repeats every few epochs)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = lg.Procedure.Test(ds, model, 0)
        torch.cuda.synchronize(); first_ms = 1e3 * (time.perf_counter() - t0)
        warm = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = lg.Procedure.Test(ds, model, 0)
            torch.cuda.synchronize(); warm.append(1e3 * (time.perf_counter() - t0))
        n_test_users = int(ds.test_csr()[0].numel())
        line["eval"] = {"test_ms": statistics.median(warm), "test_first_call_ms": first_ms, "users": n_test_users,
                        "recall@20": float(res['recall'][0]), "score_gflop": 2.0 * n_test_users * ds.m_items * 64 / 1e9,
                        "rows_redone_by_exact_kernel": int(getattr(model, 'last_rank_redone', 0))}
        # ---- CPU baseline on this box's host cores ---------------------------------------------------------
        if not args.no_cpu_baseline:
            r = cpu_port_bench(graph, 12, 2, budget_s=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="yelp2018", choices=["gowalla", "yelp2018", "amazon-book", "tiny"])
    ap.add_argument("--parallel", default="dp_idx", choices=["dp_idx", "dp", "rowpart"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
