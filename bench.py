#!/usr/bin/env python
"""bench.py — BPR train samples/s (+ SpMM propagation GB/s roofline) for LightGCN L=3 d=64, gowalla-shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload gowalla]

A "step" is one BPR training step on B = 2048 sampled triples: 3 forward SpMM layers, fused BPR loss + gradient,
3 backward SpMM layers, Adam.  The headline workload is BASELINE.json's metric configuration: the gowalla-SHAPE synthetic
graph (29,858 x 40,981, ~810 k train edges — the real file is reference data and does not travel; tests/ cover it),
triples from the reference's sampler stream (C sampler seeded 2020, numpy shuffle seeded 2020).

N = 1            one GPU, the whole graph.
N > 1 (torchrun) the SAME step on the SAME workload split over the N GPUs — strong scaling: `value` stays B / step time, what
                 N GPUs buy is a shorter step, not more samples per step.  Two decompositions:
                 * `large_graph` (BASELINE config 5, and anything beyond one GPU): the adjacency ROW-PARTITIONED (north_star) —
                   each rank holds only its CSR block and its rows of the Adam moments, K1's epilogue stores every finished row
                   into all replicas over NVLink (multimem.st), layers are ordered by a device-side flag barrier, one CUDA graph
                   per rank, no collective launch on the path; the one-GPU step is timed on rank 0 first.
                 * the headline and amazon-book-shape (graphs that fit one GPU; `--parallel auto`): the FEATURE partition — every
                   rank holds d/N embedding columns of every table and the whole CSR; a CSR SpMM is independent per column, so
                   the propagation exchanges NOTHING and the only cross-rank traffic of a step is K2's five dot products per
                   triple (40 KB, peer stores) around ONE device barrier.  The row partition of the same step is reported
                   beside it (`extra.headline_under_row_partition`): there every rank must ingest (N-1)/N of the table per
                   layer over NVLink, which makes a layer slower than the whole graph on one GPU.
                 `--parallel rowpart|featpart` force one; `--parallel dp_idx` (replicas, NOT a scaling measurement) is a side mode.

value    : samples/s, epoch's triples resident in HBM; R repeats of a K-step loop, every step bracketed by CUDA events
           on the launching stream, L2 flushed before every step, max over ranks per repeat, MEDIAN over repeats.
e2e      : the same through the public API utils.BPRLoss.stageOne(users, pos, neg) with pinned HOST tensors (one H2D
           + loss D2H per step inside the timed region).  `value` = from CUDA events (excludes the L2-flush scaffolding
           between steps); `wall_value` = same loop by host clock (includes the flush); `wall_value_no_flush` = a
           back-to-back loop by host clock with no flush — what a training loop sees.
roofline : dominant kernel spmm_kernel<64,...> (K1).  achieved = B_spmm / launch duration, B_spmm = 8 nnz + 4 (N+1) + 8 N d
           (one unit = one layer over the whole graph), launch duration = a CUDA graph of 2L full-graph launches
           (the step's count, ping-pong buffers) replayed with the L2 flushed before each replay, / 2L.
           spmm_share_of_step = (graph of the step's OWN 2L K1 launches) / step — both inside captured graphs.
roofline_l2 : the L2->SM gather ceiling measured on this box with a plain 256-byte random-row gather kernel
           (lgcn_debug_gather_rows) on an L2-resident table, and K1's gather-effective bytes against it.
extra    : yelp2018-shape and amazon-book-shape (north_star's >= 70 % target config): step, cold layer, frac, eval;
           L x d sweep on amazon-book-shape; epoch_ms / eval_ms through the reference-facing procedures.
cpu_baseline / --impl reference : oracle/ref_port (torch-CPU call-for-call port of the reference, asserted bit-identical
           to it by oracle/gen_golden.py) on this box's host cores, same graph, same sampled triples.
gpu_library_baseline : the same port with its tensors on cuda:0 (cuSPARSE SpMM / ATen kernels / torch Adam) — the
           library bar BASELINE.md §4.6 names.
"""
import argparse
import functools
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "BPR train samples/s, LightGCN L=3 d=64 gowalla-shape"
UNIT = "samples/s"
B = 2048
L_LAYERS, D = 3, 64


_JSON_OUT = None


def claim_stdout():
    """Keep file descriptor 1 for the ONE JSON line: everything else that writes to stdout (the procedures' own prints of
    the metric dicts — the reference prints them too, code/Procedure.py:205 —, NCCL's version banner) goes to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", 0.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1350.0, "fallback (B200_PROFILING.md)"


def workload_graph(name):
    import lgcn_b200 as lg
    return lg.synth.make_graph(name, seed=2020)


def bench_config(name, graph):
    """`config` of the JSON line — identical in both arms (the driver compares them)."""
    import numpy as np
    e = int(graph['train_user'].size)
    nnz = 2 * int(np.unique(graph['train_user'] * graph['m_items'] + graph['train_item']).size)
    return {"workload": f"{name}-shape synthetic graph ({graph['n_users']} users x {graph['m_items']} items, {e} train edges, "
                        f"adjacency nnz {nnz}), LightGCN L={L_LAYERS} d={D}, BPR batch {B} per step",
            "triples": "reference sampler stream: C sampler (sources/sampling.cpp semantics) seeded 2020, one epoch, numpy shuffle seeded 2020",
            "global_batch": B}


def epoch_triples(lg, ds, seed=2020):
    """One epoch of (user, pos, neg) exactly as Procedure.BPR_train_original draws them — used by BOTH arms."""
    import numpy as np
    import torch
    lg.utils.sampler_seed(seed)
    S = lg.utils.UniformSample_original(ds)
    np.random.seed(seed)
    perm = np.arange(S.shape[0]); np.random.shuffle(perm)
    return torch.from_numpy(np.ascontiguousarray(S[perm, :3].T)).to(torch.int64)


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

    def sample_now(self):
        """One NVML reading from the calling thread (Timing.loop calls it after a loop's steps are enqueued and before it
        waits for them, i.e. with the GPU under load — the polling thread alone was seen to get a single reading in when
        the loop enqueues 100 steps without a host-side wait in between)."""
        if self.nvml is None or self.thread is None:
            return
        nv, h = self.nvml
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self.sample_now()
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
def port_bench(graph, S, steps, warmup, device="cpu", budget_s=None, seed=2020):
    """Times oracle/ref_port stageOne on the same graph and the same sampled triples `S` (int64 [3, n]).
    device='cpu': the reference's own path on all host threads (cpu_baseline / --impl reference);
    device='cuda': the same library calls on the GPU (cuSPARSE / ATen) — gpu_library_baseline.
    The only place outside tests/smoke where oracle/ is executed: it is the baseline being measured, never the product."""
    import torch
    from oracle import ref_port
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    nu, ni = graph['n_users'], graph['m_items']
    t0 = time.perf_counter()
    g, _, _ = ref_port.build_graph(graph['train_user'], graph['train_item'], nu, ni)
    t_graph = time.perf_counter() - t0
    torch.manual_seed(seed)
    model = ref_port.RefLightGCN(nu, ni, D, L_LAYERS, g)
    if device != "cpu":
        model = model.to(device)
        model.Graph = g.to(device)
    bpr = ref_port.RefBPRLoss(model, 1e-4, 1e-3)
    Sd = S.to(device)
    n_batches = Sd.shape[1] // B
    times = []
    for s in range(warmup + steps):
        lo = (s % n_batches) * B
        u, p, n = Sd[0, lo:lo + B], Sd[1, lo:lo + B], Sd[2, lo:lo + B]
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        bpr.stageOne(u, p, n)            # ends with loss.cpu().item(): synchronises
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if budget_s is not None and s >= warmup and sum(times) > budget_s:
            break
    total = sum(times)
    where = f"torch {torch.__version__} CPU, {cores} threads" if device == "cpu" else f"torch {torch.__version__} on {device} (cuSPARSE/ATen), host-clock per step incl. the loss read-back"
    return {"value": len(times) * B / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} stageOne steps of B={B} after {warmup} warm-up ({where}); same graph and same sampled triples as the GPU arm; graph build {t_graph:.2f} s (bmat path)",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def reference_triples(graph, seed=2020):
    """The same epoch of triples as epoch_triples(), drawn WITHOUT the product: oracle/c/sampler_ref.c (the reference's
    sampling.cpp restated; the product's sampler is bit-identical to it, tests/test_host.py) + the numpy shuffle."""
    import numpy as np
    import torch
    from oracle import lightgcn_oracle as orc
    nu, ni = graph['n_users'], graph['m_items']
    key = np.unique(graph['train_user'] * ni + graph['train_item'])            # allPos: sorted unique items per user
    users, items = key // ni, (key % ni).astype(np.int32)
    indptr = np.concatenate([[0], np.cumsum(np.bincount(users, minlength=nu))]).astype(np.int64)
    S = orc.sample_negative_ref(seed, nu, ni, int(graph['train_user'].size), indptr, items, 1)
    np.random.seed(seed)
    perm = np.arange(S.shape[0]); np.random.shuffle(perm)
    return torch.from_numpy(np.ascontiguousarray(S[perm, :3].T)).to(torch.int64)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from importlib import util as _u
    spec = _u.spec_from_file_location("_synth_only", os.path.join(ROOT, "graph-and-sequential-recommendation-systems_b200", "synth.py"))
    synth = _u.module_from_spec(spec); spec.loader.exec_module(synth)         # the graph generator alone: numpy, no .so
    graph = synth.make_graph(args.workload, seed=2020)
    S = reference_triples(graph)
    r = port_bench(graph, S, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(args.workload, graph),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ measurement helpers (our arm)
class Timing:
    def __init__(self, dist, flush_buf):
        self.dist, self.flush_buf = dist, flush_buf
        self.clocks = None              # a started ClockSampler: read once per loop while the loop's steps are executing

    def flush_l2(self):
        # read 512 MiB (4x the 126 MB L2): everything the step touched is evicted and the lines left behind are
        # clean, so the timed kernels neither hit stale data nor pay for write-backs of the flush itself
        self.flush_buf.sum()

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def loop(self, step_fn, K, flush=True):
        """K steps, each bracketed by CUDA events on the launching stream -> (sum of event ms [max over ranks], wall s)."""
        import torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        self.barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            if flush:
                self.flush_l2()
            evs[i][0].record()
            step_fn(i)
            evs[i][1].record()
        if self.clocks is not None:
            self.clocks.sample_now()
        self.barrier()
        wall = time.perf_counter() - wall0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.dist is not None:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        return ms, wall

    def repeats(self, step_fn, K, R, flush=True):
        """R repeats of the K-step loop -> (median ms per loop, median wall s per loop, all ms)."""
        runs = [self.loop(step_fn, K, flush) for _ in range(R)]
        return statistics.median(r[0] for r in runs), statistics.median(r[1] for r in runs), [r[0] for r in runs]

    def median_us(self, fn, reps=11, flush=True):
        """Median duration of fn() in microseconds (events on the current stream, L2 flushed before each call)."""
        import torch
        out = []
        for _ in range(reps):
            if flush:
                self.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            out.append(1e3 * a.elapsed_time(b))
        return statistics.median(out)


def k1_roofline(lg, eng, csr, tm, step_ms, peak, peak_src, workload):
    """roofline block of the dominant kernel on this engine's graph (single GPU)."""
    import torch
    L, d = eng.L, eng.d
    alg = csr.algorithmic_bytes(d)
    gather = 8 * csr.nnz + 4 * (csr.n_rows + 1) + 4 * csr.nnz * d + 4 * csr.n_rows * d
    # (1) a CUDA graph of 2L full-graph launches, ping-pong between two scratch tables (the step's launch count)
    a, b = torch.randn_like(eng.E0), torch.empty_like(eng.E0)
    lg.ops.spmm(csr, a, b); lg.ops.spmm(csr, b, a); torch.cuda.synchronize()
    g6 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g6):
        for _ in range(L):
            lg.ops.spmm(csr, a, b); lg.ops.spmm(csr, b, a)
    t6 = tm.median_us(g6.replay)
    launch_us = t6 / (2 * L)
    cold_us = tm.median_us(lambda: lg.ops.spmm(csr, a, b))
    del g6, a, b
    # (2) the step's own 2L K1 launches (masked last forward layer, Adam epilogue on the last backward one), in a graph
    saved = [t.clone() for t in (eng.E0, eng.M, eng.V, eng.scalars)]
    gs = eng.spmm_only_graph()
    ts = tm.median_us(gs.replay)
    del gs
    eng.clear_batch_mask()
    for dst, src in zip((eng.E0, eng.M, eng.V, eng.scalars), saved):
        dst.copy_(src)
    torch.cuda.synchronize()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(workload, {}).get("spmm_dram_bytes_per_launch")
    except Exception:
        pass
    achieved = alg / (launch_us * 1e-6) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": "spmm_kernel<64,8,4,...> (K1), one launch = one layer over the whole graph",
            "algorithmic_bytes_per_launch": alg, "launch_us": launch_us,
            "how": f"CUDA graph of {2 * L} full-graph launches replayed 11x with the L2 flushed before each replay, median / {2 * L}",
            "cold_launch_us": cold_us, "cold_frac": alg / (cold_us * 1e-6) / 1e9 / peak,
            "peak_source": peak_src,
            "step_k1_graph_us": ts, "spmm_share_of_step": ts / (1e3 * step_ms),
            "gather_bytes_per_launch": gather, "l2_gather_gbs": gather / (launch_us * 1e-6) / 1e9}


def l2_gather_ceiling(lg, tm, n_rows, nnz):
    """roofline_l2: plain random-row gather (256-byte rows, table of this workload's size, L2-resident) on this box."""
    import torch
    X = torch.randn(n_rows, D, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, n_rows, (max(nnz, 1 << 24),), device="cuda", generator=gen, dtype=torch.int32)
    best, best_v = None, None
    out = None
    for variant in range(4):
        out = lg.ops.gather_probe(X, idx, run=32, variant=variant, out=out)
        us = tm.median_us(lambda: lg.ops.gather_probe(X, idx, run=32, variant=variant, out=out), reps=9, flush=False)
        gbs = idx.numel() * D * 4 / (us * 1e-6) / 1e9
        if best is None or gbs > best:
            best, best_v = gbs, variant
    return {"gather_ceiling_gbs": best, "variant": best_v, "table_mb": n_rows * D * 4 / 1e6, "gathers": int(idx.numel()),
            "kernel": "gather_probe_kernel (lgcn_debug_gather_rows): uniform random 256-B rows, runs of 32, no values/epilogue",
            "script": "scripts/l2_gather_ceiling.py"}


def setup_workload(lg, name, cfg, graph=None):
    import torch
    graph = graph if graph is not None else workload_graph(name)
    ds = lg.InteractionDataset(graph['n_users'], graph['m_items'], graph['train_user'], graph['train_item'],
                               graph['test_user'], graph['test_item'], config=cfg, name=name)
    lg.utils.set_seed(2020)
    model = lg.LightGCN(cfg, ds)
    bpr = lg.utils.BPRLoss(model, cfg)
    S_host = epoch_triples(lg, ds).pin_memory()
    return graph, ds, model, bpr, S_host


def resident_step_fn(eng, S_host):
    """Step function over an HBM-resident epoch (full batches only, wraps around)."""
    n_epoch_steps = eng.begin_epoch(S_host.cuda())
    pos = [0]

    def step(i):
        if pos[0] >= n_epoch_steps - 1:
            eng.rewind_epoch(); pos[0] = 0
        eng.epoch_step(); pos[0] += 1
    return step


def rank_record(lg, tm, ds, model):
    """The ranking alone (K3: score GEMM on tcgen05 + mask + top-k), CUDA events, warm tables: the tensor roofline of §8d."""
    _, tc_peak, _ = peaks()
    n_test = int(ds.test_csr()[0].numel())
    lg.Procedure.rank_all(ds, model, 20, user_tile=65536)
    rank_us = tm.median_us(lambda: lg.Procedure.rank_all(ds, model, 20, user_tile=65536), reps=7, flush=False)
    score_flop = 2.0 * n_test * ds.m_items * D
    return {"rank_all_ms": rank_us / 1e3, "users": n_test, "items": ds.m_items, "rank_useful_tflops": score_flop / (rank_us * 1e-6) / 1e12,
            "rank_frac_of_bf16_dense_peak": score_flop / (rank_us * 1e-6) / 1e12 / tc_peak if tc_peak else None,
            "rows_redone_by_exact_kernel": int(getattr(model, 'last_rank_redone', 0)),
            "rank_note": "useful flops 2*U*M*d counted once; the kernel runs the TF32 GEMM 1.5x (sampled pass + full pass) and TF32 peaks at half the bf16 rate; "
                         "ncu tensor-pipe activity of the two passes: profiles/r2_evaltc_launches_ncu.csv"}


def shape_record(lg, tm, name, cfg, peak, K, R):
    """extra.<shape>: step, cold K1 layer + roofline fraction, eval — one GPU."""
    import torch
    graph, ds, model, bpr, S_host = setup_workload(lg, name, cfg)
    eng, csr = model._engine, ds.getCSRGraph()
    step = resident_step_fn(eng, S_host)
    for i in range(5):
        step(i)
    ms, _, _ = tm.repeats(step, K, R)
    step_ms = ms / K
    roof = k1_roofline(lg, eng, csr, tm, step_ms, peak, "", name)
    prop_us = tm.median_us(lambda: eng.forward())
    lg.Procedure.Test(ds, model, 0)
    ev = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = lg.Procedure.Test(ds, model, 0)
        torch.cuda.synchronize(); ev.append(1e3 * (time.perf_counter() - t0))
    n_test = int(ds.test_csr()[0].numel())
    rec = {"workload": bench_config(name, graph)["workload"], "step_ms": step_ms, "samples_per_s": B / (step_ms * 1e-3),
           **{k: v for k, v in rank_record(lg, tm, ds, model).items() if k in ("rank_all_ms", "rank_useful_tflops", "rank_frac_of_bf16_dense_peak")},
           "spmm_launch_us": roof["launch_us"], "spmm_frac": roof["frac"], "spmm_cold_launch_us": roof["cold_launch_us"],
           "spmm_cold_frac": roof["cold_frac"], "spmm_alg_gbs": roof["achieved"], "spmm_l2_gather_gbs": roof["l2_gather_gbs"],
           "spmm_share_of_step": roof["spmm_share_of_step"], "propagation_L3_cold_us": prop_us,
           "hbm_roofline_step_ms": (2 * L_LAYERS * roof["algorithmic_bytes_per_launch"] + 28 * csr.n_rows * D + 24 * B * D) / (peak * 1e9) * 1e3,
           "eval_ms": statistics.median(ev), "eval_users": n_test, "score_gflop": 2.0 * n_test * ds.m_items * D / 1e9,
           "eval_rows_redone_by_exact_kernel": int(getattr(model, 'last_rank_redone', 0)), "recall@20": float(res['recall'][0])}
    return rec, (graph, ds, model)


def shape_record_partitioned(lg, tm, name, cfg, K, R, dist, world):
    """extra.<shape> at N > 1 (BASELINE config 3 names 1/2/4/8 GPUs for amazon-book-shape): the same step, epoch and full-ranking
    evaluation under the row partition.  Collective: every rank calls it; times are max over ranks."""
    import torch
    graph, ds, model, bpr, S_host = setup_workload(lg, name, cfg)
    eng = model._engine
    step = resident_step_fn(eng, S_host)
    for i in range(5):
        step(i)
    ms, _, _ = tm.repeats(step, K, R)
    step_ms = ms / K
    if eng._barrier is not None:
        eng._barrier.check()
    proc = procedures_record(lg, ds, model, bpr, dist)
    t = torch.tensor([proc["epoch_ms"], proc["eval_ms"], torch.cuda.max_memory_allocated() / 1e9], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    csr_nnz = 2 * int(__import__('numpy').unique(graph['train_user'] * graph['m_items'] + graph['train_item']).size)
    return {"workload": bench_config(name, graph)["workload"], "n_gpus": world, "parallelism": eng.dist_mode,
            "step_ms": step_ms, "samples_per_s": B / (step_ms * 1e-3), "epoch_ms": float(t[0].item()), "epoch_steps": proc["epoch_steps"],
            "eval_ms": float(t[1].item()), "rows_per_rank": eng.r1 - eng.r0, "adjacency_nnz": csr_nnz, "mem_gb_max_over_ranks": float(t[2].item()),
            "how": "step: median over repeats of the K-step loop, CUDA events, L2 flushed, max over ranks; epoch/eval: wall clock around "
                   "Procedure.BPR_train_original / Procedure.Test, best of 2, max over ranks",
            "note": ("feature partition: d/N columns per rank, no exchange in the propagation; " if eng.dist_mode == 'featpart' else
                     "row partition of an L2-resident graph: the exchanged layer is latency/NVLink-ingest bound (DESIGN.md §5); ")
                    + "the one-GPU numbers are in the N=1 line's extra.amazon_book"}


def sweep_record(lg, tm, graph, peak):
    """BASELINE config 4: L = 1..4 x d = 64/128/256 propagation on amazon-book-shape (cold L2), K1 roofline fraction."""
    import torch
    out = {}
    for d in (64, 128, 256):
        cfg = dict(lg.world.config); cfg.update(latent_dim_rec=d, cuda_graph=False)
        per_L = {}
        for L in (1, 2, 3, 4):
            cfg.update(lightGCN_n_layers=L)
            ds = lg.InteractionDataset(graph['n_users'], graph['m_items'], graph['train_user'], graph['train_item'],
                                       graph['test_user'], graph['test_item'], config=cfg)
            lg.utils.set_seed(2020)
            m = lg.LightGCN(cfg, ds)
            eng = m._engine
            eng.forward(); torch.cuda.synchronize()
            us = tm.median_us(lambda: eng.forward(), reps=7)
            alg = L * ds.getCSRGraph().algorithmic_bytes(d)
            per_L[f"L{L}"] = {"propagation_us": us, "alg_gbs": alg / (us * 1e-6) / 1e9, "frac": alg / (us * 1e-6) / 1e9 / peak}
            del m, eng, ds
        out[f"d{d}"] = per_L
    return out


def procedures_record(lg, ds, model, bpr, dist):
    """epoch_ms / eval_ms through the reference-facing procedures (BPR_train_original, Test) — at any N."""
    import torch
    lg.utils.set_seed(2020); lg.utils.sampler_seed(2020)
    lg.Procedure.BPR_train_original(ds, model, bpr, 0)                    # warm (graphs captured, mask caches built)
    lg.Procedure.Test(ds, model, 0)
    ep, evs = [], []
    for e in range(2):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = lg.Procedure.BPR_train_original(ds, model, bpr, e + 1)
        torch.cuda.synchronize(); ep.append(1e3 * (time.perf_counter() - t0))
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lg.Procedure.Test(ds, model, e + 1)
        torch.cuda.synchronize(); evs.append(1e3 * (time.perf_counter() - t0))
    sample_s = float(info.split('Sample:')[1].split('|')[0]) if 'Sample:' in info else None
    n_triples = ds.trainDataSize // ds.n_users * ds.n_users
    return {"epoch_ms": min(ep), "epoch_steps": (n_triples + B - 1) // B, "eval_ms": min(evs),
            "host_sampler_s_of_last_epoch": sample_s,
            "how": "wall clock around Procedure.BPR_train_original (host sampler + shuffle + H2D of the epoch + all steps + loss read) and Procedure.Test (propagate + rank all test users + metrics); best of 2"}


def large_graph_record(lg, dist, rank, world, args, peak):
    """BASELINE config 5: power-law 10 M x 2 M, 500 M edges.  Rank 0 first times the ONE-GPU step (whole graph on one GPU),
    then all N ranks run the memory-partitioned row partition on the same graph and the same triples."""
    import torch
    sc = args.large_scale
    nu, ni, ne = int(10_000_000 * sc), int(2_000_000 * sc), int(500_000_000 * sc)
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = dict(lg.world.config)
    cfg.update(bpr_batch_size=B, lightGCN_n_layers=L_LAYERS, latent_dim_rec=D)
    rec = {"config": f"power-law synthetic graph {nu} users x {ni} items, {ne} edges (BASELINE config 5 x {sc:g}), LightGCN L={L_LAYERS} d={D}, batch {B}",
           "n_gpus": world}

    def triples():
        tu, ti = next(iter(lg.synth.powerlaw_chunks(nu, ni, min(ne, 1 << 22), seed=2020, device=dev, chunk=1 << 22)))
        gen = torch.Generator(device=dev).manual_seed(7)
        n = min(tu.numel(), 64 * B)
        return torch.stack([tu[:n], ti[:n], torch.randint(0, ni, (n,), device=dev, generator=gen)]).contiguous()

    def time_steps(eng, S, n_timed):
        nb = S.shape[1] // B

        def step(i):
            lo = (i % nb) * B
            eng.step(S[0, lo:lo + B], S[1, lo:lo + B], S[2, lo:lo + B])
        for i in range(3):
            step(i)
        if dist is not None and eng.dist_mode is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ts = []
        for i in range(n_timed):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(3 + i); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        t = torch.tensor([statistics.median(ts)], device="cuda", dtype=torch.float64)
        if dist is not None and eng.dist_mode is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def local_k1_ms(eng, graph_):
        Y = torch.empty((graph_.n_rows, D), dtype=torch.float32, device=dev)
        lg.ops.spmm(graph_, eng.E0, Y); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lg.ops.spmm(graph_, eng.E0, Y); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    # ---- one GPU (rank 0), whole graph
    one = None
    if rank == 0 and not args.large_skip_1gpu:
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        tu, ti = lg.synth.make_powerlaw_device(nu, ni, ne, seed=2020, device=dev)
        ds1 = lg.synth.DeviceGraphDataset(nu, ni, tu, ti)
        del tu, ti
        g1 = ds1.getCSRGraph()
        torch.cuda.synchronize(); t_build = time.perf_counter() - t0
        torch.cuda.empty_cache()
        c1 = dict(cfg); c1.update(dist_mode=None)
        lg.utils.set_seed(2020)
        m1 = lg.LightGCN(c1, ds1)
        S = triples()
        ms1 = time_steps(m1._engine, S, 3)
        k1 = local_k1_ms(m1._engine, g1)
        alg = g1.algorithmic_bytes(D)
        one = {"ms_per_step": ms1, "samples_per_s": B / (ms1 * 1e-3), "k1_layer_ms": k1, "k1_alg_gbs": alg / (k1 * 1e-3) / 1e9,
               "k1_frac": alg / (k1 * 1e-3) / 1e9 / peak, "nnz": g1.nnz, "generate_and_build_s": t_build,
               "mem_gb": torch.cuda.max_memory_allocated() / 1e9, "loss": float(m1._engine.loss_to_host()[2])}
        del m1, ds1, g1, S
        torch.cuda.empty_cache()
    rec["one_gpu"] = one
    rec["ms_per_step_1gpu"] = one["ms_per_step"] if one else None
    if world == 1:
        if one:
            rec.update(ms_per_step=one["ms_per_step"], speedup=1.0, mem_gb_per_rank=one["mem_gb"])
        return rec
    # ---- N GPUs: memory-partitioned row partition of the same graph
    dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    dsn = lg.synth.DeviceGraphDataset(nu, ni, chunks=functools.partial(lg.synth.powerlaw_chunks, nu, ni, ne, 2020, dev), n_edges=ne)
    cn = dict(cfg); cn.update(dist_mode='rowpart')
    lg.utils.set_seed(2020)
    mn = lg.LightGCN(cn, dsn)
    torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    eng = mn._engine
    S = triples()
    msn = time_steps(eng, S, 7)
    k1 = local_k1_ms(eng, eng.local)
    # one exchanged layer (K1 + row stores into every replica + device barrier), and the barrier alone
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize(); dist.barrier()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        return statistics.median(ts)
    layer = timed(lambda: eng._layer(eng.E0, eng.X[0], 1.0, 0.0, None))
    bar = timed(eng._rank_barrier, reps=20) if eng._barrier is not None else 0.0
    mine = torch.tensor([k1, layer, bar, torch.cuda.max_memory_allocated() / 1e9, eng.local.nnz, eng.r1 - eng.r0], device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    allr = [t.tolist() for t in allr]
    if eng._barrier is not None:
        eng._barrier.check()
    k1_max, layer_max = max(r[0] for r in allr), max(r[1] for r in allr)
    alg_local = [8 * r[4] + 4 * (r[5] + 1) + 4 * eng.N * D + 4 * r[5] * D for r in allr]
    rec.update(ms_per_step=msn, samples_per_s=B / (msn * 1e-3), speedup=(one["ms_per_step"] / msn) if one else None,
               strong_scaling_efficiency=(one["ms_per_step"] / msn / world) if one else None,
               local_k1_ms_per_rank=[round(r[0], 3) for r in allr], exchanged_layer_ms_per_rank=[round(r[1], 3) for r in allr],
               barrier_us=round(1e3 * max(r[2] for r in allr), 1), mem_gb_per_rank=[round(r[3], 2) for r in allr],
               nnz_per_rank=[int(r[4]) for r in allr], rows_per_rank=[int(r[5]) for r in allr],
               local_k1_alg_gbs_per_rank=[round(a / (r[0] * 1e-3) / 1e9, 1) for a, r in zip(alg_local, allr)],
               setup_s=t_setup, cuda_graph=bool(eng.use_graph), multicast=bool(eng._mc), device_barrier=eng._barrier is not None,
               loss=float(eng.loss_to_host()[2]),
               limiter=("local K1: X (N x d x 4 = %.1f GB) is far larger than L2, the gathers run at HBM speed" % (eng.N * D * 4 / 1e9)
                        if layer_max < 1.2 * k1_max else
                        "exchange tail: the exchanged layer (%.2f ms) exceeds the slowest local K1 (%.2f ms) by the row stores into %d replicas + barrier" % (layer_max, k1_max, world)))
    del mn, dsn, eng
    torch.cuda.empty_cache()
    return rec


def guarded(name, fn, line, dist=None):
    """Optional sub-records must never cost the headline line: a failure is recorded in place of the record.
    (Collective sections are only guarded when they fail on every rank alike, e.g. on an argument error; a rank-local
    failure inside a collective would hang the others either way, so those sections keep their own timeouts.)"""
    try:
        line[name] = fn()
    except Exception as e:                                  # noqa: BLE001
        import traceback
        line[name] = {"error": f"{type(e).__name__}: {e}", "where": traceback.format_exc().strip().splitlines()[-3:]}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
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
    lg.world.configure(device=f"cuda:{local_rank}", checkpoint_dir=f"/tmp/lgcn_b200_bench/r{rank}")
    cfg = dict(lg.world.config)
    cfg.update(lightGCN_n_layers=L_LAYERS, latent_dim_rec=D, bpr_batch_size=B)
    mode = args.parallel if world > 1 else None
    if mode == 'auto':
        # graphs that fit one GPU (every --workload here): the feature partition — embedding columns over the ranks, no exchange in
        # the propagation; the row partition is what the large_graph record (and anything beyond one GPU's memory) runs
        mode = 'featpart' if (D % world == 0 and D // world in (8, 16, 32)) else 'rowpart'
    cfg.update(dist_mode=mode)
    peak, tc_peak, peak_src = peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda").view(torch.int64)
    tm = Timing(dist, flush)
    K, W = args.steps, args.warmup

    graph = workload_graph(args.workload)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    graph, ds, model, bpr, S_host = setup_workload(lg, args.workload, cfg, graph)
    torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    eng = model._engine
    n_batches = S_host.shape[1] // B
    replicated = mode in ('dp', 'dp_idx')
    if replicated:      # replicas: every rank feeds its own shard of a GLOBAL batch of N x 2048 (not a scaling measurement)
        S_host = S_host[:, rank * (S_host.shape[1] // world):(rank + 1) * (S_host.shape[1] // world)].contiguous().pin_memory()
        n_batches = S_host.shape[1] // B

    # ---- leg 1: inputs resident in HBM ----------------------------------------------------------
    if not replicated:
        step1 = resident_step_fn(eng, S_host)
    else:
        S_dev = S_host.cuda()
        B_glob = B * world if mode == 'dp' else 0

        def step1(i):
            lo = (i % n_batches) * B
            eng.step(S_dev[0, lo:lo + B], S_dev[1, lo:lo + B], S_dev[2, lo:lo + B], B_global=B_glob)
    for i in range(W):
        step1(i)
    probe_ms, _ = tm.loop(step1, K)
    R = max(5, min(40, int(60.0 / max(probe_ms, 1e-3)) + 1))           # >= 5 repeats, ~60 ms of timed steps in total
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        tm.clocks = clocks
    ms_dev, _, all_ms = tm.repeats(step1, K, R)
    tm.clocks = None
    clk = clocks.stop() if rank == 0 else None
    samples_per_step = B * (world if replicated else 1)
    value = samples_per_step * K / (ms_dev * 1e-3)

    # ---- leg 2: end to end through utils.BPRLoss.stageOne with pinned host batches ---------------------
    B_glob = B * world if mode == 'dp' else 0

    def step2(i):
        lo = (i % n_batches) * B
        if not replicated:
            return bpr.stageOne(S_host[0, lo:lo + B], S_host[1, lo:lo + B], S_host[2, lo:lo + B])
        eng.step(S_host[0, lo:lo + B], S_host[1, lo:lo + B], S_host[2, lo:lo + B], B_global=B_glob)
        return float(eng.loss_to_host()[2])
    for i in range(W):
        step2(i)
    ms_e2e, wall_e2e, _ = tm.repeats(step2, K, R)
    _, wall_nf, _ = tm.repeats(step2, K, R, flush=False)
    last_loss = step2(0)
    if eng._barrier is not None:
        eng._barrier.check()

    n_k1 = 2 * L_LAYERS
    # adam_tick + [batch_masks: pruning, not under the row partition] + K2 + clear_rows + [batch_advance: resident epoch]
    # + 2L x K1 + [2L x rank_barrier: fused exchange]
    # step_begin (Adam tick + window advance | host-batch pull) + [batch_masks] + K2 + 2L x K1 + [2L x rank_barrier]
    # + [clear_rows: only when the Adam-epilogue K1 cannot zero G itself, i.e. under the row partition]
    # feature partition: K2 is two kernels around ONE rank_barrier, nothing else is exchanged
    if mode == 'featpart':
        launches_per_step = 1 + (1 if eng.prune else 0) + 3 + n_k1
    else:
        launches_per_step = 1 + (1 if eng.prune else 0) + 1 + n_k1 + (n_k1 if eng._barrier is not None else 0) + (1 if mode == 'rowpart' else 0)
    launches_per_step_e2e = launches_per_step
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "weak" if replicated else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.workload, graph),
            "setup": {"parallelism": mode or "single", "cuda_graph": bool(eng.use_graph),
                      "l2": "flushed before every timed step (512 MiB read outside the event brackets)",
                      "repeats": R, "statistic": "median over repeats of the K-step loop (each: sum of per-step CUDA-event times, max over ranks)",
                      "ms_per_step_all_repeats": [round(m / K, 5) for m in all_ms],
                      "rows_per_rank": (eng.r1 - eng.r0), "columns_per_rank": eng.d, "partition_memory": bool(eng._mv_local) or mode == 'featpart',
                      "device_barrier": eng._barrier is not None, "multicast": bool(eng._mc),
                      "model_and_graph_setup_s": t_setup},
            "e2e": {"value": samples_per_step * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 + 3 * B * 8, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / K, "clock": "CUDA events around each stageOne call (H2D + step + loss D2H); the L2 flush between steps is outside the brackets",
                    "transfers": ("no memcpy calls: the head kernel of the captured step pulls the batch out of the pinned host block and K2 writes the loss "
                                  "record into pinned host memory (%d kernels per step)" % launches_per_step_e2e) if eng.zero_copy else "cudaMemcpyAsync",
                    "wall_value": samples_per_step * K / wall_e2e, "wall_ms_per_step": 1e3 * wall_e2e / K,
                    "wall_note": "host clock over the same loop: includes the 512 MiB L2-flush read between steps (bench scaffolding)",
                    "wall_value_no_flush": samples_per_step * K / wall_nf, "wall_ms_per_step_no_flush": 1e3 * wall_nf / K,
                    "wall_no_flush_note": "host clock over a back-to-back stageOne loop, no flush: what a training loop sees"},
            "gpu_launches": K * launches_per_step, "gpu_launches_per_step": launches_per_step,
            "clocks": clk, "loss": last_loss}
    nnz_full = 2 * int(__import__('numpy').unique(graph['train_user'] * graph['m_items'] + graph['train_item']).size)
    n_nodes = graph['n_users'] + graph['m_items']
    step_bytes = 2 * L_LAYERS * (8 * nnz_full + 4 * (n_nodes + 1) + 8 * n_nodes * D) + 28 * n_nodes * D + 24 * B * D
    hbm_ms = step_bytes / (peak * 1e9 * world) * 1e3
    line["step_roofline"] = {"algorithmic_bytes_per_step": step_bytes, "formula": "2L*B_spmm + 28*N*d + 24*B*d (SURVEY.md §8d)",
                             "hbm_ms_at_peak": hbm_ms, "frac": hbm_ms / (ms_dev / K), "peak_gbs": peak * world,
                             "epoch_ms_at_peak": None, "note": "fraction of the memory roofline of the whole step on the %d GPU(s) used" % world}
    if replicated:
        line["setup"]["note"] = "replicated_throughput: N replicas each run the full-graph step on the all-gathered global batch — NOT a scaling measurement"

    # ---- procedures at any N: epoch and evaluation through the reference-facing API -----------------------
    if not replicated:
        guarded("procedures", lambda: procedures_record(lg, ds, model, bpr, dist), line)
        if isinstance(line.get("procedures"), dict) and "epoch_steps" in line["procedures"]:
            line["step_roofline"]["epoch_ms_at_peak"] = hbm_ms * line["procedures"]["epoch_steps"]
            line["procedures"]["epoch_frac_of_hbm_roofline"] = line["step_roofline"]["epoch_ms_at_peak"] / line["procedures"]["epoch_ms"]

    if world == 1:
        csr = ds.getCSRGraph()
        line["roofline"] = k1_roofline(lg, eng, csr, tm, ms_dev / K, peak, peak_src, args.workload)

        def _l2():
            r = l2_gather_ceiling(lg, tm, csr.n_rows, csr.nnz)
            r.update(k1_gather_gbs=line["roofline"]["l2_gather_gbs"], k1_frac_of_gather_ceiling=line["roofline"]["l2_gather_gbs"] / r["gather_ceiling_gbs"])
            return r
        guarded("roofline_l2", _l2, line)
        guarded("eval", lambda: rank_record(lg, tm, ds, model), line)
        # ---- the library bar: the reference's own calls with tensors on the GPU (cuSPARSE / ATen) -------------
        if not args.no_baselines:
            try:
                r = port_bench(graph, S_host, 30, 5, device="cuda:0")
                line["gpu_library_baseline"] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "kind": "port on cuda", "sample": r["sample"]}
            except Exception as e:                                     # noqa: BLE001
                line["gpu_library_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"}
        del model, bpr, eng, ds
        torch.cuda.empty_cache()
        # ---- the other shapes north_star names ------------------------------------------------------------------
        if not args.no_extra:
            c1 = dict(cfg); c1.update(dist_mode=None)
            extra = {}
            ab = {}

            def _yelp():
                return shape_record(lg, tm, "yelp2018", c1, peak, K, 5)[0]

            def _amazon():
                rec, (ab["graph"], _, _) = shape_record(lg, tm, "amazon-book", c1, peak, K, 5)
                rec["target"] = "north_star: 3-layer d=64 propagation at >= 70 % of HBM bandwidth on this shape"
                return rec
            guarded("yelp2018", _yelp, extra)
            guarded("amazon_book", _amazon, extra)
            guarded("sweep_amazon_book", lambda: sweep_record(lg, tm, ab.get("graph") or workload_graph("amazon-book"), peak), extra)
            line["extra"] = extra
            torch.cuda.empty_cache()
    if world > 1 and mode in ('rowpart', 'featpart') and not args.no_extra:
        # BASELINE config 3: amazon-book-shape at N GPUs, through the same partition (collective section, same code path as the headline)
        del model, bpr, eng, ds
        torch.cuda.empty_cache()
        extra = {}
        guarded("amazon_book", lambda: shape_record_partitioned(lg, tm, "amazon-book", cfg, K, 5, dist, world), extra)
        if mode == 'featpart':
            # the same headline step under the ROW partition (north_star's decomposition), for comparison: on graphs whose table
            # is L2-resident every rank must ingest (N-1)/N of the table per layer over NVLink, which bounds it
            def _rowpart():
                c2 = dict(cfg); c2.update(dist_mode='rowpart')
                _, ds2, model2, _, S2 = setup_workload(lg, args.workload, c2, graph)
                e2 = model2._engine
                st2 = resident_step_fn(e2, S2)
                for i in range(5):
                    st2(i)
                ms2, _, _ = tm.repeats(st2, K, 5)
                if e2._barrier is not None:
                    e2._barrier.check()
                return {"parallelism": "rowpart", "ms_per_step": ms2 / K, "samples_per_s": B * K / (ms2 * 1e-3), "rows_per_rank": e2.r1 - e2.r0,
                        "multicast": bool(e2._mc), "device_barrier": e2._barrier is not None}
            guarded("headline_under_row_partition", _rowpart, extra)
        line["extra"] = extra
        torch.cuda.empty_cache()
    if not args.no_large_graph and not replicated:
        guarded("large_graph", lambda: large_graph_record(lg, dist, rank, world, args, peak), line)
    if rank == 0 and world == 1 and not args.no_baselines:
        # ---- CPU baseline on this box's host cores (same graph, same triples) ------------------------------------
        def _cpu():
            r = port_bench(graph, S_host, 12, 2, budget_s=25.0)
            return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        guarded("cpu_baseline", _cpu, line)
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gowalla", choices=["gowalla", "yelp2018", "amazon-book", "tiny"])
    ap.add_argument("--parallel", default="auto", choices=["auto", "featpart", "rowpart", "dp_idx", "dp"],
                    help="N > 1: auto = featpart (embedding columns over the ranks) for the workloads here, which fit one GPU; rowpart = rows of "
                         "the adjacency over the ranks (always used for large_graph); dp/dp_idx = replicas (not a scaling measurement)")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline and gpu_library_baseline")
    ap.add_argument("--no-extra", action="store_true", help="skip the yelp2018 / amazon-book / sweep records")
    ap.add_argument("--no-large-graph", action="store_true", help="skip the BASELINE config-5 record")
    ap.add_argument("--large-scale", type=float, default=1.0, help="scale of the config-5 graph (1.0 = 10 M x 2 M, 500 M edges)")
    ap.add_argument("--large-skip-1gpu", action="store_true", help="do not time the one-GPU step of the large graph on rank 0")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
