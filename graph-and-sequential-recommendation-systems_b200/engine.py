"""Propagation / training-step engine: owns the HBM-resident buffers and sequences the kernels.

What the reference does with torch ops per step (code/utils.py:53-64 -> code/model.py:162-231):
cat -> L x sparse.mm -> stack -> mean -> 3 gathers -> ~10 elementwise kernels -> autograd backward
(L more sparse.mm on a transposed COO, index_put, stack/mean/cat backward) -> Adam.  Here a step is

    head         Adam step scalars + (advance of the resident batch window | pull of a pinned HOST batch)   [1 kernel]
    masks        bitmap of the <= 3B batch rows (dead-row pruning of the last forward layer)
    K1 x (L-1)   X_k = A X_{k-1}
    K1           out = (A X_{L-1} + X_0 + ... + X_{L-1}) / (L+1)        (layer mean fused; batch rows only)
    K2           loss, reg, G = d(loss + decay*reg)/d(out)              (closed form, scatter-add; clears the bitmap,
                                                                         writes the loss record to pinned host memory)
    K1 x (L-1)   g_{L-1} = s (G + A G);  g_k = s G + A g_{k+1}           (s = 1/(L+1); A symmetric)
    K1 + Adam    g_0 = s G + A g_1  consumed by the Adam epilogue, never written; zeroes the rows of G as it reads them

= 2L + 3 kernels on one stream, captured in a CUDA graph (under the row partition: + 2L rank barriers + clear_rows).  Buffers: E0 (parameters, [users;items]),
L-1 layer buffers (re-used as backward ping-pong), out, G, Adam M/V — (L+4) * N * d * 4 bytes.

Multi-GPU (one process per GPU, torch.distributed):
  mode 'dp'       graph replicated, batch sharded, G all-reduced  (weak scaling of the batch)
  mode 'dp_idx'   same replicas, but the ranks all-gather their 49 KB index batches instead of
                  all-reducing the 18 MB gradient: G is a function of (out, indices) and out is replicated,
                  so every rank scatters the global batch itself (K2 is ~10 us).  Same result, and the
                  compute part stays inside the single-GPU CUDA graph.
  mode 'rowpart'  rows of A partitioned in contiguous nnz-balanced blocks; K2 runs redundantly on the full
                  batch so no gradient collective is needed; Adam only touches owned rows.  Given an
                  ops.RowBlockBuilder instead of a whole CSRGraph, a rank holds ONLY its CSR block and only its rows of
                  the Adam moments M, V; replicated per rank: the exchanged tables E0, X_k, out and the (sparse) G.
                  Exchange of every layer's output block: FUSED into K1 (p2p=True, default on one NVSwitch box) —
                  K1's epilogue stores each finished row into all peers over NVLink while the gathers of the
                  following rows are in flight; the layers are ordered by a DEVICE-SIDE flag barrier
                  (ops.RankBarrier, csrc/exchange.cu), so the step has no collective launch at all and is captured
                  in one CUDA graph like the single-GPU step.  p2p=False falls back to NCCL broadcasts after the
                  kernel (eager, any backend — what the gloo tests exercise).
                  multicast=True (default): the exchanged tables live in torch symmetric memory and the epilogue
                  issues ONE NVSwitch multicast store (multimem.st) per 16 bytes instead of world-1 unicast stores.
                  rebalance=3 (default, rounds): the block boundaries are moved to equal measured time at start-up —
                  item rows gather from the larger table, so equal nnz (+ row cost) is not equal time.
  mode 'featpart' COLUMNS of the embedding tables partitioned: rank p holds columns [p*d/P, (p+1)*d/P) of E0, X_k, out, G, M, V
                  and the whole CSR.  A CSR SpMM is independent per column, so the 2L products of a step need NO exchange; the
                  only cross-rank dependence is K2's five dot products per triple: every rank stores its partial sums into all
                  ranks (40 KB per step, peer stores), ONE device barrier, then K2 runs on the slice with the summed records
                  (ops.FeatExchange, csrc/bpr.cu).  This is the mode for graphs that FIT one GPU (the three named shapes), where
                  the row partition is bound by the NVLink ingest of every layer's table; it does not partition the CSR.
                  Every other line of the step is the single-GPU step at width d/P, captured in one CUDA graph per rank.
"""
import torch

from . import ops


def balanced_row_bounds(indptr_cpu, parts, row_cost=0):
    """Contiguous row blocks with ~equal cost = nnz + row_cost * rows: returns parts+1 boundaries (host int list).
    row_cost expresses what publishing one output row to the peers costs in units of one gathered non-zero (with the
    fused NVLink exchange a 256-byte row sent to 7 peers is worth ~64 gathers); 0 balances non-zeros only."""
    n_rows = indptr_cpu.numel() - 1
    cost = indptr_cpu.to(torch.int64) + int(row_cost) * torch.arange(n_rows + 1, dtype=torch.int64)
    total = int(cost[-1])
    bounds = [0]
    for p in range(1, parts):
        target = total * p // parts
        r = int(torch.searchsorted(cost, torch.tensor(target, dtype=torch.int64), right=False))
        r = max(bounds[-1], min(r, n_rows))
        bounds.append(r)
    bounds.append(n_rows)
    return bounds


def rebalance_by_time(indptr_cpu, bounds, times, row_cost=0):
    """Move the boundaries so that every block would take the same time, given the time each current block took.
    Within a block the time is taken to be proportional to the block's cost (nnz + row_cost * rows): the cumulative
    time over the rows is then piecewise linear in the cumulative cost, and the new boundaries are where it reaches
    k/parts of the total.  (User rows gather item embeddings and item rows gather user embeddings — tables of very
    different size and cache residency — so equal cost is not equal time.)"""
    parts = len(bounds) - 1
    n_rows = indptr_cpu.numel() - 1
    cost = indptr_cpu.to(torch.int64) + int(row_cost) * torch.arange(n_rows + 1, dtype=torch.int64)
    t = [max(float(x), 1e-9) for x in times]
    total_t = sum(t)
    c_at = [int(cost[b]) for b in bounds]
    new = [0]
    acc_t, blk = 0.0, 0
    for k in range(1, parts):
        target = total_t * k / parts
        while blk < parts - 1 and acc_t + t[blk] < target:
            acc_t += t[blk]; blk += 1
        frac = (target - acc_t) / t[blk]
        c_target = c_at[blk] + frac * (c_at[blk + 1] - c_at[blk])
        r = int(torch.searchsorted(cost, torch.tensor(int(c_target), dtype=torch.int64), right=False))
        new.append(max(new[-1], min(r, n_rows)))
    new.append(n_rows)
    return new


def allgather_rows(buf, bounds, group=None):
    """Uneven in-place all-gather: rank p broadcasts rows [bounds[p], bounds[p+1]) of `buf`.
    Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    for p in range(world):
        b0, b1 = bounds[p], bounds[p + 1]
        if b1 > b0:
            src = dist.get_global_rank(group, p) if group is not None else p
            dist.broadcast(buf[b0:b1], src=src, group=group)
    return buf


def map_peer_buffers(t, group=None):
    """Every rank exports `t` (CUDA IPC handle of its allocation) and opens the others' in ITS OWN device context, so
    that kernels on this GPU may load/store the peers' memory directly over NVLink.  Returns a list indexed by rank
    (own entry = t itself).  torch is used for the handle exchange only."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    meta = t.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, (meta, t.storage_offset(), tuple(t.shape), tuple(t.stride())), group=group)
    out = []
    for p, (m, off, shape, stride) in enumerate(objs):
        if p == rank:
            out.append(t)
            continue
        st = torch.UntypedStorage._new_shared_cuda(dev, *m[1:])       # opened in my context: lazy peer access
        out.append(torch.empty(0, dtype=t.dtype, device=t.device).set_(st, off, shape, stride))
    return out


def symmetric_tables(n, shape, device, group, mc_out):
    """n zeroed fp32 tables in torch symmetric memory, rendezvoused over `group`; mc_out[data_ptr] = multicast address.
    Returns [None]*n when the box has no NVSwitch multicast (or torch no symmetric memory): the caller then uses
    ordinary tensors and unicast peer stores."""
    try:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        g = dist.group.WORLD if group is None else group
        tabs, ptrs = [], []
        for _ in range(n):
            t = symm_mem.empty(shape, dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, g)
            if not h.multicast_ptr:
                return [None] * n
            t.zero_()
            tabs.append(t); ptrs.append(int(h.multicast_ptr))
        for t, mp in zip(tabs, ptrs):
            mc_out[t.data_ptr()] = mp
        return tabs
    except Exception as e:          # noqa: BLE001 — any failure here means "no multicast", not an error
        import warnings
        warnings.warn(f"rowpart: symmetric memory unavailable ({type(e).__name__}: {e}); using unicast peer stores")
        mc_out.clear()
        return [None] * n


def shard_batch(n, rank, world):
    """Contiguous shard [lo,hi) of a batch of n triples for data-parallel rank `rank`."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def link_feat_engines(engines, timeout_ms=20000):
    """Wire together feature-partition ranks emulated in ONE process (Engine(..., dist_mode='featpart', feat=(rank, world)),
    one CUDA stream per engine): the 'peer' buffers are plain tensors of the same device.  The step sequence, the record
    exchange and the device barrier are the ones the multi-process mode runs."""
    world = len(engines)
    engines = sorted(engines, key=lambda e: e.rank)
    if [e.rank for e in engines] != list(range(world)) or any(e.world != world or not e._feat_virtual for e in engines):
        raise RuntimeError("link_feat_engines: need exactly one emulated engine per rank")
    # Load every kernel of the step first (a throw-away step per engine on a one-rank exchange): with lazy module loading the
    # first launch of a kernel can wait for the device to go idle — while another emulated rank's barrier is spinning on it and
    # this host thread, the only one, has not yet enqueued the rank it waits for.  (Separate processes cannot block each other.)
    for e in engines:
        rec1 = torch.zeros(2 * e.B_cap * 8, dtype=torch.float32, device=e.device)
        flg1 = torch.zeros(64, dtype=torch.int32, device=e.device)
        e.xchg = ops.FeatExchange(rec1, [rec1], flg1, [flg1], 0, 1, e.B_cap, e.scalars)
        e._warm_kernels(e.bu, e.bp, e.bn, e.ctl)
    torch.cuda.synchronize()
    recs, flags = [e._feat_records for e in engines], [e._feat_flags for e in engines]
    for e in engines:
        e.xchg = ops.FeatExchange(e._feat_records, recs, e._feat_flags, flags, e.rank, world, e.B_cap, e.scalars, timeout_ms=timeout_ms)
        e._barrier = e.xchg.barrier
        e._graphs = {}


class Engine:
    def __init__(self, csr, n_users, m_items, d, n_layers, device, *, lr=1e-3, decay=1e-4, B_cap=2048,
                 deterministic=False, use_graph=True, dist_mode=None, group=None, prune=True, p2p=True, row_cost=None, multicast=True, rebalance=3,
                 feat=None):
        """feat=(rank, world): dist_mode='featpart' WITHOUT a process group — several ranks emulated in one process (one
        engine and one CUDA stream each; tests): the caller wires them together with link_feat_engines()."""
        if n_layers > 8:
            raise RuntimeError("lightGCN_n_layers > 8 is not supported (LGCN_MAX_Z)")
        if feat is not None and dist_mode != 'featpart':
            raise RuntimeError("feat=(rank, world) goes with dist_mode='featpart' only")
        self.builder = csr if isinstance(csr, ops.RowBlockBuilder) else None
        if self.builder is not None and dist_mode != 'rowpart':
            raise RuntimeError("a RowBlockBuilder is the graph source of dist_mode='rowpart' only")
        self.csr = None if self.builder is not None else csr
        self.nu, self.ni, self.N, self.d, self.L = n_users, m_items, n_users + m_items, d, n_layers
        self.device = device
        self.decay, self.lr = float(decay), float(lr)
        self.deterministic = bool(deterministic)
        self.zero_copy = torch.device(device).type == 'cuda'
        self.fuse_small = True              # fold the step's small kernels (see _enqueue_step)
        self.B_cap = int(B_cap)
        self.dist_mode = dist_mode
        self.group = group
        self.rank, self.world = 0, 1
        self._feat_virtual = feat is not None
        if feat is not None:
            self.rank, self.world = int(feat[0]), int(feat[1])
        elif dist_mode is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.d_full, self.c0 = d, 0
        self.xchg = None
        self.pre_step_hook = None           # set by the model under the feature partition: push externally written parameters into the slice
        self.full_param_sync = None         #   "    : refresh the model's full parameter table from the ranks' slices (collective)
        if dist_mode == 'featpart':
            if d % self.world or (d // self.world) not in (8, 16, 32, 64):
                raise RuntimeError(f"dist_mode='featpart': latent_dim {d} over {self.world} ranks gives a slice of {d / self.world:g} columns; supported slices: 8, 16, 32, 64")
            if isinstance(csr, ops.RowBlockBuilder) or csr is None:
                raise RuntimeError("dist_mode='featpart' replicates the CSR: hand it the whole CSRGraph (use 'rowpart' for graphs beyond one GPU)")
            d = d // self.world                       # everything below is the single-GPU engine at the slice width
            self.d, self.c0 = d, self.rank * d
            if d == 16:                               # 4-lane groups walk 4 non-zeros per round trip: shorter segments bound the longest item
                self.csr = self.csr.rows(0, self.csr.n_rows, seg_len=min(self.csr.seg_len, 64))     # (18.8 vs 24.1 us per layer, gowalla shape)
        # fused exchange (K1 stores into the peers + device-side barrier): one NVSwitch box, CUDA only
        self.p2p = bool(p2p) and dist_mode == 'rowpart' and self.world > 1 and self.world - 1 <= 7 and torch.device(device).type == 'cuda'
        self.use_graph = bool(use_graph) and (dist_mode in (None, 'dp_idx', 'featpart') or (dist_mode == 'rowpart' and (self.p2p or self.world == 1)))
        if dist_mode == 'dp_idx':
            self.B_cap = self.B_cap * self.world          # the static batch buffers hold the GLOBAL batch
            self.deterministic = True                     # replicas are never re-synchronised: they must not drift in rounding
        N, L = self.N, self.L
        f32 = dict(dtype=torch.float32, device=device)
        # exchanged tables of the row partition: NVSwitch multicast (symmetric memory) when the box offers it — one
        # multimem store per 16 bytes reaches every replica instead of world-1 unicast stores
        self._mc = {}
        exchanged = [None] * (2 + max(L - 1, 0))
        if self.p2p and multicast:
            exchanged = symmetric_tables(len(exchanged), (N, d), device, group, self._mc)
        self.E0 = exchanged[0] if exchanged[0] is not None else torch.zeros((N, d), **f32)
        self.X = [exchanged[2 + i] if exchanged[2 + i] is not None else torch.zeros((N, d), **f32) for i in range(max(L - 1, 0))]
        self.out = exchanged[1] if exchanged[1] is not None else torch.zeros((N, d), **f32)
        self.G = torch.zeros((N, d), **f32)
        self._gradE0 = None
        self._scratch = None
        self.loss_out = torch.zeros(4, **f32)
        self._host_step = 0
        self.param_epoch = 0
        # dead-row pruning of the training step (see _enqueue_step): bitmaps over the N nodes
        # (row partition: the bitmap covers the rank's own rows, indexed by local row; (re)allocated in _take_block)
        self.prune = bool(prune) and dist_mode in (None, 'dp_idx', 'rowpart', 'featpart')
        words = (N + 31) // 32
        self.m0 = torch.zeros(words, dtype=torch.int32, device=device)
        # row partition
        self.r0, self.r1 = 0, N
        self.bounds = [0, N]
        self.local = self.csr
        self._mv_local = False
        if dist_mode == 'rowpart':
            if row_cost is None:
                row_cost = d if (self.p2p and self.world > 1) else 0
            cost_cpu = self.builder.cost_prefix if self.builder is not None else self.csr.indptr.cpu()
            self.bounds = balanced_row_bounds(cost_cpu, self.world, row_cost)
            self._take_block()
            for _ in range(int(rebalance) if self.world > 1 else 0):      # equal cost is not equal time: measure, move the boundaries
                new = self._rebalance(cost_cpu, row_cost)
                if new != self.bounds:
                    self.bounds = new
                    self._take_block()
            self._mv_local = self.builder is not None
        rows_mv = (self.r1 - self.r0) if self._mv_local else N
        self.M = torch.zeros((rows_mv, d), **f32)
        self.V = torch.zeros((rows_mv, d), **f32)
        self.scalars = ops.adam_scalars(device, self.lr)
        # fused exchange: map the peers' copies of every exchanged buffer, and the flags of the device-side barrier
        self._peer = {}
        self._e0_synced = False
        self._barrier = None
        if self.p2p:
            import torch.distributed as dist
            from . import _lib
            for p_dev in range(torch.cuda.device_count()):
                _lib.load().lgcn_enable_peer_access(p_dev)
            if not self._mc:
                for buf in [self.E0, self.out] + self.X:
                    self._peer[buf.data_ptr()] = map_peer_buffers(buf, group)
            self._flags = torch.zeros(64, dtype=torch.int32, device=device)
            torch.cuda.synchronize()
            self._barrier = ops.RankBarrier(self._flags, map_peer_buffers(self._flags, group), self.rank, self.world)
            dist.barrier(group)
        # batch staging: [ctl(4 x int32) | users | pos | neg] in one block so that a host batch is one H2D
        self.pg = None              # popularity-gate variant (enable_popgate): MLP parameter block + its Adam state
        self.i2i = None             # item-item smoothing variant (enable_i2i)
        self._alloc_batch(self.B_cap)
        self._epoch = None          # (S tensor [3,cap], ctl) for epoch-resident mode
        self._graphs = {}

    def enable_popgate(self, item_pop, params, pop_hidden, gate_hidden, temperature, entropy_coeff):
        """Train the popularity-gate variant (code/model.py:139-183) through the fused step: K2 is replaced by the pop-gate
        BPR kernel (csrc/popgate.cu), the MLP parameter block gets its own dense Adam launch.  Single GPU only."""
        if self.dist_mode is not None:
            raise NotImplementedError("the fused pop-gate step is single-GPU (the MLP gradients come from float atomics, so "
                                      "replicas would drift); use bpr_loss().backward() under dist_mode")
        self.pg = dict(pop=item_pop, params=params, H1=int(pop_hidden), H2=int(gate_hidden), temp=float(temperature), coeff=float(entropy_coeff),
                       grad=torch.zeros_like(params), M=torch.zeros_like(params), V=torch.zeros_like(params),
                       ws=ops.popgate_workspace(self.B_cap, self.device))
        self._graphs = {}

    def enable_i2i(self, i2i, i2i_t, alpha):
        """Train the item-item smoothing variant (code/model.py:228-229: items <- items + alpha * I2I @ items on the
        propagated item table) through the fused step: one more K1 product with the explicit-value item CSR after the
        propagation, and one with its transpose before the backward chain.  Single GPU; dead-row pruning is off (the
        smoothing reads the propagated rows of every neighbour in the item graph)."""
        if self.dist_mode is not None:
            raise NotImplementedError("the fused item-item step is single-GPU; use bpr_loss().backward() under dist_mode")
        f32 = dict(dtype=torch.float32, device=self.device)
        self.i2i = dict(A=i2i, At=i2i_t, alpha=float(alpha), outS=torch.zeros((self.N, self.d), **f32), G2=torch.zeros((self.N, self.d), **f32))
        self.prune = False
        self._graphs = {}

    def _take_block(self):
        """(Re)build this rank's row block for the current bounds: its own CSR from the builder (memory-partitioned), or
        views into the replicated CSR when the caller handed the whole graph."""
        self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.local = None
        self.local = self.builder.build(self.r0, self.r1) if self.builder is not None else self.csr.rows(self.r0, self.r1)
        self.m0 = torch.zeros((self.r1 - self.r0 + 31) // 32 + 1, dtype=torch.int32, device=self.device)

    def adam_state_full(self):
        """Full (N,d) Adam moments on every rank (collective in the memory-partitioned row partition: checkpoints)."""
        if self.dist_mode == 'featpart' and self.world > 1:
            return self.gather_columns(self.M), self.gather_columns(self.V)
        if not self._mv_local:
            if self.dist_mode == 'rowpart' and self.world > 1:
                M, V = self.M.clone(), self.V.clone()
                allgather_rows(M, self.bounds, self.group); allgather_rows(V, self.bounds, self.group)
                return M, V
            return self.M, self.V
        full = []
        for loc in (self.M, self.V):
            t = torch.zeros((self.N, self.d), dtype=torch.float32, device=self.device)
            t[self.r0:self.r1].copy_(loc)
            allgather_rows(t, self.bounds, self.group)
            full.append(t)
        return full[0], full[1]

    def load_adam_state(self, M_full, V_full):
        if self.dist_mode == 'featpart' and self.world > 1:
            c0, c1 = self.c0, self.c0 + self.d
            self.M.copy_(M_full[:, c0:c1]); self.V.copy_(V_full[:, c0:c1])
        elif self._mv_local:
            self.M.copy_(M_full[self.r0:self.r1]); self.V.copy_(V_full[self.r0:self.r1])
        else:
            self.M.copy_(M_full); self.V.copy_(V_full)

    def _rebalance(self, indptr_cpu, row_cost):
        """One round of time-based re-partitioning: every rank times its local product (same launch the layers use,
        without the exchange), the times are all-gathered, the boundaries recomputed identically on every rank."""
        import torch.distributed as dist
        Y = self.out[self.r0:self.r1]
        ops.spmm(self.local, self.E0, Y)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            ops.spmm(self.local, self.E0, Y)
        b.record(); b.synchronize()
        mine = torch.tensor([a.elapsed_time(b) / 3], dtype=torch.float64, device=self.device)
        allt = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allt, mine, group=self.group)
        self.out.zero_()
        times = [float(x.item()) for x in allt]
        if max(times) < 1.03 * min(times):          # already balanced to within the timing noise
            return list(self.bounds)
        return rebalance_by_time(indptr_cpu, self.bounds, times, row_cost)

    # ------------------------------------------------------------------ batch staging
    def _alloc_batch(self, B_cap):
        self.B_cap = int(B_cap)
        self._blk = torch.zeros(2 + 3 * self.B_cap, dtype=torch.int64, device=self.device)
        self.ctl = self._blk[:2].view(torch.int32)
        self.bu = self._blk[2:2 + self.B_cap]
        self.bp = self._blk[2 + self.B_cap:2 + 2 * self.B_cap]
        self.bn = self._blk[2 + 2 * self.B_cap:2 + 3 * self.B_cap]
        # two pinned staging buffers: a buffer is rewritten only after the H2D that read it has finished
        self._stage = [[torch.zeros(2 + 3 * self.B_cap, dtype=torch.int64).pin_memory(), None, None] for _ in range(2)]
        self._stage_i = 0
        self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.bpr_ws = ops.bpr_workspace(self.B_cap, self.d, self.device)
        self._graphs = {}
        if self.dist_mode == 'featpart' and self.world > 1:
            self._setup_feat_exchange()
        # zero-copy staging: a host batch is read straight out of the pinned block by a kernel at the head of the captured
        # step, and the loss is written into pinned memory by a kernel at its tail (no copy-engine hops, one graph launch)
        self._zc_slot = None
        self._loss_host_valid = False
        if self.zero_copy:
            ops.copy_words(self._loss_host, self.loss_out, 16, dst_is_host=True)      # loads the kernel before any capture
            torch.cuda.current_stream().synchronize()

    def _setup_feat_exchange(self):
        """Feature partition: this rank's record buffer and barrier flags, and (between processes) the peers' mappings.
        Collective when a process group is used — _alloc_batch is entered by every rank with the same B_cap."""
        self._feat_records = torch.zeros(2 * self.world * self.B_cap * 8, dtype=torch.float32, device=self.device)
        self._feat_flags = torch.zeros(64, dtype=torch.int32, device=self.device)
        self.xchg, self._barrier = None, None
        if self._feat_virtual:
            return                                  # link_feat_engines() wires the emulated ranks together
        import torch.distributed as dist
        from . import _lib
        if self.world - 1 > 7:
            raise RuntimeError("dist_mode='featpart' serves the <= 8 GPUs of one NVSwitch box")
        for p_dev in range(torch.cuda.device_count()):
            _lib.load().lgcn_enable_peer_access(p_dev)
        torch.cuda.synchronize()
        self.xchg = ops.FeatExchange(self._feat_records, map_peer_buffers(self._feat_records, self.group), self._feat_flags,
                                     map_peer_buffers(self._feat_flags, self.group), self.rank, self.world, self.B_cap, self.scalars)
        self._barrier = self.xchg.barrier
        dist.barrier(self.group)

    def gather_columns(self, local, out_full=None):
        """(N, d) table on every rank from the ranks' (N, d/P) column slices (collective; evaluation and checkpoints)."""
        if self.dist_mode != 'featpart' or self.world == 1:
            return local
        import torch.distributed as dist
        n = local.shape[0]
        parts = torch.empty((self.world * n, self.d), dtype=local.dtype, device=local.device)     # rank blocks stacked along dim 0 (NCCL and gloo)
        dist.all_gather_into_tensor(parts, local.contiguous(), group=self.group)
        if out_full is None:
            out_full = torch.empty((n, self.d_full), dtype=local.dtype, device=local.device)
        out_full.view(n, self.world, self.d).copy_(parts.view(self.world, n, self.d).permute(1, 0, 2))
        return out_full

    def _stage_batch(self, users, pos, neg, B_global=0):
        """Put the batch where the step reads it.  Device tensors: device-side copies into the staging block.  Host tensors:
        fill one of the two pinned blocks; with zero_copy the captured step itself pulls it in (returns the slot), else one
        H2D memcpy is enqueued here."""
        B = int(users.numel())
        if B > self.B_cap:
            torch.cuda.current_stream().synchronize()
            self._alloc_batch(B)
        self._stage_i ^= 1
        slot = self._stage[self._stage_i]
        if slot[1] is not None:
            slot[1].synchronize()                   # the step that read this pinned block two calls ago has consumed it
        h = slot[0]
        self._zc_slot = None
        if users.is_cuda:
            self.bu[:B].copy_(users.to(torch.int64), non_blocking=True)
            self.bp[:B].copy_(pos.to(torch.int64), non_blocking=True)
            self.bn[:B].copy_(neg.to(torch.int64), non_blocking=True)
            hc = h[:2].view(torch.int32)
            hc[0], hc[1], hc[2], hc[3] = 0, B, B, B_global
            self._blk[:2].copy_(h[:2], non_blocking=True)
            if slot[1] is None:
                slot[1] = torch.cuda.Event()
            slot[1].record()
            return B
        hc = h[:2].view(torch.int32)
        if slot[2] != (B, B_global):                # the control words rarely change: skip four scalar tensor writes per step
            hc[0], hc[1], hc[2], hc[3] = 0, B, B, B_global
            slot[2] = (B, B_global)
        cap = self.B_cap
        h[2:2 + B].copy_(users)
        h[2 + cap:2 + cap + B].copy_(pos)
        h[2 + 2 * cap:2 + 2 * cap + B].copy_(neg)
        if self.zero_copy:
            self._zc_slot = self._stage_i           # _enqueue_step copies the block in with a kernel; the event is recorded after the step
            return B
        if B == cap:
            self._blk.copy_(h, non_blocking=True)                       # one H2D for the whole batch
        else:
            self._blk[:2 + B].copy_(h[:2 + B], non_blocking=True)
            self._blk[2 + cap:2 + cap + B].copy_(h[2 + cap:2 + cap + B], non_blocking=True)
            self._blk[2 + 2 * cap:2 + 2 * cap + B].copy_(h[2 + 2 * cap:2 + 2 * cap + B], non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record()
        return B

    def stage_batch_now(self, users, pos, neg, B_global=0):
        """_stage_batch for callers that launch kernels on the batch themselves (the autograd path): the batch is on the
        device when this returns (stream-ordered)."""
        B = self._stage_batch(users, pos, neg, B_global)
        if self._zc_slot is not None:
            slot = self._stage[self._zc_slot]
            ops.copy_words(self._blk, slot[0])
            if slot[1] is None:
                slot[1] = torch.cuda.Event()
            slot[1].record()
            self._zc_slot = None
        return B

    # ------------------------------------------------------------------ collectives
    def _allgather_rows(self, buf):
        """Every rank broadcasts its row block of `buf` (uneven all-gather, in place)."""
        allgather_rows(buf, self.bounds, self.group)

    def _rank_barrier(self):
        """Stream-ordered rendezvous of the ranks ON THE DEVICE (flag barrier in peer-mapped memory, csrc/exchange.cu):
        every rank's stores into my buffers are complete and visible once the kernels enqueued after it start.  No
        collective launch, capturable in the step's CUDA graph; does not block the host."""
        self._barrier()

    def _layer(self, X, Y, alpha, beta, zs, row_mask=None, col_mask=None):
        r0, r1 = self.r0, self.r1
        if self.dist_mode == 'rowpart':
            peers = self._peer.get(Y.data_ptr()) if self.p2p else None
            mc = self._mc.get(Y.data_ptr(), 0) if self.p2p else 0
            if mc:
                ops.spmm(self.local, X, Y[r0:r1], alpha, beta, [z[r0:r1] for z in zs] if zs else None, row_mask=row_mask, mc_y=mc + r0 * self.d * 4)
                self._rank_barrier()
            elif peers is not None:
                ops.spmm(self.local, X, Y[r0:r1], alpha, beta, [z[r0:r1] for z in zs] if zs else None, row_mask=row_mask,
                         peer_y=[peers[p][r0:r1] for p in range(self.world) if p != self.rank])
                self._rank_barrier()
            else:
                ops.spmm(self.local, X, Y[r0:r1], alpha, beta, [z[r0:r1] for z in zs] if zs else None, row_mask=row_mask)
                self._allgather_rows(Y)
        else:
            ops.spmm(self.csr, X, Y, alpha, beta, zs, row_mask=row_mask, col_mask=col_mask)

    # ------------------------------------------------------------------ propagation
    def forward(self, masks=None):
        """out = mean_{k<=L} A^k E0  (reference code/model.py:201-222).
        masks=(m0, m1): only the rows a training step reads are produced — `out` on the batch rows m0, X_{L-1}
        on m0 + neighbours m1; earlier layers are complete.  The skipped rows are simply not written."""
        L, s = self.L, 1.0 / (self.L + 1)
        if self.dist_mode == 'rowpart' and self.world > 1 and not (self.p2p and self._e0_synced):
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("internal: parameter rows must be synchronised before the step is captured")
            self._allgather_rows(self.E0)       # owners publish their updated parameter rows
            self._e0_synced = self.p2p          # from here on the Adam epilogue publishes them itself
        if L == 0:
            self.out.copy_(self.E0)
            return self.out
        cur = self.E0
        for k in range(L - 1):
            self._layer(cur, self.X[k], 1.0, 0.0, None, row_mask=masks[1] if (masks and masks[1] is not None and k == L - 2) else None)
            cur = self.X[k]
        self._layer(cur, self.out, s, s, [self.E0] + self.X[:L - 1], row_mask=masks[0] if masks else None)
        return self.out

    def _backward_chain(self, G, last, g_rows=None):
        """g_0 = dL/dE0 given G = dL/dout.  `last(X, alpha, beta, zs, col_mask)` runs the final product.
        g_rows: bitmap of the rows of G that can be non-zero (the first product never reads the others)."""
        L, s = self.L, 1.0 / (self.L + 1)
        if L == 1:
            return last(G, s, s, [G], g_rows)
        bufs = self.X if L >= 3 else [self.X[0], None]
        if L >= 3 and len(bufs) < 2:
            raise RuntimeError("internal: missing ping-pong buffers")
        cur = bufs[0]
        self._layer(G, cur, s, s, [G], col_mask=g_rows)
        for _ in range(L - 2):
            nxt = bufs[1] if cur is bufs[0] else bufs[0]
            self._layer(cur, nxt, 1.0, s, [G])
            cur = nxt
        return last(cur, 1.0, s, [G], None)

    def backward_to(self, G, grad_out):
        """grad_out (N,d) = dL/dE0 for an arbitrary dense G (generic autograd path)."""
        if self.L == 0:
            grad_out.copy_(G)
            return grad_out
        if self.dist_mode == 'rowpart':
            r0, r1 = self.r0, self.r1

            def last(X, alpha, beta, zs, col_mask):
                ops.spmm(self.local, X, grad_out[r0:r1], alpha, beta, [z[r0:r1] for z in zs])
                self._allgather_rows(grad_out)
        else:
            def last(X, alpha, beta, zs, col_mask):
                ops.spmm(self.csr, X, grad_out, alpha, beta, zs, col_mask=col_mask)
        self._backward_chain(G, last)
        return grad_out

    def grad_buffer(self):
        if self._gradE0 is None:
            self._gradE0 = torch.zeros((self.N, self.d), dtype=torch.float32, device=self.device)
        return self._gradE0

    def scratch(self):
        if self._scratch is None:
            self._scratch = torch.zeros((self.N, self.d), dtype=torch.float32, device=self.device)
        return self._scratch

    # ------------------------------------------------------------------ fused training step
    def set_lr(self, lr):
        if float(lr) != self.lr:
            self.lr = float(lr)
            ops.adam_reinit(self.scalars, self.lr, step=self._host_step)

    def set_adam_step(self, step):
        self._host_step = int(step)
        ops.adam_reinit(self.scalars, self.lr, step=self._host_step)

    def _enqueue_step(self, users, pos, neg, ctl, only_spmm=False, advance=False):
        """Everything between 'batch is in the staging block' and 'loss_out is written'.
        advance: first move the resident epoch's batch window (ctl) by one batch.
        only_spmm (measurement): just the step's 2L K1 launches (with their exchanges), on whatever the buffers hold.
        Small kernels are folded (fuse_small): ONE head kernel (Adam tick + window advance | pull of the pinned host batch);
        K2 clears the bits of the batch-row bitmap it is the last reader of (no memset next step) and writes the loss into
        pinned host memory; the Adam-epilogue K1 zeroes the rows of G as it reads them (no clear_rows kernel)."""
        if only_spmm:
            return self._enqueue_spmm_only()
        zc, fuse = self._zc_slot, self.fuse_small
        local_g = self.dist_mode in (None, 'dp_idx', 'featpart')           # G, m0 are whole-graph and every row is visited by this rank
        if fuse:
            ops.step_begin(self.scalars, self.B_cap, advance_ctl=ctl if advance else None,
                           stage_dst=self._blk if zc is not None else None, stage_src=self._stage[zc][0] if zc is not None else None)
        else:
            if zc is not None:                      # host batch: pull the pinned staging block in (kernel, part of the graph)
                ops.copy_words(self._blk, self._stage[zc][0])
            if advance:
                ops.batch_advance(ctl, self.B_cap)
            ops.adam_tick(self.scalars)
        k2_clears_mask = fuse and self.prune and local_g and self.pg is None
        k2_writes_host = fuse and zc is not None and self.pg is None and self.dist_mode != 'dp'
        k1_clears_g = fuse and local_g and self.i2i is None and self.L > 0
        masks = None
        if self.prune:
            # Dead-row pruning: the loss reads `out` only on the <= 3B batch rows (bitmap m0), so the last forward
            # layer is evaluated there and nowhere else.  Same values as the full pass on every row that is read;
            # the other rows of `out` are not consumed before the next step overwrites them.  (Measured on the
            # yelp2018 shape: the batch rows hold ~60 % of the non-zeros because positives are popularity-biased,
            # so this saves ~15 us of the 44 us layer; pruning X_{L-1} to the 1-hop set m1 (90 % of the rows) and
            # masking the first backward product by m0 did not pay and are not used — profiles/README.md.)
            if self.dist_mode == 'rowpart':      # each rank prunes to the batch rows it owns (local bitmap)
                ops.batch_masks_rows(users, pos, neg, self.B_cap, ctl, self.nu, self.r0, self.r1, self.m0)
            else:
                # (the bitmap starts all-zero and, when K2 clears the bits it set, is all-zero again after every step)
                ops.batch_masks(users, pos, neg, self.B_cap, ctl, self.nu, self.csr, self.m0, None, clear_first=not k2_clears_mask)
            masks = (self.m0, None)
        self.forward(masks)
        out, G_chain = self.out, self.G
        if self.i2i is not None:
            ii, nu = self.i2i, self.nu
            out = ii['outS']
            out[:nu].copy_(self.out[:nu])
            ops.spmm(ii['A'], self.out[nu:], out[nu:], ii['alpha'], 1.0, [self.out[nu:]])      # items + alpha * I2I @ items
        if self.dist_mode == 'dp':
            import torch.distributed as dist
            ops.bpr_fwd_bwd(self.out, users, pos, neg, self.B_cap, ctl, self.nu, self.ni, 0.0, self.decay, 1.0,
                            self.decay, self.loss_out, self.G, self.bpr_ws, deterministic=self.deterministic)
            dist.all_reduce(self.G, group=self.group)
            dist.all_reduce(self.loss_out[:3], group=self.group)
        elif self.pg is not None:
            pg = self.pg
            ops.popgate_bpr_fwd_bwd(out, users, pos, neg, self.B_cap, ctl, self.nu, self.ni, pg['pop'], pg['params'], pg['H1'], pg['H2'],
                                    pg['temp'], pg['coeff'], self.decay, self.loss_out, self.G, pg['grad'], pg['ws'])
        elif self.dist_mode == 'featpart' and self.world > 1:
            if self.xchg is None:
                raise RuntimeError("featpart: the emulated ranks were not linked (link_feat_engines)")
            # the one exchange of the step: partial dot products into every rank, barrier, K2 on the slice with the sums
            ops.bpr_feat_partial(out, users, pos, neg, self.B_cap, ctl, self.nu, self.ni, self.xchg, self.bpr_ws)
            self.xchg.barrier()
            ops.bpr_feat_finish(out, users, pos, neg, self.B_cap, ctl, self.nu, self.ni, 0.0, self.decay, 1.0, self.decay,
                                self.loss_out, self.G, self.xchg, self.bpr_ws, deterministic=self.deterministic,
                                clear_mask=self.m0 if k2_clears_mask else None, loss_host=self._loss_host if k2_writes_host else None)
        else:
            ops.bpr_fwd_bwd(out, users, pos, neg, self.B_cap, ctl, self.nu, self.ni, 0.0, self.decay, 1.0,
                            self.decay, self.loss_out, self.G, self.bpr_ws, deterministic=self.deterministic,
                            clear_mask=self.m0 if k2_clears_mask else None, loss_host=self._loss_host if k2_writes_host else None)
        if self.i2i is not None:    # back through the smoothing: dL/d(items) = G' + alpha * I2I^T @ G'
            ii, nu = self.i2i, self.nu
            G_chain = ii['G2']
            G_chain[:nu].copy_(self.G[:nu])
            ops.spmm(ii['At'], self.G[nu:], G_chain[nu:], ii['alpha'], 1.0, [self.G[nu:]])
        r0, r1 = self.r0, self.r1
        Mo, Vo = (self.M, self.V) if self._mv_local else (self.M[r0:r1], self.V[r0:r1])
        if self.L == 0:
            ops.adam(self.E0[r0:r1], Mo, Vo, G_chain[r0:r1], self.scalars)
        else:
            g = self.local if self.dist_mode == 'rowpart' else self.csr

            peers_e0 = self._peer.get(self.E0.data_ptr()) if self.p2p else None
            mc_e0 = self._mc.get(self.E0.data_ptr(), 0) if self.p2p else 0

            def last(X, alpha, beta, zs, col_mask):
                ops.spmm_adam(g, X, self.E0[r0:r1], Mo, Vo, self.scalars, alpha, beta,
                              [z[r0:r1] for z in zs], col_mask=col_mask,
                              peer_p=None if (peers_e0 is None or mc_e0) else [peers_e0[p][r0:r1] for p in range(self.world) if p != self.rank],
                              mc_p=(mc_e0 + r0 * self.d * 4) if mc_e0 else 0, clear_z0=k1_clears_g)
                if peers_e0 is not None or mc_e0:        # the updated parameter rows are already in every replica
                    self._rank_barrier()
            self._backward_chain(G_chain, last)
        if self.pg is not None:     # the 8 MLP tensors are one flat block: one dense Adam launch, same step scalars
            ops.adam(self.pg['params'], self.pg['M'], self.pg['V'], self.pg['grad'], self.scalars)
            self.pg['grad'].zero_()
        if self.dist_mode == 'dp':
            self.G.zero_()          # the all-reduced G is dense in the rows any rank touched
        elif not k1_clears_g:
            ops.bpr_clear_rows(self.G, users, pos, neg, self.B_cap, ctl, self.nu)
        if zc is not None and not k2_writes_host:   # ... and push the loss into pinned memory: loss_to_host() only has to wait
            ops.copy_words(self._loss_host, self.loss_out, 16, dst_is_host=True)

    def _enqueue_spmm_only(self):
        self.forward((self.m0, None) if self.prune else None)
        r0, r1 = self.r0, self.r1
        Mo, Vo = (self.M, self.V) if self._mv_local else (self.M[r0:r1], self.V[r0:r1])
        g = self.local if self.dist_mode == 'rowpart' else self.csr
        peers_e0 = self._peer.get(self.E0.data_ptr()) if self.p2p else None
        mc_e0 = self._mc.get(self.E0.data_ptr(), 0) if self.p2p else 0

        def last(X, alpha, beta, zs, col_mask):
            ops.spmm_adam(g, X, self.E0[r0:r1], Mo, Vo, self.scalars, alpha, beta, [z[r0:r1] for z in zs], col_mask=col_mask,
                          peer_p=None if (peers_e0 is None or mc_e0) else [peers_e0[p][r0:r1] for p in range(self.world) if p != self.rank],
                          mc_p=(mc_e0 + r0 * self.d * 4) if mc_e0 else 0)
            if peers_e0 is not None or mc_e0:
                self._rank_barrier()
        if self.L > 0:
            self._backward_chain(self.G, last)

    def spmm_only_graph(self):
        """A CUDA graph holding exactly the 2L K1 launches of a training step (same arguments, same row mask as the last
        step) — bench.py times its replays to get K1's launch duration under the conditions of the captured step.
        Replays run the Adam epilogue on whatever G holds: callers save and restore E0/M/V/scalars around them."""
        self._sync_params_across_ranks()
        if self.prune and self.dist_mode in (None, 'dp_idx', 'featpart'):
            # K2 leaves the batch-row bitmap all-zero after every step: set the bits of the current batch window again
            # (they stay set: these replays have no K2), so that the masked layer does the work it does in a real step
            if self._epoch is not None:
                S, ctl = self._epoch
                ops.batch_masks(S[0], S[1], S[2], self.B_cap, ctl, self.nu, self.csr, self.m0, None, clear_first=True)
            else:
                ops.batch_masks(self.bu, self.bp, self.bn, self.B_cap, self.ctl, self.nu, self.csr, self.m0, None, clear_first=True)
        self._enqueue_spmm_only()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue_spmm_only()
        return g

    def clear_batch_mask(self):
        """Back to the all-zero bitmap the captured steps expect (after spmm_only_graph measurements)."""
        self.m0.zero_()

    def _warm_kernels(self, users, pos, neg, ctl, advance=False):
        """Run the step once on throw-away state so every kernel is loaded before graph capture."""
        state = (self.E0, self.M, self.V, self.scalars, self.loss_out, ctl)
        if self.pg is not None:
            state = state + (self.pg['params'], self.pg['M'], self.pg['V'])
        saved = [t.clone() for t in state]
        ctl.zero_()                                   # B = 0: the batch kernels touch nothing
        self._enqueue_step(users, pos, neg, ctl, advance=advance)
        for dst, src in zip(state, saved):
            dst.copy_(src)
        self.G.zero_()
        torch.cuda.current_stream().synchronize()

    def _run(self, key, users, pos, neg, ctl, advance=False):
        if not self.use_graph:
            self._enqueue_step(users, pos, neg, ctl, advance=advance)
            return
        g = self._graphs.get(key)
        if g is None:
            self._warm_kernels(users, pos, neg, ctl, advance=advance)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_step(users, pos, neg, ctl, advance=advance)
            self._graphs[key] = g
        g.replay()

    def step(self, users, pos, neg, B_global=0):
        """One BPR step on a batch given as host or device int64 tensors. Returns nothing; the loss is
        in self.loss_out (device).  Host batches cost one H2D copy."""
        if self.pre_step_hook is not None:
            self.pre_step_hook()
        if self.dist_mode == 'dp' and not B_global:
            B_global = int(users.numel()) * self.world          # equal shards unless the caller says otherwise
        if self.dist_mode == 'dp_idx':
            users, pos, neg = self._gather_indices(users, pos, neg)
            B_global = 0
        self._stage_batch(users, pos, neg, B_global)
        self._sync_params_across_ranks()
        zc = self._zc_slot
        self._run('direct' if zc is None else ('host', zc), self.bu, self.bp, self.bn, self.ctl)
        if zc is not None:
            slot = self._stage[zc]
            if slot[1] is None:
                slot[1] = torch.cuda.Event()
            slot[1].record()                        # this pinned block may be refilled once the step has run
        self._loss_host_valid = zc is not None
        self._zc_slot = None
        self._host_step += 1
        self.param_epoch += 1

    def _sync_params_across_ranks(self):
        """Row partition with the fused exchange: after the parameters were set from outside (init, load_state_dict) every
        rank's replica of E0 must be the same before the (captured) step sequence starts; afterwards the Adam epilogue
        keeps the replicas in step by itself."""
        if self.dist_mode == 'rowpart' and self.p2p and not self._e0_synced:
            self._allgather_rows(self.E0)
            self._e0_synced = True

    def sync_params_for_read(self):
        """Make this rank's replica of the parameter table complete (collective; no-op when the fused exchange already
        keeps the replicas in step or when nothing is partitioned)."""
        if self.full_param_sync is not None:
            self.full_param_sync()
        if self.dist_mode == 'rowpart' and self.world > 1 and not (self.p2p and self._e0_synced):
            self._allgather_rows(self.E0)
            self._e0_synced = self.p2p

    def _gather_indices(self, users, pos, neg):
        """dp_idx: all-gather every rank's (users,pos,neg) shard -> the global batch on every rank."""
        import torch.distributed as dist
        Bl = int(users.numel())
        loc = torch.stack([users.to(self.device, torch.int64, non_blocking=True), pos.to(self.device, torch.int64, non_blocking=True),
                           neg.to(self.device, torch.int64, non_blocking=True)])                     # [3, Bl]
        allb = torch.empty((self.world, 3, Bl), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allb, loc, group=self.group)
        g = allb.permute(1, 0, 2).reshape(3, self.world * Bl)
        return g[0], g[1], g[2]

    def loss_to_host(self):
        """{bpr, reg, total, running sum} on the host; synchronises the stream.  After a zero-copy step the values are already
        in pinned memory (written by the step's last kernel); otherwise one 16-byte D2H."""
        if not self._loss_host_valid:
            self._loss_host.copy_(self.loss_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._loss_host

    # epoch-resident mode: the whole epoch's shuffled triples live on the device, each step only moves
    # the device-side window (lgcn_batch_advance) and replays the graph.
    def begin_epoch(self, S_dev):
        """S_dev int64[3, n] on the device (users, pos, neg rows)."""
        n = S_dev.shape[1]
        if self._epoch is None or self._epoch[0].shape[1] < n:
            cap = max(n, 1)
            S = torch.zeros((3, cap), dtype=torch.int64, device=self.device)
            ctl = torch.zeros(4, dtype=torch.int32, device=self.device)
            self._epoch = (S, ctl)
            self._graphs.pop('epoch', None)
        S, ctl = self._epoch
        S[:, :n].copy_(S_dev)
        ctl.copy_(torch.tensor([0, 0, n, 0], dtype=torch.int32), non_blocking=False)
        self.loss_out.zero_()
        self._epoch_pos = 0
        self._epoch_steps = (n + self.B_cap - 1) // self.B_cap
        return self._epoch_steps

    def rewind_epoch(self):
        """Start the resident epoch over (same triples) without touching the data: only the window is reset."""
        S, ctl = self._epoch
        ctl[:2].zero_()
        self._epoch_pos = 0

    def epoch_step(self):
        S, ctl = self._epoch
        if self.pre_step_hook is not None:
            self.pre_step_hook()
        self._sync_params_across_ranks()
        self._zc_slot = None
        self._loss_host_valid = False
        self._run('epoch', S[0], S[1], S[2], ctl, advance=True)
        self._host_step += 1
        self.param_epoch += 1
