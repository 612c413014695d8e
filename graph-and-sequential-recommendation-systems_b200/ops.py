"""Tensor-level wrappers over the C ABI (include/lgcn_b200.h).

torch is used for device memory and streams only; every computation below is a kernel of
liblgcn_b200.so launched on torch's current CUDA stream.  Nothing here falls back to torch ops.
"""
import ctypes
import os
from ctypes import byref, c_void_p

import torch

from . import _lib

DEFAULT_SEG_LEN = 128
# Column-slab blocking of K1 (csrc/spmm.cu) for gathered tables that do not fit L2: slabs of at most SLAB_BYTES stay resident
# in the 126 MB L2 next to the streams of one launch.  OFF by default (LGCN_BLOCKING=1 turns it on for tables larger than
# BLOCK_THRESHOLD_BYTES; CSRGraph.block_plans(d, slab_bytes=...) forces it): measured on BASELINE config 5 it cuts the DRAM
# traffic of a layer from 245 GB to 155 GB but the layer takes the same 35-36 ms — every gathered byte still has to cross
# L2 -> SM, and gathers + fills + running sums together saturate the L2 slices (profiles/README.md, round 2).
SLAB_BYTES = int(os.environ.get('LGCN_SLAB_MB', '64')) << 20
BLOCK_THRESHOLD_BYTES = int(os.environ.get('LGCN_BLOCK_THRESHOLD_MB', '112')) << 20
BLOCKING_DEFAULT = os.environ.get('LGCN_BLOCKING', '0') == '1'
# L2 eviction hints for K1's gathers when the gathered table does not fit L2 (csrc/spmm.cu, HINTED): the rows of the hottest
# columns — as many as fill HOT_BYTES — are gathered evict-last, everything else evict-first.  OFF by default (LGCN_L2_HINTS=1,
# or CSRGraph.hint_indices(d, hot_bytes=...)): measured on BASELINE config 5 it changes nothing — 36.2 ms per layer without,
# 35.9-36.4 ms with hot sets of 16-96 MB covering 13-28 % of the non-zeros (profiles/r2_hint_probe.jsonl): the hardware's own
# replacement already keeps the frequently re-read rows, and what misses are the rows that are read once.
HINTS_DEFAULT = os.environ.get('LGCN_L2_HINTS', '0') == '1'
HOT_BYTES = int(os.environ.get('LGCN_HOT_MB', '48')) << 20


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _need(t, dtype, name, dim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if dim is not None and t.dim() != dim:
        raise RuntimeError(f"{name}: expected {dim} dimensions, got {t.dim()}")
    return t


def device_info():
    out = (ctypes.c_int32 * 3)()
    _lib.check(_lib.load().lgcn_device_info(out), "device_info")
    return {"sm_count": out[0], "max_smem_optin": out[1], "cc": out[2]}


class CSRGraph:
    """int32 CSR on the device + the SpMM segment plan for its long rows."""

    def __init__(self, indptr, indices, vals, n_cols, deg=None, dinv=None, seg_len=DEFAULT_SEG_LEN, col_weight=None):
        self.col_weight = col_weight          # optional int32[n_cols]: how often a column is gathered (its degree) — L2 hints
        self._hinted = None
        self._hint_bytes = None
        self.indptr = _need(indptr, torch.int32, "indptr", 1)
        self.indices = _need(indices, torch.int32, "indices", 1)
        self.vals = _need(vals, torch.float32, "vals", 1)
        self.n_rows = indptr.numel() - 1
        self.n_cols = int(n_cols)
        self.nnz = indices.numel()
        self.deg, self.dinv = deg, dinv
        self.device = indptr.device
        self._build_plan(seg_len)

    def _build_plan(self, seg_len):
        """Degree-binned work items (descending length) + segments of the long rows."""
        lib = _lib.load()
        self.seg_len = int(seg_len)
        counts = torch.zeros(4, dtype=torch.int32, device=self.device)
        _lib.check(lib.lgcn_spmm_plan_count(_p(self.indptr), self.n_rows, self.seg_len, _p(counts), _stream()), "spmm_plan_count")
        n_long, n_segs, max_len, _ = (int(v) for v in counts.cpu().tolist())      # one-time setup sync
        self.max_item_len = max_len
        self.n_long, self.n_segs = n_long, n_segs
        self.n_items = self.n_rows - n_long + n_segs
        self._items = torch.zeros(max(self.n_items, 1) * 4, dtype=torch.int32, device=self.device)
        self._seginfo = torch.zeros(max(n_segs, 1) * 4, dtype=torch.int32, device=self.device)
        self._counters = torch.zeros(max(n_long, 1), dtype=torch.int32, device=self.device)
        ws_bytes = lib.lgcn_spmm_plan_workspace_bytes(max_len)
        ws = torch.zeros((ws_bytes + 3) // 4, dtype=torch.int32, device=self.device)
        _lib.check(lib.lgcn_spmm_plan_fill(_p(self.indptr), self.n_rows, self.seg_len, max_len, _p(self._items), _p(self._seginfo),
                                           _p(ws), ws.numel() * 4, _stream()), "spmm_plan_fill")
        self._partials = None
        self._plan = None
        self.use_plan = True
        self.use_blocking = True      # honoured only for plans that exist: automatic planning needs BLOCKING_DEFAULT
        self._blocks = None           # (d_max, [SpmmPlan per non-empty slab], keep-alive tensors)

    def block_plans(self, d, slab_bytes=None):
        """Column-slab plans for gathers from an (n_cols, d) fp32 table that does not fit L2, else None.
        One layer = one launch per returned plan, in order (ascending columns)."""
        if not (self.use_plan and self.use_blocking) or self.n_rows == 0:
            return None
        if self._blocks is not None and self._blocks[0] >= d and slab_bytes is None:
            return self._blocks[1]                    # planned before (automatically, or forced with an explicit slab size)
        if slab_bytes is None and (not BLOCKING_DEFAULT or self.n_cols * d * 4 <= BLOCK_THRESHOLD_BYTES):
            return None
        lib = _lib.load()
        slab_bytes = int(slab_bytes or SLAB_BYTES)
        n_slabs = max(1, -(-self.n_cols * d * 4 // slab_bytes))
        slab_cols = -(-self.n_cols // n_slabs)
        dev = self.device
        acc = torch.empty(self.n_rows * d, dtype=torch.float32, device=dev)
        counts = torch.zeros(4, dtype=torch.int32, device=dev)
        plans, keep, max_long, max_segs = [], [acc], 1, 1
        specs = []
        for sidx in range(n_slabs):
            lo, hi = sidx * slab_cols, min(self.n_cols, (sidx + 1) * slab_cols)
            _lib.check(lib.lgcn_spmm_plan_count_slab(_p(self.indptr), _p(self.indices), self.n_rows, self.seg_len, lo, hi, int(sidx == 0),
                                                     _p(counts), _stream()), "spmm_plan_count_slab")
            n_long, n_segs, max_len, n_with = (int(v) for v in counts.cpu().tolist())          # setup-time sync, once per slab
            n_items = n_with - n_long + n_segs
            if n_items == 0:
                continue
            items = torch.empty(n_items * 4, dtype=torch.int32, device=dev)
            seginfo = torch.zeros(max(n_segs, 1) * 4, dtype=torch.int32, device=dev)
            ws_bytes = lib.lgcn_spmm_plan_workspace_bytes(max_len)
            ws = torch.zeros((ws_bytes + 3) // 4, dtype=torch.int32, device=dev)
            _lib.check(lib.lgcn_spmm_plan_fill_slab(_p(self.indptr), _p(self.indices), self.n_rows, self.seg_len, max_len, lo, hi, int(sidx == 0),
                                                    _p(items), _p(seginfo), _p(ws), ws.numel() * 4, _stream()), "spmm_plan_fill_slab")
            keep += [items, seginfo]
            max_long, max_segs = max(max_long, n_long), max(max_segs, n_segs)
            specs.append((n_long, n_segs, n_items, items, seginfo))
        counters = torch.zeros(max_long, dtype=torch.int32, device=dev)
        partials = torch.empty(max_segs * d, dtype=torch.float32, device=dev)
        keep += [counters, partials]
        for n_long, n_segs, n_items, items, seginfo in specs:
            pl = _lib.SpmmPlan()
            pl.seg_len, pl.n_long, pl.n_segs, pl.n_items, pl.d_max = self.seg_len, n_long, n_segs, n_items, d
            pl.items, pl.seginfo = items.data_ptr(), seginfo.data_ptr()
            pl.counters, pl.partials, pl.acc = counters.data_ptr(), partials.data_ptr(), acc.data_ptr()
            plans.append(pl)
        self._blocks = (d, plans, keep)
        self.n_slabs, self.n_block_items = n_slabs, sum(sp[2] for sp in specs)
        return plans

    def clear_blocking(self):
        self._blocks = None

    def hint_indices(self, d, hot_bytes=None):
        """Hinted copy of the column indices for gathers from an (n_cols, d) table that does not fit L2 (bit 31 = hot column),
        or None.  hot_bytes forces it (and its size) regardless of the table size."""
        if hot_bytes is None:
            if not HINTS_DEFAULT or self.col_weight is None or self.n_cols * d * 4 <= BLOCK_THRESHOLD_BYTES or self.nnz == 0:
                return None
            hot_bytes = HOT_BYTES
        if self.col_weight is None:
            raise RuntimeError("hint_indices: this graph carries no column weights")
        if self._hinted is not None and self._hint_bytes == (int(hot_bytes), d):
            return self._hinted
        k = max(1, min(self.n_cols, int(hot_bytes) // (4 * d)))
        thr = int(torch.topk(self.col_weight, k).values[-1].item())          # setup: degree of the k-th hottest column
        thr = max(thr, 2)                                                      # a column read once has nothing to keep
        out = torch.empty_like(self.indices)
        _lib.check(_lib.load().lgcn_spmm_hint_indices(_p(self.indices), self.nnz, _p(self.col_weight), thr, _p(out), _stream()), "spmm_hint_indices")
        self._hinted, self._hint_bytes, self._plan = out, (int(hot_bytes), d), None
        return out

    def clear_hints(self):
        self._hinted, self._hint_bytes, self._plan = None, None, None

    def plan(self, d):
        if not self.use_plan:
            return None
        if self._plan is None or self._plan.d_max < d:
            self._partials = torch.empty(max(self.n_segs, 1) * d, dtype=torch.float32, device=self.device)
            pl = _lib.SpmmPlan()
            pl.seg_len, pl.n_long, pl.n_segs, pl.n_items, pl.d_max = self.seg_len, self.n_long, self.n_segs, self.n_items, d
            pl.items = self._items.data_ptr()
            pl.seginfo = self._seginfo.data_ptr()
            pl.counters = self._counters.data_ptr()
            pl.partials = self._partials.data_ptr()
            h = self.hint_indices(d) if self._hinted is None else self._hinted
            pl.hinted_indices = h.data_ptr() if h is not None else None
            self._plan = pl
        return self._plan

    def _plan_ref(self, d):
        pl = self.plan(d)
        return None if pl is None else byref(pl)

    def rows(self, begin, end, seg_len=None):
        """Row block [begin,end) as its own CSRGraph (indptr rebased; shares indices/vals storage)."""
        ip = (self.indptr[begin:end + 1] - self.indptr[begin]).contiguous()
        lo, hi = int(self.indptr[begin]), int(self.indptr[end])
        return CSRGraph(ip, self.indices[lo:hi], self.vals[lo:hi], self.n_cols,
                        seg_len=self.seg_len if seg_len is None else seg_len, col_weight=self.col_weight)

    def to_torch_sparse_csr(self):
        return torch.sparse_csr_tensor(self.indptr, self.indices, self.vals, size=(self.n_rows, self.n_cols),
                                       check_invariants=False)

    def algorithmic_bytes(self, d):
        """B_spmm = 8*nnz + 4*(N+1) + 8*N*d."""
        return 8 * self.nnz + 4 * (self.n_rows + 1) + 8 * self.n_rows * d


def csr_build(train_user, train_item, n_users, m_items, seg_len=DEFAULT_SEG_LEN):
    """K4: (trainUser, trainItem) int64 device arrays -> normalised symmetric bipartite CSRGraph."""
    lib = _lib.load()
    tu = _need(train_user, torch.int64, "train_user", 1)
    ti = _need(train_item, torch.int64, "train_item", 1)
    E = tu.numel()
    if ti.numel() != E:
        raise RuntimeError("csr_build: trainUser and trainItem differ in length")
    dev = tu.device
    N = n_users + m_items
    indptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(2 * E, 1), dtype=torch.int32, device=dev)
    vals = torch.empty(max(2 * E, 1), dtype=torch.float32, device=dev)
    deg = torch.empty(N, dtype=torch.float32, device=dev)
    dinv = torch.empty(N, dtype=torch.float32, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.lgcn_csr_build_workspace_bytes(E, n_users, m_items)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    _lib.check(lib.lgcn_csr_build(_p(tu), _p(ti), E, n_users, m_items, _p(indptr), _p(indices), _p(vals),
                                  _p(deg), _p(dinv), _p(nnz), _p(status), c_void_p(ws_ptr), ws_bytes, _stream()), "csr_build")
    nnz_h, status_h = int(nnz.item()), int(status.item())
    if status_h != 0:
        raise RuntimeError("csr_build: a user or item id is outside [0,n_users) x [0,m_items)")
    del ws
    return CSRGraph(indptr, indices[:nnz_h], vals[:nnz_h], N, deg=deg, dinv=dinv, seg_len=seg_len,
                    col_weight=deg.to(torch.int32) if N * 64 * 4 > BLOCK_THRESHOLD_BYTES else None)


class RowBlockBuilder:
    """Row blocks of the normalised adjacency straight from an edge stream (multi-GPU row partition):
    a rank never holds more of the CSR than the rows it owns.  `chunks` is a zero-argument callable returning an
    iterable of (train_user, train_item) int64 CUDA tensor pairs — the whole edge list in one piece, or pieces that are
    re-generated on every pass — and is walked once for the degrees and once per build().  Only setup code: a few host
    syncs (key counts) are fine here."""

    def __init__(self, n_users, m_items, chunks, seg_len=DEFAULT_SEG_LEN, device=None):
        lib = _lib.load()
        self.n_users, self.m_items, self.N = int(n_users), int(m_items), int(n_users) + int(m_items)
        self.chunks, self.seg_len = chunks, int(seg_len)
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.counts = torch.zeros(self.N, dtype=torch.int32, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_edges = 0
        for tu, ti in chunks():
            _need(tu, torch.int64, "train_user", 1), _need(ti, torch.int64, "train_item", 1)
            _lib.check(lib.lgcn_degree_accumulate(_p(tu), _p(ti), tu.numel(), self.n_users, self.m_items, _p(self.counts), _p(status), _stream()),
                       "degree_accumulate")
            self.n_edges += tu.numel()
        if int(status.item()) != 0:
            raise RuntimeError("row-block build: a user or item id is outside [0,n_users) x [0,m_items)")
        self.deg = torch.empty(self.N, dtype=torch.float32, device=self.device)
        self.dinv = torch.empty(self.N, dtype=torch.float32, device=self.device)
        _lib.check(lib.lgcn_degree_finalize(_p(self.counts), self.N, _p(self.deg), _p(self.dinv), _stream()), "degree_finalize")
        c = self.counts.cpu().to(torch.int64)
        # cumulative stored entries per row (duplicates counted): what the partition is balanced on — plays the role of indptr
        self.cost_prefix = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(c, 0)])
        self.n_rows = self.n_cols = self.N

    def build(self, r0, r1):
        """CSRGraph of rows [r0, r1): local indptr, GLOBAL column ids, n_cols = N."""
        lib = _lib.load()
        r0, r1 = int(r0), int(r1)
        cap = int(self.cost_prefix[r1] - self.cost_prefix[r0])
        ws_bytes = lib.lgcn_csr_rows_workspace_bytes(cap, r1 - r0)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=self.device)
        ws_ptr = c_void_p((ws.data_ptr() + 255) // 256 * 256)
        cursor = torch.zeros(1, dtype=torch.int64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        for tu, ti in self.chunks():
            _lib.check(lib.lgcn_csr_rows_emit(_p(tu), _p(ti), tu.numel(), self.n_users, self.m_items, r0, r1, cap, _p(cursor), _p(status),
                                              ws_ptr, ws_bytes, _stream()), "csr_rows_emit")
        n_keys, st = int(cursor.item()), int(status.item())
        if st != 0 or n_keys != cap:
            raise RuntimeError(f"row-block build: edge stream changed between passes (status {st}, {n_keys} keys, expected {cap})")
        indptr = torch.empty(r1 - r0 + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(max(n_keys, 1), dtype=torch.int32, device=self.device)
        vals = torch.empty(max(n_keys, 1), dtype=torch.float32, device=self.device)
        nnz = torch.zeros(1, dtype=torch.int64, device=self.device)
        _lib.check(lib.lgcn_csr_rows_finish(n_keys, cap, self.n_users, self.m_items, r0, r1, _p(self.dinv), _p(indptr), _p(indices), _p(vals),
                                            _p(nnz), ws_ptr, ws_bytes, _stream()), "csr_rows_finish")
        nnz_h = int(nnz.item())
        del ws
        if nnz_h < int(0.9 * n_keys):           # many duplicates: do not keep the slack alive
            indices, vals = indices[:nnz_h].clone(), vals[:nnz_h].clone()
        return CSRGraph(indptr, indices[:nnz_h], vals[:nnz_h], self.N, seg_len=self.seg_len, col_weight=self.counts)


class RankBarrier:
    """Device-side rendezvous of the ranks (lgcn_rank_barrier): flags in peer-mapped memory, no collective launch, capturable
    in a CUDA graph.  `peer_flags` = every rank's flags tensor as seen from this process (own entry included)."""

    def __init__(self, flags_local, peer_flags, rank, world, timeout_ms=20000):
        self.flags, self.peers, self.rank, self.world = flags_local, list(peer_flags), int(rank), int(world)
        self.timeout_ms = int(timeout_ms)
        dev = flags_local.device
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._arr = (c_void_p * self.world)()
        for p, t in enumerate(self.peers):
            self._arr[p] = t.data_ptr()
        self.count = 0

    def __call__(self):
        _lib.check(_lib.load().lgcn_rank_barrier(_p(self.flags), self._arr, self.rank, self.world, _p(self.epoch), _p(self.err),
                                                 self.timeout_ms, _stream()), "rank_barrier")
        self.count += 1

    def check(self):
        """Host-side (synchronising) check that no barrier timed out."""
        e = int(self.err.item())
        if e:
            raise RuntimeError(f"rank {self.rank}: device barrier timed out waiting for rank {e - 1}")


def coo_to_csr(rows, cols, vals, n_rows, n_cols, seg_len=DEFAULT_SEG_LEN):
    """Row-major-sorted COO (torch coalesced layout) -> CSRGraph."""
    lib = _lib.load()
    rows = _need(rows, torch.int64, "rows", 1)
    cols = _need(cols, torch.int64, "cols", 1)
    vals = _need(vals.to(torch.float32), torch.float32, "vals", 1)
    nnz = rows.numel()
    indptr = torch.empty(n_rows + 1, dtype=torch.int32, device=rows.device)
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=rows.device)
    _lib.check(lib.lgcn_coo_to_csr(_p(rows), _p(cols), nnz, n_rows, _p(indptr), _p(indices), _stream()), "coo_to_csr")
    return CSRGraph(indptr, indices[:nnz], vals, n_cols, seg_len=seg_len)


def _z_array(zs):
    zs = list(zs or [])
    if len(zs) > _lib.MAX_Z:
        raise RuntimeError(f"spmm: at most {_lib.MAX_Z} own-row addends (n_layers <= {_lib.MAX_Z})")
    arr = (c_void_p * max(len(zs), 1))()
    for i, z in enumerate(zs):
        arr[i] = z.data_ptr()
    return arr, len(zs)


def _peers_struct(peer_y=None, peer_p=None, mc_y=0, mc_p=0):
    """peer_y / peer_p: the peers' copies of the destination row block (tensors, unicast stores);
    mc_y / mc_p: instead, the NVSwitch multicast ADDRESS of the row block (int): one store reaches every replica."""
    if mc_y or mc_p:
        st = _lib.SpmmPeers()
        st.n_peers, st.multicast = 1, 1
        st.y[0] = int(mc_y) if mc_y else None
        st.p[0] = int(mc_p) if mc_p else None
        return byref(st)
    ys, ps = list(peer_y or []), list(peer_p or [])
    n = max(len(ys), len(ps))
    if n == 0:
        return None
    st = _lib.SpmmPeers()
    st.n_peers, st.multicast = n, 0
    for i in range(n):
        st.y[i] = ys[i].data_ptr() if i < len(ys) and ys[i] is not None else None
        st.p[i] = ps[i].data_ptr() if i < len(ps) and ps[i] is not None else None
    return byref(st)


def spmm(g, X, Y, alpha=1.0, beta=0.0, zs=None, row_mask=None, col_mask=None, peer_y=None, mc_y=0):
    """Y = alpha * (A @ X) + beta * sum(zs)  — K1.  X is indexed by column id, Y/zs by local row."""
    lib = _lib.load()
    d = X.shape[1]
    _need(X, torch.float32, "X", 2), _need(Y, torch.float32, "Y", 2)
    if X.shape[0] < g.n_cols or Y.shape[0] < g.n_rows or Y.shape[1] != d:
        raise RuntimeError(f"spmm: shape mismatch X{tuple(X.shape)} Y{tuple(Y.shape)} graph {g.n_rows}x{g.n_cols}")
    if Y.data_ptr() == X.data_ptr():
        raise RuntimeError("spmm: Y must not alias X")
    arr, nz = _z_array(zs)
    for z in (zs or []):
        _need(z, torch.float32, "z", 2)
    blocks = g.block_plans(d) if col_mask is None else None
    if blocks:        # the gathered table does not fit L2: one launch per column slab, running sums carried in plan.acc
        peers = _peers_struct(peer_y, None, mc_y)
        for pl in blocks:
            _lib.check(lib.lgcn_spmm_f32(_p(g.indptr), _p(g.indices), _p(g.vals), g.n_rows, d, _p(X), _p(Y),
                                         float(alpha), float(beta), arr, nz, byref(pl), _p(row_mask), None, peers, _stream()), "spmm")
        return Y
    _lib.check(lib.lgcn_spmm_f32(_p(g.indptr), _p(g.indices), _p(g.vals), g.n_rows, d, _p(X), _p(Y),
                                 float(alpha), float(beta), arr, nz, g._plan_ref(d), _p(row_mask), _p(col_mask), _peers_struct(peer_y, None, mc_y), _stream()), "spmm")
    return Y


def spmm_adam(g, X, P, M, V, scalars, alpha=1.0, beta=0.0, zs=None, Y=None, row_mask=None, col_mask=None, peer_p=None, mc_p=0, clear_z0=False):
    """K1 with the Adam epilogue: grad = alpha*(A@X) + beta*sum(zs); P,M,V updated in place.
    clear_z0: the non-zero rows of zs[0] (= G) are zeroed as they are read (its last reader in the step)."""
    lib = _lib.load()
    d = X.shape[1]
    for t, n in ((X, "X"), (P, "P"), (M, "M"), (V, "V")):
        _need(t, torch.float32, n, 2)
    arr, nz = _z_array(zs)
    blocks = g.block_plans(d) if col_mask is None else None
    if blocks:
        peers = _peers_struct(None, peer_p, 0, mc_p)
        for pl in blocks:
            _lib.check(lib.lgcn_spmm_adam_f32(_p(g.indptr), _p(g.indices), _p(g.vals), g.n_rows, d, _p(X), _p(Y),
                                              float(alpha), float(beta), arr, nz, _p(P), _p(M), _p(V), _p(scalars),
                                              byref(pl), _p(row_mask), None, peers, int(bool(clear_z0)), _stream()), "spmm_adam")
        return
    _lib.check(lib.lgcn_spmm_adam_f32(_p(g.indptr), _p(g.indices), _p(g.vals), g.n_rows, d, _p(X), _p(Y),
                                      float(alpha), float(beta), arr, nz, _p(P), _p(M), _p(V), _p(scalars),
                                      g._plan_ref(d), _p(row_mask), _p(col_mask), _peers_struct(None, peer_p, 0, mc_p), int(bool(clear_z0)), _stream()), "spmm_adam")


def gather_probe(X, idx, run=32, variant=0, out=None):
    """Measurement hook: sum the rows X[idx] in runs of `run` (lgcn_debug_gather_rows) -> float32 [ceil(n/run), d]."""
    _need(X, torch.float32, "X", 2), _need(idx, torch.int32, "idx", 1)
    n, d = idx.numel(), X.shape[1]
    groups = (n + run - 1) // run
    if out is None:
        out = torch.empty((groups, d), dtype=torch.float32, device=X.device)
    _lib.check(_lib.load().lgcn_debug_gather_rows(_p(X), _p(idx), n, d, int(run), int(variant), _p(out), _stream()), "debug_gather_rows")
    return out


def copy_words(dst, src, n_bytes=None, dst_is_host=False):
    """Kernel copy between device memory and mapped pinned host memory (lgcn_copy_words); capturable in a CUDA graph."""
    n = int(n_bytes if n_bytes is not None else src.numel() * src.element_size())
    _lib.check(_lib.load().lgcn_copy_words(c_void_p(dst.data_ptr()), c_void_p(src.data_ptr()), n, int(bool(dst_is_host)), _stream()), "copy_words")


def adam_scalars(device, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=0):
    s = torch.zeros(ctypes.sizeof(_lib.AdamScalars) // 4, dtype=torch.int32, device=device)
    _lib.check(_lib.load().lgcn_adam_init(_p(s), lr, beta1, beta2, eps, int(step), _stream()), "adam_init")
    return s


def adam_reinit(scalars, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=0):
    _lib.check(_lib.load().lgcn_adam_init(_p(scalars), lr, beta1, beta2, eps, int(step), _stream()), "adam_init")


def step_begin(scalars, B_cap, advance_ctl=None, stage_dst=None, stage_src=None):
    """Head of a captured step: Adam tick + (advance of the resident batch window | pull of a pinned host batch)."""
    n = 0 if stage_src is None else stage_src.numel() * stage_src.element_size()
    _lib.check(_lib.load().lgcn_step_begin(_p(scalars), _p(advance_ctl), int(B_cap), _p(stage_dst), _p(stage_src), n, _stream()), "step_begin")


def adam_tick(scalars):
    _lib.check(_lib.load().lgcn_adam_tick(_p(scalars), _stream()), "adam_tick")


def adam_step_count(scalars):
    return int(scalars[8].item())


def adam(P, M, V, G, scalars):
    _lib.check(_lib.load().lgcn_adam_f32(_p(P), _p(M), _p(V), _p(G), P.numel(), _p(scalars), _stream()), "adam")


def bpr_workspace(B_cap, d, device):
    nbytes = _lib.load().lgcn_bpr_workspace_bytes(B_cap, d)
    return torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)


def bpr_fwd_bwd(out, users, pos, neg, B_cap, ctl, n_users, m_items, inv_norm, decay, c_bpr, c_reg,
                loss_out, G, workspace, own=(0, None), deterministic=False, clear_mask=None, loss_host=None):
    """K2: loss_out[0..2] = (bpr, reg, bpr+decay*reg); G += closed-form gradient (if G is not None)."""
    lib = _lib.load()
    d = out.shape[1]
    _need(out, torch.float32, "out", 2)
    for t, n in ((users, "users"), (pos, "pos"), (neg, "neg")):
        _need(t, torch.int64, n, 1)
    _need(ctl, torch.int32, "ctl", 1), _need(loss_out, torch.float32, "loss_out", 1)
    own_end = out.shape[0] if own[1] is None else own[1]
    _lib.check(lib.lgcn_bpr_fwd_bwd(_p(out), _p(users), _p(pos), _p(neg), B_cap, _p(ctl), n_users, m_items, d,
                                    float(inv_norm), float(decay), float(c_bpr), float(c_reg), _p(loss_out), _p(G),
                                    int(own[0]), int(own_end), int(bool(deterministic)), _p(workspace),
                                    workspace.numel() * 4, _p(clear_mask), _p(loss_host), _stream()), "bpr_fwd_bwd")


class FeatExchange:
    """Cross-rank state of the feature partition's K2 (include/lgcn_b200.h, lgcn_bpr_feat_*): this rank's record buffer
    float32[2 * world * B_cap * 8], every rank's buffer as seen from this process, and the device barrier between the two
    halves.  `peer_records` / `peer_flags`: lists indexed by rank (own entries included) — CUDA-IPC mappings between
    processes, or plain tensors of one device when several ranks are emulated on one GPU (tests)."""

    def __init__(self, records, peer_records, flags, peer_flags, rank, world, B_cap, scalars, timeout_ms=20000):
        if records.numel() != 2 * world * B_cap * 8 or records.dtype != torch.float32:
            raise ValueError("FeatExchange: records must be float32[2 * world * B_cap * 8]")
        self.records, self.peer_records = records, list(peer_records)
        self.rank, self.world, self.B_cap = int(rank), int(world), int(B_cap)
        self.barrier = RankBarrier(flags, peer_flags, rank, world, timeout_ms=timeout_ms)
        self.feat = _lib.BprFeat()
        self.feat.n_parts, self.feat.part = self.world, self.rank
        self.feat.scalars_dev = scalars.data_ptr()
        self.feat.records_local = records.data_ptr()
        for q, t in enumerate(self.peer_records):
            self.feat.records_peer[q] = t.data_ptr()
        self._keep = scalars


def bpr_feat_partial(out, users, pos, neg, B_cap, ctl, n_users, m_items, xchg, workspace):
    """First half of K2 under the feature partition: this rank's partial dot products into every rank's record buffer."""
    _need(out, torch.float32, "out", 2)
    _lib.check(_lib.load().lgcn_bpr_feat_partial(_p(out), _p(users), _p(pos), _p(neg), int(B_cap), _p(ctl), int(n_users), int(m_items), out.shape[1],
                                                 ctypes.byref(xchg.feat), _p(workspace), workspace.numel() * 4, _stream()), "bpr_feat_partial")


def bpr_feat_finish(out, users, pos, neg, B_cap, ctl, n_users, m_items, inv_norm, decay, c_bpr, c_reg, loss_out, G, xchg, workspace,
                    deterministic=False, clear_mask=None, loss_host=None):
    """Second half (after the barrier): K2 on the column slice with the dot products taken from the summed records."""
    _need(out, torch.float32, "out", 2)
    _lib.check(_lib.load().lgcn_bpr_feat_finish(_p(out), _p(users), _p(pos), _p(neg), int(B_cap), _p(ctl), int(n_users), int(m_items), out.shape[1],
                                                float(inv_norm), float(decay), float(c_bpr), float(c_reg), _p(loss_out), _p(G),
                                                int(bool(deterministic)), ctypes.byref(xchg.feat), _p(workspace), workspace.numel() * 4,
                                                _p(clear_mask), _p(loss_host), _stream()), "bpr_feat_finish")


def popgate_param_count(d, pop_hidden, gate_hidden):
    return int(_lib.load().lgcn_popgate_param_count(int(d), int(pop_hidden), int(gate_hidden)))


def popgate_fuse(out, n_users, m_items, item_pop, params, pop_hidden, gate_hidden, temperature=1.0, want_gate=False):
    """fused item table [m_items, d] (+ gates [m_items]) from the propagated table `out` [N, d]  (code/model.py:139-157)."""
    _need(out, torch.float32, "out", 2), _need(item_pop, torch.float32, "item_pop", 1), _need(params, torch.float32, "params", 1)
    d = out.shape[1]
    fused = torch.empty((m_items, d), dtype=torch.float32, device=out.device)
    gate = torch.empty(m_items, dtype=torch.float32, device=out.device) if want_gate else None
    _lib.check(_lib.load().lgcn_popgate_fuse(_p(out), int(n_users), int(m_items), d, _p(item_pop), _p(params), int(pop_hidden), int(gate_hidden),
                                             float(temperature), _p(fused), _p(gate), _stream()), "popgate_fuse")
    return (fused, gate) if want_gate else fused


def popgate_workspace(B_cap, device):
    nbytes = _lib.load().lgcn_popgate_bpr_workspace_bytes(B_cap)
    return torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)


def popgate_bpr_fwd_bwd(out, users, pos, neg, B_cap, ctl, n_users, m_items, item_pop, params, pop_hidden, gate_hidden, temperature,
                        entropy_coeff, decay, loss_out, G, params_grad, workspace):
    """K2 of the pop-gate variant: loss_out = {bpr - coeff*entropy, reg, total, running}; G += d total/d out; params_grad += d total/d MLPs."""
    _need(out, torch.float32, "out", 2)
    _lib.check(_lib.load().lgcn_popgate_bpr_fwd_bwd(_p(out), _p(users), _p(pos), _p(neg), B_cap, _p(ctl), int(n_users), int(m_items), out.shape[1],
                                                    _p(item_pop), _p(params), int(pop_hidden), int(gate_hidden), float(temperature),
                                                    float(entropy_coeff), float(decay), _p(loss_out), _p(G), _p(params_grad),
                                                    _p(workspace), workspace.numel() * 4, _stream()), "popgate_bpr_fwd_bwd")


def bpr_clear_rows(G, users, pos, neg, B_cap, ctl, n_users):
    _lib.check(_lib.load().lgcn_bpr_clear_rows(_p(G), _p(users), _p(pos), _p(neg), B_cap, _p(ctl), n_users,
                                               G.shape[1], _stream()), "bpr_clear_rows")


def batch_masks(users, pos, neg, B_cap, ctl, n_users, g, m0, m1, clear_first=True):
    """Bitmaps of the rows a step needs: m0 = batch rows, m1 (optional) = m0 + their neighbours.
    clear_first=False: m0 is known to be all-zero (the previous step's K2 cleared its bits)."""
    _lib.check(_lib.load().lgcn_batch_masks(_p(users), _p(pos), _p(neg), B_cap, _p(ctl), n_users, g.n_rows, _p(g.indptr),
                                            _p(g.indices), _p(m0), _p(m1), int(bool(clear_first)), _stream()), "batch_masks")


def batch_masks_rows(users, pos, neg, B_cap, ctl, n_users, row_begin, row_end, m0_local):
    """Bitmap of the batch rows inside [row_begin, row_end), indexed by local row."""
    _lib.check(_lib.load().lgcn_batch_masks_rows(_p(users), _p(pos), _p(neg), B_cap, _p(ctl), n_users, int(row_begin), int(row_end),
                                                 _p(m0_local), _stream()), "batch_masks_rows")


def batch_advance(ctl, B_cap):
    _lib.check(_lib.load().lgcn_batch_advance(_p(ctl), B_cap, _stream()), "batch_advance")


def score_topk(users_emb, items_emb, users, k, mask_indptr=None, mask_indices=None, mask_col_offset=0):
    """K3: per-row top-k item ids/scores with train items masked to -1024 (lowest id wins ties)."""
    lib = _lib.load()
    _need(users_emb, torch.float32, "users_emb", 2), _need(items_emb, torch.float32, "items_emb", 2)
    Bt = users_emb.shape[0] if users is None else users.numel()
    m_items, d = items_emb.shape
    dev = items_emb.device
    if users is not None:
        _need(users, torch.int64, "users", 1)
    idx = torch.empty((Bt, k), dtype=torch.int64, device=dev)
    val = torch.empty((Bt, k), dtype=torch.float32, device=dev)
    if Bt == 0:
        return idx, val
    ws_bytes = lib.lgcn_score_topk_workspace_bytes(Bt, m_items, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.lgcn_score_topk(_p(users_emb), _p(items_emb), _p(users), Bt, m_items, d, _p(mask_indptr),
                                   _p(mask_indices), int(mask_col_offset), int(k), _p(idx), _p(val), _p(ws),
                                   ws_bytes, _stream()), "score_topk")
    return idx, val


last_tc_flags = None
last_tc_workspace = None
TC_MIN_ITEMS = 16384      # the tensor-core path needs >= 128 item tiles (interleaved layout, row threshold from tile maxima)
_tc_mask_cache = {}


def tc_mask_positions(mask_indptr, mask_indices, mask_col_offset, n_users, m_items):
    """The train-item mask in the position space of the tensor-core kernel: a CSR over the user rows whose column
    entries are n_users + position, ascending per row.  Built once per (mask, m_items) with the library's own
    kernels (lgcn_score_topk_tc_item_positions + lgcn_csr_build over (user, position) pairs) and cached."""
    key = (mask_indptr.data_ptr(), mask_indices.data_ptr(), int(mask_col_offset), int(n_users), int(m_items))
    hit = _tc_mask_cache.get(key)
    if hit is not None and hit[0]() is mask_indices:
        return hit[1], hit[2]
    lib = _lib.load()
    deg = (mask_indptr[1:n_users + 1] - mask_indptr[:n_users]).to(torch.int64)
    nnz = int(mask_indptr[n_users].item()) - int(mask_indptr[0].item())
    rows = torch.repeat_interleave(torch.arange(n_users, device=mask_indices.device, dtype=torch.int64), deg, output_size=nnz)
    start = int(mask_indptr[0].item())
    items = (mask_indices[start:start + nnz].to(torch.int64) - int(mask_col_offset)).contiguous()
    pos = torch.empty_like(items)
    _lib.check(lib.lgcn_score_topk_tc_item_positions(_p(items), nnz, int(m_items), _p(pos), _stream()), "score_topk_tc_item_positions")
    g = csr_build(rows, pos, int(n_users), int(lib.lgcn_score_topk_tc_position_space(int(m_items))))
    import weakref
    if len(_tc_mask_cache) > 8:
        _tc_mask_cache.clear()
    _tc_mask_cache[key] = (weakref.ref(mask_indices), g.indptr, g.indices)
    return g.indptr, g.indices


def score_topk_tc(users_emb, items_emb, users, k, mask_indptr=None, mask_indices=None, mask_col_offset=0, min_items=TC_MIN_ITEMS):
    """K3 on tensor cores: same result as score_topk, bit for bit.  Returns (idx, val, n_rows_redone);
    item tables smaller than `min_items` go to the exact kernel directly (every row counts as redone).
    The mask is the same id-space CSR score_topk takes; its position-space form is derived and cached."""
    lib = _lib.load()
    _need(users_emb, torch.float32, "users_emb", 2), _need(items_emb, torch.float32, "items_emb", 2)
    Bt = users_emb.shape[0] if users is None else users.numel()
    m_items, d = items_emb.shape
    if not lib.lgcn_score_topk_tc_supported(d, k) or Bt == 0 or m_items < max(min_items, TC_MIN_ITEMS):
        idx, val = score_topk(users_emb, items_emb, users, k, mask_indptr, mask_indices, mask_col_offset)
        return idx, val, Bt
    dev = items_emb.device
    if users is not None:
        _need(users, torch.int64, "users", 1)
    pm_indptr = pm_indices = None
    n_users = users_emb.shape[0]
    if mask_indptr is not None:
        pm_indptr, pm_indices = tc_mask_positions(mask_indptr, mask_indices, mask_col_offset, n_users, m_items)
    idx = torch.empty((Bt, k), dtype=torch.int64, device=dev)
    val = torch.empty((Bt, k), dtype=torch.float32, device=dev)
    flags = torch.empty(Bt, dtype=torch.int32, device=dev)
    n_flagged = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.lgcn_score_topk_tc_workspace_bytes(Bt, m_items, k)
    ws = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
    _lib.check(lib.lgcn_score_topk_tc(_p(users_emb), _p(items_emb), _p(users), Bt, m_items, d, _p(pm_indptr), _p(pm_indices),
                                      int(n_users) if pm_indptr is not None else 0, int(k), _p(idx), _p(val), _p(flags), _p(n_flagged),
                                      c_void_p(ws_ptr), ws_bytes, _stream()), "score_topk_tc")
    n = int(n_flagged.item())
    global last_tc_flags, last_tc_workspace
    # diagnostics (lgcn_score_topk_tc_debug_layout): the workspace is several hundred MB, so it is only kept on request
    last_tc_workspace = (ws, ws_ptr - ws.data_ptr(), Bt, m_items) if os.environ.get("LGCN_TC_KEEP_WORKSPACE") else None
    last_tc_flags = flags            # per row: 0 certified, 1 certificate failed, 2 fewer than k candidates, 3 candidate list overflowed
    if n:   # rows that could not be certified: exact path, scattered back
        rows = torch.nonzero(flags, as_tuple=False).flatten()
        sub = rows if users is None else users[rows].contiguous()
        idx2, val2 = score_topk(users_emb, items_emb, sub, k, mask_indptr, mask_indices, mask_col_offset)
        idx[rows] = idx2
        val[rows] = val2
    return idx, val, n


def score_dense(users_emb, items_emb, users):
    lib = _lib.load()
    _need(users_emb, torch.float32, "users_emb", 2), _need(items_emb, torch.float32, "items_emb", 2)
    Bt = users_emb.shape[0] if users is None else users.numel()
    m_items, d = items_emb.shape
    out = torch.empty((Bt, m_items), dtype=torch.float32, device=items_emb.device)
    if Bt:
        if users is not None:
            _need(users, torch.int64, "users", 1)
        _lib.check(lib.lgcn_score_dense(_p(users_emb), _p(items_emb), _p(users), Bt, m_items, d, _p(out), _stream()), "score_dense")
    return out


def score_dense_tc(users_emb, items_emb, users):
    """getUsersRating on tensor cores (3xTF32, fp32-class accuracy); d = 64 — other widths go to score_dense."""
    lib = _lib.load()
    _need(users_emb, torch.float32, "users_emb", 2), _need(items_emb, torch.float32, "items_emb", 2)
    Bt = users_emb.shape[0] if users is None else users.numel()
    m_items, d = items_emb.shape
    if not lib.lgcn_score_dense_tc_supported(d) or Bt == 0:
        return score_dense(users_emb, items_emb, users)
    if users is not None:
        _need(users, torch.int64, "users", 1)
    out = torch.empty((Bt, m_items), dtype=torch.float32, device=items_emb.device)
    ws_bytes = lib.lgcn_score_dense_tc_workspace_bytes(Bt, m_items)
    ws = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=items_emb.device)
    ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
    _lib.check(lib.lgcn_score_dense_tc(_p(users_emb), _p(items_emb), _p(users), Bt, m_items, d, _p(out), c_void_p(ws_ptr), ws_bytes, _stream()), "score_dense_tc")
    return out


def rank_metrics(topk_idx, test_indptr, test_indices, ks):
    """Sums over rows of precision/recall/ndcg at each k in ks -> float64 tensor [len(ks), 3]."""
    lib = _lib.load()
    _need(topk_idx, torch.int64, "topk_idx", 2)
    Bt, k_max = topk_idx.shape
    ks_t = torch.tensor(list(ks), dtype=torch.int32, device=topk_idx.device)
    sums = torch.zeros(len(ks) * 3, dtype=torch.float64, device=topk_idx.device)
    ws_bytes = lib.lgcn_rank_metrics_workspace_bytes(Bt, len(ks))
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=topk_idx.device)
    _lib.check(lib.lgcn_rank_metrics(_p(topk_idx), Bt, k_max, _p(_need(test_indptr, torch.int32, "test_indptr", 1)),
                                     _p(_need(test_indices, torch.int32, "test_indices", 1)), _p(ks_t), len(ks),
                                     _p(sums), _p(ws), ws_bytes, _stream()), "rank_metrics")
    return sums.view(len(ks), 3)


def sample_bpr(g, n_users, m_items, train_num, seed, epoch, out=None, check=True):
    """Device-side sampler + shuffle: int64 tensor [3, n] (users, pos, neg), n = (train_num // n_users) * n_users."""
    n = (int(train_num) // int(n_users)) * int(n_users)
    if out is None or out.shape[1] < n:
        out = torch.empty((3, max(n, 1)), dtype=torch.int64, device=g.device)
    S = out[:, :n]
    status = torch.zeros(1, dtype=torch.int32, device=g.device)
    _lib.check(_lib.load().lgcn_sample_bpr(_p(g.indptr), _p(g.indices), n_users, m_items, int(train_num), int(seed) & (2**64 - 1),
                                           int(epoch), _p(out[0]), _p(out[1]), _p(out[2]), _p(status), _stream()), "sample_bpr")
    if check:
        st = int(status.item())             # one sync per epoch
        if st & 1:
            raise RuntimeError("sample_bpr: a user has no train item (the reference's sampler cannot serve such a dataset either)")
        if st & 2:
            raise RuntimeError("sample_bpr: a user has no admissible negative item")
    return S
