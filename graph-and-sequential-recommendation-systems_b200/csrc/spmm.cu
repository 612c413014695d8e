// spmm.cu — K1: CSR SpMM over the normalised bipartite adjacency with fused epilogues.
//
// Replaces torch.sparse.mm(g, x) (reference code/model.py:216-218), the stack/mean layer
// combination (code/model.py:221-222), its autograd transpose product (code/utils.py:61; A_hat is
// symmetric so the same CSR is used) and, in the last backward layer, torch.optim.Adam.step()
// (code/utils.py:62).
//
// Schedule (round-1 measurements in profiles/README.md drove every choice below)
//   * WORK ITEMS.  A plan (lgcn_spmm_plan_*) turns the rows into items {row, start, end}: a row with at
//     most seg_len non-zeros is one item, a longer row is cut into equal segments.  Items are binned
//     by exact length and laid out in DESCENDING length (degree-binned load balancing): the groups
//     that share a warp/CTA do the same amount of work, the long items start first, and the item
//     descriptor is ONE coalesced 16-byte load instead of the row_order -> indptr -> indptr chain.
//   * one GROUP of LANES lanes owns one item; every lane holds d/4/LANES float4 accumulators
//     (d=64, LANES=8: a 256-B embedding row is two 16-byte loads per lane, four items per warp — half the
//     shuffles per gathered byte of the 16-lane layout, which is what the L1 data pipe was spending its time on);
//   * the item's (col,val) pairs are read LANES at a time with one coalesced streaming load each and
//     handed round the group with shuffles; lanes past the end re-read the last valid column with
//     weight 0, so the gather loop has no predicates; the next chunk is prefetched;
//   * UNROLL independent 16-byte gathers are in flight per lane before the first FMA; registers are
//     capped at 64 so that 32 warps/SM are resident (the kernel is latency-, not bandwidth-bound:
//     ncu showed L2 at 16 % of peak with 16 resident warps);
//   * segments write their partial row to plan.partials; the segment that arrives last (atomic
//     counter, self-resetting) adds the partials in part order — a fixed summation order — and runs
//     the epilogue.  No float atomics, one launch, results independent of scheduling.
//
// HBM roofline: B_spmm = 8*nnz + 4*(N+1) + 8*N*d bytes per layer.
#include "common.cuh"

namespace lgcn {

struct SpmmArgs {
    const int4* items; int n_items;            // sorted work items {row, start, end, seg_ref}; NULL -> rows of indptr
    const int4* seginfo;                       // {part, n_parts, slot_base, long_id}
    const int* indptr; const int* indices; const float* vals;
    const float4* X; float4* Y;
    float alpha, beta;
    int nz;
    const float4* z[LGCN_MAX_Z];
    int* counters; float4* partials;
    float4* P; float4* M; float4* V; const lgcn_adam_scalars_t* sc;   // adam epilogue
    const unsigned* row_mask;   // optional bitmap: items whose row bit is 0 are skipped (output not written)
    const unsigned* col_mask;   // optional bitmap: columns whose bit is 0 are known-zero rows of X (never read)
    // fused exchange (row partition over GPUs): every finished row is also stored into the peers' copies of Y
    // (and of P in the Adam epilogue) over NVLink, so no separate all-gather pass re-reads and re-sends it
    int n_peers; float4* peerY[LGCN_MAX_PEERS]; float4* peerP[LGCN_MAX_PEERS];
    int multicast;      // peerY[0] / peerP[0] are NVSwitch multimem addresses: ONE store reaches every replica
};

__device__ __forceinline__ bool mask_bit(const unsigned* m, int i) { return (__ldg(m + (i >> 5)) >> (i & 31)) & 1u; }

static int g_variant = 0;      // tuning variant of the d=64 kernels (lgcn_debug_spmm_variant, profiling hook)

__device__ __forceinline__ float4 gather_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int D, int LANES, int UNROLL>
__device__ __forceinline__ void accumulate_item(const SpmmArgs& a, int start, int end, int lane,
                                                unsigned gmask, float4 (&acc)[D / 4 / LANES]) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    static_assert(LANES % UNROLL == 0, "UNROLL must divide the group width");
    if (start >= end) return;
    int cnt = min(LANES, end - start);
    int j = start + min(lane, cnt - 1);                     // lanes past the end: last valid entry, weight 0
    int c_nxt = ld_stream_i32(a.indices + j);
    float v_nxt = lane < cnt ? ld_stream_f32(a.vals + j) : 0.f;
    for (int base = start; base < end; base += LANES) {
        const int c = c_nxt; const float v = v_nxt;
        const int cur = cnt;
        if (base + LANES < end) {
            cnt = min(LANES, end - base - LANES);
            j = base + LANES + min(lane, cnt - 1);
            c_nxt = ld_stream_i32(a.indices + j);
            v_nxt = lane < cnt ? ld_stream_f32(a.vals + j) : 0.f;
        }
#pragma unroll 1
        for (int t = 0; t < cur; t += UNROLL) {
            float4 x[UNROLL][VPL]; float w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(gmask, c, t + u, LANES);
                w[u] = __shfl_sync(gmask, v, t + u, LANES);
                const float4* src = a.X + (size_t)cc * VEC + lane;
#pragma unroll
                for (int p = 0; p < VPL; ++p) x[u][p] = gather_f4(src + p * LANES);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int p = 0; p < VPL; ++p) f4_fma(acc[p], w[u], x[u][p]);
        }
    }
}

// Same product when most rows of X are known to be zero (first backward layer: X = G has at most 3B
// non-zero rows): every lane tests the bitmap for its own column once per chunk, the group ballots, and
// only the surviving entries are gathered.  Skipped entries contribute exact zeros, so the result is
// bit-identical to the unmasked kernel.
template <int D, int LANES, int UNROLL>
__device__ __forceinline__ void accumulate_item_masked(const SpmmArgs& a, int start, int end, int lane,
                                                       unsigned gmask, int gshift, float4 (&acc)[D / 4 / LANES]) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    for (int base = start; base < end; base += LANES) {
        const int j = base + lane;
        int c = 0; float v = 0.f; bool keep = false;
        if (j < end) {
            c = ld_stream_i32(a.indices + j);
            keep = mask_bit(a.col_mask, c);
            if (keep) v = ld_stream_f32(a.vals + j);
        }
        unsigned km = __ballot_sync(gmask, keep) >> gshift;
        if (LANES < 32) km &= (1u << LANES) - 1u;
        while (km) {
            float4 x[UNROLL][VPL]; float w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const bool on = km != 0;
                const int idx = on ? __ffs(km) - 1 : 0;
                km &= km - 1;                                   // 0 stays 0
                const int cc = __shfl_sync(gmask, c, idx, LANES);
                const float ww = __shfl_sync(gmask, v, idx, LANES);
                w[u] = on ? ww : 0.f;
                const float4* src = a.X + (size_t)cc * VEC + lane;
#pragma unroll
                for (int p = 0; p < VPL; ++p) x[u][p] = on ? gather_f4(src + p * LANES) : f4_zero();
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int p = 0; p < VPL; ++p) f4_fma(acc[p], w[u], x[u][p]);
        }
    }
}

template <int D, int LANES, bool ADAM>
__device__ __forceinline__ void epilogue(const SpmmArgs& a, int row, int lane, const float4 (&acc)[D / 4 / LANES]) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
        const size_t off = (size_t)row * VEC + lane + p * LANES;
        float4 g = make_float4(a.alpha * acc[p].x, a.alpha * acc[p].y, a.alpha * acc[p].z, a.alpha * acc[p].w);
        if (a.nz > 0) {
            float4 zs = ld_once_f4(a.z[0] + off);
#pragma unroll 1
            for (int t = 1; t < a.nz; ++t) f4_add(zs, ld_once_f4(a.z[t] + off));
            f4_fma(g, a.beta, zs);
        }
        if constexpr (ADAM) {
            // torch.optim.Adam single-tensor arithmetic (betas/eps/step scalars live on the device)
            const float b2 = a.sc->beta2, eps = a.sc->eps, w1 = a.sc->w1, w2 = a.sc->w2;
            const float step_size = a.sc->step_size, bc2s = a.sc->bc2_sqrt;
            float4 pw = a.P[off], m = a.M[off], vv = a.V[off];
            adam_update1(pw.x, m.x, vv.x, g.x, b2, w1, w2, step_size, bc2s, eps);
            adam_update1(pw.y, m.y, vv.y, g.y, b2, w1, w2, step_size, bc2s, eps);
            adam_update1(pw.z, m.z, vv.z, g.z, b2, w1, w2, step_size, bc2s, eps);
            adam_update1(pw.w, m.w, vv.w, g.w, b2, w1, w2, step_size, bc2s, eps);
            a.P[off] = pw; a.M[off] = m; a.V[off] = vv;
            if (a.multicast) { if (a.peerP[0]) st_multicast_f4(a.peerP[0] + off, pw); }
            else for (int q = 0; q < a.n_peers; ++q) if (a.peerP[q]) st_stream_f4(a.peerP[q] + off, pw);
            if (a.Y != nullptr) {
                st_stream_f4(a.Y + off, g);
                if (a.multicast) { if (a.peerY[0]) st_multicast_f4(a.peerY[0] + off, g); }
                else for (int q = 0; q < a.n_peers; ++q) if (a.peerY[q]) st_stream_f4(a.peerY[q] + off, g);
            }
        } else {
            st_stream_f4(a.Y + off, g);
            if (a.multicast) st_multicast_f4(a.peerY[0] + off, g);
            else for (int q = 0; q < a.n_peers; ++q) st_stream_f4(a.peerY[q] + off, g);
        }
    }
}

template <int D, int LANES, int UNROLL, bool ADAM, int THREADS, int MINB, bool MASKED = false>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_kernel(const __grid_constant__ SpmmArgs a) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    constexpr int GROUPS = THREADS / LANES;
    static_assert(VEC % LANES == 0 && LANES <= 32 && VPL >= 1, "bad group width");
    const int lane = threadIdx.x % LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const long long gidx = (long long)blockIdx.x * GROUPS + threadIdx.x / LANES;
    if (gidx >= a.n_items) return;
    int row, start, end, seg_ref;
    if (a.items != nullptr) {
        const int4 it = __ldg(a.items + gidx);
        row = it.x; start = it.y; end = it.z; seg_ref = it.w;
    } else {
        row = (int)gidx; start = __ldg(a.indptr + row); end = __ldg(a.indptr + row + 1); seg_ref = -1;
    }
    if (a.row_mask != nullptr && !mask_bit(a.row_mask, row)) return;     // dead row: nobody reads it this step
    float4 acc[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
    if constexpr (MASKED) accumulate_item_masked<D, LANES, UNROLL>(a, start, end, lane, gmask, (threadIdx.x & 31) / LANES * LANES, acc);
    else accumulate_item<D, LANES, UNROLL>(a, start, end, lane, gmask, acc);
    if (seg_ref < 0) {
        epilogue<D, LANES, ADAM>(a, row, lane, acc);
        return;
    }
    const int4 si = __ldg(a.seginfo + seg_ref);
    const int part = si.x, n_parts = si.y, slot_base = si.z, long_id = si.w;
    float4* mine = a.partials + (size_t)(slot_base + part) * VEC + lane;
#pragma unroll
    for (int p = 0; p < VPL; ++p) mine[p * LANES] = acc[p];
    __threadfence();
    __syncwarp(gmask);
    int prev = 0;
    if (lane == 0) prev = atomicAdd(a.counters + long_id, 1);
    prev = __shfl_sync(gmask, prev, 0, LANES);
    if (prev != n_parts - 1) return;
    __threadfence();
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
    // fixed summation order (part 0, 1, 2, ...), but the loads of several parts are in flight together: with one part per
    // iteration every partial cost an L2 round trip, which made the tail of a hub row proportional to its part count
    constexpr int RB = (VPL >= 4) ? 2 : (VPL == 2 ? 4 : 8);    // 8 float4 of partials per lane in flight
    int q = 0;
#pragma unroll 1
    for (; q + RB <= n_parts; q += RB) {
        float4 t[RB][VPL];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const float4* src = a.partials + (size_t)(slot_base + q + u) * VEC + lane;
#pragma unroll
            for (int p = 0; p < VPL; ++p) t[u][p] = ld_cg_f4(src + p * LANES);
        }
#pragma unroll
        for (int u = 0; u < RB; ++u)
#pragma unroll
            for (int p = 0; p < VPL; ++p) f4_add(acc[p], t[u][p]);
    }
#pragma unroll 1
    for (; q < n_parts; ++q) {
        const float4* src = a.partials + (size_t)(slot_base + q) * VEC + lane;
#pragma unroll
        for (int p = 0; p < VPL; ++p) f4_add(acc[p], ld_cg_f4(src + p * LANES));
    }
    epilogue<D, LANES, ADAM>(a, row, lane, acc);
    if (lane == 0) a.counters[long_id] = 0;                  // ready for the next launch
}

// ---- plan kernels -------------------------------------------------------------------------
// ws ints: [0]=long cursor, [1]=segment cursor, then bins[seg_len+1], offsets[seg_len+1], cursors[seg_len+1]
// A hub row is cut into at most kMaxParts segments.  One 8-lane group walks a segment (~0.25 us per non-zero with four
// gathers in flight) and the last-arriving segment adds the partials in part order (~0.08 us each, 8 loads in flight):
// for a 1.5 M-nnz item row 2048 parts of 732 cost ~0.18 + 0.15 ms; 256 parts of 5.9 k cost 1.5 ms (the tail that made a
// row-partitioned layer 3x slower than its local SpMM), 12 k parts of 128 cost ~1 ms of serial adds.
constexpr int kMaxParts = 2048;

__device__ __forceinline__ void split_row(int deg, int seg_len, int& n_parts, int& len) {
    n_parts = (deg + seg_len - 1) / seg_len;
    if (n_parts > kMaxParts) n_parts = kMaxParts;
    len = (deg + n_parts - 1) / n_parts;                     // equal-length parts, never empty
    n_parts = (deg + len - 1) / len;
}

// counts: {n_long, n_segs, longest item}
__global__ void plan_count_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, int* counts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int deg = indptr[r + 1] - indptr[r];
    if (deg > seg_len) {
        int n_parts, len; split_row(deg, seg_len, n_parts, len);
        atomicAdd(counts + 0, 1);
        atomicAdd(counts + 1, n_parts);
        atomicMax(counts + 2, len);
    } else {
        atomicMax(counts + 2, deg);
    }
}

__global__ void plan_hist_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, int* bins) {   // bins[0..max_len]
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int deg = indptr[r + 1] - indptr[r];
    if (deg <= seg_len) { atomicAdd(bins + deg, 1); return; }
    int n_parts, len; split_row(deg, seg_len, n_parts, len);
    for (int p = 0; p < n_parts; ++p) {
        const int l = min(len, deg - p * len);
        atomicAdd(bins + l, 1);
    }
}

// offsets[l] = number of items longer than l  (descending-length layout)
__global__ void plan_offsets_kernel(const int* __restrict__ bins, int max_len, int* offsets) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int run = 0;
    for (int l = max_len; l >= 0; --l) { offsets[l] = run; run += bins[l]; }
}

__global__ void plan_scatter_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, const int* __restrict__ offsets,
                                    int* cursors, int* long_cursor, int4* items, int4* seginfo) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int s = indptr[r], deg = indptr[r + 1] - s;
    if (deg <= seg_len) {
        const int pos = offsets[deg] + atomicAdd(cursors + deg, 1);
        items[pos] = make_int4(r, s, s + deg, -1);
        return;
    }
    int n_parts, len; split_row(deg, seg_len, n_parts, len);
    const int long_id = atomicAdd(long_cursor + 0, 1);
    const int slot_base = atomicAdd(long_cursor + 1, n_parts);
    for (int p = 0; p < n_parts; ++p) {
        const int b = s + p * len, l = min(len, deg - p * len);
        seginfo[slot_base + p] = make_int4(p, n_parts, slot_base, long_id);
        const int pos = offsets[l] + atomicAdd(cursors + l, 1);
        items[pos] = make_int4(r, b, b + l, slot_base + p);
    }
}

template <int D, int LANES, int UNROLL, bool ADAM, int THREADS, int MINB>
static int launch_cfg(const SpmmArgs& a, cudaStream_t st) {
    constexpr int GROUPS = THREADS / LANES;
    if (a.n_items == 0) return 0;
    const long long blocks = ((long long)a.n_items + GROUPS - 1) / GROUPS;
    if (blocks > 0x7fffffffLL) return fail("spmm: grid too large");
    if (a.col_mask != nullptr) {
        constexpr int UM = UNROLL < 4 ? UNROLL : 4;
        spmm_kernel<D, LANES, UM, ADAM, THREADS, MINB, true><<<(unsigned)blocks, THREADS, 0, st>>>(a);
        LGCN_CHECK_LAUNCH("spmm_kernel<masked>");
        return 0;
    }
    spmm_kernel<D, LANES, UNROLL, ADAM, THREADS, MINB><<<(unsigned)blocks, THREADS, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("spmm_kernel");
    return 0;
}

template <bool ADAM>
static int dispatch_spmm(int d, const SpmmArgs& a, cudaStream_t st) {
    switch (d) {
        case 16:  return launch_cfg<16, 4, 4, ADAM, 128, 8>(a, st);
        case 32:  return launch_cfg<32, 4, 4, ADAM, 128, 8>(a, st);
        case 64:
            if (!ADAM) switch (g_variant) {                       // tuning variants, same results
                case 1:  return launch_cfg<64, 16, 8, false, 128, 8>(a, st);
                case 2:  return launch_cfg<64, 16, 4, false, 128, 8>(a, st);
                case 3:  return launch_cfg<64, 8, 4, false, 64, 16>(a, st);
                case 4:  return launch_cfg<64, 8, 4, false, 256, 4>(a, st);
                case 5:  return launch_cfg<64, 4, 2, false, 128, 8>(a, st);
                case 6:  return launch_cfg<64, 4, 4, false, 128, 4>(a, st);
                case 7:  return launch_cfg<64, 8, 2, false, 128, 10>(a, st);
                case 8:  return launch_cfg<64, 4, 2, false, 64, 16>(a, st);
                case 9:  return launch_cfg<64, 8, 8, false, 128, 4>(a, st);
                default: break;
            }
            return launch_cfg<64, 8, 4, ADAM, 128, 8>(a, st);
        case 128:
            if (!ADAM && g_variant == 1) return launch_cfg<128, 32, 8, false, 128, 8>(a, st);
            if (!ADAM && g_variant == 2) return launch_cfg<128, 8, 2, false, 128, 8>(a, st);
            return launch_cfg<128, 16, 4, ADAM, 128, 8>(a, st);
        case 256:
            if (!ADAM && g_variant == 1) return launch_cfg<256, 16, 2, false, 128, 8>(a, st);
            if (!ADAM && g_variant == 2) return launch_cfg<256, 32, 2, false, 128, 8>(a, st);
            return launch_cfg<256, 32, 4, ADAM, 128, 8>(a, st);
        default:  return fail("spmm: d=%d unsupported (16,32,64,128,256)", d);
    }
}

static int fill_args(SpmmArgs& a, const int32_t* indptr, const int32_t* indices, const float* vals,
                     int32_t n_rows, int32_t d, const float* X, float* Y, float alpha, float beta,
                     const float* const* z_host, int32_t nz, const lgcn_spmm_plan_t* plan,
                     const uint32_t* row_mask, const uint32_t* col_mask, const lgcn_spmm_peers_t* peers) {
    LGCN_CHECK_ARG(X, "spmm: null X");
    LGCN_CHECK_ARG(n_rows >= 0, "spmm: n_rows < 0");
    LGCN_CHECK_ARG(nz >= 0 && nz <= LGCN_MAX_Z, "spmm: nz=%d out of range (max %d)", nz, LGCN_MAX_Z);
    LGCN_CHECK_ARG(nz == 0 || z_host, "spmm: nz>0 but z_host is null");
    LGCN_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0, "spmm: X/Y must be 16-byte aligned");
    a.indptr = indptr; a.indices = indices; a.vals = vals;
    a.X = reinterpret_cast<const float4*>(X); a.Y = reinterpret_cast<float4*>(Y);
    a.alpha = alpha; a.beta = beta; a.nz = nz;
    for (int t = 0; t < LGCN_MAX_Z; ++t) {
        a.z[t] = (t < nz) ? reinterpret_cast<const float4*>(z_host[t]) : nullptr;
        LGCN_CHECK_ARG(t >= nz || (z_host[t] && ((uintptr_t)z_host[t] % 16) == 0), "spmm: z[%d] null or misaligned", t);
    }
    if (plan) {
        LGCN_CHECK_ARG(plan->n_items >= 0 && (plan->n_items == 0 || plan->items), "spmm: plan.items missing");
        LGCN_CHECK_ARG(((uintptr_t)plan->items % 16) == 0 && ((uintptr_t)plan->seginfo % 16) == 0, "spmm: plan arrays must be 16-byte aligned");
        LGCN_CHECK_ARG(plan->n_segs == 0 || (plan->seginfo && plan->counters && plan->partials), "spmm: plan segment buffers missing");
        LGCN_CHECK_ARG(plan->n_segs == 0 || plan->d_max >= d, "spmm: plan.d_max=%d < d=%d", plan->d_max, d);
        a.items = reinterpret_cast<const int4*>(plan->items); a.n_items = plan->n_items;
        a.seginfo = reinterpret_cast<const int4*>(plan->seginfo);
        a.counters = plan->counters; a.partials = reinterpret_cast<float4*>(plan->partials);
    } else {
        LGCN_CHECK_ARG(indptr || n_rows == 0, "spmm: indptr is null and no plan was given");
        a.items = nullptr; a.n_items = n_rows; a.seginfo = nullptr; a.counters = nullptr; a.partials = nullptr;
    }
    a.P = nullptr; a.M = nullptr; a.V = nullptr; a.sc = nullptr;
    a.row_mask = row_mask; a.col_mask = col_mask;
    a.n_peers = 0; a.multicast = 0;
    for (int q = 0; q < LGCN_MAX_PEERS; ++q) { a.peerY[q] = nullptr; a.peerP[q] = nullptr; }
    if (peers) {
        LGCN_CHECK_ARG(peers->n_peers >= 0 && peers->n_peers <= LGCN_MAX_PEERS, "spmm: n_peers=%d out of range", peers->n_peers);
        LGCN_CHECK_ARG(!peers->multicast || peers->n_peers == 1, "spmm: a multicast destination is ONE address (n_peers=%d)", peers->n_peers);
        a.n_peers = peers->n_peers; a.multicast = peers->multicast ? 1 : 0;
        for (int q = 0; q < a.n_peers; ++q) {
            LGCN_CHECK_ARG(((uintptr_t)peers->y[q] % 16) == 0 && ((uintptr_t)peers->p[q] % 16) == 0, "spmm: peer pointers must be 16-byte aligned");
            a.peerY[q] = reinterpret_cast<float4*>(peers->y[q]); a.peerP[q] = reinterpret_cast<float4*>(peers->p[q]);
        }
    }
    return 0;
}

// ---- L2 -> SM gather ceiling probe (bench.py `roofline_l2`) -------------------------------------------------
// The plainest kernel with K1's access pattern and none of its overheads: every LANES-lane group walks a contiguous run of
// row ids (coalesced index loads) and gathers the d-float rows they name, 16 bytes per lane, UNROLL rows in flight, summing
// them; one row per group is written at the end so nothing is optimised away.  No values, no shuffles beyond the index
// hand-out, no epilogue, perfectly balanced runs.  What it reaches on an L2-resident table is the bandwidth K1's gathers
// are bounded by on this box.
template <int D, int LANES, int UNROLL>
__global__ void __launch_bounds__(128, 8)
gather_probe_kernel(const float4* __restrict__ X, const int* __restrict__ idx, long long n_idx, int run, float4* __restrict__ out) {
    constexpr int VEC = D / 4, VPL = VEC / LANES, GROUPS = 128 / LANES;
    const int lane = threadIdx.x % LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const long long g = (long long)blockIdx.x * GROUPS + threadIdx.x / LANES;
    const long long begin = g * run;
    if (begin >= n_idx) return;
    const long long end = begin + run < n_idx ? begin + run : n_idx;
    float4 acc[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
    for (long long base = begin; base < end; base += LANES) {
        const long long j = base + lane < end ? base + lane : end - 1;
        const int c = ld_stream_i32(idx + j);
        const int cnt = (int)(end - base < LANES ? end - base : LANES);
#pragma unroll 1
        for (int t = 0; t < cnt; t += UNROLL) {
            float4 x[UNROLL][VPL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(gmask, c, (t + u) % LANES, LANES);
                const float4* src = X + (size_t)cc * VEC + lane;
#pragma unroll
                for (int p = 0; p < VPL; ++p) x[u][p] = gather_f4(src + p * LANES);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int p = 0; p < VPL; ++p) f4_add(acc[p], x[u][p]);
        }
    }
#pragma unroll
    for (int p = 0; p < VPL; ++p) out[(size_t)g * VEC + lane + p * LANES] = acc[p];
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_debug_gather_rows(const float* X, const int32_t* idx, int64_t n_idx, int32_t d, int32_t run,
                                      int32_t variant, float* out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(X && idx && out && n_idx > 0 && run > 0, "debug_gather_rows: bad arguments");
    LGCN_CHECK_ARG(d == 64, "debug_gather_rows: d=%d unsupported (64)", d);
    const long long groups = (n_idx + run - 1) / run;
    const float4* X4 = reinterpret_cast<const float4*>(X); float4* o4 = reinterpret_cast<float4*>(out);
    cudaStream_t st = as_stream(stream);
    switch (variant) {
        case 1:  gather_probe_kernel<64, 16, 8><<<(unsigned)((groups + 7) / 8), 128, 0, st>>>(X4, idx, n_idx, run, o4); break;
        case 2:  gather_probe_kernel<64, 8, 8><<<(unsigned)((groups + 15) / 16), 128, 0, st>>>(X4, idx, n_idx, run, o4); break;
        case 3:  gather_probe_kernel<64, 16, 4><<<(unsigned)((groups + 7) / 8), 128, 0, st>>>(X4, idx, n_idx, run, o4); break;
        default: gather_probe_kernel<64, 8, 4><<<(unsigned)((groups + 15) / 16), 128, 0, st>>>(X4, idx, n_idx, run, o4); break;
    }
    LGCN_CHECK_LAUNCH("gather_probe_kernel");
    return 0;
}

extern "C" int lgcn_debug_spmm_variant(int variant) { const int old = g_variant; g_variant = variant; return old; }

extern "C" int lgcn_spmm_plan_count(const int32_t* indptr, int32_t n_rows, int32_t seg_len,
                                    int32_t* counts_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && counts_out && seg_len > 0 && n_rows >= 0, "spmm_plan_count: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts_out, 0, 4 * sizeof(int32_t), st);
    if (n_rows > 0) plan_count_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(indptr, n_rows, seg_len, counts_out);
    LGCN_CHECK_LAUNCH("plan_count_kernel");
    return 0;
}

extern "C" size_t lgcn_spmm_plan_workspace_bytes(int32_t max_len) {
    if (max_len < 0) return 0;
    return sizeof(int32_t) * (4 + 3 * ((size_t)max_len + 1));
}

extern "C" int lgcn_spmm_plan_fill(const int32_t* indptr, int32_t n_rows, int32_t seg_len, int32_t max_len,
                                   int32_t* items_out, int32_t* seginfo_out,
                                   void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && items_out && seg_len > 0 && n_rows >= 0 && max_len >= 0, "spmm_plan_fill: bad arguments");
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_spmm_plan_workspace_bytes(max_len), "spmm_plan_fill: workspace too small");
    LGCN_CHECK_ARG(((uintptr_t)items_out % 16) == 0 && ((uintptr_t)seginfo_out % 16) == 0, "spmm_plan_fill: outputs must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    int* ws = static_cast<int*>(workspace);
    int* long_cursor = ws; int* bins = ws + 4; int* offsets = bins + max_len + 1; int* cursors = offsets + max_len + 1;
    cudaMemsetAsync(workspace, 0, lgcn_spmm_plan_workspace_bytes(max_len), st);
    if (n_rows == 0) return 0;
    const unsigned nb = (n_rows + 255) / 256;
    plan_hist_kernel<<<nb, 256, 0, st>>>(indptr, n_rows, seg_len, bins);
    LGCN_CHECK_LAUNCH("plan_hist_kernel");
    plan_offsets_kernel<<<1, 32, 0, st>>>(bins, max_len, offsets);
    LGCN_CHECK_LAUNCH("plan_offsets_kernel");
    plan_scatter_kernel<<<nb, 256, 0, st>>>(indptr, n_rows, seg_len, offsets, cursors, long_cursor,
                                            reinterpret_cast<int4*>(items_out), reinterpret_cast<int4*>(seginfo_out));
    LGCN_CHECK_LAUNCH("plan_scatter_kernel");
    return 0;
}

extern "C" int lgcn_spmm_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                             int32_t n_rows, int32_t d, const float* X, float* Y,
                             float alpha, float beta, const float* const* z_host, int32_t nz,
                             const lgcn_spmm_plan_t* plan_host, const uint32_t* row_mask, const uint32_t* col_mask,
                             const lgcn_spmm_peers_t* peers_host, lgcn_stream_t stream) {
    SpmmArgs a;
    if (int rc = fill_args(a, indptr, indices, vals, n_rows, d, X, Y, alpha, beta, z_host, nz, plan_host, row_mask, col_mask, peers_host)) return rc;
    LGCN_CHECK_ARG(Y, "spmm: Y is null");
    for (int q = 0; q < a.n_peers; ++q) LGCN_CHECK_ARG(a.peerY[q], "spmm: peer Y pointer %d is null", q);
    return dispatch_spmm<false>(d, a, as_stream(stream));
}

extern "C" int lgcn_spmm_adam_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                                  int32_t n_rows, int32_t d, const float* X, float* Y,
                                  float alpha, float beta, const float* const* z_host, int32_t nz,
                                  float* P, float* M, float* V, const lgcn_adam_scalars_t* scalars_dev,
                                  const lgcn_spmm_plan_t* plan_host, const uint32_t* row_mask, const uint32_t* col_mask,
                                  const lgcn_spmm_peers_t* peers_host, lgcn_stream_t stream) {
    SpmmArgs a;
    if (int rc = fill_args(a, indptr, indices, vals, n_rows, d, X, Y, alpha, beta, z_host, nz, plan_host, row_mask, col_mask, peers_host)) return rc;
    LGCN_CHECK_ARG(P && M && V && scalars_dev, "spmm_adam: null P/M/V/scalars");
    LGCN_CHECK_ARG(((uintptr_t)P % 16) == 0 && ((uintptr_t)M % 16) == 0 && ((uintptr_t)V % 16) == 0, "spmm_adam: P/M/V must be 16-byte aligned");
    a.P = reinterpret_cast<float4*>(P); a.M = reinterpret_cast<float4*>(M); a.V = reinterpret_cast<float4*>(V);
    a.sc = scalars_dev;
    return dispatch_spmm<true>(d, a, as_stream(stream));
}
