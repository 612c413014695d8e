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
//   * COLUMN-SLAB BLOCKING (graphs whose gathered table does not fit the 126 MB L2 — BASELINE config 5: N*d*4 = 3 GB).
//     Unblocked, every non-zero is a 256-byte gather from HBM: nnz*256 B = 256 GB per layer against 14 GB algorithmic
//     (18x amplification, measured 36 ms = HBM speed on the gathers).  Blocked: the columns are cut into slabs of
//     <= ~64 MB of X; the layer is one launch per slab over the (row, slab) segments (the row's non-zeros are sorted by
//     column, so a slab is a contiguous piece of the row, found by binary search at plan time); a segment starts from the
//     row's running sum (plan.acc, 256 B read) unless it is the row's first, and writes it back unless it is the row's
//     last, which runs the epilogue.  X is then read from HBM once per layer and the gathers hit L2; what is paid
//     instead is the running-sum traffic, 512 B per (row, slab) pair.  The per-row summation order is unchanged
//     (bit-identical to the unblocked kernel except for rows long enough to be segmented inside a slab).  Streams
//     (indices, values, descriptors, running sums) carry an L2 evict-first policy so that they do not displace the slab.
//
// HBM roofline: B_spmm = 8*nnz + 4*(N+1) + 8*N*d bytes per layer.
#include "common.cuh"
#include <stdlib.h>

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
    int hinted;         // a.indices is the hinted copy (bit 31 = hot column): spmm_kernel<..., HINTED>
    int clear_z0;       // Adam epilogue: z[0] (= G, the batch gradient: zero except <= 3B rows) is zeroed where it was non-zero — this
                        // launch is its last reader in the step, so no separate clear_rows kernel is needed
    // column-slab blocking: items carry flags (bit0 = not the row's first segment: start from acc[row]; bit1 = not its last:
    // store the running sum to acc[row] instead of running the epilogue); acc == NULL -> every item is a whole row
    float4* acc;
};

// L2 eviction policy for data that is touched once per launch (so that it does not displace the gathered slab)
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ld_f4_policy(const float4* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_f4_policy(float4* p, const float4& v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// Programmatic dependent launch (PDL): a K1 launched with the programmatic-serialization attribute may START while its
// predecessor is still draining — its CTAs read their work item and prefetch the first (col,val) chunk (static data: the
// plan and the CSR are never written inside a step), then wait here until the predecessor has completed and its writes are
// visible.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool mask_bit(const unsigned* m, int i) { return (__ldg(m + (i >> 5)) >> (i & 31)) & 1u; }

static int g_variant = 0;      // tuning variant of the d=64 kernels (lgcn_debug_spmm_variant, profiling hook)
static int g_pdl = -1;              // programmatic dependent launch of K1 (LGCN_PDL=0 disables, lgcn_debug_spmm_variant(200/201) toggles)
static bool pdl_enabled() {
    if (g_pdl < 0) { const char* e = getenv("LGCN_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
    return g_pdl == 1;
}
static int g_blocked_variant = 0;   // items per group of the column-blocked kernel: 0 -> 4 (shipped), 1 -> 1, 2 -> 8, 3 -> 2 (variant 100 + x)

__device__ __forceinline__ float4 gather_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ float4 gather_f4_policy(const float4* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ unsigned long long policy_evict_normal() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ int ld_stream_i32_policy(const int* p, unsigned long long pol) {
    int v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ float ld_stream_f32_policy(const float* p, unsigned long long pol) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol)); return v;
}

// HINT (gathered table larger than L2, BASELINE config 5): a.indices is the HINTED copy of the column indices — bit 31 set
// for a hot column (one of the highest-degree columns whose rows together fill ~half of L2).  Hot rows are gathered with an
// L2 evict-last policy, the others with evict-first, and so is the (col,val) stream: the rows that are re-read most stay
// resident instead of being displaced by rows that are read once (plain LRU keeps ~6 % of the gathers in L2 on config 5).
template <int D, int LANES, int UNROLL, bool HINT>
__device__ __forceinline__ void accumulate_item(const SpmmArgs& a, int start, int end, int lane,
                                                unsigned gmask, float4 (&acc)[D / 4 / LANES]) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    static_assert(LANES % UNROLL == 0, "UNROLL must divide the group width");
    if (start >= end) return;
    unsigned long long pol = 0, pol_hot = 0;
    if constexpr (HINT) { pol = policy_evict_first(); pol_hot = policy_evict_last(); }
    auto ld_c = [&](const int* p) { if constexpr (HINT) return ld_stream_i32_policy(p, pol); else return ld_stream_i32(p); };
    auto ld_v = [&](const float* p) { if constexpr (HINT) return ld_stream_f32_policy(p, pol); else return ld_stream_f32(p); };
    int cnt = min(LANES, end - start);
    int j = start + min(lane, cnt - 1);                     // lanes past the end: last valid entry, weight 0
    int c_nxt = ld_c(a.indices + j);
    float v_nxt = lane < cnt ? ld_v(a.vals + j) : 0.f;
    griddep_wait();                                         // X (and everything the epilogue touches) comes from the previous kernel
    for (int base = start; base < end; base += LANES) {
        const int c = c_nxt; const float v = v_nxt;
        const int cur = cnt;
        if (base + LANES < end) {
            cnt = min(LANES, end - base - LANES);
            j = base + LANES + min(lane, cnt - 1);
            c_nxt = ld_c(a.indices + j);
            v_nxt = lane < cnt ? ld_v(a.vals + j) : 0.f;
        }
#pragma unroll 1
        for (int t = 0; t < cur; t += UNROLL) {
            float4 x[UNROLL][VPL]; float w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(gmask, c, t + u, LANES);
                w[u] = __shfl_sync(gmask, v, t + u, LANES);
                if constexpr (HINT) {
                    const float4* src = a.X + (size_t)(cc & 0x7fffffff) * VEC + lane;
                    const unsigned long long pg = cc < 0 ? pol_hot : pol;
#pragma unroll
                    for (int p = 0; p < VPL; ++p) x[u][p] = gather_f4_policy(src + p * LANES, pg);
                } else {
                    const float4* src = a.X + (size_t)cc * VEC + lane;
#pragma unroll
                    for (int p = 0; p < VPL; ++p) x[u][p] = gather_f4(src + p * LANES);
                }
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
            if constexpr (ADAM) {
                if (a.clear_z0 && (zs.x != 0.f || zs.y != 0.f || zs.z != 0.f || zs.w != 0.f))
                    const_cast<float4*>(a.z[0])[off] = f4_zero();
            }
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

// Everything after an item's products are summed.  Whole-row item: epilogue.  Blocked plans (BLOCKED): a whole segment that is
// not its row's last stores the running sum to a.acc instead.  Part of a segmented (hub) row/segment: the partial goes to
// a.partials and the part that arrives last adds them in part order (after the running sum, if any) and finishes the row.
template <int D, int LANES, bool ADAM, bool BLOCKED>
__device__ __forceinline__ void finish_item(const SpmmArgs& a, int row, int lane, unsigned gmask, int seg_ref, float4 (&acc)[D / 4 / LANES]) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    if (seg_ref < 0) {
        if (BLOCKED && ((-1 - seg_ref) & 2)) {
            const unsigned long long pol = policy_evict_first();
#pragma unroll
            for (int p = 0; p < VPL; ++p) st_f4_policy(a.acc + (size_t)row * VEC + lane + p * LANES, acc[p], pol);
        } else {
            epilogue<D, LANES, ADAM>(a, row, lane, acc);
        }
        return;
    }
    const int4 si = __ldg(a.seginfo + seg_ref);
    const int part = si.x, n_parts = si.y, slot_base = si.z, long_id = si.w & 0x0fffffff, sflags = BLOCKED ? (int)((unsigned)si.w >> 28) : 0;
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
    if (sflags & 1) {                                                     // the row's running sum comes first, then the parts in order
#pragma unroll
        for (int p = 0; p < VPL; ++p) acc[p] = ld_cg_f4(a.acc + (size_t)row * VEC + lane + p * LANES);
    } else {
#pragma unroll
        for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
    }
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
    if (sflags & 2) {
#pragma unroll
        for (int p = 0; p < VPL; ++p) a.acc[(size_t)row * VEC + lane + p * LANES] = acc[p];
    } else {
        epilogue<D, LANES, ADAM>(a, row, lane, acc);
    }
    if (lane == 0) a.counters[long_id] = 0;                  // ready for the next launch
}

template <int D, int LANES, int UNROLL, bool ADAM, int THREADS, int MINB, bool MASKED = false, bool HINTED = false>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_kernel(const __grid_constant__ SpmmArgs a) {
    constexpr int VEC = D / 4, VPL = VEC / LANES;
    constexpr int GROUPS = THREADS / LANES;
    static_assert(VEC % LANES == 0 && LANES <= 32 && VPL >= 1, "bad group width");
    const int lane = threadIdx.x % LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    // Items are laid out by descending length.  With the fused exchange (finished rows are stored into every replica over
    // NVLink) taking them in that order bunches the stores at the end of the kernel — most rows are short — and the
    // launch ends with a drain tail (~1 ms of a 5.5 ms layer on BASELINE config 5, 8 GPUs).  Then the CTAs alternate
    // between the long end and the short end of the list, so rows complete at an even rate.
    long long cta = blockIdx.x;
    if (a.n_peers > 0) cta = (blockIdx.x & 1) ? (long long)gridDim.x - 1 - (blockIdx.x >> 1) : (long long)(blockIdx.x >> 1);
    const long long gidx = cta * GROUPS + threadIdx.x / LANES;
    if (gidx >= a.n_items) return;
    int row, start, end, seg_ref;
    if (a.items != nullptr) {
        const int4 it = __ldg(a.items + gidx);
        row = it.x; start = it.y; end = it.z; seg_ref = it.w;
    } else {
        row = (int)gidx; start = __ldg(a.indptr + row); end = __ldg(a.indptr + row + 1); seg_ref = -1;
    }
    if (a.row_mask != nullptr) {
        griddep_wait();                                                   // the bitmap was written earlier in this step
        if (!mask_bit(a.row_mask, row)) return;                           // dead row: nobody reads it this step
    }
    float4 acc[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
    if constexpr (MASKED) { griddep_wait(); accumulate_item_masked<D, LANES, UNROLL>(a, start, end, lane, gmask, (threadIdx.x & 31) / LANES * LANES, acc); }
    else accumulate_item<D, LANES, UNROLL, HINTED>(a, start, end, lane, gmask, acc);
    griddep_wait();                                                       // (empty items skip the wait inside accumulate_item)
    griddep_launch_dependents();
    finish_item<D, LANES, ADAM, false>(a, row, lane, gmask, seg_ref, acc);
}

// ---- narrow tables (d = 8 / 16 / 32: the column slices of the feature partition, dist_mode='featpart') ---------------------
// With a 32..128-byte row the lane layout of spmm_kernel (a group of lanes shares ONE row, the (col,val) pairs go round by
// shuffle) pays two shuffles, an address and a request per lane for 16 useful bytes: measured on the gowalla shape a d = 8
// layer took 29-49 us against 35 us at d = 64 — no faster for an eighth of the bytes
// (profiles/r2_feat_probe_row_per_group_kernel.jsonl).  Here a
// lane owns whole NON-ZEROS: the GL lanes of a group stride over the item's entries (their (col,val) loads are coalesced, no
// shuffles), each lane gathers the whole row of its entry with 256-bit loads (one request per 32 bytes: LDG.E.256) and keeps a
// full-width partial sum; the group's partial sums meet in a fixed xor tree at the end.  Deterministic, but the summation
// order differs from spmm_kernel's (entry j goes to lane j mod GL), so the two agree to rounding, not bit for bit.
// What bounds it (ncu, profiles/r2_k1_slice_widths_ncu_raw.csv): per-NON-ZERO costs, not bytes — the L1 data pipe is 65-74 % busy at every
// width (a wavefront per gathered row + the (col,val) stream), ~2 SM cycles per non-zero; wider slices (d = 16, 32) are served by
// spmm_kernel, whose cooperating lanes fetch a 64/128-byte row with one request.
__device__ __forceinline__ void ld_row256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

template <int D, int GL, int UNROLL, bool ADAM, int MINB>
__global__ void __launch_bounds__(128, MINB)
spmm_narrow_kernel(const __grid_constant__ SpmmArgs a) {
    constexpr int VEC = D / 4, GROUPS = 128 / GL;
    static_assert(VEC >= 2 && VEC <= GL && VEC % 2 == 0 && GL <= 32, "narrow kernel: 8 <= d <= 4 * GL, d a multiple of 8");
    const int lane = threadIdx.x % GL;
    const int gshift = (threadIdx.x & 31) / GL * GL;
    const unsigned gmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << gshift);
    const long long gidx = (long long)blockIdx.x * GROUPS + threadIdx.x / GL;
    if (gidx >= a.n_items) return;
    int row, start, end, seg_ref;
    if (a.items != nullptr) {
        const int4 it = __ldg(a.items + gidx);
        row = it.x; start = it.y; end = it.z; seg_ref = it.w;
    } else {
        row = (int)gidx; start = __ldg(a.indptr + row); end = __ldg(a.indptr + row + 1); seg_ref = -1;
    }
    if (a.row_mask != nullptr) {
        griddep_wait();
        if (!mask_bit(a.row_mask, row)) return;
    }
    float4 acc[VEC];
#pragma unroll
    for (int p = 0; p < VEC; ++p) acc[p] = f4_zero();
    // entries start + lane, + GL, + 2 GL, ...: UNROLL of them per round; entries past the end re-read the last one with weight 0
    int c[UNROLL]; float v[UNROLL];
    auto fetch = [&](int base) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int j = base + u * GL + lane;
            const bool ok = j < end;
            c[u] = ld_stream_i32(a.indices + (ok ? j : end - 1));
            v[u] = ok ? ld_stream_f32(a.vals + j) : 0.f;
        }
    };
    if (start < end) fetch(start);                          // static data: before the wait for the previous kernel
    griddep_wait();
    for (int base = start; base < end; base += GL * UNROLL) {
        float4 x[UNROLL][VEC]; float w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float4* src = a.X + (size_t)c[u] * VEC;
            w[u] = v[u];
#pragma unroll
            for (int p = 0; p < VEC; p += 2) ld_row256(src + p, x[u][p], x[u][p + 1]);
        }
        if (base + GL * UNROLL < end) fetch(base + GL * UNROLL);     // next round's (col,val) while the rows are in flight
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int p = 0; p < VEC; ++p) f4_fma(acc[p], w[u], x[u][p]);
    }
    griddep_launch_dependents();
#pragma unroll
    for (int o = GL >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int p = 0; p < VEC; ++p) {
            acc[p].x += __shfl_xor_sync(gmask, acc[p].x, o, GL); acc[p].y += __shfl_xor_sync(gmask, acc[p].y, o, GL);
            acc[p].z += __shfl_xor_sync(gmask, acc[p].z, o, GL); acc[p].w += __shfl_xor_sync(gmask, acc[p].w, o, GL);
        }
    }
    // the first VEC lanes of the group finish the row, one float4 each (finish_item / epilogue in their VEC-lane layout)
    if (lane >= VEC) return;
    float4 mine[1] = {acc[0]};
#pragma unroll
    for (int p = 1; p < VEC; ++p) if (lane == p) mine[0] = acc[p];
    const unsigned emask = ((1u << VEC) - 1u) << gshift;
    finish_item<D, VEC, ADAM, false>(a, row, lane, emask, seg_ref, mine);
}

// ---- column-slab blocked launches --------------------------------------------------------------------------------
// The items of a slab launch are SHORT (a row's entries inside one 64 MB slab of columns: ~3-40) and each one is a dependent
// chain  descriptor -> (columns, values, running sum) -> gathers -> store.  With one item per group (spmm_kernel) a slab launch
// is latency-bound: measured 4.3 TB/s of DRAM traffic on BASELINE config 5, no faster than the unblocked kernel although it
// moves 40 % fewer bytes (profiles/r2_blocked_v1_ncu_dram.csv).  Here a group walks IPG consecutive items (same length: the
// list is sorted by length) as a software pipeline: while item k is gathered, the descriptor of item k+2 and the first
// (col,val) chunk of item k+1 are already in flight.
template <int D, int LANES, int UNROLL, bool ADAM, int IPG>
__global__ void __launch_bounds__(128, 8)
spmm_blocked_kernel(const __grid_constant__ SpmmArgs a) {
    constexpr int VEC = D / 4, VPL = VEC / LANES, GROUPS = 128 / LANES;
    static_assert(VEC % LANES == 0 && LANES <= 32 && VPL >= 1 && LANES % UNROLL == 0, "bad group width");
    const int lane = threadIdx.x % LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const long long first = ((long long)blockIdx.x * GROUPS + threadIdx.x / LANES) * IPG;
    if (first >= a.n_items) return;
    const int n_mine = (int)min((long long)IPG, (long long)a.n_items - first);
    const unsigned long long pol = policy_evict_first();
    const int4* items = a.items + first;
    int4 it = __ldg(items);
    int4 it_n = n_mine > 1 ? __ldg(items + 1) : it;
    // first (col,val) chunk of item 0
    int cnt = min(LANES, it.z - it.y);
    int c = 0; float v = 0.f;
    if (cnt > 0) { const int j = it.y + min(lane, cnt - 1); c = ld_stream_i32_policy(a.indices + j, pol); v = lane < cnt ? ld_stream_f32_policy(a.vals + j, pol) : 0.f; }
#pragma unroll 1
    for (int k = 0; k < n_mine; ++k) {
        // stage A(k+2): descriptor;  stage B(k+1): first chunk of the next item
        const int4 it_nn = (k + 2 < n_mine) ? __ldg(items + k + 2) : it_n;
        int cnt_n = 0, c_n = 0; float v_n = 0.f;
        if (k + 1 < n_mine) {
            cnt_n = min(LANES, it_n.z - it_n.y);
            if (cnt_n > 0) { const int j = it_n.y + min(lane, cnt_n - 1); c_n = ld_stream_i32_policy(a.indices + j, pol); v_n = lane < cnt_n ? ld_stream_f32_policy(a.vals + j, pol) : 0.f; }
        }
        // stage C(k)
        const int row = it.x, start = it.y, end = it.z, seg_ref = it.w;
        if (a.row_mask == nullptr || mask_bit(a.row_mask, row)) {
            float4 acc[VPL];
            if (seg_ref < 0 && ((-1 - seg_ref) & 1)) {
#pragma unroll
                for (int p = 0; p < VPL; ++p) acc[p] = ld_f4_policy(a.acc + (size_t)row * VEC + lane + p * LANES, pol);
            } else {
#pragma unroll
                for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();
            }
            int cc = c; float vv = v; int cur = cnt;
            for (int base = start; base < end; base += LANES) {
                int c2 = 0; float v2 = 0.f; int cnt2 = 0;
                if (base + LANES < end) {                      // items longer than one chunk: in-item prefetch, as in spmm_kernel
                    cnt2 = min(LANES, end - base - LANES);
                    const int j = base + LANES + min(lane, cnt2 - 1);
                    c2 = ld_stream_i32_policy(a.indices + j, pol);
                    v2 = lane < cnt2 ? ld_stream_f32_policy(a.vals + j, pol) : 0.f;
                }
#pragma unroll 1
                for (int t = 0; t < cur; t += UNROLL) {
                    float4 x[UNROLL][VPL]; float w[UNROLL];
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        const int col = __shfl_sync(gmask, cc, t + u, LANES);
                        w[u] = __shfl_sync(gmask, vv, t + u, LANES);
                        const float4* src = a.X + (size_t)col * VEC + lane;
#pragma unroll
                        for (int p = 0; p < VPL; ++p) x[u][p] = gather_f4(src + p * LANES);
                    }
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                        for (int p = 0; p < VPL; ++p) f4_fma(acc[p], w[u], x[u][p]);
                }
                cc = c2; vv = v2; cur = cnt2;
            }
            finish_item<D, LANES, ADAM, true>(a, row, lane, gmask, seg_ref, acc);
        }
        it = it_n; it_n = it_nn; c = c_n; v = v_n; cnt = cnt_n;
    }
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

// Which piece of every row a plan covers: the whole row (indices == nullptr), or — column-slab blocking — the entries whose
// column lies in [col_lo, col_hi) (contiguous, because a row's entries are sorted by column).
struct Slab { const int* indices; int col_lo, col_hi, is_first; };

__device__ __forceinline__ int lower_bound_col(const int* __restrict__ idx, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(idx + mid) < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// false: row r has no item in this plan.  flags: bit0 = the segment is not the row's first, bit1 = not its last.
__device__ __forceinline__ bool slab_segment(const int* __restrict__ indptr, int r, const Slab& sl, int& s, int& e, int& flags) {
    const int rb = indptr[r], re = indptr[r + 1];
    if (sl.indices == nullptr) { s = rb; e = re; flags = 0; return true; }
    s = lower_bound_col(sl.indices, rb, re, sl.col_lo);
    e = lower_bound_col(sl.indices, s, re, sl.col_hi);
    flags = (s > rb ? 1 : 0) | (e < re ? 2 : 0);
    if (e > s) return true;
    return rb == re && sl.is_first;          // an empty row still needs its epilogue: once, in the first slab
}

// counts: {n_long, n_segs, longest item, rows with an item}
__global__ void plan_count_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, Slab sl, int* counts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int s, e, flags;
    if (!slab_segment(indptr, r, sl, s, e, flags)) return;
    atomicAdd(counts + 3, 1);
    const int deg = e - s;
    if (deg > seg_len) {
        int n_parts, len; split_row(deg, seg_len, n_parts, len);
        atomicAdd(counts + 0, 1);
        atomicAdd(counts + 1, n_parts);
        atomicMax(counts + 2, len);
    } else {
        atomicMax(counts + 2, deg);
    }
}

__global__ void plan_hist_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, Slab sl, int* bins) {   // bins[0..max_len]
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int s, e, flags;
    if (!slab_segment(indptr, r, sl, s, e, flags)) return;
    const int deg = e - s;
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

__global__ void plan_scatter_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, Slab sl, const int* __restrict__ offsets,
                                    int* cursors, int* long_cursor, int4* items, int4* seginfo) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int s, e, flags;
    if (!slab_segment(indptr, r, sl, s, e, flags)) return;
    const int deg = e - s;
    if (deg <= seg_len) {
        const int pos = offsets[deg] + atomicAdd(cursors + deg, 1);
        items[pos] = make_int4(r, s, s + deg, -1 - flags);
        return;
    }
    int n_parts, len; split_row(deg, seg_len, n_parts, len);
    const int long_id = atomicAdd(long_cursor + 0, 1);
    const int slot_base = atomicAdd(long_cursor + 1, n_parts);
    for (int p = 0; p < n_parts; ++p) {
        const int b = s + p * len, l = min(len, deg - p * len);
        seginfo[slot_base + p] = make_int4(p, n_parts, slot_base, long_id | (flags << 28));
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
    if (a.hinted) {
        constexpr int UH = UNROLL < 4 ? UNROLL : 4;
        spmm_kernel<D, LANES, UH, ADAM, THREADS, MINB, false, true><<<(unsigned)blocks, THREADS, 0, st>>>(a);
        LGCN_CHECK_LAUNCH("spmm_kernel<hinted>");
        return 0;
    }
    if (pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, spmm_kernel<D, LANES, UNROLL, ADAM, THREADS, MINB>, a);
        if (e != cudaSuccess) return fail("spmm_kernel (programmatic dependent launch): %s", cudaGetErrorString(e));
        return 0;
    }
    spmm_kernel<D, LANES, UNROLL, ADAM, THREADS, MINB><<<(unsigned)blocks, THREADS, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("spmm_kernel");
    return 0;
}

template <int D, int GL, int UNROLL, bool ADAM, int MINB>
static int launch_narrow(const SpmmArgs& a, cudaStream_t st) {
    constexpr int GROUPS = 128 / GL;
    if (a.n_items == 0) return 0;
    const long long blocks = ((long long)a.n_items + GROUPS - 1) / GROUPS;
    if (blocks > 0x7fffffffLL) return fail("spmm: grid too large");
    if (pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, spmm_narrow_kernel<D, GL, UNROLL, ADAM, MINB>, a);
        if (e != cudaSuccess) return fail("spmm_narrow_kernel (programmatic dependent launch): %s", cudaGetErrorString(e));
        return 0;
    }
    spmm_narrow_kernel<D, GL, UNROLL, ADAM, MINB><<<(unsigned)blocks, 128, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("spmm_narrow_kernel");
    return 0;
}

template <int D, int LANES, int UNROLL, bool ADAM>
static int launch_blocked(const SpmmArgs& a, cudaStream_t st) {
    if (a.n_items == 0) return 0;
    if (a.col_mask != nullptr) return fail("spmm: col_mask is not supported with a column-blocked plan");
    auto grid = [&](int ipg) { return (unsigned)((((long long)a.n_items + ipg - 1) / ipg + (128 / LANES) - 1) / (128 / LANES)); };
    if (g_blocked_variant == 1) spmm_blocked_kernel<D, LANES, UNROLL, ADAM, 1><<<grid(1), 128, 0, st>>>(a);
    else if (g_blocked_variant == 2) spmm_blocked_kernel<D, LANES, UNROLL, ADAM, 8><<<grid(8), 128, 0, st>>>(a);
    else if (g_blocked_variant == 3) spmm_blocked_kernel<D, LANES, UNROLL, ADAM, 2><<<grid(2), 128, 0, st>>>(a);
    else spmm_blocked_kernel<D, LANES, UNROLL, ADAM, 4><<<grid(4), 128, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("spmm_blocked_kernel");
    return 0;
}

template <bool ADAM>
static int dispatch_spmm(int d, const SpmmArgs& a, cudaStream_t st) {
    if (a.acc != nullptr) switch (d) {            // column-slab blocked plan: the pipelined short-item kernel
        case 16:  return launch_blocked<16, 4, 2, ADAM>(a, st);
        case 32:  return launch_blocked<32, 4, 2, ADAM>(a, st);
        case 64:  return launch_blocked<64, 8, 2, ADAM>(a, st);
        case 128: return launch_blocked<128, 16, 2, ADAM>(a, st);
        case 256: return launch_blocked<256, 32, 2, ADAM>(a, st);
        default:  return fail("spmm: d=%d unsupported (16,32,64,128,256)", d);
    }
    // 32-byte rows (d = 8, the slices of an 8-way feature partition): a lane per non-zero (spmm_narrow_kernel) unless a column mask,
    // hints or a fused exchange ask for the row-per-group kernel (variant 50 forces that kernel: measurement).  Measured per layer
    // inside a captured graph (profiles/r2_feat_probe_k1_slice_widths.jsonl, gowalla / amazon-book shape): 13.8 / 34.3 us against
    // 25-39 / 47-56 us; at d = 16 and 32 the row-per-group kernel is the faster one (18.8 / 35.3 and 29.7 / 53.6 us) and is kept.
    if (d == 8 && a.col_mask == nullptr && !a.hinted && a.n_peers == 0 && g_variant != 50 && ((uintptr_t)a.X % 32) == 0)
        return launch_narrow<8, 8, 2, ADAM, 8>(a, st);
    switch (d) {
        case 8:   return launch_cfg<8, 2, 2, ADAM, 128, 8>(a, st);       // feature partition over 8 ranks: 32-byte rows
        case 16:  return launch_cfg<16, 4, 4, ADAM, 128, 8>(a, st);
        case 32:
            // 8 lanes x one float4: a 128-byte row is ONE request (4 lanes x 2 float4 made it two): 18.9 / 43.1 us per layer inside a
            // captured graph on the gowalla / amazon-book shape against 29.8 / 54.3 (profiles/r2_feat_probe_k1_d32_lane_configs.jsonl)
            if (!ADAM && g_variant == 51) return launch_cfg<32, 8, 4, false, 128, 8>(a, st);
            if (!ADAM && g_variant == 54) return launch_cfg<32, 4, 4, false, 128, 8>(a, st);
            return launch_cfg<32, 8, 8, ADAM, 128, 8>(a, st);
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
        default:  return fail("spmm: d=%d unsupported (8,16,32,64,128,256)", d);
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
    a.indptr = indptr; a.indices = indices; a.vals = vals; a.hinted = 0;
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
        LGCN_CHECK_ARG(((uintptr_t)plan->acc % 16) == 0, "spmm: plan.acc must be 16-byte aligned");
        a.acc = reinterpret_cast<float4*>(plan->acc);
        if (plan->hinted_indices) {
            LGCN_CHECK_ARG(!plan->acc && !col_mask, "spmm: hinted indices go with the whole-row plan only");
            a.indices = plan->hinted_indices; a.hinted = 1;
        }
    } else {
        LGCN_CHECK_ARG(indptr || n_rows == 0, "spmm: indptr is null and no plan was given");
        a.items = nullptr; a.n_items = n_rows; a.seginfo = nullptr; a.counters = nullptr; a.partials = nullptr; a.acc = nullptr;
    }
    a.P = nullptr; a.M = nullptr; a.V = nullptr; a.sc = nullptr; a.clear_z0 = 0;
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

// hinted copy of the column indices: bit 31 set where the column's weight (degree) reaches the threshold
__global__ void hint_indices_kernel(const int* __restrict__ indices, long long nnz, const int* __restrict__ col_weight, int threshold,
                                    int* __restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int c = indices[j];
    out[j] = (__ldg(col_weight + c) >= threshold) ? (c | (int)0x80000000u) : c;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_spmm_hint_indices(const int32_t* indices, int64_t nnz, const int32_t* col_weight, int32_t threshold,
                                      int32_t* hinted_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indices && col_weight && hinted_out && nnz >= 0, "spmm_hint_indices: bad arguments");
    if (nnz == 0) return 0;
    hint_indices_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, as_stream(stream)>>>(indices, nnz, col_weight, threshold, hinted_out);
    LGCN_CHECK_LAUNCH("hint_indices_kernel");
    return 0;
}

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

extern "C" int lgcn_debug_spmm_variant(int variant) {
    if (variant == 200 || variant == 201) { const int old = 200 + (pdl_enabled() ? 1 : 0); g_pdl = variant - 200; return old; }
    if (variant >= 100) { const int old = 100 + g_blocked_variant; g_blocked_variant = variant - 100; return old; }
    const int old = g_variant; g_variant = variant; return old;
}

static int plan_count_impl(const int32_t* indptr, int32_t n_rows, int32_t seg_len, const Slab& sl, int32_t* counts_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && counts_out && seg_len > 0 && n_rows >= 0, "spmm_plan_count: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts_out, 0, 4 * sizeof(int32_t), st);
    if (n_rows > 0) plan_count_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(indptr, n_rows, seg_len, sl, counts_out);
    LGCN_CHECK_LAUNCH("plan_count_kernel");
    return 0;
}

extern "C" int lgcn_spmm_plan_count(const int32_t* indptr, int32_t n_rows, int32_t seg_len,
                                    int32_t* counts_out, lgcn_stream_t stream) {
    return plan_count_impl(indptr, n_rows, seg_len, Slab{nullptr, 0, 0, 1}, counts_out, stream);
}

extern "C" int lgcn_spmm_plan_count_slab(const int32_t* indptr, const int32_t* indices, int32_t n_rows, int32_t seg_len,
                                         int32_t col_lo, int32_t col_hi, int32_t is_first_slab, int32_t* counts_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indices && col_lo <= col_hi, "spmm_plan_count_slab: bad arguments");
    return plan_count_impl(indptr, n_rows, seg_len, Slab{indices, col_lo, col_hi, is_first_slab ? 1 : 0}, counts_out, stream);
}

extern "C" size_t lgcn_spmm_plan_workspace_bytes(int32_t max_len) {
    if (max_len < 0) return 0;
    return sizeof(int32_t) * (4 + 3 * ((size_t)max_len + 1));
}

static int plan_fill_impl(const int32_t* indptr, int32_t n_rows, int32_t seg_len, int32_t max_len, const Slab& sl,
                          int32_t* items_out, int32_t* seginfo_out, void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && items_out && seg_len > 0 && n_rows >= 0 && max_len >= 0, "spmm_plan_fill: bad arguments");
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_spmm_plan_workspace_bytes(max_len), "spmm_plan_fill: workspace too small");
    LGCN_CHECK_ARG(((uintptr_t)items_out % 16) == 0 && ((uintptr_t)seginfo_out % 16) == 0, "spmm_plan_fill: outputs must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    int* ws = static_cast<int*>(workspace);
    int* long_cursor = ws; int* bins = ws + 4; int* offsets = bins + max_len + 1; int* cursors = offsets + max_len + 1;
    cudaMemsetAsync(workspace, 0, lgcn_spmm_plan_workspace_bytes(max_len), st);
    if (n_rows == 0) return 0;
    const unsigned nb = (n_rows + 255) / 256;
    plan_hist_kernel<<<nb, 256, 0, st>>>(indptr, n_rows, seg_len, sl, bins);
    LGCN_CHECK_LAUNCH("plan_hist_kernel");
    plan_offsets_kernel<<<1, 32, 0, st>>>(bins, max_len, offsets);
    LGCN_CHECK_LAUNCH("plan_offsets_kernel");
    plan_scatter_kernel<<<nb, 256, 0, st>>>(indptr, n_rows, seg_len, sl, offsets, cursors, long_cursor,
                                            reinterpret_cast<int4*>(items_out), reinterpret_cast<int4*>(seginfo_out));
    LGCN_CHECK_LAUNCH("plan_scatter_kernel");
    return 0;
}

extern "C" int lgcn_spmm_plan_fill(const int32_t* indptr, int32_t n_rows, int32_t seg_len, int32_t max_len,
                                   int32_t* items_out, int32_t* seginfo_out,
                                   void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    return plan_fill_impl(indptr, n_rows, seg_len, max_len, Slab{nullptr, 0, 0, 1}, items_out, seginfo_out, workspace, workspace_bytes, stream);
}

extern "C" int lgcn_spmm_plan_fill_slab(const int32_t* indptr, const int32_t* indices, int32_t n_rows, int32_t seg_len, int32_t max_len,
                                        int32_t col_lo, int32_t col_hi, int32_t is_first_slab,
                                        int32_t* items_out, int32_t* seginfo_out,
                                        void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indices && col_lo <= col_hi, "spmm_plan_fill_slab: bad arguments");
    return plan_fill_impl(indptr, n_rows, seg_len, max_len, Slab{indices, col_lo, col_hi, is_first_slab ? 1 : 0}, items_out, seginfo_out,
                          workspace, workspace_bytes, stream);
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
                                  const lgcn_spmm_peers_t* peers_host, int32_t clear_first_addend, lgcn_stream_t stream) {
    SpmmArgs a;
    if (int rc = fill_args(a, indptr, indices, vals, n_rows, d, X, Y, alpha, beta, z_host, nz, plan_host, row_mask, col_mask, peers_host)) return rc;
    LGCN_CHECK_ARG(P && M && V && scalars_dev, "spmm_adam: null P/M/V/scalars");
    LGCN_CHECK_ARG(!clear_first_addend || (nz >= 1 && row_mask == nullptr && X != z_host[0]), "spmm_adam: clear_first_addend needs an addend that is not the gathered table, and no row mask");
    a.clear_z0 = clear_first_addend ? 1 : 0;
    LGCN_CHECK_ARG(((uintptr_t)P % 16) == 0 && ((uintptr_t)M % 16) == 0 && ((uintptr_t)V % 16) == 0, "spmm_adam: P/M/V must be 16-byte aligned");
    a.P = reinterpret_cast<float4*>(P); a.M = reinterpret_cast<float4*>(M); a.V = reinterpret_cast<float4*>(V);
    a.sc = scalars_dev;
    return dispatch_spmm<true>(d, a, as_stream(stream));
}
