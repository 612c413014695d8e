// spmm.cu — K1: CSR SpMM over the normalised bipartite adjacency with fused epilogues.
//
// Replaces torch.sparse.mm(g, x) (reference code/model.py:216-218), the stack/mean layer
// combination (code/model.py:221-222), its autograd transpose product (code/utils.py:61; A_hat is
// symmetric so the same CSR is used) and, in the last backward layer, torch.optim.Adam.step()
// (code/utils.py:62).
//
// Schedule
//   * one GROUP of LANES = min(32, d/4) lanes owns one output row; every lane holds d/4/LANES
//     float4 accumulators, so one embedding row (d=64: 256 B) is one 16-byte load per lane;
//   * the row's (col,val) pairs are read LANES at a time with one coalesced streaming load each and
//     handed round the group with shuffles (register staging of the row segment); the next chunk is
//     prefetched while the current one is being gathered;
//   * UNROLL independent 16-byte gathers are in flight per lane before the first FMA;
//   * degree binning: rows with more than plan.seg_len non-zeros are cut into equal segments that
//     are scheduled as independent groups (appended after the n_rows short-row groups).  Each
//     segment writes its partial row to plan.partials; the segment that arrives last (atomic
//     counter) adds the partials in part order — a fixed summation order — and runs the epilogue.
//
// HBM roofline: B_spmm = 8*nnz + 4*(N+1) + 8*N*d bytes per layer.
#include "common.cuh"

namespace lgcn {

struct SpmmArgs {
    const int* indptr; const int* indices; const float* vals;
    int n_rows;
    const float4* X; float4* Y;
    float alpha, beta;
    int nz;
    const float4* z[LGCN_MAX_Z];
    // plan
    int seg_len; int n_segs;
    const int4* segs; int* counters; float4* partials; const int* row_order;
    // adam epilogue
    float4* P; float4* M; float4* V; const lgcn_adam_scalars_t* sc;
};

template <int D> struct Geo {
    static constexpr int VEC = D / 4;
    static constexpr int LANES = VEC < 32 ? VEC : 32;
    static constexpr int VPL = VEC / LANES;
    static_assert(D % 16 == 0 && VPL >= 1, "d must be a multiple of 16");
};

constexpr int kThreads = 256;

template <int D, int UNROLL>
__device__ __forceinline__ void accumulate_segment(const SpmmArgs& a, int start, int end, int lane,
                                                   unsigned gmask, float4 (&acc)[Geo<D>::VPL]) {
    constexpr int LANES = Geo<D>::LANES, VPL = Geo<D>::VPL, VEC = Geo<D>::VEC;
    static_assert(LANES % UNROLL == 0, "UNROLL must divide the group width");
    int c_nxt = 0; float v_nxt = 0.f;
    if (start + lane < end) { c_nxt = ld_stream_i32(a.indices + start + lane); v_nxt = ld_stream_f32(a.vals + start + lane); }
    for (int base = start; base < end; base += LANES) {
        const int c = c_nxt; const float v = v_nxt;
        const int jn = base + LANES + lane;
        c_nxt = 0; v_nxt = 0.f;
        if (jn < end) { c_nxt = ld_stream_i32(a.indices + jn); v_nxt = ld_stream_f32(a.vals + jn); }
        const int cnt = min(LANES, end - base);
        for (int t = 0; t < cnt; t += UNROLL) {
            float4 x[UNROLL][VPL]; float w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(gmask, c, t + u, LANES);
                w[u] = __shfl_sync(gmask, v, t + u, LANES);
                if (t + u < cnt) {
                    const float4* src = a.X + (size_t)cc * VEC + lane;
#pragma unroll
                    for (int p = 0; p < VPL; ++p) x[u][p] = ld_gather_f4(src + p * LANES);
                } else {
#pragma unroll
                    for (int p = 0; p < VPL; ++p) x[u][p] = f4_zero();
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int p = 0; p < VPL; ++p) f4_fma(acc[p], w[u], x[u][p]);
        }
    }
}

template <int D, bool ADAM>
__device__ __forceinline__ void epilogue(const SpmmArgs& a, int row, int lane, const float4 (&acc)[Geo<D>::VPL]) {
    constexpr int LANES = Geo<D>::LANES, VPL = Geo<D>::VPL, VEC = Geo<D>::VEC;
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
        const size_t off = (size_t)row * VEC + lane + p * LANES;
        float4 g = make_float4(a.alpha * acc[p].x, a.alpha * acc[p].y, a.alpha * acc[p].z, a.alpha * acc[p].w);
        if (a.nz > 0) {
            float4 zs = ld_once_f4(a.z[0] + off);
            for (int t = 1; t < a.nz; ++t) f4_add(zs, ld_once_f4(a.z[t] + off));
            f4_fma(g, a.beta, zs);
        }
        if constexpr (ADAM) {
            // torch.optim.Adam single-tensor arithmetic (betas/eps/step scalars live on the device)
            const float b1 = a.sc->beta1, b2 = a.sc->beta2, eps = a.sc->eps;
            const float step_size = a.sc->step_size, bc2s = a.sc->bc2_sqrt;
            float4 pw = a.P[off], m = a.M[off], vv = a.V[off];
            const float w1 = 1.f - b1, w2 = 1.f - b2;
#define LGCN_ADAM1(c) \
            m.c = m.c + w1 * (g.c - m.c); \
            vv.c = vv.c * b2 + w2 * g.c * g.c; \
            pw.c = pw.c - step_size * (m.c / (sqrtf(vv.c) / bc2s + eps));
            LGCN_ADAM1(x) LGCN_ADAM1(y) LGCN_ADAM1(z) LGCN_ADAM1(w)
#undef LGCN_ADAM1
            a.P[off] = pw; a.M[off] = m; a.V[off] = vv;
            if (a.Y != nullptr) st_stream_f4(a.Y + off, g);
        } else {
            st_stream_f4(a.Y + off, g);
        }
    }
}

template <int D, int UNROLL, bool ADAM>
__global__ void __launch_bounds__(kThreads)
spmm_kernel(const __grid_constant__ SpmmArgs a) {
    constexpr int LANES = Geo<D>::LANES, VPL = Geo<D>::VPL, VEC = Geo<D>::VEC;
    constexpr int GROUPS = kThreads / LANES;
    const int lane = threadIdx.x % LANES;
    const int group_in_cta = threadIdx.x / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const long long gidx = (long long)blockIdx.x * GROUPS + group_in_cta;

    float4 acc[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = f4_zero();

    if (gidx < a.n_rows) {
        const int row = a.row_order ? __ldg(a.row_order + gidx) : (int)gidx;
        const int s = __ldg(a.indptr + row), e = __ldg(a.indptr + row + 1);
        if (e - s > a.seg_len) return;                       // handled by its segments
        accumulate_segment<D, UNROLL>(a, s, e, lane, gmask, acc);
        epilogue<D, ADAM>(a, row, lane, acc);
        return;
    }
    const long long seg = gidx - a.n_rows;
    if (seg >= a.n_segs) return;
    const int4 s0 = __ldg(a.segs + 2 * seg), s1 = __ldg(a.segs + 2 * seg + 1);
    const int row = s0.x, start = s0.y, end = s0.z, part = s0.w;
    const int n_parts = s1.x, slot_base = s1.y, long_id = s1.z;
    accumulate_segment<D, UNROLL>(a, start, end, lane, gmask, acc);
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
    for (int q = 0; q < n_parts; ++q) {
        const float4* src = a.partials + (size_t)(slot_base + q) * VEC + lane;
#pragma unroll
        for (int p = 0; p < VPL; ++p) f4_add(acc[p], ld_cg_f4(src + p * LANES));
    }
    epilogue<D, ADAM>(a, row, lane, acc);
    if (lane == 0) a.counters[long_id] = 0;                  // ready for the next launch
}

// ---- plan kernels -------------------------------------------------------------------------
__global__ void plan_count_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, int* counts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int deg = indptr[r + 1] - indptr[r];
    if (deg > seg_len) {
        atomicAdd(counts + 0, 1);
        atomicAdd(counts + 1, (deg + seg_len - 1) / seg_len);
    }
}

__global__ void plan_fill_kernel(const int* __restrict__ indptr, int n_rows, int seg_len, int* segs, int* cursor) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int s = indptr[r], deg = indptr[r + 1] - s;
    if (deg <= seg_len) return;
    const int n_parts = (deg + seg_len - 1) / seg_len;
    const int len = (deg + n_parts - 1) / n_parts;          // equal-length parts
    const int long_id = atomicAdd(cursor + 0, 1);
    const int slot_base = atomicAdd(cursor + 1, n_parts);
    for (int p = 0; p < n_parts; ++p) {
        int* o = segs + (size_t)(slot_base + p) * 8;
        const int b = s + p * len;
        int e = b + len; if (e > s + deg) e = s + deg;
        o[0] = r; o[1] = b; o[2] = e; o[3] = p; o[4] = n_parts; o[5] = slot_base; o[6] = long_id; o[7] = 0;
    }
}

template <int D, bool ADAM>
static int launch_spmm(const SpmmArgs& a, cudaStream_t st) {
    constexpr int LANES = Geo<D>::LANES;
    constexpr int GROUPS = kThreads / LANES;
    constexpr int UNROLL = (LANES >= 8) ? 8 : 4;
    const long long groups = (long long)a.n_rows + a.n_segs;
    if (groups == 0) return 0;
    const long long blocks = (groups + GROUPS - 1) / GROUPS;
    if (blocks > 0x7fffffffLL) return fail("spmm: grid too large");
    spmm_kernel<D, UNROLL, ADAM><<<(unsigned)blocks, kThreads, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("spmm_kernel");
    return 0;
}

template <bool ADAM>
static int dispatch_spmm(int d, const SpmmArgs& a, cudaStream_t st) {
    switch (d) {
        case 16:  return launch_spmm<16, ADAM>(a, st);
        case 32:  return launch_spmm<32, ADAM>(a, st);
        case 64:  return launch_spmm<64, ADAM>(a, st);
        case 128: return launch_spmm<128, ADAM>(a, st);
        case 256: return launch_spmm<256, ADAM>(a, st);
        default:  return fail("spmm: d=%d unsupported (16,32,64,128,256)", d);
    }
}

static int fill_args(SpmmArgs& a, const int32_t* indptr, const int32_t* indices, const float* vals,
                     int32_t n_rows, int32_t d, const float* X, float* Y, float alpha, float beta,
                     const float* const* z_host, int32_t nz, const lgcn_spmm_plan_t* plan) {
    LGCN_CHECK_ARG(indptr && X, "spmm: null indptr/X");
    LGCN_CHECK_ARG(n_rows >= 0, "spmm: n_rows < 0");
    LGCN_CHECK_ARG(nz >= 0 && nz <= LGCN_MAX_Z, "spmm: nz=%d out of range (max %d)", nz, LGCN_MAX_Z);
    LGCN_CHECK_ARG(nz == 0 || z_host, "spmm: nz>0 but z_host is null");
    LGCN_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0, "spmm: X/Y must be 16-byte aligned");
    a.indptr = indptr; a.indices = indices; a.vals = vals; a.n_rows = n_rows;
    a.X = reinterpret_cast<const float4*>(X); a.Y = reinterpret_cast<float4*>(Y);
    a.alpha = alpha; a.beta = beta; a.nz = nz;
    for (int t = 0; t < LGCN_MAX_Z; ++t) {
        a.z[t] = (t < nz) ? reinterpret_cast<const float4*>(z_host[t]) : nullptr;
        LGCN_CHECK_ARG(t >= nz || (z_host[t] && ((uintptr_t)z_host[t] % 16) == 0), "spmm: z[%d] null or misaligned", t);
    }
    if (plan) {
        LGCN_CHECK_ARG(plan->seg_len > 0, "spmm: plan.seg_len must be > 0");
        LGCN_CHECK_ARG(plan->n_segs == 0 || (plan->segs && plan->counters && plan->partials), "spmm: plan buffers missing");
        LGCN_CHECK_ARG(plan->n_segs == 0 || plan->d_max >= d, "spmm: plan.d_max=%d < d=%d", plan->d_max, d);
        a.seg_len = plan->seg_len; a.n_segs = plan->n_segs;
        a.segs = reinterpret_cast<const int4*>(plan->segs); a.counters = plan->counters;
        a.partials = reinterpret_cast<float4*>(plan->partials); a.row_order = plan->row_order;
    } else {
        a.seg_len = 0x7fffffff; a.n_segs = 0; a.segs = nullptr; a.counters = nullptr; a.partials = nullptr; a.row_order = nullptr;
    }
    a.P = nullptr; a.M = nullptr; a.V = nullptr; a.sc = nullptr;
    return 0;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_spmm_plan_count(const int32_t* indptr, int32_t n_rows, int32_t seg_len,
                                    int32_t* counts_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && counts_out && seg_len > 0 && n_rows >= 0, "spmm_plan_count: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts_out, 0, 2 * sizeof(int32_t), st);
    if (n_rows > 0) plan_count_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(indptr, n_rows, seg_len, counts_out);
    LGCN_CHECK_LAUNCH("plan_count_kernel");
    return 0;
}

extern "C" int lgcn_spmm_plan_fill(const int32_t* indptr, int32_t n_rows, int32_t seg_len,
                                   int32_t* segs, int32_t* cursor, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && segs && cursor && seg_len > 0 && n_rows >= 0, "spmm_plan_fill: bad arguments");
    LGCN_CHECK_ARG(((uintptr_t)segs % 16) == 0, "spmm_plan_fill: segs must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (n_rows > 0) plan_fill_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(indptr, n_rows, seg_len, segs, cursor);
    LGCN_CHECK_LAUNCH("plan_fill_kernel");
    return 0;
}

extern "C" int lgcn_spmm_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                             int32_t n_rows, int32_t d, const float* X, float* Y,
                             float alpha, float beta, const float* const* z_host, int32_t nz,
                             const lgcn_spmm_plan_t* plan_host, lgcn_stream_t stream) {
    SpmmArgs a;
    if (int rc = fill_args(a, indptr, indices, vals, n_rows, d, X, Y, alpha, beta, z_host, nz, plan_host)) return rc;
    LGCN_CHECK_ARG(Y, "spmm: Y is null");
    return dispatch_spmm<false>(d, a, as_stream(stream));
}

extern "C" int lgcn_spmm_adam_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                                  int32_t n_rows, int32_t d, const float* X, float* Y,
                                  float alpha, float beta, const float* const* z_host, int32_t nz,
                                  float* P, float* M, float* V, const lgcn_adam_scalars_t* scalars_dev,
                                  const lgcn_spmm_plan_t* plan_host, lgcn_stream_t stream) {
    SpmmArgs a;
    if (int rc = fill_args(a, indptr, indices, vals, n_rows, d, X, Y, alpha, beta, z_host, nz, plan_host)) return rc;
    LGCN_CHECK_ARG(P && M && V && scalars_dev, "spmm_adam: null P/M/V/scalars");
    LGCN_CHECK_ARG(((uintptr_t)P % 16) == 0 && ((uintptr_t)M % 16) == 0 && ((uintptr_t)V % 16) == 0, "spmm_adam: P/M/V must be 16-byte aligned");
    a.P = reinterpret_cast<float4*>(P); a.M = reinterpret_cast<float4*>(M); a.V = reinterpret_cast<float4*>(V);
    a.sc = scalars_dev;
    return dispatch_spmm<true>(d, a, as_stream(stream));
}
