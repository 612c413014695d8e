// bpr.cu — K2: fused BPR forward + closed-form gradient scatter.
//
// Replaces, in one launch, the reference's getEmbedding gathers (code/model.py:125-134), the BPR and
// L2 arithmetic (code/model.py:162-173, on the PROPAGATED embeddings), the scaling by `decay`
// (code/utils.py:55-57) and the autograd backward of all of them (IndexBackward = index_put with
// accumulate, code/utils.py:61).
//
//   z_b   = <U_b,P_b> - <U_b,N_b>                      (pos_scores - neg_scores)
//   bpr   = (1/B) sum_b softplus(-z_b)                 (= -mean(logsigmoid(z)))
//   reg   = (1/B) sum_b 0.5 (|U_b|^2+|P_b|^2+|N_b|^2)
//   s_b   = sigmoid(-z_b)
//   G[u_b]      += c_bpr*s_b*(N_b-P_b)/B + c_reg*U_b/B
//   G[nu+p_b]   += -c_bpr*s_b*U_b/B      + c_reg*P_b/B
//   G[nu+n_b]   +=  c_bpr*s_b*U_b/B      + c_reg*N_b/B
//
// One group of LANES = min(32,d/4) lanes per triple, float4 row gathers, shuffle reductions, one
// vectorised red.global.add.v4.f32 per lane per row.  Loss/reg partial sums are reduced in a fixed
// order (per-CTA partials + last-arriving CTA), so the scalars are run-to-run deterministic; the
// deterministic mode also replaces the float atomics by an owner-computes row reduction.
#include "common.cuh"

namespace lgcn {

constexpr int kBprThreads = 256;

template <int D> struct BGeo {
    static constexpr int VEC = D / 4;
    static constexpr int LANES = VEC < 32 ? VEC : 32;
    static constexpr int VPL = VEC / LANES;
};

struct BprArgs {
    const float4* out; const long long* users; const long long* pos; const long long* neg;
    int B_cap; const int* ctl; int n_users; int m_items;
    float inv_norm, decay, c_bpr, c_reg;
    float* loss_out; float4* G; int own_begin, own_end;
    int* counter; float* partials; float* coef;   // workspace
    int write_coef; int scatter;
    unsigned* clear_mask;       // optional: the batch-row bitmap of lgcn_batch_masks — its bits are cleared here (its last reader
                                // ran before this kernel), so the next step's batch_masks needs no memset
    float* loss_host;           // optional: mapped pinned host float[4] that receives a copy of loss_out
    // feature partition (dist_mode='featpart': every rank holds d/P COLUMNS of all tables, `out` here is this rank's (N, D)
    // slice): the five dot products of a triple are sums over the ranks' partial dot products.  n_parts > 1:
    //   bpr_feat_partial_kernel  stores this rank's partials into slot `part` of EVERY rank's record buffer (peer stores),
    //   [lgcn_rank_barrier]
    //   bpr_kernel               replaces its local sums by the sum of the n_parts records in rank order — every rank gets
    //                            the same bits — and carries on: loss, coefficients, gradient rows of ITS columns.
    // Records are double-buffered by the parity of the Adam step counter (sc->step, device-resident): a rank can only
    // overwrite a buffer two steps later, i.e. after a barrier every rank reached having finished reading it.
    int n_parts, part;
    const lgcn_adam_scalars_t* sc;
    float4* xchg_local; float4* xchg_peer[LGCN_MAX_PEERS + 1];      // [2][n_parts][B_cap][2] float4: {pu,nu,uu,pp},{nn,0,0,0}
};

__device__ __forceinline__ size_t feat_record(const BprArgs& a, int part, int t) {
    return (((size_t)(a.sc->step & 1) * a.n_parts + part) * a.B_cap + t) * 2;
}

// 1/B of the means: explicit when > 0, else from the device-resident batch descriptor
// (ctl[3] = global batch size when the batch is sharded over ranks, else ctl[1])
__device__ __forceinline__ float eff_inv_norm(const BprArgs& a, int B) {
    if (a.inv_norm > 0.f) return a.inv_norm;
    const int g = a.ctl[3] > 0 ? a.ctl[3] : B;
    return 1.f / (float)(g > 0 ? g : 1);
}

__device__ __forceinline__ float group_sum(float v, unsigned gmask, int lanes) {
    for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, lanes);
    return v;
}

template <int D>
__global__ void __launch_bounds__(kBprThreads)
bpr_kernel(const __grid_constant__ BprArgs a) {
    constexpr int LANES = BGeo<D>::LANES, VPL = BGeo<D>::VPL, VEC = BGeo<D>::VEC;
    constexpr int GROUPS = kBprThreads / LANES;
    __shared__ float s_loss[GROUPS], s_reg[GROUPS];
    __shared__ int s_last;
    const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const int off = a.ctl[0];
    const int B = min(a.ctl[1], a.B_cap);
    const int t = blockIdx.x * GROUPS + grp;
    const float inv_norm = eff_inv_norm(a, B);
    float loss_t = 0.f, reg_t = 0.f;
    if (t < B) {
        const long long u = a.users[off + t], p = a.pos[off + t] + a.n_users, n = a.neg[off + t] + a.n_users;
        const float4* ur = a.out + (size_t)u * VEC + lane;
        const float4* pr = a.out + (size_t)p * VEC + lane;
        const float4* nr = a.out + (size_t)n * VEC + lane;
        float4 U[VPL], P[VPL], Nn[VPL];
        float pu = 0.f, nu = 0.f, uu = 0.f, pp = 0.f, nn = 0.f;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            U[q] = __ldg(ur + q * LANES); P[q] = __ldg(pr + q * LANES); Nn[q] = __ldg(nr + q * LANES);
            pu += f4_dot(U[q], P[q]); nu += f4_dot(U[q], Nn[q]);
            uu += f4_dot(U[q], U[q]); pp += f4_dot(P[q], P[q]); nn += f4_dot(Nn[q], Nn[q]);
        }
        if (a.n_parts > 1) {
            // the ranks' partial dot products, added in rank order (written by peers: L2-coherent loads, not the L1)
            pu = 0.f; nu = 0.f; uu = 0.f; pp = 0.f; nn = 0.f;
            for (int q = 0; q < a.n_parts; ++q) {
                const float4 r0 = ld_cg_f4(a.xchg_local + feat_record(a, q, t));
                const float4 r1 = ld_cg_f4(a.xchg_local + feat_record(a, q, t) + 1);
                pu += r0.x; nu += r0.y; uu += r0.z; pp += r0.w; nn += r1.x;
            }
        } else {
            pu = group_sum(pu, gmask, LANES); nu = group_sum(nu, gmask, LANES);
            uu = group_sum(uu, gmask, LANES); pp = group_sum(pp, gmask, LANES); nn = group_sum(nn, gmask, LANES);
        }
        const float z = pu - nu;
        const float ez = expf(-fabsf(z));
        loss_t = fmaxf(-z, 0.f) + log1pf(ez);                 // softplus(-z) = -logsigmoid(z)
        const float s = (z >= 0.f) ? ez / (1.f + ez) : 1.f / (1.f + ez);   // sigmoid(-z)
        reg_t = 0.5f * (uu + pp + nn);
        const float ca = a.c_bpr * s * inv_norm, cr = a.c_reg * inv_norm;
        if (a.write_coef && lane == 0) a.coef[t] = ca;
        if (a.clear_mask != nullptr && lane == 0) {
            atomicAnd(a.clear_mask + (u >> 5), ~(1u << (u & 31)));
            atomicAnd(a.clear_mask + (p >> 5), ~(1u << (p & 31)));
            atomicAnd(a.clear_mask + (n >> 5), ~(1u << (n & 31)));
        }
        if (a.scatter) {
            const bool own_u = (u >= a.own_begin && u < a.own_end);
            const bool own_p = (p >= a.own_begin && p < a.own_end);
            const bool own_n = (n >= a.own_begin && n < a.own_end);
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
                float4 gu, gp, gn;
#define LGCN_BPR1(c) \
                gu.c = ca * (Nn[q].c - P[q].c) + cr * U[q].c; \
                gp.c = -ca * U[q].c + cr * P[q].c; \
                gn.c = ca * U[q].c + cr * Nn[q].c;
                LGCN_BPR1(x) LGCN_BPR1(y) LGCN_BPR1(z) LGCN_BPR1(w)
#undef LGCN_BPR1
                if (own_u) red_add_f4(a.G + (size_t)u * VEC + lane + q * LANES, gu);
                if (own_p) red_add_f4(a.G + (size_t)p * VEC + lane + q * LANES, gp);
                if (own_n) red_add_f4(a.G + (size_t)n * VEC + lane + q * LANES, gn);
            }
        }
    }
    // ---- deterministic scalar reduction: CTA partial -> last CTA sums partials in block order ---
    if (lane == 0) { s_loss[grp] = loss_t; s_reg[grp] = reg_t; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f, r = 0.f;
        for (int g = 0; g < GROUPS; ++g) { l += s_loss[g]; r += s_reg[g]; }
        a.partials[2 * blockIdx.x] = l; a.partials[2 * blockIdx.x + 1] = r;
        __threadfence();
        s_last = (atomicAdd(a.counter, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        // 32 lanes stride over the block partials, then a fixed-shape butterfly
        float l = 0.f, r = 0.f;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) {
            l += __ldcg(a.partials + 2 * b); r += __ldcg(a.partials + 2 * b + 1);
        }
        for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); r += __shfl_xor_sync(0xffffffffu, r, o); }
        if (threadIdx.x == 0) {
            const float bpr = l * inv_norm, reg = r * inv_norm;
            const float total = bpr + a.decay * reg;
            a.loss_out[0] = bpr; a.loss_out[1] = reg; a.loss_out[2] = total;
            const float run = a.loss_out[3] + total;           // running sum (epoch average), reset by the host
            a.loss_out[3] = run;
            if (a.loss_host != nullptr) {                      // the host reads the loss without a D2H copy
                a.loss_host[0] = bpr; a.loss_host[1] = reg; a.loss_host[2] = total; a.loss_host[3] = run;
                __threadfence_system();
            }
            *a.counter = 0;
        }
    }
}

// Feature partition, first half of K2: this rank's share of the five dot products of every triple, stored into every rank.
template <int D>
__global__ void __launch_bounds__(kBprThreads)
bpr_feat_partial_kernel(const __grid_constant__ BprArgs a) {
    constexpr int LANES = BGeo<D>::LANES, VPL = BGeo<D>::VPL, VEC = BGeo<D>::VEC;
    constexpr int GROUPS = kBprThreads / LANES;
    const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const int off = a.ctl[0];
    const int B = min(a.ctl[1], a.B_cap);
    const int t = blockIdx.x * GROUPS + grp;
    if (t >= B) return;
    const long long u = a.users[off + t], p = a.pos[off + t] + a.n_users, n = a.neg[off + t] + a.n_users;
    const float4* ur = a.out + (size_t)u * VEC + lane;
    const float4* pr = a.out + (size_t)p * VEC + lane;
    const float4* nr = a.out + (size_t)n * VEC + lane;
    float pu = 0.f, nu = 0.f, uu = 0.f, pp = 0.f, nn = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        const float4 U = __ldg(ur + q * LANES), P = __ldg(pr + q * LANES), Nn = __ldg(nr + q * LANES);
        pu += f4_dot(U, P); nu += f4_dot(U, Nn);
        uu += f4_dot(U, U); pp += f4_dot(P, P); nn += f4_dot(Nn, Nn);
    }
    pu = group_sum(pu, gmask, LANES); nu = group_sum(nu, gmask, LANES);
    uu = group_sum(uu, gmask, LANES); pp = group_sum(pp, gmask, LANES); nn = group_sum(nn, gmask, LANES);
    if (lane == 0) {
        const size_t at = feat_record(a, a.part, t);
        const float4 r0 = make_float4(pu, nu, uu, pp), r1 = make_float4(nn, 0.f, 0.f, 0.f);
        for (int q = 0; q < a.n_parts; ++q) { st_stream_f4(a.xchg_peer[q] + at, r0); st_stream_f4(a.xchg_peer[q] + at + 1, r1); }
    }
}

// Deterministic gradient assembly: one group per (role, triple) entry e of the 3B-long list
// [users | pos | neg].  The group whose entry is the FIRST occurrence of its row adds up the
// contributions of every later entry with the same row in entry order and does a plain
// read-modify-write of that row of G (no other group touches it).
template <int D>
__global__ void __launch_bounds__(kBprThreads)
bpr_rowreduce_kernel(const __grid_constant__ BprArgs a) {
    constexpr int LANES = BGeo<D>::LANES, VPL = BGeo<D>::VPL, VEC = BGeo<D>::VEC;
    constexpr int GROUPS = kBprThreads / LANES;
    const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    const int off = a.ctl[0];
    const int B = min(a.ctl[1], a.B_cap);
    const int e = blockIdx.x * GROUPS + grp;
    if (e >= 3 * B) return;
    const long long* users = a.users + off; const long long* pos = a.pos + off; const long long* neg = a.neg + off;
    auto row_of = [&](int q) -> long long {
        return q < B ? users[q] : (q < 2 * B ? pos[q - B] + a.n_users : neg[q - 2 * B] + a.n_users);
    };
    const long long row = row_of(e);
    if (row < a.own_begin || row >= a.own_end) return;
    // user rows only collide with user entries, item rows with pos/neg entries
    const int lo = (e < B) ? 0 : B, hi = (e < B) ? B : 3 * B;
    // (1) is there an earlier entry with the same row?
    int dup_before = 0;
    for (int q = lo + lane; q < e; q += LANES) dup_before |= (row_of(q) == row);
    for (int o = LANES >> 1; o > 0; o >>= 1) dup_before |= __shfl_xor_sync(gmask, dup_before, o, LANES);
    if (dup_before) return;
    // (2) leader: accumulate entries e, then every later duplicate in increasing order
    const float cr = a.c_reg * eff_inv_norm(a, B);
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
    auto add_entry = [&](int q) {
        const int role = q < B ? 0 : (q < 2 * B ? 1 : 2);
        const int t = q - role * B;
        const float ca = a.coef[t];
        const float4* ur = a.out + (size_t)users[t] * VEC + lane;
        const float4* pr = a.out + (size_t)(pos[t] + a.n_users) * VEC + lane;
        const float4* nr = a.out + (size_t)(neg[t] + a.n_users) * VEC + lane;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 U = __ldg(ur + v * LANES);
            float4 g;
            if (role == 0) {
                const float4 P = __ldg(pr + v * LANES), Nn = __ldg(nr + v * LANES);
                g.x = ca * (Nn.x - P.x) + cr * U.x; g.y = ca * (Nn.y - P.y) + cr * U.y;
                g.z = ca * (Nn.z - P.z) + cr * U.z; g.w = ca * (Nn.w - P.w) + cr * U.w;
            } else if (role == 1) {
                const float4 P = __ldg(pr + v * LANES);
                g.x = -ca * U.x + cr * P.x; g.y = -ca * U.y + cr * P.y; g.z = -ca * U.z + cr * P.z; g.w = -ca * U.w + cr * P.w;
            } else {
                const float4 Nn = __ldg(nr + v * LANES);
                g.x = ca * U.x + cr * Nn.x; g.y = ca * U.y + cr * Nn.y; g.z = ca * U.z + cr * Nn.z; g.w = ca * U.w + cr * Nn.w;
            }
            f4_add(acc[v], g);
        }
    };
    add_entry(e);
    for (int base = e + 1; base < hi; base += LANES) {
        const int q = base + lane;
        const bool hit = (q < hi) && (row_of(q) == row);
        unsigned m = __ballot_sync(gmask, hit);
        m >>= ((threadIdx.x & 31) / LANES * LANES);           // bits relative to the group
        if (LANES < 32) m &= ((1u << LANES) - 1u);
        while (m) { const int b = __ffs(m) - 1; m &= m - 1; add_entry(base + b); }
    }
    float4* dst = a.G + (size_t)row * VEC + lane;
#pragma unroll
    for (int v = 0; v < VPL; ++v) { float4 cur = dst[v * LANES]; f4_add(cur, acc[v]); dst[v * LANES] = cur; }
}

template <int D>
__global__ void __launch_bounds__(kBprThreads)
bpr_clear_rows_kernel(float4* G, const long long* users, const long long* pos, const long long* neg,
                      int B_cap, const int* ctl, int n_users) {
    constexpr int LANES = BGeo<D>::LANES, VPL = BGeo<D>::VPL, VEC = BGeo<D>::VEC;
    constexpr int GROUPS = kBprThreads / LANES;
    const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
    const int off = ctl[0];
    const int B = min(ctl[1], B_cap);
    const int e = blockIdx.x * GROUPS + grp;
    if (e >= 3 * B) return;
    const long long row = e < B ? users[off + e] : (e < 2 * B ? pos[off + e - B] + n_users : neg[off + e - 2 * B] + n_users);
    float4* dst = G + (size_t)row * VEC + lane;
#pragma unroll
    for (int v = 0; v < VPL; ++v) dst[v * LANES] = f4_zero();
}

// ctl = {offset, B, total, -}: move to the next batch of an epoch that is resident on the device
__global__ void batch_advance_kernel(int* ctl, int B_cap) {
    const int off = ctl[0] + ctl[1];
    int B = ctl[2] - off; if (B > B_cap) B = B_cap; if (B < 0) B = 0;
    ctl[0] = off; ctl[1] = B;
}

// Row sets a training step can be restricted to (engine.py: dead-row pruning).  One warp per batch entry:
//   m0 = rows the loss reads (the 3B batch rows), m1 = m0 + their neighbours (rows of X_{L-1} that
//   out[m0] depends on; also the non-zero rows of the first backward product).
__global__ void __launch_bounds__(256)
batch_masks_kernel(const long long* users, const long long* pos, const long long* neg, int B_cap, const int* ctl,
                   int n_users, const int* __restrict__ indptr, const int* __restrict__ indices, unsigned* m0, unsigned* m1) {
    const int off = ctl[0];
    const int B = min(ctl[1], B_cap);
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= 3 * B) return;
    const int row = (int)(e < B ? users[off + e] : (e < 2 * B ? pos[off + e - B] + n_users : neg[off + e - 2 * B] + n_users));
    if (lane == 0) { atomicOr(m0 + (row >> 5), 1u << (row & 31)); if (m1) atomicOr(m1 + (row >> 5), 1u << (row & 31)); }
    if (m1 == nullptr) return;
    const int s = indptr[row], t = indptr[row + 1];
    for (int j = s + lane; j < t; j += 32) { const int c = indices[j]; atomicOr(m1 + (c >> 5), 1u << (c & 31)); }
}

// the same bitmap restricted to a row block [row_begin, row_end) and indexed by LOCAL row (row partition: a rank prunes
// the last forward layer to the batch rows it owns)
__global__ void __launch_bounds__(256)
batch_masks_rows_kernel(const long long* users, const long long* pos, const long long* neg, int B_cap, const int* ctl,
                        int n_users, int row_begin, int row_end, unsigned* m0) {
    const int off = ctl[0];
    const int B = min(ctl[1], B_cap);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * B) return;
    const int row = (int)(e < B ? users[off + e] : (e < 2 * B ? pos[off + e - B] + n_users : neg[off + e - 2 * B] + n_users));
    if (row < row_begin || row >= row_end) return;
    const int r = row - row_begin;
    atomicOr(m0 + (r >> 5), 1u << (r & 31));
}

static size_t bpr_blocks(int B_cap, int d) {
    const int vec = d / 4, lanes = vec < 32 ? vec : 32, groups = kBprThreads / lanes;
    return (size_t)(B_cap + groups - 1) / groups;
}

template <int D>
static int launch_bpr(BprArgs& a, int deterministic, cudaStream_t st) {
    constexpr int GROUPS = kBprThreads / BGeo<D>::LANES;
    const unsigned blocks = (unsigned)((a.B_cap + GROUPS - 1) / GROUPS);
    const bool want_grad = (a.G != nullptr);
    a.scatter = want_grad && !deterministic;
    a.write_coef = want_grad && deterministic;
    bpr_kernel<D><<<blocks, kBprThreads, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("bpr_kernel");
    if (want_grad && deterministic) {
        const unsigned blocks3 = (unsigned)((3LL * a.B_cap + GROUPS - 1) / GROUPS);
        bpr_rowreduce_kernel<D><<<blocks3, kBprThreads, 0, st>>>(a);
        LGCN_CHECK_LAUNCH("bpr_rowreduce_kernel");
    }
    return 0;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" size_t lgcn_bpr_workspace_bytes(int32_t B_cap, int32_t d) {
    if (B_cap <= 0 || d < 8) return 0;
    // [counter:16 B][partials: 2 floats per CTA][coef: B_cap floats]
    return 16 + align_up(2 * sizeof(float) * bpr_blocks(B_cap, d), 16) + align_up(sizeof(float) * (size_t)B_cap, 16);
}

static int bpr_run(const float* out, const int64_t* users, const int64_t* pos, const int64_t* neg,
                   int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items,
                   int32_t d, float inv_norm, float decay, float c_bpr, float c_reg,
                   float* loss_out, float* G, int32_t own_begin, int32_t own_end,
                   int32_t deterministic, void* workspace, size_t workspace_bytes,
                   uint32_t* clear_mask, float* loss_host_mapped, const lgcn_bpr_feat_t* feat, int feat_phase, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(out && users && pos && neg && batch_ctl_dev && loss_out, "bpr: null argument");
    LGCN_CHECK_ARG(B_cap > 0, "bpr: B_cap must be > 0");
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_bpr_workspace_bytes(B_cap, d), "bpr: workspace too small");
    LGCN_CHECK_ARG(((uintptr_t)out % 16) == 0 && ((uintptr_t)G % 16) == 0 && ((uintptr_t)workspace % 16) == 0, "bpr: out/G/workspace must be 16-byte aligned");
    BprArgs a;
    a.out = reinterpret_cast<const float4*>(out);
    a.users = reinterpret_cast<const long long*>(users);
    a.pos = reinterpret_cast<const long long*>(pos);
    a.neg = reinterpret_cast<const long long*>(neg);
    a.B_cap = B_cap; a.ctl = batch_ctl_dev; a.n_users = n_users; a.m_items = m_items;
    a.inv_norm = inv_norm; a.decay = decay; a.c_bpr = c_bpr; a.c_reg = c_reg;
    a.loss_out = loss_out; a.G = reinterpret_cast<float4*>(G); a.own_begin = own_begin; a.own_end = own_end;
    a.clear_mask = clear_mask; a.loss_host = loss_host_mapped;
    a.n_parts = 1; a.part = 0; a.sc = nullptr; a.xchg_local = nullptr;
    for (int q = 0; q <= LGCN_MAX_PEERS; ++q) a.xchg_peer[q] = nullptr;
    if (feat) {
        LGCN_CHECK_ARG(feat->n_parts >= 1 && feat->n_parts <= LGCN_MAX_PEERS + 1 && feat->part >= 0 && feat->part < feat->n_parts,
                       "bpr_feat: part %d of %d out of range", feat->part, feat->n_parts);
        LGCN_CHECK_ARG(feat->scalars_dev && feat->records_local && ((uintptr_t)feat->records_local % 16) == 0, "bpr_feat: scalars / record buffer missing or misaligned");
        a.n_parts = feat->n_parts; a.part = feat->part; a.sc = feat->scalars_dev;
        a.xchg_local = reinterpret_cast<float4*>(feat->records_local);
        for (int q = 0; q < feat->n_parts; ++q) {
            LGCN_CHECK_ARG(feat_phase != 1 || (feat->records_peer[q] && ((uintptr_t)feat->records_peer[q] % 16) == 0), "bpr_feat: record buffer of rank %d not mapped", q);
            a.xchg_peer[q] = reinterpret_cast<float4*>(feat->records_peer[q]);
        }
    }
    char* w = static_cast<char*>(workspace);
    a.counter = reinterpret_cast<int*>(w);
    a.partials = reinterpret_cast<float*>(w + 16);
    a.coef = reinterpret_cast<float*>(w + 16 + align_up(2 * sizeof(float) * bpr_blocks(B_cap, d), 16));
    cudaStream_t st = as_stream(stream);
    if (feat_phase == 1) {
#define LGCN_FEAT1(DD) { constexpr int GR = kBprThreads / BGeo<DD>::LANES; \
        bpr_feat_partial_kernel<DD><<<(unsigned)((B_cap + GR - 1) / GR), kBprThreads, 0, st>>>(a); } break
        switch (d) {
            case 8: LGCN_FEAT1(8); case 16: LGCN_FEAT1(16); case 32: LGCN_FEAT1(32); case 64: LGCN_FEAT1(64);
            default: return fail("bpr_feat_partial: local width %d unsupported (8,16,32,64)", d);
        }
#undef LGCN_FEAT1
        LGCN_CHECK_LAUNCH("bpr_feat_partial_kernel");
        return 0;
    }
    switch (d) {
        case 8:   return launch_bpr<8>(a, deterministic, st);
        case 16:  return launch_bpr<16>(a, deterministic, st);
        case 32:  return launch_bpr<32>(a, deterministic, st);
        case 64:  return launch_bpr<64>(a, deterministic, st);
        case 128: return launch_bpr<128>(a, deterministic, st);
        case 256: return launch_bpr<256>(a, deterministic, st);
        default:  return fail("bpr: d=%d unsupported (8,16,32,64,128,256)", d);
    }
}

extern "C" int lgcn_bpr_fwd_bwd(const float* out, const int64_t* users, const int64_t* pos, const int64_t* neg,
                                int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items,
                                int32_t d, float inv_norm, float decay, float c_bpr, float c_reg,
                                float* loss_out, float* G, int32_t own_begin, int32_t own_end,
                                int32_t deterministic, void* workspace, size_t workspace_bytes,
                                uint32_t* clear_mask, float* loss_host_mapped, lgcn_stream_t stream) {
    return bpr_run(out, users, pos, neg, B_cap, batch_ctl_dev, n_users, m_items, d, inv_norm, decay, c_bpr, c_reg, loss_out, G,
                   own_begin, own_end, deterministic, workspace, workspace_bytes, clear_mask, loss_host_mapped, nullptr, 0, stream);
}

extern "C" int lgcn_bpr_feat_partial(const float* out_slice, const int64_t* users, const int64_t* pos, const int64_t* neg,
                                     int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items, int32_t d_local,
                                     const lgcn_bpr_feat_t* feat, void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(feat, "bpr_feat_partial: feat is null");
    float dummy_loss;      // never dereferenced by the partial kernel
    return bpr_run(out_slice, users, pos, neg, B_cap, batch_ctl_dev, n_users, m_items, d_local, 0.f, 0.f, 0.f, 0.f, &dummy_loss, nullptr,
                   0, 0, 0, workspace, workspace_bytes, nullptr, nullptr, feat, 1, stream);
}

extern "C" int lgcn_bpr_feat_finish(const float* out_slice, const int64_t* users, const int64_t* pos, const int64_t* neg,
                                    int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items,
                                    int32_t d_local, float inv_norm, float decay, float c_bpr, float c_reg,
                                    float* loss_out, float* G_slice, int32_t deterministic, const lgcn_bpr_feat_t* feat,
                                    void* workspace, size_t workspace_bytes,
                                    uint32_t* clear_mask, float* loss_host_mapped, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(feat, "bpr_feat_finish: feat is null");
    return bpr_run(out_slice, users, pos, neg, B_cap, batch_ctl_dev, n_users, m_items, d_local, inv_norm, decay, c_bpr, c_reg, loss_out, G_slice,
                   0, n_users + m_items, deterministic, workspace, workspace_bytes, clear_mask, loss_host_mapped, feat, 2, stream);
}

extern "C" int lgcn_bpr_clear_rows(float* G, const int64_t* users, const int64_t* pos, const int64_t* neg,
                                   int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t d,
                                   lgcn_stream_t stream) {
    LGCN_CHECK_ARG(G && users && pos && neg && batch_ctl_dev && B_cap > 0, "bpr_clear_rows: bad arguments");
    cudaStream_t st = as_stream(stream);
    float4* g4 = reinterpret_cast<float4*>(G);
    const long long* u = reinterpret_cast<const long long*>(users);
    const long long* p = reinterpret_cast<const long long*>(pos);
    const long long* n = reinterpret_cast<const long long*>(neg);
#define LGCN_CLR(DD) { constexpr int GR = kBprThreads / BGeo<DD>::LANES; \
        bpr_clear_rows_kernel<DD><<<(unsigned)((3LL * B_cap + GR - 1) / GR), kBprThreads, 0, st>>>(g4, u, p, n, B_cap, batch_ctl_dev, n_users); } break
    switch (d) {
        case 8: LGCN_CLR(8); case 16: LGCN_CLR(16); case 32: LGCN_CLR(32); case 64: LGCN_CLR(64); case 128: LGCN_CLR(128); case 256: LGCN_CLR(256);
        default: return fail("bpr_clear_rows: d=%d unsupported", d);
    }
#undef LGCN_CLR
    LGCN_CHECK_LAUNCH("bpr_clear_rows_kernel");
    return 0;
}

extern "C" int lgcn_batch_advance(int32_t* batch_ctl_dev, int32_t B_cap, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(batch_ctl_dev && B_cap > 0, "batch_advance: bad arguments");
    batch_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(batch_ctl_dev, B_cap);
    LGCN_CHECK_LAUNCH("batch_advance_kernel");
    return 0;
}

extern "C" int lgcn_batch_masks(const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t B_cap,
                                const int32_t* batch_ctl_dev, int32_t n_users, int32_t n_nodes,
                                const int32_t* indptr, const int32_t* indices, uint32_t* m0, uint32_t* m1,
                                int32_t clear_first, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(users && pos && neg && batch_ctl_dev && m0 && B_cap > 0 && n_nodes > 0, "batch_masks: bad arguments");
    LGCN_CHECK_ARG(m1 == nullptr || (indptr && indices), "batch_masks: m1 needs the CSR");
    cudaStream_t st = as_stream(stream);
    const size_t words = ((size_t)n_nodes + 31) / 32;
    if (clear_first) cudaMemsetAsync(m0, 0, words * 4, st);
    if (m1) cudaMemsetAsync(m1, 0, words * 4, st);
    const unsigned blocks = (unsigned)((3LL * B_cap * 32 + 255) / 256);
    batch_masks_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(users), reinterpret_cast<const long long*>(pos),
                                               reinterpret_cast<const long long*>(neg), B_cap, batch_ctl_dev, n_users, indptr, indices, m0, m1);
    LGCN_CHECK_LAUNCH("batch_masks_kernel");
    return 0;
}

extern "C" int lgcn_batch_masks_rows(const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t B_cap,
                                     const int32_t* batch_ctl_dev, int32_t n_users, int32_t row_begin, int32_t row_end,
                                     uint32_t* m0_local, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(users && pos && neg && batch_ctl_dev && m0_local && B_cap > 0 && row_begin >= 0 && row_end >= row_begin, "batch_masks_rows: bad arguments");
    cudaStream_t st = as_stream(stream);
    const size_t words = ((size_t)(row_end - row_begin) + 31) / 32;
    if (words == 0) return 0;
    cudaMemsetAsync(m0_local, 0, words * 4, st);
    const unsigned blocks = (unsigned)((3LL * B_cap + 255) / 256);
    batch_masks_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(users), reinterpret_cast<const long long*>(pos),
                                                    reinterpret_cast<const long long*>(neg), B_cap, batch_ctl_dev, n_users, row_begin, row_end, m0_local);
    LGCN_CHECK_LAUNCH("batch_masks_rows_kernel");
    return 0;
}
