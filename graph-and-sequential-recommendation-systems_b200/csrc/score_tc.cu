// score_tc.cu — K3 (tensor-core path): user x item scores on tcgen05 (TF32, TMEM accumulators, TMA-fed
// shared memory), fused with the train-item mask and a per-row candidate filter, followed by an EXACT
// fp32 rescoring of the candidates with a certificate — so the result is bit-identical to the exact
// path (score_topk.cu) and to the CPU oracle, while > 99 % of the 2*U*M*d flops run on tensor cores.
//
// Replaces torch.matmul (reference code/model.py:122), the -(1<<10) mask (code/Procedure.py:177-181)
// and torch.topk (code/Procedure.py:183).
//
//   phase A  score_tc_kernel: CTA = 128 users x (a split of) all items, 6 warps:
//              warp 0   TMA producer   cp.async.bulk.tensor.2d (128B swizzle) -> 2-stage ring of 256-item B tiles
//              warp 1   MMA issuer     8 x tcgen05.mma.kind::tf32 (M=128,N=256,K=8) per tile into one of two
//                                      256-column TMEM accumulators; tcgen05.commit frees the smem stage and
//                                      publishes the accumulator
//              warps 2-5 epilogue      tcgen05.ld 32x32b.x32: thread t owns row 32*(warp%4)+t, compares 256
//                                      approximate scores per tile with its running K'-th best and, on a hit,
//                                      binary-searches the user's CSR row (train-item mask) before inserting
//            -> per (row, split) the K' = 32 best approximate candidates.
//   phase B  rescore_kernel: warp per row recomputes the candidates' scores as the fp32 FMA chain of the
//            exact contract, selects the top-k (score desc, item id asc) and CERTIFIES it: every item that
//            was filtered out has approx <= t (the smallest kept approx of its split), hence
//            exact <= t + eps with eps = c * |u| * max|v|  (TF32 truncation bound, Cauchy-Schwarz);
//            if t + eps < (k-th exact score) nothing outside the candidate set can enter the top-k.
//            Rows that fail the certificate (or have < k unmasked items) are flagged and re-done by the
//            exact kernel — the caller sees one bit-exact result either way.
//
// Only d = 64 and k <= 24 take this path (A stays resident in shared memory: 32 KB; B ring 128 KB; lists 32 KB).
#include "common.cuh"
#include <cuda.h>
#include <float.h>

namespace lgcn {

constexpr int TC_M = 128;            // users per CTA
constexpr int TC_N = 256;            // items per MMA tile
constexpr int TC_D = 64;             // embedding width handled by this path
constexpr int TC_KP = 32;            // candidates kept per (row, split)
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int TC_KEEP = 16;            // entries a buffer is compacted to
constexpr int TC_ATOM_K = 32;        // fp32 elements per 128-byte swizzle atom row
constexpr int TC_A_ATOM_BYTES = TC_M * 128;           // 16 KB
constexpr int TC_B_ATOM_BYTES = TC_N * 128;           // 32 KB
constexpr int TC_A_BYTES = 2 * TC_A_ATOM_BYTES;       // 32 KB
constexpr int TC_B_STAGE_BYTES = 2 * TC_B_ATOM_BYTES; // 64 KB
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_A_BYTES + TC_STAGES * TC_B_STAGE_BYTES + 2 * (8 * 32 * TC_KP) * 4 + TC_M * 4 + 128;
constexpr float TC_EPS_C = 0.0025f;  // > 2^-9 (both operands truncated to 10 mantissa bits) + accumulation slack

struct TcArgs {
    const long long* users; int Bt; int m_items;
    const int* mask_indptr; const int* mask_indices; int mask_col_offset;
    int tiles_per_split; int n_splits;
    float* cand_val; int* cand_idx;      // [Bt][2*n_splits][TC_KP]
    float* cand_tau;                     // [Bt][2*n_splits]: everything the sub-stream dropped is <= this
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (long long spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1LL << 28)) __trap();          // a protocol bug must fail, not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand, 128-byte swizzle, rows packed at 128 B (8-row groups 1024 B apart):
// start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 >> 4 | version 1 (sm_100) | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// order-preserving float <-> int (for atomicMax on thresholds that may be negative)
__device__ __forceinline__ int tc_enc(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float tc_dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct RowState { int cnt; float tau; };

// Warp-cooperative compaction of the candidate buffers named in `need` (one bit per lane/row): every lane takes one
// entry of the row, ranks it with 32 shuffles, the best TC_KEEP are written back and the row's threshold becomes
// the TC_KEEP-th best.  Returns the calling lane's updated (cnt, tau).
__device__ __noinline__ RowState tc_compact_rows(unsigned need, float* bv, int* bi, int* tau_row, int lane, int cnt, float tau) {
    while (need) {
        const int L = __ffs(need) - 1; need &= need - 1;
        const int n = __shfl_sync(0xffffffffu, cnt, L);
        const int slot = L * TC_KP + ((lane + L) & 31);
        const float sv = (lane < n) ? bv[slot] : -FLT_MAX;
        const int sid = (lane < n) ? bi[slot] : 0x7fffffff;
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            const float so = __shfl_sync(0xffffffffu, sv, o);
            rank += (so > sv || (so == sv && o < lane)) ? 1 : 0;
        }
        __syncwarp();
        if (rank < TC_KEEP) { const int d = L * TC_KP + ((rank + L) & 31); bv[d] = sv; bi[d] = sid; }
        const int src = __ffs(__ballot_sync(0xffffffffu, rank == TC_KEEP - 1)) - 1;
        const float t16 = __shfl_sync(0xffffffffu, sv, src);
        if (lane == L) { cnt = TC_KEEP; tau = fmaxf(tau, t16); atomicMax(tau_row + L, tc_enc(t16)); }
        __syncwarp();
    }
    RowState r; r.cnt = cnt; r.tau = tau;
    return r;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024-B alignment
    uint8_t* sA = smem;
    uint8_t* sB = sA + TC_A_BYTES;
    float* lv = reinterpret_cast<float*>(sB + TC_STAGES * TC_B_STAGE_BYTES);      // [8 warps][32 rows][TC_KP]
    int* li = reinterpret_cast<int*>(lv + 8 * 32 * TC_KP);
    int* tau_sh = li + 8 * 32 * TC_KP;                                             // [TC_M] row thresholds shared by the two halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(tau_sh + TC_M);
    // bars: 0 a_full | 1,2 b_full | 3,4 b_empty | 5,6 tmem_full | 7,8 tmem_empty ; then the TMEM base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ub = blockIdx.x * TC_M;
    const int n_item_tiles = (a.m_items + TC_N - 1) / TC_N;
    const int t_begin = blockIdx.y * a.tiles_per_split;
    const int t_end = min(n_item_tiles, t_begin + a.tiles_per_split);
    const int n_tiles = t_end - t_begin;

    if (threadIdx.x == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(BAR(1 + s), 1); mbar_init(BAR(3 + s), 1); }
        for (int c = 0; c < 2; ++c) { mbar_init(BAR(5 + c), 1); mbar_init(BAR(7 + c), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < TC_M) tau_sh[threadIdx.x] = tc_enc(-FLT_MAX);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(BAR(0), TC_A_BYTES);
            tma_load_2d(smem_u32(sA), &map_a, BAR(0), 0, ub);
            tma_load_2d(smem_u32(sA + TC_A_ATOM_BYTES), &map_a, BAR(0), TC_ATOM_K, ub);
            for (int it = 0; it < n_tiles; ++it) {
                const int s = it % TC_STAGES, r = it / TC_STAGES;
                mbar_wait(BAR(3 + s), (r & 1) ^ 1);
                mbar_expect_tx(BAR(1 + s), TC_B_STAGE_BYTES);
                const int ib = (t_begin + it) * TC_N;
                uint8_t* dst = sB + s * TC_B_STAGE_BYTES;
                tma_load_2d(smem_u32(dst), &map_b, BAR(1 + s), 0, ib);
                tma_load_2d(smem_u32(dst + TC_B_ATOM_BYTES), &map_b, BAR(1 + s), TC_ATOM_K, ib);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(BAR(0), 0);
            for (int it = 0; it < n_tiles; ++it) {
                const int s = it % TC_STAGES, r = it / TC_STAGES, acc = it & 1, ra = it >> 1;
                mbar_wait(BAR(1 + s), r & 1);                 // B tile landed
                mbar_wait(BAR(7 + acc), (ra & 1) ^ 1);        // accumulator drained by the epilogue
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + s * TC_B_STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < TC_D / 8; ++j) {
                    const uint32_t koff = (j >> 2) * 0 + (j & 3) * 32;         // 32 bytes per K=8 step inside the atom
                    const uint64_t da = umma_desc_k_sw128(a0 + (j >> 2) * TC_A_ATOM_BYTES + koff);
                    const uint64_t db = umma_desc_k_sw128(b0 + (j >> 2) * TC_B_ATOM_BYTES + koff);
                    tc_mma_tf32(tmem_base + acc * TC_N, da, db, TC_IDESC, j > 0 ? 1u : 0u);
                }
                tc_commit(BAR(3 + s));                         // smem stage free when these MMAs retire
                tc_commit(BAR(5 + acc));                       // accumulator ready
            }
        }
    } else {
        // ================= epilogue: 8 warps, thread-per-row threshold filter =================
        // Warp e reads TMEM lane quarter q = warp%4 (hardware rule) and every other 32-column chunk (half h), so a
        // row is watched by two threads with private candidate buffers.  A thread only compares: max of 4 scores
        // against its row threshold tau; hits are APPENDED to a 32-slot buffer (no sorting, no cooperation, lanes
        // proceed independently).  When a buffer holds >= 24 entries the warp compacts it cooperatively (one entry
        // per lane, rank by 32 shuffles) to its 16 best and raises tau to the 16th.  Everything that was ever
        // dropped or skipped is <= the final tau, which is what phase B needs for its certificate.
        const int e = warp - 2, q = warp & 3, h = e >> 2;
        const int r = q * 32 + lane;                           // row of the tile owned by this thread
        const bool live = (ub + r) < a.Bt;
        const int* mrow = nullptr; int m_cur = 0, m_end = 0, m_next = 0x7fffffff;
        if (live && a.mask_indptr) {
            const long long my_user = a.users ? a.users[ub + r] : (long long)(ub + r);
            const int lo = __ldg(a.mask_indptr + my_user), hi = __ldg(a.mask_indptr + my_user + 1);
            mrow = a.mask_indices + lo; m_end = hi - lo;
            const int first_item = t_begin * TC_N + a.mask_col_offset;        // first column id this CTA scores
            int l = 0, hh = m_end;                                            // lower_bound: cursor into the sorted row
            while (l < hh) { const int mid = (l + hh) >> 1; if (__ldg(mrow + mid) < first_item) l = mid + 1; else hh = mid; }
            m_cur = l;
            if (m_cur < m_end) m_next = __ldg(mrow + m_cur) - a.mask_col_offset;
        }
        float* bv = lv + e * (32 * TC_KP);                     // this warp's 32 buffers: slot(row, p) = row*32 + ((p+row)&31)
        int* bi = li + e * (32 * TC_KP);
        int cnt = 0; float tau = live ? -FLT_MAX : FLT_MAX; bool lost = false;
        int* tau_row = tau_sh + q * 32;
        for (int it = 0; it < n_tiles; ++it) {
            const int acc = it & 1, ra = it >> 1;
            mbar_wait(BAR(5 + acc), ra & 1);
            tc_fence_after();
            const int ib = (t_begin + it) * TC_N;
#pragma unroll 1
            for (int ch = h; ch < TC_N / 32; ch += 2) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_N + ch * 32), v);
                const int i0 = ib + ch * 32;
                // (2) train items of this row inside [i0, i0+32): advance the cursor through the sorted CSR row
                unsigned mbits = 0;
                while (m_next < i0 + 32) {
                    if (m_next >= i0) mbits |= 1u << (m_next - i0);
                    ++m_cur;
                    m_next = (m_cur < m_end) ? __ldg(mrow + m_cur) - a.mask_col_offset : 0x7fffffff;
                }
                unsigned bad = mbits;                                       // masked or out-of-range columns: rejected at append time
                if (i0 + 32 > a.m_items) bad |= (i0 < a.m_items) ? ~((1u << (a.m_items - i0)) - 1u) : 0xffffffffu;
                tau = fmaxf(tau, tc_dec(tau_row[lane]));                    // the other half may have raised the row's threshold
                // (3) filter: 4 scores at a time
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    // make room first: a group appends at most 4 entries, so a buffer with <= 28 can never overflow
                    const unsigned need = __ballot_sync(0xffffffffu, cnt > TC_KP - 4);
                    if (need) { const RowState st = tc_compact_rows(need, bv, bi, tau_row, lane, cnt, tau); cnt = st.cnt; tau = st.tau; }
                    const float m4 = fmaxf(fmaxf(v[4 * g4], v[4 * g4 + 1]), fmaxf(v[4 * g4 + 2], v[4 * g4 + 3]));
                    if (m4 > tau) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const float sc = v[4 * g4 + x];
                            if (sc > tau && !((bad >> (4 * g4 + x)) & 1u)) {
                                if (cnt < TC_KP) {
                                    const int d = lane * TC_KP + ((cnt + lane) & 31);
                                    bv[d] = sc; bi[d] = i0 + 4 * g4 + x; ++cnt;
                                } else lost = true;             // > 8 hits in one chunk on a nearly full buffer: give the row up
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(7 + acc));
        }
        __syncwarp();
        if (live) {
            const int n_sub = a.n_splits * 2, sub = blockIdx.y * 2 + h;
            const size_t o = ((size_t)(ub + r) * n_sub + sub) * TC_KP;
            for (int p = 0; p < TC_KP; ++p) {
                const int d = lane * TC_KP + ((p + lane) & 31);
                a.cand_val[o + p] = (p < cnt) ? bv[d] : -FLT_MAX;
                a.cand_idx[o + p] = (p < cnt) ? bi[d] : 0x7fffffff;
            }
            a.cand_tau[(size_t)(ub + r) * n_sub + sub] = lost ? FLT_MAX : fmaxf(tau, tc_dec(tau_row[lane]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
    }
}

// A operand: the batch's user rows, gathered and zero-padded to a multiple of 128 rows
__global__ void gather_rows_kernel(const float4* __restrict__ U, const long long* __restrict__ users, int Bt, int Bt_pad, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // one float4 per thread, 16 per row (d = 64)
    if (i >= Bt_pad * 16) return;
    const int r = i >> 4, c = i & 15;
    float4 v = f4_zero();
    if (r < Bt) { const long long u = users ? users[r] : (long long)r; v = __ldg(U + (size_t)u * 16 + c); }
    out[i] = v;
}

__global__ void item_norm_max_kernel(const float4* __restrict__ V, int m_items, int* __restrict__ vmax_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float n2 = 0.f;
    if (i < m_items) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { const float4 x = __ldg(V + (size_t)i * 16 + c); n2 += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w; }
    }
    float nrm = sqrtf(n2) * 1.0001f;
    for (int o = 16; o > 0; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
    if ((threadIdx.x & 31) == 0) atomicMax(vmax_bits, __float_as_int(nrm));        // non-negative floats order like ints
}

// phase B: warp per row — exact rescoring, top-k, certificate
constexpr int RS_WARPS = 4;
__global__ void __launch_bounds__(RS_WARPS * 32)
rescore_kernel(const float* __restrict__ U, const float* __restrict__ V, const long long* __restrict__ users, int Bt, int n_splits, int k,
               const float* __restrict__ cand_val, const int* __restrict__ cand_idx, const float* __restrict__ cand_tau, const int* __restrict__ vmax_bits,
               long long* __restrict__ idx_out, float* __restrict__ val_out, int* __restrict__ flags, int* __restrict__ n_flagged) {
    __shared__ float s_sc[RS_WARPS][32 * TC_KP];
    __shared__ int s_id[RS_WARPS][32 * TC_KP];
    __shared__ float s_u[RS_WARPS][TC_D];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * RS_WARPS + w;
    if (b >= Bt) return;
    const long long u = users ? users[b] : (long long)b;
    const float* ur = U + (size_t)u * TC_D;
    float un2 = 0.f;
    for (int c = lane; c < TC_D; c += 32) { const float x = __ldg(ur + c); s_u[w][c] = x; un2 += x * x; }
    for (int o = 16; o > 0; o >>= 1) un2 += __shfl_xor_sync(0xffffffffu, un2, o);
    __syncwarp();
    const int ncand = n_splits * TC_KP;
    const size_t base = (size_t)b * ncand;
    const float eps = TC_EPS_C * sqrtf(un2) * __int_as_float(*vmax_bits);
    // (1) approximate scores: find the k-th best; only candidates within 2*eps of it can be in the exact top-k
    for (int c = lane; c < ncand; c += 32) { s_sc[w][c] = cand_val[base + c]; s_id[w][c] = cand_idx[base + c]; }
    __syncwarp();
    unsigned taken = 0; float kth_apx = -FLT_MAX;
    for (int qq = 0; qq < k; ++qq) {
        float bv = -FLT_MAX; int bc = -1;
        for (int c = lane, t = 0; c < ncand; c += 32, ++t) {
            const float s = s_sc[w][c];
            if (!((taken >> t) & 1u) && s_id[w][c] != 0x7fffffff && s > bv) { bv = s; bc = c; }
        }
        float mv = bv; int mc = bc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, mv, o); const int oc = __shfl_xor_sync(0xffffffffu, mc, o);
            if (ov > mv || (ov == mv && oc > mc)) { mv = ov; mc = oc; }
        }
        if (mc >= 0 && (mc & 31) == lane) taken |= 1u << (mc >> 5);
        kth_apx = mv;
    }
    const float keep_above = kth_apx - 2.f * eps;
    // (2) exact fp32 FMA chain for the survivors
    int valid = 0;
    for (int c = lane; c < ncand; c += 32) {
        const int id = s_id[w][c];
        float s = -FLT_MAX;
        if (id != 0x7fffffff && s_sc[w][c] >= keep_above) {
            const float4* vr = reinterpret_cast<const float4*>(V + (size_t)id * TC_D);
            s = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < TC_D / 4; ++c4) {                 // the exact contract: one fp32 FMA chain in k order
                const float4 x = __ldg(vr + c4);
                s = fmaf(s_u[w][4 * c4 + 0], x.x, s); s = fmaf(s_u[w][4 * c4 + 1], x.y, s);
                s = fmaf(s_u[w][4 * c4 + 2], x.z, s); s = fmaf(s_u[w][4 * c4 + 3], x.w, s);
            }
            ++valid;
        } else {
            s_id[w][c] = 0x7fffffff;
        }
        s_sc[w][c] = s;
    }
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    // bound for everything that was filtered out: the largest final threshold of the row's sub-streams
    float tmin = -FLT_MAX;
    for (int s = lane; s < n_splits; s += 32) tmin = fmaxf(tmin, cand_tau[(size_t)b * n_splits + s]);
    for (int o = 16; o > 0; o >>= 1) tmin = fmaxf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    __syncwarp();
    float kth = -FLT_MAX;
    for (int q = 0; q < k; ++q) {
        float bv = -FLT_MAX; int bi = 0x7fffffff, bc = -1;
        for (int c = lane; c < ncand; c += 32) {
            const float s = s_sc[w][c]; const int id = s_id[w][c];
            if (id != 0x7fffffff && (s > bv || (s == bv && id < bi))) { bv = s; bi = id; bc = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bc = oc; }
        }
        if (bc >= 0 && (bc & 31) == lane) s_id[w][bc] = 0x7fffffff;       // taken
        if (lane == 0) { idx_out[(size_t)b * k + q] = bi; val_out[(size_t)b * k + q] = bv; }
        kth = bv;
        __syncwarp();
    }
    if (lane == 0) {
        const bool ok = (valid >= k) && (tmin == -FLT_MAX || (tmin < FLT_MAX && tmin + eps < kth));
        flags[b] = ok ? 0 : 1;
        if (!ok) atomicAdd(n_flagged, 1);
    }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map(CUtensorMap* m, const float* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail("score_tc: cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {(cuuint64_t)TC_D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_D * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_ATOM_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("score_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static int tc_pick_splits(int Bt, int m_items) {
    const int row_tiles = (Bt + TC_M - 1) / TC_M, item_tiles = (m_items + TC_N - 1) / TC_N;
    int want = (2 * sm_count() + row_tiles - 1) / row_tiles;
    if (want < 2) want = 2;                         // >= 4 sub-streams per row: no single one can hold most of the top-k
    if (want > 16) want = 16;                       // 2 sub-streams per split, 32 per row at most
    if (want > item_tiles) want = item_tiles;
    if (want < 1) want = 1;
    return want;
}

}  // namespace lgcn

using namespace lgcn;

// workspace: [A_gathered: Bt_pad*64 floats][cand_val: Bt*32*KP floats][cand_idx: Bt*32*KP ints][vmax: 16 B]
extern "C" size_t lgcn_score_topk_tc_workspace_bytes(int32_t Bt, int32_t m_items, int32_t k) {
    if (Bt <= 0 || m_items <= 0 || k <= 0) return 0;
    const size_t bt_pad = ((size_t)Bt + TC_M - 1) / TC_M * TC_M;
    return align_up(bt_pad * TC_D * 4, 1024) + 2 * align_up((size_t)Bt * 32 * TC_KP * 4, 256) + align_up((size_t)Bt * 32 * 4, 256) + 256;
}

extern "C" int lgcn_score_topk_tc_supported(int32_t d, int32_t k) { return (d == TC_D && k >= 1 && k <= TC_KP - 8) ? 1 : 0; }

extern "C" int lgcn_score_topk_tc(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                                  int32_t m_items, int32_t d, const int32_t* mask_indptr, const int32_t* mask_indices,
                                  int32_t mask_col_offset, int32_t k, int64_t* idx_out, float* val_out,
                                  int32_t* flags_out, int32_t* n_flagged_out,
                                  void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(users_emb && items_emb && idx_out && val_out && flags_out && n_flagged_out, "score_topk_tc: null argument");
    LGCN_CHECK_ARG(lgcn_score_topk_tc_supported(d, k), "score_topk_tc: only d=%d and k<=%d take the tensor-core path", TC_D, TC_KP - 8);
    LGCN_CHECK_ARG(Bt > 0 && m_items >= k, "score_topk_tc: Bt=%d m_items=%d k=%d", Bt, m_items, k);
    LGCN_CHECK_ARG((mask_indptr == nullptr) == (mask_indices == nullptr), "score_topk_tc: mask arrays must both be set or both null");
    LGCN_CHECK_ARG(workspace && ((uintptr_t)workspace % 1024) == 0 && workspace_bytes >= lgcn_score_topk_tc_workspace_bytes(Bt, m_items, k),
                   "score_topk_tc: workspace too small or not 1024-byte aligned");
    LGCN_CHECK_ARG(((uintptr_t)items_emb % 16) == 0 && ((uintptr_t)users_emb % 16) == 0, "score_topk_tc: tables must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int bt_pad = (Bt + TC_M - 1) / TC_M * TC_M;
    char* w = static_cast<char*>(workspace);
    float* A = reinterpret_cast<float*>(w);
    const size_t cand_bytes = align_up((size_t)Bt * 32 * TC_KP * 4, 256);
    float* cand_val = reinterpret_cast<float*>(w + align_up((size_t)bt_pad * TC_D * 4, 1024));
    int* cand_idx = reinterpret_cast<int*>(reinterpret_cast<char*>(cand_val) + cand_bytes);
    float* cand_tau = reinterpret_cast<float*>(reinterpret_cast<char*>(cand_idx) + cand_bytes);
    int* vmax = reinterpret_cast<int*>(reinterpret_cast<char*>(cand_tau) + align_up((size_t)Bt * 32 * 4, 256));

    gather_rows_kernel<<<(bt_pad * 16 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(users_emb),
                                                                 reinterpret_cast<const long long*>(users), Bt, bt_pad, reinterpret_cast<float4*>(A));
    LGCN_CHECK_LAUNCH("gather_rows_kernel");
    cudaMemsetAsync(vmax, 0, 16, st);
    cudaMemsetAsync(n_flagged_out, 0, sizeof(int32_t), st);
    item_norm_max_kernel<<<(m_items + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(items_emb), m_items, vmax);
    LGCN_CHECK_LAUNCH("item_norm_max_kernel");

    CUtensorMap map_a, map_b;
    if (int rc = make_map(&map_a, A, (uint64_t)bt_pad, TC_M)) return rc;
    if (int rc = make_map(&map_b, items_emb, (uint64_t)m_items, TC_N)) return rc;
    TcArgs a;
    a.users = reinterpret_cast<const long long*>(users); a.Bt = Bt; a.m_items = m_items;
    a.mask_indptr = mask_indptr; a.mask_indices = mask_indices; a.mask_col_offset = mask_col_offset;
    const int item_tiles = (m_items + TC_N - 1) / TC_N;
    a.n_splits = tc_pick_splits(Bt, m_items);
    a.tiles_per_split = (item_tiles + a.n_splits - 1) / a.n_splits;
    a.n_splits = (item_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.cand_val = cand_val; a.cand_idx = cand_idx; a.cand_tau = cand_tau;
    cudaError_t e = cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return fail("score_topk_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid(bt_pad / TC_M, a.n_splits);
    score_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(map_a, map_b, a);
    LGCN_CHECK_LAUNCH("score_tc_kernel");
    rescore_kernel<<<(Bt + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, st>>>(users_emb, items_emb, reinterpret_cast<const long long*>(users), Bt,
        2 * a.n_splits, k, cand_val, cand_idx, cand_tau, vmax, reinterpret_cast<long long*>(idx_out), val_out, flags_out, n_flagged_out);
    LGCN_CHECK_LAUNCH("rescore_kernel");
    return 0;
}
