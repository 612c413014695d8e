// score_tc.cu — K3 (tensor-core path): user x item scores on tcgen05 (TF32, TMEM accumulators, TMA-fed
// shared memory), fused with the train-item mask and a per-row candidate filter, followed by an EXACT
// fp32 rescoring of the candidates with a certificate — so the result is bit-identical to the exact
// path (score_topk.cu) and to the CPU oracle, while the 2*U*M*d flops run on tensor cores.
//
// Replaces torch.matmul (reference code/model.py:122), the -(1<<10) mask (code/Procedure.py:177-181)
// and torch.topk (code/Procedure.py:183).
//
// The B x M score matrix is produced twice on the tensor pipe and never written.  What sets the pace (measured,
// profiles/README.md): with both operands in shared memory an M128 N128 K8 tf32 MMA needs the whole 128 B/clk of
// shared memory (100 cycles per MMA instead of 64), and pass 2 reads every score out of TMEM (64 B/clk).
//
//   pass 1   score_tc_kernel<1>: CTA = 256 users (two 128-row MMA tiles sharing every B tile) x a split of the
//            item tiles.  warp 0 = TMA producer (4-stage ring of 128-item tiles), warp 1 = MMA issuer
//            (16 x tcgen05.mma.kind::tf32 M128 N128 K8 per tile, double-buffered TMEM accumulators), warps 2-17 =
//            epilogue.  Pass 1 only needs a SAMPLE to place the threshold, so it runs the GEMM on every other tile
//            (global tile index even — with the interleaved item layout any fixed set of tiles is a fair sample of
//            the ids; round 1 ran all tiles and read half of each: same 50 % sample, twice the MMA work): a thread
//            owns one row and one 64-item half of every sampled tile (tcgen05.ld 32x32b.x32) and reduces it to ONE
//            number, the largest approximate score among those items that are NOT train items of the row (a cursor
//            walks the sorted position-space mask row) -> sample maxima Mx[row][tile + half].
//   select   tc_select_kernel: warp per row, radix select: the KSEL-th largest sample maximum is the row threshold
//            tau.  KSEL unmasked sampled items score >= tau, hence about 2*KSEL +- sqrt(2*KSEL) items overall (any
//            tau is SAFE — the certificate below does not depend on how it was chosen; a poor tau only costs time).
//   pass 2   score_tc_kernel<2>: the same GEMM; 16 epilogue warps, a thread owns one row and one 64-item half of
//            every tile and compares 4 scores at a time against the now FIXED tau.  A hit (rare per lane) stores the
//            4 scores, their position and the mask bits of the group as one event in the list of its (row, split,
//            half) in global memory.  No sorting, no cooperation between lanes, no running threshold.
//   rescore  rescore_kernel: warp per row keeps the event scores that reach tau and are not train items, recomputes
//            them as the fp32 FMA chain of the exact contract (item rows staged through shared memory), ranks by
//            counting (score desc, item id asc) and CERTIFIES: every item that is not a candidate is masked or
//            has approx < tau, hence exact < tau + eps with eps = c * |u| * max|v| (TF32 truncation bound,
//            Cauchy-Schwarz); if tau + eps < (k-th exact score) nothing outside the candidate set can enter the
//            top-k.  Rows that fail the certificate, overflow an event list or have < k candidates are flagged
//            and re-done by the exact kernel — the caller sees one bit-exact result either way.
//
// Item layout.  A tile holds the items {b*T + r : b = 0..127} of one residue r (T = number of tiles), block b sitting
// in slot (b >> 1) + 64*((b + r) & 1), and residue r is tile (r * mul) mod T with mul ~ 0.618 T: neighbouring item ids
// land in tiles far apart (different splits), and the sampled half (slots 0-63) alternates from one id to the next.  With items in id
// order a trained model puts most of a user's best items — popular items with small, adjacent ids — into a few
// 64-item blocks of one split; each block contributes ONE maximum, tau came out far too low and 70 % of the rows
// overflowed their lists (real gowalla after 10 epochs).  The interleaved layout makes a block a spread-out sample.
// A permuted copy of the item table is made per call (m_items * 256 B, a few us); the train-item mask has to be
// given in the same position space, sorted per row (lgcn_score_topk_tc_item_positions + lgcn_csr_build).
//
// Only d = 64, k <= 24 and m_items >= 16384 take this path.
#include "common.cuh"
#include <cuda.h>
#include <float.h>
#include <stdlib.h>

namespace lgcn {

constexpr int TC_M = 128;            // rows per MMA
constexpr int TC_RT = 2;             // row tiles per CTA (both consume the same B tile from shared memory)
constexpr int TC_ROWS = TC_M * TC_RT;
constexpr int TC_N = 128;            // items per tile
constexpr int TC_D = 64;             // embedding width handled by this path
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_WARPS = 16;     // lane quarter = warp % 4 (hardware rule) x 2 row tiles x 2 column halves of the item tile
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
constexpr int TC_HALF = TC_N / 2;    // items per epilogue warp and tile
constexpr int TC_ATOM_K = 32;        // fp32 elements per 128-byte swizzle atom row
constexpr int TC_A_ATOM_BYTES = TC_M * 128;                 // 16 KB: one row tile, one K atom
constexpr int TC_A_BYTES = TC_RT * 2 * TC_A_ATOM_BYTES;     // 64 KB
constexpr int TC_B_ATOM_BYTES = TC_N * 128;                 // 16 KB
constexpr int TC_B_STAGE_BYTES = 2 * TC_B_ATOM_BYTES;       // 32 KB
constexpr int TC_STG = 16;                                  // pass 1: maxima staged per row before a flush (64-byte row segments)
constexpr int TC_STG_FLOATS = TC_EPI_WARPS * 32 * TC_STG;   // per epilogue warp 32 rows x TC_STG
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_A_BYTES + TC_STAGES * TC_B_STAGE_BYTES + TC_STG_FLOATS * 4 + 256;
constexpr int TC_CAP = 32;           // hit events (4 neighbouring scores each) per (row, split, column half)
constexpr int TC_KSEL = 28;          // tau = KSEL-th largest maximum of the 50 % sample: 56 +- 7.5 items reach it (k <= 24)
constexpr int TC_MAX_SPLITS = 16;
constexpr int TC_MIN_ITEMS = 16384;  // >= 128 tiles: the holes of the interleaved layout all fall into block 127, and tau needs KSEL tiles
constexpr float TC_EPS_C = 0.0025f;  // > 2^-9 (both operands truncated to 10 mantissa bits) + accumulation slack

struct TcOrder { int T, mul, inv; };       // item layout, see tc_pos_of_item: inv = mul^-1 mod T

struct TcArgs {
    int Bt; int m_items;
    TcOrder order;
    int hole_from_residue;               // tiles of residue >= this have no item of block 127 (128*T - m_items < 128 holes): slot 127 (even residue) or 63 (odd)
    const long long* users;
    const int* mask_indptr; const int* mask_indices; int mask_col_offset;
    int tiles_per_split; int n_splits;
    float* tile_max; int tile_stride;    // pass 1 out: [bt_pad][tile_stride]: maxima over the first 64 items of each tile
    const float* tau;                    // pass 2 in : [bt_pad]
    float4* cand_val; int* cand_idx;     // pass 2 out: [Bt][2*n_splits][TC_CAP] events: 4 scores | first item id + (masked bits << 28)
    int* cand_cnt;                       //             [Bt][2*n_splits]  (-1: the list overflowed)
};

// position (tile * 128 + slot) <-> item id.  T = number of item tiles; item i belongs to "residue" r = i % T and block
// b = i / T; it sits in tile (r * mul) % T (mul coprime to T, about 0.618 T: neighbouring ids land far apart, so a
// run of popular ids is spread over all the splits of the tile range) and slot (b >> 1) + 64 * ((b + r) & 1): the
// sampled half of a tile (slots 0-63) is the items with b + r even, so whether an item is sampled alternates from one
// id to the next (with (b & 1) alone it was constant over runs of T ids — a model whose best items all had ids in
// [T, 2T) would have had none of them in the sample).  Block 127 is the only one with holes (ids >= m_items): slot 127
// in the tiles of even residue, slot 63 in those of odd residue.
__host__ __device__ __forceinline__ int tc_pos_of_item(int i, TcOrder o) {
    const int b = i / o.T, r = i - b * o.T;
    return (int)(((long long)r * o.mul) % o.T) * TC_N + (b >> 1) + 64 * ((b ^ r) & 1);
}
__host__ __device__ __forceinline__ int tc_residue_of_tile(int tile, TcOrder o) { return (int)(((long long)tile * o.inv) % o.T); }
__host__ __device__ __forceinline__ int tc_item_of_pos(int p, TcOrder o) {
    const int s = p & (TC_N - 1), r = tc_residue_of_tile(p >> 7, o);
    return (2 * (s & 63) + (((s >> 6) ^ r) & 1)) * o.T + r;
}
static TcOrder tc_order(int m_items) {
    TcOrder o; o.T = (m_items + TC_N - 1) / TC_N;
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    o.mul = (int)(0.6180339887 * o.T); if (o.mul < 1) o.mul = 1;
    while (gcd(o.mul, o.T) != 1) ++o.mul;
    long long r0 = o.T, r1 = o.mul, t0 = 0, t1 = 1;                 // extended Euclid: t1 * mul == r1 (mod T)
    while (r1 != 1) { const long long q = r0 / r1, r2 = r0 - q * r1, t2 = t0 - q * t1; r0 = r1; r1 = r2; t0 = t1; t1 = t2; }
    o.inv = (int)(((t1 % o.T) + o.T) % o.T);
    return o;
}

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
    for (unsigned spin = 0; !done; ++spin) {
        // the suspend-time hint lets a waiting warp sleep in hardware instead of spinning through issue slots
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (spin > (1u << 24)) __trap();           // a protocol bug must fail, not hang the GPU
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
// 32 lanes x 32 consecutive columns; the load is only ISSUED here — tmem_ld_wait() makes the registers valid
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
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
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ float tc_max32(const uint32_t (&r)[32]) {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        m[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
    return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}
// columns named in `bad` (bit i = column i of the chunk) take no part
__device__ __forceinline__ void tc_kill_columns(uint32_t (&r)[32], unsigned bad) {
#pragma unroll
    for (int i = 0; i < 32; ++i) if ((bad >> i) & 1u) r[i] = __float_as_uint(-FLT_MAX);
}
// ties the registers of an in-flight tcgen05.ld to the preceding tcgen05.wait::ld for the compiler (no instruction)
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// the row's train items ("mask"), walked in step with the item tiles: bits of the current 128-item tile
struct TcMaskCursor {
    const int* row; int cur, end, next, next2, off;
    __device__ __forceinline__ void init(const TcArgs& a, bool live, int grow, int first_col) {
        row = nullptr; cur = 0; end = 0; next = 0x7fffffff; next2 = 0x7fffffff; off = a.mask_col_offset;
        if (!(live && a.mask_indptr)) return;
        const long long my_user = a.users ? a.users[grow] : (long long)grow;
        const int lo = __ldg(a.mask_indptr + my_user), hi = __ldg(a.mask_indptr + my_user + 1);
        row = a.mask_indices + lo; end = hi - lo;
        const int first_item = first_col + off;                            // first column id this CTA scores
        int l = 0, hh = end;                                               // lower_bound into the sorted row
        while (l < hh) { const int mid = (l + hh) >> 1; if (__ldg(row + mid) < first_item) l = mid + 1; else hh = mid; }
        cur = l;
        if (cur < end) next = __ldg(row + cur) - off;
        if (cur + 1 < end) next2 = __ldg(row + cur + 1) - off;
    }
    // bits [0,64) and [64,128) of the tile starting at item ib; the entry after next is always already in flight
    __device__ __forceinline__ void tile_bits(int ib, unsigned long long& b0, unsigned long long& b1) {
        b0 = 0ull; b1 = 0ull;
        while (next < ib + TC_N) {
            const int o = next - ib;
            if (o >= 64) b1 |= 1ull << (o - 64); else if (o >= 0) b0 |= 1ull << o;
            ++cur; next = next2;
            next2 = (cur + 1 < end) ? __ldg(row + cur + 1) - off : 0x7fffffff;
        }
    }
};

// pass 2: compare one 32-column chunk against the row's fixed threshold, 4 scores per compare.  A hit (rare per lane,
// but some lane of the warp hits in every few groups) costs two stores: the 4 scores and where they came from —
// which of them count is sorted out by the rescore kernel, so that the divergent path stays a handful of instructions.
struct TcHits { float4* val; int* idx; int cnt; float tau; };
__device__ __forceinline__ void tc_collect(const uint32_t (&r)[32], int i0, unsigned bad, TcHits& hs) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float v0 = __uint_as_float(r[4 * g4]), v1 = __uint_as_float(r[4 * g4 + 1]);
        const float v2 = __uint_as_float(r[4 * g4 + 2]), v3 = __uint_as_float(r[4 * g4 + 3]);
        if (fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) >= hs.tau) {
            // (events whose only hits are train items are stored too — rescore drops them; testing for that here was
            // measured at +70 us per pass on the divergent path, for nothing: the lists did not overflow either way)
            if (hs.cnt < TC_CAP) {
                hs.val[hs.cnt] = make_float4(v0, v1, v2, v3);
                hs.idx[hs.cnt] = (i0 + 4 * g4) | (int)(((bad >> (4 * g4)) & 0xfu) << 28);
            }
            ++hs.cnt;                                                 // > TC_CAP at the end: the row is given up (flagged by rescore)
        }
    }
}

template <int PASS>
__global__ void __launch_bounds__(TC_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024-B alignment
    uint8_t* sA = smem;                                    // [row tile][K atom][128 rows][128 B]
    uint8_t* sB = sA + TC_A_BYTES;                         // [stage][K atom][128 items][128 B]
    float* stg = reinterpret_cast<float*>(sB + TC_STAGES * TC_B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + TC_STG_FLOATS);
    // bars: 0 a_full | 1..4 b_full | 5..8 b_empty | 9,10 tmem_full | 11,12 tmem_empty ; then the TMEM base address
    constexpr int B_FULL = 1, B_EMPTY = 1 + TC_STAGES, T_FULL = 1 + 2 * TC_STAGES, T_EMPTY = 3 + 2 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T_EMPTY + 2);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ub = blockIdx.x * TC_ROWS;
    const int n_item_tiles = (a.m_items + TC_N - 1) / TC_N;
    const int t_begin = blockIdx.y * a.tiles_per_split;
    const int t_end = min(n_item_tiles, t_begin + a.tiles_per_split);
    const int n_tiles = t_end - t_begin;
    // tiles this launch processes: all of them (pass 2), or those with an even GLOBAL index (pass 1: the sample)
    const int it_first = (PASS == 1) ? (t_begin & 1) : 0, it_step = (PASS == 1) ? 2 : 1;
    const int n_proc = (n_tiles > it_first) ? (n_tiles - it_first + it_step - 1) / it_step : 0;

    if (threadIdx.x == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        for (int c = 0; c < 2; ++c) { mbar_init(BAR(T_FULL + c), 1); mbar_init(BAR(T_EMPTY + c), 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;                 // accumulator (buffer b, row tile rt) = columns (2b + rt) * 128 ...

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(BAR(0), TC_A_BYTES);
            for (int rt = 0; rt < TC_RT; ++rt)
                for (int ka = 0; ka < 2; ++ka)
                    tma_load_2d(smem_u32(sA + (rt * 2 + ka) * TC_A_ATOM_BYTES), &map_a, BAR(0), ka * TC_ATOM_K, ub + rt * TC_M);
            for (int j = 0; j < n_proc; ++j) {
                const int it = it_first + j * it_step;
                const int s = j % TC_STAGES, r = j / TC_STAGES;
                mbar_wait(BAR(B_EMPTY + s), (r & 1) ^ 1);
                mbar_expect_tx(BAR(B_FULL + s), TC_B_STAGE_BYTES);
                const int ib = (t_begin + it) * TC_N;
                uint8_t* dst = sB + s * TC_B_STAGE_BYTES;
                tma_load_2d(smem_u32(dst), &map_b, BAR(B_FULL + s), 0, ib);
                tma_load_2d(smem_u32(dst + TC_B_ATOM_BYTES), &map_b, BAR(B_FULL + s), TC_ATOM_K, ib);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop, so that descriptors and barrier addresses are computed in uniform registers; only
        // tcgen05.mma / tcgen05.commit are issued by one lane.  (Inside `if (lane == 0)` every MMA carries a chain of vector ALU
        // ops + R2UR, and this warp shares its scheduler with four epilogue warps.)
        const bool leader = (lane == 0);
        mbar_wait(BAR(0), 0);
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(sA));
        for (int it = 0; it < n_proc; ++it) {                      // it = index among the processed tiles
            const int s = it % TC_STAGES, r = it / TC_STAGES, acc = it & 1, ra = it >> 1;
            mbar_wait(BAR(B_FULL + s), r & 1);                     // B tile landed
            mbar_wait(BAR(T_EMPTY + acc), (ra & 1) ^ 1);           // accumulator pair drained by the epilogue
            tc_fence_after();
            const uint64_t db0 = umma_desc_k_sw128(smem_u32(sB + s * TC_B_STAGE_BYTES));
#pragma unroll
            for (int rt = 0; rt < TC_RT; ++rt) {
#pragma unroll
                for (int j = 0; j < TC_D / 8; ++j) {
                    // K step j: 32 bytes further inside the swizzle atom, second atom for j >= 4; bases are 1024-aligned and
                    // below 256 KB, so the offset can be added to the encoded start address without a carry
                    const uint64_t da = da0 + (uint64_t)(((rt * 2 + (j >> 2)) * TC_A_ATOM_BYTES + (j & 3) * 32) >> 4);
                    const uint64_t db = db0 + (uint64_t)(((j >> 2) * TC_B_ATOM_BYTES + (j & 3) * 32) >> 4);
                    if (leader) tc_mma_tf32(tmem_base + (acc * TC_RT + rt) * TC_N, da, db, TC_IDESC, j > 0 ? 1u : 0u);
                }
            }
            if (leader) {
                tc_commit(BAR(B_EMPTY + s));                       // smem stage free when these MMAs retire
                tc_commit(BAR(T_FULL + acc));                      // accumulator pair ready
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: a thread owns one row; pass 2: one 64-item half of every tile, pass 1: the first half of every
        // other tile (group g = e / 8 takes the tiles with it % 2 == g, i.e. always the same accumulator buffer) =================
        const int e = warp - 2, q = warp & 3, rt = (e >> 2) & 1, grp = e >> 3;    // TMEM lane quarter q = warp % 4 is a hardware rule
        const int half = grp;
        const int row = ub + rt * TC_M + q * 32 + lane;
        const bool live = row < a.Bt;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rt * TC_N + half * TC_HALF);
        uint32_t ra_[32], rb_[32];

        if (PASS == 3) {                                           // debug: MMA-only pacing (accumulators released unread)
            for (int it = 0; it < n_proc; ++it) {
                const int acc = it & 1, rph = it >> 1;
                mbar_wait(BAR(T_FULL + acc), rph & 1);
                tc_fence_after();
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + acc));
            }
        } else         if (PASS == 1) {
            // ---- maxima over the warp's 64-item half of every SAMPLED tile, staged TC_STG at a time so that the global
            // stores are row segments; entry (global tile + half) of the row: consecutive staged entries are 2 apart ----
            float* my = stg + e * (32 * TC_STG);                   // slot(row l, p) = l*TC_STG + ((p + l) % TC_STG)
            int nbuf = 0;
            float* out_rows = a.tile_max + (size_t)(ub + rt * TC_M + q * 32) * a.tile_stride;
            auto flush = [&](int t0) {
                __syncwarp();
                for (int rr = (lane >> 4); rr < 32; rr += 2) {                       // two rows per step, 16 lanes each
                    const int p = lane & (TC_STG - 1);
                    if (p < nbuf) out_rows[(size_t)rr * a.tile_stride + t0 + 2 * p] = my[rr * TC_STG + ((p + rr) & (TC_STG - 1))];
                }
                __syncwarp();
            };
            TcMaskCursor mc; mc.init(a, live, row, t_begin * TC_N);
            int first_staged = t_begin + it_first + half;
            for (int j = 0; j < n_proc; ++j) {
                const int it = it_first + 2 * j;
                const int acc = j & 1, rph = j >> 1;
                const int ib = (t_begin + it) * TC_N;
                unsigned long long mb0, mb1;
                mc.tile_bits(ib, mb0, mb1);                        // before the wait: overlaps the MMA of this tile (entries of the skipped tile are dropped)
                const unsigned long long mb = half ? mb1 : mb0;
                mbar_wait(BAR(T_FULL + acc), rph & 1);
                tc_fence_after();
                const uint32_t tb = t_row + (uint32_t)(acc * TC_RT * TC_N);
                tmem_ld32_issue(tb, ra_); tmem_ld32_issue(tb + 32, rb_);
                tmem_ld_wait(ra_); tmem_ld_fence(rb_);
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + acc));                   // accumulator free: this warp's part is in registers
                // train items of the row take no part; block 127 is missing in this tile: slot 127 (half 1) if the residue is even,
                // slot 63 (half 0) if it is odd
                const int res1 = tc_residue_of_tile(t_begin + it, a.order);
                const unsigned hole1 = (res1 >= a.hole_from_residue && ((res1 & 1) ^ half)) ? 0x80000000u : 0u;
                const unsigned bad0 = (unsigned)mb, bad1 = (unsigned)(mb >> 32) | hole1;
                if (bad0) tc_kill_columns(ra_, bad0);
                if (bad1) tc_kill_columns(rb_, bad1);
                my[lane * TC_STG + ((nbuf + lane) & (TC_STG - 1))] = fmaxf(tc_max32(ra_), tc_max32(rb_));
                if (++nbuf == TC_STG) { flush(first_staged); nbuf = 0; first_staged = t_begin + it + 2 + half; }
            }
            if (nbuf) flush(first_staged);
        } else {
            // ---- fixed threshold: collect everything >= tau that is not a train item ----
            TcMaskCursor mc; mc.init(a, live, row, t_begin * TC_N);
            TcHits hs;
            hs.cnt = 0; hs.tau = live ? a.tau[row] : FLT_MAX;
            const size_t list = live ? ((size_t)row * a.n_splits + blockIdx.y) * 2 + half : 0;
            hs.val = a.cand_val + list * TC_CAP; hs.idx = a.cand_idx + list * TC_CAP;
            int residue = tc_residue_of_tile(t_begin, a.order);
            for (int it = 0; it < n_tiles; ++it) {
                const int acc = it & 1, rph = it >> 1;
                const int ib = (t_begin + it) * TC_N;
                unsigned long long mb0, mb1;
                mc.tile_bits(ib, mb0, mb1);
                const unsigned long long mb = half ? mb1 : mb0;
                mbar_wait(BAR(T_FULL + acc), rph & 1);
                tc_fence_after();
                const uint32_t tb = t_row + (uint32_t)(acc * TC_RT * TC_N);
                const int i0 = ib + half * TC_HALF;
                tmem_ld32_issue(tb, ra_); tmem_ld32_issue(tb + 32, rb_);
                tmem_ld_wait(ra_); tmem_ld_fence(rb_);
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + acc));
                // block 127 is missing in this tile: slot 127 (half 1) if the residue is even, slot 63 (half 0) if it is odd
                const unsigned hole = (residue >= a.hole_from_residue && ((residue & 1) ^ half)) ? 0x80000000u : 0u;
                residue += a.order.inv; if (residue >= a.order.T) residue -= a.order.T;          // residue of the next tile
                tc_collect(ra_, i0, (unsigned)mb, hs);
                tc_collect(rb_, i0 + 32, (unsigned)(mb >> 32) | hole, hs);
            }
            if (live) a.cand_cnt[list] = (hs.cnt > TC_CAP) ? -1 : hs.cnt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
    }
}

// A operand: the batch's user rows, gathered and zero-padded to a multiple of 256 rows
__global__ void gather_rows_kernel(const float4* __restrict__ U, const long long* __restrict__ users, int Bt, int Bt_pad, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // one float4 per thread, 16 per row (d = 64)
    if (i >= Bt_pad * 16) return;
    const int r = i >> 4, c = i & 15;
    float4 v = f4_zero();
    if (r < Bt) { const long long u = users ? users[r] : (long long)r; v = __ldg(U + (size_t)u * 16 + c); }
    out[i] = v;
}

// B operand: the item table in the interleaved layout (position p holds item tc_item_of_pos(p); holes are zero rows)
__global__ void permute_items_kernel(const float4* __restrict__ V, int m_items, TcOrder o, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // one float4 per thread, 16 per row (d = 64)
    if (i >= o.T * TC_N * 16) return;
    const int p = i >> 4, c = i & 15;
    const int item = tc_item_of_pos(p, o);
    out[i] = (item < m_items) ? __ldg(V + (size_t)item * 16 + c) : f4_zero();
}

__global__ void tc_item_positions_kernel(const long long* __restrict__ items, long long n, TcOrder o, long long* __restrict__ pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = tc_pos_of_item((int)items[i], o);
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

// order-preserving float -> uint (larger float <=> larger key)
__device__ __forceinline__ unsigned tc_key(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float tc_unkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// select: warp per row — tau = KSEL-th largest sample maximum, to 24 significant key bits (tau is the lower edge of
// the bucket the KSEL-th largest falls into, so at least KSEL sample maxima reach it; 16 bits were too coarse for
// trained models, whose best scores lie within a fraction of a percent of each other); radix select, 3 x 8 bits
constexpr int SEL_WARPS = 4;
__global__ void __launch_bounds__(SEL_WARPS * 32)
tc_select_kernel(int Bt, int n_item_tiles, const float* __restrict__ tile_max, int tile_stride, int ksel, float* __restrict__ tau_out) {
    __shared__ int s_hist[SEL_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * SEL_WARPS + w;
    if (b >= Bt) return;
    const float* M0 = tile_max + (size_t)b * tile_stride;
    const unsigned dead = tc_key(-FLT_MAX);                               // nothing but train items in the sample
    unsigned prefix = 0; int remaining = ksel; bool enough = true;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = lane; i < 256; i += 32) s_hist[w][i] = 0;
        __syncwarp();
        const unsigned hi_mask = pass ? (0xffffffffu << (shift + 8)) : 0u;
        for (int c = lane; c < n_item_tiles; c += 32) {
            const unsigned key = tc_key(__ldg(M0 + c));
            if (key != dead && ((key ^ prefix) & hi_mask) == 0u) atomicAdd(&s_hist[w][(key >> shift) & 255u], 1);
        }
        __syncwarp();
        // lane l owns bins 255-8l .. 248-8l (descending); inclusive scan over lanes from the top
        int mine = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mine += s_hist[w][255 - 8 * lane - i];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const unsigned hit = __ballot_sync(0xffffffffu, incl >= remaining);
        if (hit == 0u) { enough = false; break; }                        // fewer than ksel usable tiles
        const int L = __ffs(hit) - 1;
        int digit = 0, rem2 = 0;
        if (lane == L) {
            int before = incl - mine;
            for (int i = 0; i < 8; ++i) {
                const int bin = 255 - 8 * lane - i, cnt = s_hist[w][bin];
                if (before + cnt >= remaining) { digit = bin; rem2 = remaining - before; break; }
                before += cnt;
            }
        }
        digit = __shfl_sync(0xffffffffu, digit, L); remaining = __shfl_sync(0xffffffffu, rem2, L);
        prefix |= (unsigned)digit << shift;
        __syncwarp();
    }
    // smallest float whose key starts with the 24 selected bits (for negative values the low key bits run the other way)
    if (lane == 0) tau_out[b] = enough ? tc_unkey(prefix) : -FLT_MAX;
}

// rescore: warp per row — exact rescoring of every candidate, top-k by rank counting, certificate
constexpr int RS_WARPS = 4;
constexpr int RS_CAP = 192;
constexpr int RS_BATCH = 16;         // item rows staged per step (32 was measured: +8 % on the whole call — fewer resident warps — profiles/r2_eval_probe_rescore_batch32.jsonl)
__global__ void __launch_bounds__(RS_WARPS * 32)
rescore_kernel(const float* __restrict__ U, const float* __restrict__ V, const long long* __restrict__ users, int Bt, int n_lists, int k, TcOrder order,
               const float4* __restrict__ cand_val, const int* __restrict__ cand_idx, const int* __restrict__ cand_cnt,
               const float* __restrict__ tau_row, const int* __restrict__ vmax_bits,
               long long* __restrict__ idx_out, float* __restrict__ val_out, int* __restrict__ flags, int* __restrict__ n_flagged) {
    __shared__ unsigned long long s_key[RS_WARPS][RS_CAP];           // (ordered exact score << 32) | (0x7fffffff - item id): larger is better
    __shared__ int s_id[RS_WARPS][RS_CAP];
    __shared__ int s_pre[RS_WARPS][32];
    __shared__ __align__(16) float s_u[RS_WARPS][TC_D];
    __shared__ __align__(16) float s_rows[RS_WARPS][RS_BATCH * 68];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * RS_WARPS + w;
    if (b >= Bt) return;
    const long long u = users ? users[b] : (long long)b;
    const float* ur = U + (size_t)u * TC_D;
    float un2 = 0.f;
    for (int c = lane; c < TC_D; c += 32) { const float x = __ldg(ur + c); s_u[w][c] = x; un2 += x * x; }
    for (int o = 16; o > 0; o >>= 1) un2 += __shfl_xor_sync(0xffffffffu, un2, o);
    const float eps = TC_EPS_C * sqrtf(un2) * __int_as_float(*vmax_bits);
    const float tau = tau_row[b];
    // the row's hit events live in n_lists (<= 32) lists; flatten them with a prefix over the list lengths
    const int my_cnt = (lane < n_lists) ? cand_cnt[(size_t)b * n_lists + lane] : 0;
    bool lost = __any_sync(0xffffffffu, my_cnt < 0);
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int T = __shfl_sync(0xffffffffu, incl, 31);
    s_pre[w][lane] = incl - my_cnt;
    __syncwarp();
    int n = 0;
    if (!lost) {
        for (int f0 = 0; f0 < T; f0 += 32) {
            const int f = f0 + lane;
            const bool have = f < T;
            int li = 0;                                                // last list whose first event is <= f
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) if (s_pre[w][li + st] <= f) li += st;
            float4 v = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX); int meta = 0;
            if (have) { const size_t at = ((size_t)b * n_lists + li) * TC_CAP + (f - s_pre[w][li]); v = cand_val[at]; meta = cand_idx[at]; }
            const float vv[4] = {v.x, v.y, v.z, v.w};
            const int first = meta & 0x0fffffff;                       // position of the first of the 4 scores
#pragma unroll
            for (int x = 0; x < 4; ++x) {                              // keep the scores that reach tau and are not train items
                const bool take = have && vv[x] >= tau && !((meta >> (28 + x)) & 1);
                const unsigned m = __ballot_sync(0xffffffffu, take);
                const int pos = n + __popc(m & ((1u << lane) - 1u));
                if (take && pos < RS_CAP) s_id[w][pos] = tc_item_of_pos(first + x, order);
                n += __popc(m);
            }
        }
        if (n > RS_CAP) lost = true;
    }
    if (lost || n < k) {                                             // the exact kernel redoes this row (flag = why: 3 list overflow, 2 too few)
        if (lane == 0) { flags[b] = lost ? 3 : 2; atomicAdd(n_flagged, 1); }
        return;
    }
    __syncwarp();
    // exact scores, RS_BATCH candidates at a time: the item rows are staged through shared memory with coalesced loads
    // (two rows per warp instruction); a lane then runs the FMA chain of the exact contract over its own candidate's
    // row.  Rows are 68 floats apart: 16-byte accesses of 8 consecutive lanes hit 32 different banks.
    const float4* V4 = reinterpret_cast<const float4*>(V);
    for (int c0 = 0; c0 < n; c0 += RS_BATCH) {
        const int nb = min(RS_BATCH, n - c0);
#pragma unroll
        for (int t = 0; t < RS_BATCH / 2; ++t) {
            const int r = 2 * t + (lane >> 4);
            if (r < nb) *reinterpret_cast<float4*>(&s_rows[w][r * 68 + 4 * (lane & 15)]) = __ldg(V4 + (size_t)s_id[w][c0 + r] * 16 + (lane & 15));
        }
        __syncwarp();
        if (lane < nb) {
            const float* vr = &s_rows[w][lane * 68];
            float sc = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < TC_D / 4; ++c4) {                   // one fp32 FMA chain in k order
                const float4 x = *reinterpret_cast<const float4*>(vr + 4 * c4);
                const float4 y = *reinterpret_cast<const float4*>(&s_u[w][4 * c4]);
                sc = fmaf(y.x, x.x, sc); sc = fmaf(y.y, x.y, sc); sc = fmaf(y.z, x.z, sc); sc = fmaf(y.w, x.w, sc);
            }
            s_key[w][c0 + lane] = ((unsigned long long)tc_key(sc) << 32) | (unsigned)(0x7fffffff - s_id[w][c0 + lane]);
        }
        __syncwarp();
    }
    // rank of a candidate = how many others beat it (score desc, item id asc); ranks < k are the answer, in place
    unsigned long long kth_key = 0ull;
    for (int c0 = 0; c0 < n; c0 += 64) {
        const int ca = c0 + lane, cb = c0 + 32 + lane;
        const unsigned long long ka = ca < n ? s_key[w][ca] : 0ull, kb = cb < n ? s_key[w][cb] : 0ull;
        int rka = 0, rkb = 0;
        for (int j = 0; j < n; ++j) {
            const unsigned long long o = s_key[w][j];                 // broadcast read
            rka += (o > ka) ? 1 : 0; rkb += (o > kb) ? 1 : 0;
        }
        if (ca < n && rka < k) { idx_out[(size_t)b * k + rka] = 0x7fffffff - (int)(unsigned)ka; val_out[(size_t)b * k + rka] = tc_unkey((unsigned)(ka >> 32)); }
        if (cb < n && rkb < k) { idx_out[(size_t)b * k + rkb] = 0x7fffffff - (int)(unsigned)kb; val_out[(size_t)b * k + rkb] = tc_unkey((unsigned)(kb >> 32)); }
        if (ca < n && rka == k - 1) kth_key = ka;
        if (cb < n && rkb == k - 1) kth_key = kb;
    }
    unsigned kth_hi = (unsigned)(kth_key >> 32);
    for (int o = 16; o > 0; o >>= 1) kth_hi = max(kth_hi, __shfl_xor_sync(0xffffffffu, kth_hi, o));
    if (lane == 0) {
        // everything that is not a candidate is a train item or has approx < tau  =>  exact < tau + eps
        const float kth = tc_unkey(kth_hi);
        const bool ok = (tau == -FLT_MAX || tau + eps < kth);
        flags[b] = ok ? 0 : 1;
        if (!ok) atomicAdd(n_flagged, 1);
    }
}

// =====================================================================================================================
// Dense scores on tensor cores: getUsersRating (reference code/model.py:114-123, torch.matmul(u_emb, i_emb.t())).
// The caller gets the Bt x M matrix, so nothing can be filtered and the accuracy has to be fp32's: 3xTF32.  Every operand is
// split exactly into hi (the 10 mantissa bits tcgen05.mma.kind::tf32 reads) and lo = x - hi; the product is
// lo_a*hi_b + hi_a*lo_b + hi_a*hi_b, small terms first, all accumulated in the fp32 TMEM accumulator (the dropped lo*lo
// term is below 2^-20 of |a||b|).  CTA = 128 user rows x a split of the 128-item tiles (natural item order); warp 0 = TMA
// producer (2-stage ring, a stage = hi and lo of one item tile), warp 1 = MMA issuer (24 MMAs M128 N128 K8 per tile,
// double-buffered accumulators), warps 2-9 = epilogue: TMEM -> registers -> shared-memory transpose -> coalesced row
// segments of the output.  The work is bound by writing Bt x M x 4 bytes; the exact CUDA-core kernel it replaces for this
// call was compute-bound at 6.5-11 TFLOP/s (slower than cuBLAS SGEMM).
constexpr int DT_STAGES = 2;
constexpr int DT_EPI_WARPS = 8;                              // lane quarter (warp % 4) x 2 column halves
constexpr int DT_THREADS = 64 + 32 * DT_EPI_WARPS;
constexpr int DT_A_BYTES = 2 * 2 * TC_A_ATOM_BYTES;          // hi, lo x 2 K atoms x 16 KB = 64 KB
constexpr int DT_B_STAGE_BYTES = 2 * TC_B_STAGE_BYTES;       // hi + lo of one tile = 64 KB
constexpr int DT_STG_FLOATS = DT_EPI_WARPS * 32 * 33;        // per epilogue warp: 32 rows x 32 columns (+1 pad)
constexpr int DT_SMEM_BYTES = 1024 + DT_A_BYTES + DT_STAGES * DT_B_STAGE_BYTES + DT_STG_FLOATS * 4 + 256;

struct DtArgs { int Bt; int m_items; int tiles_per_split; float* out; };

__global__ void __launch_bounds__(DT_THREADS, 1)
score_dense_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                      const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, const __grid_constant__ DtArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                    // [hi|lo][K atom][128 rows][128 B]
    uint8_t* sB = sA + DT_A_BYTES;                         // [stage][hi|lo][K atom][128 items][128 B]
    float* stg = reinterpret_cast<float*>(sB + DT_STAGES * DT_B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + DT_STG_FLOATS);
    constexpr int B_FULL = 1, B_EMPTY = 1 + DT_STAGES, T_FULL = 1 + 2 * DT_STAGES, T_EMPTY = 3 + 2 * DT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T_EMPTY + 2);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ub = blockIdx.x * TC_M;
    const int n_item_tiles = (a.m_items + TC_N - 1) / TC_N;
    const int t_begin = blockIdx.y * a.tiles_per_split;
    const int n_tiles = min(n_item_tiles, t_begin + a.tiles_per_split) - t_begin;

    if (threadIdx.x == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < DT_STAGES; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        for (int c = 0; c < 2; ++c) { mbar_init(BAR(T_FULL + c), 1); mbar_init(BAR(T_EMPTY + c), 32 * DT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(BAR(0), DT_A_BYTES);
            for (int ka = 0; ka < 2; ++ka) {
                tma_load_2d(smem_u32(sA + ka * TC_A_ATOM_BYTES), &map_ah, BAR(0), ka * TC_ATOM_K, ub);
                tma_load_2d(smem_u32(sA + (2 + ka) * TC_A_ATOM_BYTES), &map_al, BAR(0), ka * TC_ATOM_K, ub);
            }
            for (int it = 0; it < n_tiles; ++it) {
                const int s = it % DT_STAGES, r = it / DT_STAGES;
                mbar_wait(BAR(B_EMPTY + s), (r & 1) ^ 1);
                mbar_expect_tx(BAR(B_FULL + s), DT_B_STAGE_BYTES);
                const int ib = (t_begin + it) * TC_N;
                uint8_t* dst = sB + s * DT_B_STAGE_BYTES;
                for (int ka = 0; ka < 2; ++ka) {
                    tma_load_2d(smem_u32(dst + ka * TC_B_ATOM_BYTES), &map_bh, BAR(B_FULL + s), ka * TC_ATOM_K, ib);
                    tma_load_2d(smem_u32(dst + (2 + ka) * TC_B_ATOM_BYTES), &map_bl, BAR(B_FULL + s), ka * TC_ATOM_K, ib);
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = (lane == 0);
        mbar_wait(BAR(0), 0);
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(sA));
        for (int it = 0; it < n_tiles; ++it) {
            const int s = it % DT_STAGES, r = it / DT_STAGES, acc = it & 1, ra = it >> 1;
            mbar_wait(BAR(B_FULL + s), r & 1);
            mbar_wait(BAR(T_EMPTY + acc), (ra & 1) ^ 1);
            tc_fence_after();
            const uint64_t db0 = umma_desc_k_sw128(smem_u32(sB + s * DT_B_STAGE_BYTES));
#pragma unroll
            for (int term = 0; term < 3; ++term) {               // lo*hi, hi*lo, then hi*hi
                const int a_part = (term == 0) ? 1 : 0, b_part = (term == 1) ? 1 : 0;
#pragma unroll
                for (int j = 0; j < TC_D / 8; ++j) {
                    const uint64_t da = da0 + (uint64_t)(((a_part * 2 + (j >> 2)) * TC_A_ATOM_BYTES + (j & 3) * 32) >> 4);
                    const uint64_t db = db0 + (uint64_t)(((b_part * 2 + (j >> 2)) * TC_B_ATOM_BYTES + (j & 3) * 32) >> 4);
                    if (leader) tc_mma_tf32(tmem_base + acc * TC_N, da, db, TC_IDESC, (term > 0 || j > 0) ? 1u : 0u);
                }
            }
            if (leader) { tc_commit(BAR(B_EMPTY + s)); tc_commit(BAR(T_FULL + acc)); }
            __syncwarp();
        }
    } else {
        const int e = warp - 2, q = warp & 3, half = e >> 2;         // TMEM lane quarter q = warp % 4 (hardware rule)
        float* my = stg + e * (32 * 33);
        uint32_t r_[32];
        for (int it = 0; it < n_tiles; ++it) {
            const int acc = it & 1, rph = it >> 1;
            const int ib = (t_begin + it) * TC_N + half * TC_HALF;
            mbar_wait(BAR(T_FULL + acc), rph & 1);
            tc_fence_after();
            const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_N + half * TC_HALF);
#pragma unroll 1
            for (int c32 = 0; c32 < 2; ++c32) {
                tmem_ld32_issue(tb + 32 * c32, r_);
                tmem_ld_wait(r_);
                if (c32 == 1) { tc_fence_before(); mbar_arrive(BAR(T_EMPTY + acc)); }   // everything of this tile is in registers
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 32; ++c) my[lane * 33 + c] = __uint_as_float(r_[c]);     // thread = row: conflict-free (stride 33)
                __syncwarp();
                const int col = ib + 32 * c32 + lane;
                if (col < a.m_items) {
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {                                          // lane = column: one 128-byte segment per row
                        const int row = ub + q * 32 + rr;
                        if (row < a.Bt) a.out[(size_t)row * a.m_items + col] = my[rr * 33 + lane];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256) : "memory");
    }
}

// x = hi + lo exactly, hi = x with the 13 low mantissa bits cleared (what kind::tf32 reads)
__global__ void split_tf32_kernel(const float4* __restrict__ in, long long n4, float4* __restrict__ hi, float4* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = in[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
    hi[i] = h; lo[i] = l;
}

// gathered, zero-padded user rows, split in one pass
__global__ void gather_split_rows_kernel(const float4* __restrict__ U, const long long* __restrict__ users, int Bt, int Bt_pad,
                                         float4* __restrict__ hi, float4* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Bt_pad * 16) return;
    const int r = i >> 4, c = i & 15;
    float4 x = f4_zero();
    if (r < Bt) { const long long u = users ? users[r] : (long long)r; x = __ldg(U + (size_t)u * 16 + c); }
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
    hi[i] = h; lo[i] = l;
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

// Item-tile splits per 256-row block: the grid (row blocks x splits, one CTA per SM) should fill whole waves;
// a split keeps >= 8 tiles so that the A load and the pipeline fill stay amortised.
static int tc_pick_splits(int Bt, int m_items) {
    const int row_blocks = (Bt + TC_ROWS - 1) / TC_ROWS, item_tiles = (m_items + TC_N - 1) / TC_N;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int best = 1; double best_eff = 0.0;
    for (int s = 1; s <= TC_MAX_SPLITS; ++s) {
        const int per = (item_tiles + s - 1) / s;
        if (s > 1 && per < 8) break;
        const int real = (item_tiles + per - 1) / per;
        const long long ctas = (long long)row_blocks * real;
        const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = real; }
    }
    return best;
}

struct TcLayout { int bt_pad, item_tiles, tile_stride, n_splits, tiles_per_split; size_t off_vp, off_mx, off_tau, off_cv, off_ci, off_cc, off_vmax, total; };
static TcLayout tc_layout(int Bt, int m_items) {
    TcLayout L;
    L.bt_pad = (Bt + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
    L.item_tiles = (m_items + TC_N - 1) / TC_N;

    L.n_splits = tc_pick_splits(Bt, m_items);
    L.tiles_per_split = (L.item_tiles + L.n_splits - 1) / L.n_splits;
    L.n_splits = (L.item_tiles + L.tiles_per_split - 1) / L.tiles_per_split;
    L.tile_stride = (L.item_tiles + 3) & ~1;          // sample maxima: entry (even tile + half), at most item_tiles + 1 of them
    size_t o = align_up((size_t)L.bt_pad * TC_D * 4, 1024);
    L.off_vp = o;   o += align_up((size_t)L.item_tiles * TC_N * TC_D * 4, 1024);
    L.off_mx = o;   o += align_up((size_t)L.bt_pad * L.tile_stride * 4, 256);
    L.off_tau = o;  o += align_up((size_t)L.bt_pad * 4, 256);
    L.off_cv = o;   o += align_up((size_t)Bt * 2 * L.n_splits * TC_CAP * 16, 256);
    L.off_ci = o;   o += align_up((size_t)Bt * 2 * L.n_splits * TC_CAP * 4, 256);
    L.off_cc = o;   o += align_up((size_t)Bt * 2 * L.n_splits * 4, 256);
    L.off_vmax = o; o += 256;
    L.total = o;
    return L;
}

}  // namespace lgcn

using namespace lgcn;

// workspace: [A gathered][items, interleaved layout][tile maxima][tau][cand_val][cand_idx][cand_cnt][vmax]
extern "C" size_t lgcn_score_topk_tc_workspace_bytes(int32_t Bt, int32_t m_items, int32_t k) {
    if (Bt <= 0 || m_items < TC_MIN_ITEMS || k <= 0) return 0;
    return tc_layout(Bt, m_items).total;
}

extern "C" int lgcn_score_topk_tc_supported(int32_t d, int32_t k) { return (d == TC_D && k >= 1 && k <= 24) ? 1 : 0; }

extern "C" int lgcn_score_topk_tc(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                                  int32_t m_items, int32_t d, const int32_t* mask_indptr, const int32_t* mask_indices,
                                  int32_t mask_col_offset, int32_t k, int64_t* idx_out, float* val_out,
                                  int32_t* flags_out, int32_t* n_flagged_out,
                                  void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(users_emb && items_emb && idx_out && val_out && flags_out && n_flagged_out, "score_topk_tc: null argument");
    LGCN_CHECK_ARG(lgcn_score_topk_tc_supported(d, k), "score_topk_tc: only d=%d and k<=24 take the tensor-core path", TC_D);
    LGCN_CHECK_ARG(Bt > 0 && m_items >= TC_MIN_ITEMS && m_items < (1 << 27), "score_topk_tc: Bt=%d m_items=%d (needs %d <= m_items < 2^27; use lgcn_score_topk)", Bt, m_items, TC_MIN_ITEMS);
    LGCN_CHECK_ARG((mask_indptr == nullptr) == (mask_indices == nullptr), "score_topk_tc: mask arrays must both be set or both null");
    const TcLayout L = tc_layout(Bt, m_items);
    const TcOrder order = tc_order(m_items);
    LGCN_CHECK_ARG(workspace && ((uintptr_t)workspace % 1024) == 0 && workspace_bytes >= L.total,
                   "score_topk_tc: workspace too small or not 1024-byte aligned");
    LGCN_CHECK_ARG(((uintptr_t)items_emb % 16) == 0 && ((uintptr_t)users_emb % 16) == 0, "score_topk_tc: tables must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    char* w = static_cast<char*>(workspace);
    float* A = reinterpret_cast<float*>(w);
    float* Vp = reinterpret_cast<float*>(w + L.off_vp);
    float* tile_max = reinterpret_cast<float*>(w + L.off_mx);
    float* tau = reinterpret_cast<float*>(w + L.off_tau);
    float4* cand_val = reinterpret_cast<float4*>(w + L.off_cv);
    int* cand_idx = reinterpret_cast<int*>(w + L.off_ci);
    int* cand_cnt = reinterpret_cast<int*>(w + L.off_cc);
    int* vmax = reinterpret_cast<int*>(w + L.off_vmax);
    const long long* users_ll = reinterpret_cast<const long long*>(users);

    gather_rows_kernel<<<(L.bt_pad * 16 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(users_emb), users_ll, Bt, L.bt_pad,
                                                                   reinterpret_cast<float4*>(A));
    LGCN_CHECK_LAUNCH("gather_rows_kernel");
    cudaMemsetAsync(vmax, 0, 16, st);
    cudaMemsetAsync(n_flagged_out, 0, sizeof(int32_t), st);
    permute_items_kernel<<<(L.item_tiles * TC_N * 16 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(items_emb), m_items, order,
                                                                                 reinterpret_cast<float4*>(Vp));
    LGCN_CHECK_LAUNCH("permute_items_kernel");
    item_norm_max_kernel<<<(m_items + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(items_emb), m_items, vmax);
    LGCN_CHECK_LAUNCH("item_norm_max_kernel");

    CUtensorMap map_a, map_b;
    if (int rc = make_map(&map_a, A, (uint64_t)L.bt_pad, TC_M)) return rc;
    if (int rc = make_map(&map_b, Vp, (uint64_t)L.item_tiles * TC_N, TC_N)) return rc;
    TcArgs a;
    a.Bt = Bt; a.m_items = m_items; a.users = users_ll;
    a.order = order;
    a.hole_from_residue = m_items - 127 * L.item_tiles;    // block 127 = items 127*T + r: present for the residues below this
    a.mask_indptr = mask_indptr; a.mask_indices = mask_indices; a.mask_col_offset = mask_col_offset;
    a.n_splits = L.n_splits; a.tiles_per_split = L.tiles_per_split;
    a.tile_max = tile_max; a.tile_stride = L.tile_stride; a.tau = tau;
    a.cand_val = cand_val; a.cand_idx = cand_idx; a.cand_cnt = cand_cnt;
    cudaError_t e = cudaFuncSetAttribute(score_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(score_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return fail("score_topk_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid(L.bt_pad / TC_ROWS, L.n_splits);
    score_tc_kernel<1><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(map_a, map_b, a);
    LGCN_CHECK_LAUNCH("score_tc_kernel<1>");
    // one maximum per 64-item half of every even tile: 2 * ceil(T / 2) entries per row
    tc_select_kernel<<<(Bt + SEL_WARPS - 1) / SEL_WARPS, SEL_WARPS * 32, 0, st>>>(Bt, 2 * ((L.item_tiles + 1) / 2), tile_max, L.tile_stride, TC_KSEL, tau);
    LGCN_CHECK_LAUNCH("tc_select_kernel");
    score_tc_kernel<2><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(map_a, map_b, a);
    LGCN_CHECK_LAUNCH("score_tc_kernel<2>");
    rescore_kernel<<<(Bt + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, st>>>(users_emb, items_emb, users_ll, Bt, 2 * L.n_splits, k, order,
        cand_val, cand_idx, cand_cnt, tau, vmax, reinterpret_cast<long long*>(idx_out), val_out, flags_out, n_flagged_out);
    LGCN_CHECK_LAUNCH("rescore_kernel");
    if (getenv("LGCN_TC_DEBUG_MMA_ONLY")) {
        cudaFuncSetAttribute(score_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        score_tc_kernel<3><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(map_a, map_b, a);
        LGCN_CHECK_LAUNCH("score_tc_kernel<3>");
    }
    return 0;
}

extern "C" int lgcn_score_topk_tc_item_positions(const int64_t* items, int64_t n, int32_t m_items, int64_t* pos_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(items && pos_out && n >= 0, "score_topk_tc_item_positions: null argument");
    LGCN_CHECK_ARG(m_items >= TC_MIN_ITEMS && m_items < (1 << 27), "score_topk_tc_item_positions: m_items=%d", m_items);
    if (n == 0) return 0;
    tc_item_positions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(items), n, tc_order(m_items),
                                                                                          reinterpret_cast<long long*>(pos_out));
    LGCN_CHECK_LAUNCH("tc_item_positions_kernel");
    return 0;
}

extern "C" int32_t lgcn_score_topk_tc_position_space(int32_t m_items) { return m_items >= TC_MIN_ITEMS ? ((m_items + TC_N - 1) / TC_N) * TC_N : 0; }

// diagnostics: where the per-row threshold and the per-list event counts live inside the workspace of the last call shape
extern "C" int lgcn_score_topk_tc_debug_layout(int32_t Bt, int32_t m_items, int64_t* out8_host) {
    LGCN_CHECK_ARG(out8_host && Bt > 0 && m_items >= TC_MIN_ITEMS, "score_topk_tc_debug_layout: bad arguments");
    const TcLayout L = tc_layout(Bt, m_items);
    out8_host[0] = (int64_t)L.off_tau; out8_host[1] = (int64_t)L.off_cc; out8_host[2] = 2 * L.n_splits; out8_host[3] = L.item_tiles;
    out8_host[4] = L.tiles_per_split; out8_host[5] = (int64_t)L.off_mx; out8_host[6] = L.tile_stride; out8_host[7] = TC_CAP;
    return 0;
}

// host-side view of the item layout (no device needed): position of an item, item at a position (-1 for a hole)
extern "C" int32_t lgcn_score_topk_tc_host_position(int32_t item, int32_t m_items) {
    if (m_items < TC_MIN_ITEMS || item < 0 || item >= m_items) return -1;
    return tc_pos_of_item(item, tc_order(m_items));
}
extern "C" int32_t lgcn_score_topk_tc_host_item(int32_t pos, int32_t m_items) {
    if (m_items < TC_MIN_ITEMS || pos < 0 || pos >= lgcn_score_topk_tc_position_space(m_items)) return -1;
    const int item = tc_item_of_pos(pos, tc_order(m_items));
    return item < m_items ? item : -1;
}

// ---- dense scores on tensor cores (getUsersRating) ----------------------------------------------------------------
struct DtLayout { int bt_pad, item_tiles, n_splits, tiles_per_split; size_t off_al, off_bh, off_bl, total; };
static DtLayout dt_layout(int Bt, int m_items) {
    DtLayout L;
    L.bt_pad = (Bt + TC_M - 1) / TC_M * TC_M;
    L.item_tiles = (m_items + TC_N - 1) / TC_N;
    const int row_blocks = L.bt_pad / TC_M, sms = sm_count() > 0 ? sm_count() : 148;
    int splits = (2 * sms + row_blocks - 1) / row_blocks;          // ~2 CTAs per SM worth of work, >= 4 tiles per split
    if (splits > (L.item_tiles + 3) / 4) splits = (L.item_tiles + 3) / 4;
    if (splits < 1) splits = 1;
    L.tiles_per_split = (L.item_tiles + splits - 1) / splits;
    L.n_splits = (L.item_tiles + L.tiles_per_split - 1) / L.tiles_per_split;
    size_t o = align_up((size_t)L.bt_pad * TC_D * 4, 1024);
    L.off_al = o; o += align_up((size_t)L.bt_pad * TC_D * 4, 1024);
    L.off_bh = o; o += align_up((size_t)L.item_tiles * TC_N * TC_D * 4, 1024);
    L.off_bl = o; o += align_up((size_t)L.item_tiles * TC_N * TC_D * 4, 1024);
    L.total = o;
    return L;
}

extern "C" int lgcn_score_dense_tc_supported(int32_t d) { return d == TC_D ? 1 : 0; }

extern "C" size_t lgcn_score_dense_tc_workspace_bytes(int32_t Bt, int32_t m_items) {
    if (Bt <= 0 || m_items <= 0) return 0;
    return dt_layout(Bt, m_items).total;
}

extern "C" int lgcn_score_dense_tc(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                                   int32_t m_items, int32_t d, float* scores, void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(users_emb && items_emb && scores && Bt > 0 && m_items > 0, "score_dense_tc: bad arguments");
    LGCN_CHECK_ARG(d == TC_D, "score_dense_tc: only d=%d takes the tensor-core path (use lgcn_score_dense)", TC_D);
    const DtLayout L = dt_layout(Bt, m_items);
    LGCN_CHECK_ARG(workspace && ((uintptr_t)workspace % 1024) == 0 && workspace_bytes >= L.total, "score_dense_tc: workspace too small or not 1024-byte aligned");
    LGCN_CHECK_ARG(((uintptr_t)items_emb % 16) == 0 && ((uintptr_t)users_emb % 16) == 0, "score_dense_tc: tables must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    char* w = static_cast<char*>(workspace);
    float* Ah = reinterpret_cast<float*>(w); float* Al = reinterpret_cast<float*>(w + L.off_al);
    float* Bh = reinterpret_cast<float*>(w + L.off_bh); float* Bl = reinterpret_cast<float*>(w + L.off_bl);
    gather_split_rows_kernel<<<(L.bt_pad * 16 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(users_emb), reinterpret_cast<const long long*>(users),
                                                                         Bt, L.bt_pad, reinterpret_cast<float4*>(Ah), reinterpret_cast<float4*>(Al));
    LGCN_CHECK_LAUNCH("gather_split_rows_kernel");
    const long long n4 = (long long)m_items * 16, pad4 = (long long)L.item_tiles * TC_N * 16;
    split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(items_emb), n4, reinterpret_cast<float4*>(Bh), reinterpret_cast<float4*>(Bl));
    LGCN_CHECK_LAUNCH("split_tf32_kernel");
    if (pad4 > n4) {                                                   // rows of the last tile beyond m_items: zero (never stored, but finite)
        cudaMemsetAsync(Bh + n4 * 4, 0, (size_t)(pad4 - n4) * 16, st);
        cudaMemsetAsync(Bl + n4 * 4, 0, (size_t)(pad4 - n4) * 16, st);
    }
    CUtensorMap mah, mal, mbh, mbl;
    if (int rc = make_map(&mah, Ah, (uint64_t)L.bt_pad, TC_M)) return rc;
    if (int rc = make_map(&mal, Al, (uint64_t)L.bt_pad, TC_M)) return rc;
    if (int rc = make_map(&mbh, Bh, (uint64_t)L.item_tiles * TC_N, TC_N)) return rc;
    if (int rc = make_map(&mbl, Bl, (uint64_t)L.item_tiles * TC_N, TC_N)) return rc;
    DtArgs a; a.Bt = Bt; a.m_items = m_items; a.tiles_per_split = L.tiles_per_split; a.out = scores;
    cudaError_t e = cudaFuncSetAttribute(score_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
    if (e != cudaSuccess) return fail("score_dense_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid(L.bt_pad / TC_M, L.n_splits);
    score_dense_tc_kernel<<<grid, DT_THREADS, DT_SMEM_BYTES, st>>>(mah, mal, mbh, mbl, a);
    LGCN_CHECK_LAUNCH("score_dense_tc_kernel");
    return 0;
}
