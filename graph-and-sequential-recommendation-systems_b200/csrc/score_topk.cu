// score_topk.cu — K3 (exact path): user x item scores fused with the train-item mask and a
// deterministic per-row top-k; plus on-device ranking metrics.
//
// Replaces getUsersRating's torch.matmul (reference code/model.py:114-123), the host-built
// rating[exclude_idx, exclude_items] = -(1<<10) mask (code/Procedure.py:177-181) and torch.topk
// (code/Procedure.py:183); the B x M score matrix is never written.
//
// Arithmetic contract (what makes the top-k bit-reproducible on a CPU, oracle/c/score_topk_ref.c):
//   score(b,i) = fma(u[d-1], v[d-1], ... fma(u[1], v[1], fma(u[0], v[0], 0)))   in fp32
//   masked cells score exactly -1024.0f; order = score descending, then item id ascending.
//
// Tiling: CTA = 128 users x 128 items, 256 threads, 8x8 register tile per thread, K chunks of <=64
// staged in shared memory (row stride K+4 floats -> conflict-free 128-bit reads).  After each item
// tile the scores go through shared memory to a thread-per-row scan that compares against the
// row's current k-th best and only then looks the item up in the user's CSR row (binary search).
// The item range can be split across CTAs (grid.y) to fill the GPU when there are few users; a
// warp-per-row merge combines the per-split lists with the same total order.
#include "common.cuh"
#include <float.h>

namespace lgcn {

constexpr int TU = 128, TI = 128, KC_MAX = 64;
constexpr int kScoreThreads = 256;
constexpr int S_STRIDE = TI + 1;

struct ScoreArgs {
    const float* U; const float* V; const long long* users; int Bt; int m_items; int d;
    const int* mask_indptr; const int* mask_indices; int mask_col_offset; int k;
    int tiles_per_split; int n_splits;
    float* part_val; int* part_idx;     // [Bt][n_splits][k]
    float* dense;                       // [Bt][m_items] (dense variant)
};

__device__ __forceinline__ bool row_has(const int* __restrict__ idx, int lo, int hi, int key) {
    int l = lo, h = hi;
    while (l < h) { const int mid = (l + h) >> 1; if (__ldg(idx + mid) < key) l = mid + 1; else h = mid; }
    return (l < hi) && (__ldg(idx + l) == key);
}

template <bool DENSE>
__global__ void __launch_bounds__(kScoreThreads, 1)
score_kernel(const __grid_constant__ ScoreArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int kc = a.d < KC_MAX ? a.d : KC_MAX;      // k-chunk
    const int rs = kc + 4;                           // row stride (floats), multiple of 4
    float* As = smem;                                // [TU][rs]
    float* Bs = As + TU * rs;                        // [TI][rs]
    float* S = Bs + TI * rs;                         // [TU][S_STRIDE]     (top-k variant)
    float* lv = S + TU * S_STRIDE;                   // [k][TU]
    int* li = reinterpret_cast<int*>(lv + a.k * TU); // [k][TU]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ub = blockIdx.x * TU;
    const int vec_per_row = kc >> 2;
    const int n_chunks = (a.d + kc - 1) / kc;

    // per-row selection state (thread r < TU owns row r)
    int cnt = 0; float tau = -FLT_MAX;
    int m_lo = 0, m_hi = 0; long long my_user = -1;
    if (!DENSE && tid < TU && ub + tid < a.Bt) {
        my_user = a.users ? a.users[ub + tid] : (long long)(ub + tid);
        if (a.mask_indptr) { m_lo = __ldg(a.mask_indptr + my_user); m_hi = __ldg(a.mask_indptr + my_user + 1); }
    }

    const int n_item_tiles = (a.m_items + TI - 1) / TI;
    const int t_begin = blockIdx.y * a.tiles_per_split;
    const int t_end = min(n_item_tiles, t_begin + a.tiles_per_split);

    auto load_A = [&](int chunk) {
        for (int idx = tid; idx < TU * vec_per_row; idx += kScoreThreads) {
            const int r = idx / vec_per_row, c4 = idx - r * vec_per_row;
            float4 v = f4_zero();
            if (ub + r < a.Bt) {
                const long long u = a.users ? a.users[ub + r] : (long long)(ub + r);
                v = __ldg(reinterpret_cast<const float4*>(a.U + (size_t)u * a.d + chunk * kc) + c4);
            }
            *reinterpret_cast<float4*>(As + r * rs + c4 * 4) = v;
        }
    };
    auto load_B = [&](int ib, int chunk) {
        for (int idx = tid; idx < TI * vec_per_row; idx += kScoreThreads) {
            const int r = idx / vec_per_row, c4 = idx - r * vec_per_row;
            float4 v = f4_zero();
            if (ib + r < a.m_items)
                v = __ldg(reinterpret_cast<const float4*>(a.V + (size_t)(ib + r) * a.d + chunk * kc) + c4);
            *reinterpret_cast<float4*>(Bs + r * rs + c4 * 4) = v;
        }
    };

    if (n_chunks == 1) load_A(0);

    for (int tile = t_begin; tile < t_end; ++tile) {
        const int ib = tile * TI;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            __syncthreads();                         // previous users of As/Bs/S are done
            if (n_chunks > 1) load_A(chunk);
            load_B(ib, chunk);
            __syncthreads();
            for (int k4 = 0; k4 < vec_per_row; ++k4) {
                float4 av[8], bv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * rs + k4 * 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * j) * rs + k4 * 4);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float c = acc[i][j];
                        c = fmaf(av[i].x, bv[j].x, c);
                        c = fmaf(av[i].y, bv[j].y, c);
                        c = fmaf(av[i].z, bv[j].z, c);
                        c = fmaf(av[i].w, bv[j].w, c);
                        acc[i][j] = c;
                    }
            }
        }

        if constexpr (DENSE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ub + ty + 16 * i;
                if (r >= a.Bt) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = ib + tx + 16 * j;
                    if (c < a.m_items) a.dense[(size_t)r * a.m_items + c] = acc[i][j];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) S[(ty + 16 * i) * S_STRIDE + tx + 16 * j] = acc[i][j];
            __syncthreads();
            if (tid < TU && my_user >= 0) {
                const int r = tid;
                const int n_valid = min(TI, a.m_items - ib);
                for (int c = 0; c < n_valid; ++c) {
                    float s = S[r * S_STRIDE + c];
                    if (cnt < a.k || s > tau) {
                        const int item = ib + c;
                        if (m_hi > m_lo && row_has(a.mask_indices, m_lo, m_hi, a.mask_col_offset + item)) s = -1024.f;
                        if (cnt < a.k || s > tau) {
                            int pos = (cnt < a.k) ? cnt : a.k - 1;
                            while (pos > 0 && lv[(pos - 1) * TU + r] < s) {
                                lv[pos * TU + r] = lv[(pos - 1) * TU + r];
                                li[pos * TU + r] = li[(pos - 1) * TU + r];
                                --pos;
                            }
                            lv[pos * TU + r] = s; li[pos * TU + r] = item;
                            if (cnt < a.k) ++cnt;
                            if (cnt == a.k) tau = lv[(a.k - 1) * TU + r];
                        }
                    }
                }
            }
        }
    }

    if constexpr (!DENSE) {
        if (tid < TU && my_user >= 0) {
            const size_t o = ((size_t)(ub + tid) * a.n_splits + blockIdx.y) * a.k;
            for (int q = 0; q < a.k; ++q) {
                a.part_val[o + q] = (q < cnt) ? lv[q * TU + tid] : -FLT_MAX;
                a.part_idx[o + q] = (q < cnt) ? li[q * TU + tid] : 0x7fffffff;
            }
        }
    }
}

// warp per row: lane s walks split s's sorted list; k rounds of a warp arg-best
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int Bt, int n_splits, int k,
                  long long* __restrict__ idx_out, float* __restrict__ val_out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= Bt) return;
    const size_t base = ((size_t)row * n_splits + lane) * k;
    int p = 0;
    float hv = -FLT_MAX; int hi = 0x7fffffff;
    if (lane < n_splits) { hv = part_val[base]; hi = part_idx[base]; }
    for (int q = 0; q < k; ++q) {
        float bv = hv; int bi = hi; int bl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
        }
        if (lane == 0) { idx_out[(size_t)row * k + q] = bi; val_out[(size_t)row * k + q] = bv; }
        if (lane == bl) {
            ++p;
            if (p < k) { hv = part_val[base + p]; hi = part_idx[base + p]; } else { hv = -FLT_MAX; hi = 0x7fffffff; }
        }
    }
}

// Deterministic metric sums (the same bits for any launch geometry, any number of calls and — because the multi-GPU Test
// gathers the ranked lists before calling this — any number of ranks): kernel 1 writes each row's 3*nk values to the
// workspace ([3*nk][Bt], coalesced), kernel 2 adds every column in a fixed order (one CTA per column: thread t adds
// rows t, t+T, ... then a fixed shared-memory tree).
__global__ void __launch_bounds__(256)
rank_metrics_rows_kernel(const long long* __restrict__ topk, int Bt, int k_max, const int* __restrict__ t_indptr,
                         const int* __restrict__ t_indices, const int* __restrict__ ks, int nk, double* __restrict__ rows) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bt) return;
    const int lo = t_indptr[b], hi = t_indptr[b + 1], ngt = hi - lo;
    for (int m = 0; m < nk; ++m) {
        double p = 0.0, r = 0.0, nd = 0.0;
        if (ngt > 0) {
            const int k = ks[m] < k_max ? ks[m] : k_max;
            int hits = 0; double dcg = 0.0, idcg = 0.0;
            for (int q = 0; q < k; ++q) {
                const double disc = 1.0 / log2((double)(q + 2));
                if (row_has(t_indices, lo, hi, (int)topk[(size_t)b * k_max + q])) { ++hits; dcg += disc; }
                if (q < ngt) idcg += disc;
            }
            if (idcg == 0.0) idcg = 1.0;
            p = (double)hits / (double)ks[m]; r = (double)hits / (double)ngt; nd = dcg / idcg;
        }
        rows[(size_t)(3 * m + 0) * Bt + b] = p;
        rows[(size_t)(3 * m + 1) * Bt + b] = r;
        rows[(size_t)(3 * m + 2) * Bt + b] = nd;
    }
}

__global__ void __launch_bounds__(1024)
rank_metrics_sum_kernel(const double* __restrict__ rows, int Bt, double* __restrict__ sums) {
    __shared__ double sh[1024];
    const double* col = rows + (size_t)blockIdx.x * Bt;
    double acc = 0.0;
    for (int b = threadIdx.x; b < Bt; b += 1024) acc += col[b];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] += sh[0];
}

static int pick_splits(int Bt, int m_items) {
    const int row_tiles = (Bt + TU - 1) / TU, item_tiles = (m_items + TI - 1) / TI;
    int want = (4 * sm_count() + row_tiles - 1) / row_tiles;
    if (want > 32) want = 32;
    if (want > item_tiles) want = item_tiles;
    if (want < 1) want = 1;
    return want;
}

static size_t score_smem_bytes(int d, int k, bool dense) {
    const int kc = d < KC_MAX ? d : KC_MAX, rs = kc + 4;
    size_t fl = (size_t)(TU + TI) * rs;
    if (!dense) fl += (size_t)TU * S_STRIDE + 2 * (size_t)k * TU;
    return fl * sizeof(float);
}

}  // namespace lgcn

using namespace lgcn;

extern "C" size_t lgcn_score_topk_workspace_bytes(int32_t Bt, int32_t m_items, int32_t k) {
    if (Bt <= 0 || m_items <= 0 || k <= 0) return 0;
    return align_up((size_t)Bt * 32 * k * sizeof(float), 256) + align_up((size_t)Bt * 32 * k * sizeof(int), 256);
}

static int check_score_args(const float* U, const float* V, int32_t Bt, int32_t m_items, int32_t d) {
    LGCN_CHECK_ARG(U && V, "score: null embedding table");
    LGCN_CHECK_ARG(Bt > 0 && m_items > 0, "score: Bt=%d m_items=%d", Bt, m_items);
    LGCN_CHECK_ARG(d >= 4 && d % 4 == 0 && (d <= KC_MAX || d % KC_MAX == 0), "score: d=%d must be a multiple of 4 and, above 64, of 64", d);
    LGCN_CHECK_ARG(((uintptr_t)U % 16) == 0 && ((uintptr_t)V % 16) == 0, "score: tables must be 16-byte aligned");
    return 0;
}

extern "C" int lgcn_score_topk(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                               int32_t m_items, int32_t d, const int32_t* mask_indptr, const int32_t* mask_indices,
                               int32_t mask_col_offset, int32_t k, int64_t* idx_out, float* val_out,
                               void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    if (int rc = check_score_args(users_emb, items_emb, Bt, m_items, d)) return rc;
    LGCN_CHECK_ARG(k > 0 && k <= LGCN_MAX_TOPK && k <= m_items, "score_topk: k=%d out of range (1..min(%d,m_items))", k, LGCN_MAX_TOPK);
    LGCN_CHECK_ARG(idx_out && val_out, "score_topk: null output");
    LGCN_CHECK_ARG((mask_indptr == nullptr) == (mask_indices == nullptr), "score_topk: mask_indptr/mask_indices must both be set or both null");
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_score_topk_workspace_bytes(Bt, m_items, k), "score_topk: workspace too small");
    cudaStream_t st = as_stream(stream);
    ScoreArgs a;
    a.U = users_emb; a.V = items_emb; a.users = reinterpret_cast<const long long*>(users); a.Bt = Bt; a.m_items = m_items; a.d = d;
    a.mask_indptr = mask_indptr; a.mask_indices = mask_indices; a.mask_col_offset = mask_col_offset; a.k = k;
    a.n_splits = pick_splits(Bt, m_items);
    const int item_tiles = (m_items + TI - 1) / TI;
    a.tiles_per_split = (item_tiles + a.n_splits - 1) / a.n_splits;
    a.n_splits = (item_tiles + a.tiles_per_split - 1) / a.tiles_per_split;       // no empty splits
    char* w = static_cast<char*>(workspace);
    a.part_val = reinterpret_cast<float*>(w);
    a.part_idx = reinterpret_cast<int*>(w + align_up((size_t)Bt * 32 * k * sizeof(float), 256));
    a.dense = nullptr;
    const size_t smem = score_smem_bytes(d, k, false);
    LGCN_CHECK_ARG((int)smem <= max_smem_optin(), "score_topk: needs %zu B shared memory (k too large)", smem);
    cudaError_t e = cudaFuncSetAttribute(score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("score_topk: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid((Bt + TU - 1) / TU, a.n_splits);
    score_kernel<false><<<grid, kScoreThreads, smem, st>>>(a);
    LGCN_CHECK_LAUNCH("score_kernel<topk>");
    topk_merge_kernel<<<(unsigned)(((size_t)Bt * 32 + 255) / 256), 256, 0, st>>>(a.part_val, a.part_idx, Bt, a.n_splits, k,
                                                                                 reinterpret_cast<long long*>(idx_out), val_out);
    LGCN_CHECK_LAUNCH("topk_merge_kernel");
    return 0;
}

extern "C" int lgcn_score_dense(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                                int32_t m_items, int32_t d, float* scores, lgcn_stream_t stream) {
    if (int rc = check_score_args(users_emb, items_emb, Bt, m_items, d)) return rc;
    LGCN_CHECK_ARG(scores, "score_dense: null output");
    cudaStream_t st = as_stream(stream);
    ScoreArgs a;
    a.U = users_emb; a.V = items_emb; a.users = reinterpret_cast<const long long*>(users); a.Bt = Bt; a.m_items = m_items; a.d = d;
    a.mask_indptr = nullptr; a.mask_indices = nullptr; a.mask_col_offset = 0; a.k = 0;
    const int item_tiles = (m_items + TI - 1) / TI;
    a.n_splits = pick_splits(Bt, m_items);
    a.tiles_per_split = (item_tiles + a.n_splits - 1) / a.n_splits;
    a.n_splits = (item_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.part_val = nullptr; a.part_idx = nullptr; a.dense = scores;
    const size_t smem = score_smem_bytes(d, 0, true);
    cudaError_t e = cudaFuncSetAttribute(score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("score_dense: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid((Bt + TU - 1) / TU, a.n_splits);
    score_kernel<true><<<grid, kScoreThreads, smem, st>>>(a);
    LGCN_CHECK_LAUNCH("score_kernel<dense>");
    return 0;
}

extern "C" size_t lgcn_rank_metrics_workspace_bytes(int32_t Bt, int32_t nk) {
    if (Bt < 0 || nk < 0) return 0;
    return sizeof(double) * 3 * (size_t)nk * (size_t)(Bt > 0 ? Bt : 1);
}

extern "C" int lgcn_rank_metrics(const int64_t* topk_idx, int32_t Bt, int32_t k_max,
                                 const int32_t* test_indptr, const int32_t* test_indices,
                                 const int32_t* ks, int32_t nk, double* sums_out,
                                 void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(topk_idx && test_indptr && test_indices && ks && sums_out, "rank_metrics: null argument");
    LGCN_CHECK_ARG(Bt > 0 && k_max > 0 && nk > 0 && nk <= 16, "rank_metrics: Bt=%d k_max=%d nk=%d", Bt, k_max, nk);
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_rank_metrics_workspace_bytes(Bt, nk) && ((uintptr_t)workspace % 8) == 0,
                   "rank_metrics: workspace too small or misaligned");
    double* rows = static_cast<double*>(workspace);
    rank_metrics_rows_kernel<<<(Bt + 255) / 256, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const long long*>(topk_idx), Bt, k_max, test_indptr, test_indices, ks, nk, rows);
    LGCN_CHECK_LAUNCH("rank_metrics_rows_kernel");
    rank_metrics_sum_kernel<<<3 * nk, 1024, 0, as_stream(stream)>>>(rows, Bt, sums_out);
    LGCN_CHECK_LAUNCH("rank_metrics_sum_kernel");
    return 0;
}
