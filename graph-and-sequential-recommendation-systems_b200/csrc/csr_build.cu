// csr_build.cu — K4: device-side builder of the symmetric-normalised bipartite adjacency.
//
// Replaces the reference's SciPy path: UserItemNet = csr_matrix((ones,(u,i))) (duplicates summed,
// code/dataloader.py:133-136), users_D/items_D (:139-142), the dok/lil block assignment
// adj[:nu,nu:]=R, adj[nu:,:nu]=R.T -> CSR (:223-227, 83.8 s on gowalla) and the normaliser
// rowsum -> power(-0.5) -> D.A.D (:230-234).
//
// Pipeline (all on the caller's stream, no host round trip):
//   1. every edge (u,i) emits two packed keys  (u<<cb | nu+i)  and  (nu+i<<cb | u),  cb = bits(N-1)
//   2. LSD radix sort of the 2E keys, 8 bits per pass, ceil(2cb/8) passes: per-tile histogram ->
//      exclusive scan of the digit-major histogram -> stable scatter (warp match_any ranking)
//   3. head flags (key != previous key) -> exclusive scan = position of each key among the UNIQUE
//      entries; run length of equal keys = duplicate multiplicity (the summed weight)
//   4. row boundaries of the sorted keys give indptr (unique positions) and the weighted degree
//      (raw positions) without atomics; dinv = deg^-1/2 in double, rounded once
//   5. vals[e] = fl32(fl32(dinv[row]*w) * dinv[col])   — the rounding order of D.dot(A).dot(D)
// The output structure is bit-identical to SciPy's (sorted, duplicate-free rows).
#include "common.cuh"

namespace lgcn {

typedef unsigned long long u64;

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int kScanThreads = 256;
constexpr int kScanIPT = 4;
constexpr int kScanTile = kScanThreads * kScanIPT;

__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const int* in, int* out, long long n, int* block_sums) {   // in == out allowed
    __shared__ int warp_sums[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanIPT;
    int v[kScanIPT]; int tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanIPT; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; tsum += v[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;      // inclusive over warps
    }
    __syncthreads();
    int run = incl - tsum + (warp > 0 ? warp_sums[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < kScanIPT; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
    if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = warp_sums[kScanThreads / 32 - 1];
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int* __restrict__ out, long long n, const int* __restrict__ block_offsets) {
    const int off = block_offsets[blockIdx.x];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanIPT;
#pragma unroll
    for (int i = 0; i < kScanIPT; ++i) if (base + i < n) out[base + i] += off;
}

static size_t scan_tmp_ints(long long n) {
    size_t total = 0;
    long long nb = (n + kScanTile - 1) / kScanTile;
    while (nb > 1) { total += align_up((size_t)nb, 4); nb = (nb + kScanTile - 1) / kScanTile; }
    return total + 4;
}

// exclusive scan of in[0..n) into out (in == out allowed); *total = sum of all elements
static int scan_exclusive(const int* in, int* out, long long n, int* tmp, int* total, cudaStream_t st) {
    if (n <= 0) { cudaMemsetAsync(total, 0, sizeof(int), st); return 0; }
    const long long nb = (n + kScanTile - 1) / kScanTile;
    if (nb > 0x7fffffffLL) return fail("scan: too many tiles");
    if (nb == 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, total);
        LGCN_CHECK_LAUNCH("scan_tile_kernel");
        return 0;
    }
    int* sums = tmp;
    scan_tile_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, sums);
    LGCN_CHECK_LAUNCH("scan_tile_kernel");
    if (int rc = scan_exclusive(sums, sums, nb, tmp + align_up((size_t)nb, 4), total, st)) return rc;
    scan_add_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(out, n, sums);
    LGCN_CHECK_LAUNCH("scan_add_kernel");
    return 0;
}

// ------------------------------------------------------------------ LSD radix sort (u64 keys)
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const u64* __restrict__ keys, long long n, int shift, int* __restrict__ hist, int nblk) {
    __shared__ int h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortRounds; ++r) {
        const long long j = base + r * kSortThreads + threadIdx.x;
        if (j < n) atomicAdd(&h[(int)((keys[j] >> shift) & (kRadix - 1))], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];     // digit-major
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const u64* __restrict__ keys, u64* __restrict__ out, long long n, int shift,
                     const int* __restrict__ offsets, int nblk) {
    constexpr int WARPS = kSortThreads / 32;
    __shared__ int base_off[kRadix];      // global offset of this tile's first key of each digit
    __shared__ int seen[kRadix];          // keys of each digit placed in earlier rounds
    __shared__ int wcnt[WARPS][kRadix];   // per-warp digit counts of the current round
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    base_off[tid] = offsets[(size_t)tid * nblk + blockIdx.x];
    seen[tid] = 0;
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int r = 0; r < kSortRounds; ++r) {
        const long long j = base + r * kSortThreads + tid;
        if (base + (long long)r * kSortThreads >= n) break;      // uniform over the CTA
#pragma unroll
        for (int w = 0; w < WARPS; ++w) wcnt[w][tid] = 0;
        __syncthreads();
        const bool valid = j < n;
        u64 key = 0; int digit = kRadix;                         // kRadix = "no key" class
        if (valid) { key = keys[j]; digit = (int)((key >> shift) & (kRadix - 1)); }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) wcnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            int before = seen[digit];
            for (int w = 0; w < warp; ++w) before += wcnt[w][digit];
            out[(long long)base_off[digit] + before + rank_in_warp] = key;
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) add += wcnt[w][tid];
        seen[tid] += add;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ CSR assembly
struct BuildArgs {
    const long long* tu; const long long* ti; long long E; int nu; int ni; int N; int cb;
    int row_base; int n_rows;      // rows [row_base, row_base + n_rows) are assembled (a row block); keys hold LOCAL rows
    int* indptr; int* indices; float* vals; float* deg; float* dinv; long long* nnz_out; int* status;
    int* excl; int* headpos; int* erow; int* rawptr; int* total;
};

__global__ void emit_keys_kernel(BuildArgs a, u64* keys) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    long long u = a.tu[e], i = a.ti[e];
    if (u < 0 || u >= a.nu || i < 0 || i >= a.ni) {
        atomicExch(a.status, 1);
        u = u < 0 ? 0 : (u >= a.nu ? a.nu - 1 : u);
        i = i < 0 ? 0 : (i >= a.ni ? a.ni - 1 : i);
    }
    const u64 r = (u64)u, c = (u64)(a.nu + i);
    keys[2 * e] = (r << a.cb) | c;
    keys[2 * e + 1] = (c << a.cb) | r;
}

__global__ void head_flags_kernel(const u64* __restrict__ keys, long long n, int* __restrict__ flags) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

__global__ void assemble_kernel(BuildArgs a, const u64* __restrict__ keys, long long n) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u64 key = keys[j];
    const int row = (int)(key >> a.cb), col = (int)(key & ((1ull << a.cb) - 1ull));
    const bool head = (j == 0) || (keys[j - 1] != key);
    const int ui = a.excl[j] + (head ? 1 : 0) - 1;       // index among unique entries
    if (head) { a.indices[ui] = col; a.erow[ui] = row; a.headpos[ui] = (int)j; }
    const int prev_row = (j == 0) ? -1 : (int)(keys[j - 1] >> a.cb);
    for (int r = prev_row + 1; r <= row; ++r) { a.indptr[r] = ui; a.rawptr[r] = (int)j; }   // ui == excl[j] here
    if (j == n - 1) {
        const int U = *a.total;
        for (int r = row + 1; r <= a.n_rows; ++r) { a.indptr[r] = U; a.rawptr[r] = (int)n; }
        a.headpos[U] = (int)n;
        *a.nnz_out = (long long)U;
    }
}

__global__ void empty_graph_kernel(BuildArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= a.N) a.indptr[r] = 0;
    if (r < a.N) { a.deg[r] = 0.f; a.dinv[r] = 0.f; }
    if (r == 0) *a.nnz_out = 0;
}

__global__ void degree_kernel(BuildArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.N) return;
    const int dw = a.rawptr[r + 1] - a.rawptr[r];
    a.deg[r] = (float)dw;
    a.dinv[r] = dw > 0 ? (float)(1.0 / sqrt((double)dw)) : 0.f;
}

__global__ void values_kernel(BuildArgs a, long long n) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || e >= *a.total) return;
    const float w = (float)(a.headpos[e + 1] - a.headpos[e]);
    a.vals[e] = __fmul_rn(__fmul_rn(a.dinv[a.row_base + a.erow[e]], w), a.dinv[a.indices[e]]);
}

// ---- row-block build (multi-GPU row partition: every rank assembles only the rows it owns) ----------------------
// weighted degrees of ALL nodes from an edge chunk (duplicates counted, like the row sums of code/dataloader.py:230)
__global__ void degree_count_kernel(const long long* __restrict__ tu, const long long* __restrict__ ti, long long E,
                                    int nu, int ni, int* __restrict__ counts, int* status) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long u = tu[e], i = ti[e];
    if (u < 0 || u >= nu || i < 0 || i >= ni) { atomicExch(status, 1); return; }
    atomicAdd(counts + u, 1);
    atomicAdd(counts + nu + i, 1);
}

__global__ void degree_finalize_kernel(const int* __restrict__ counts, int N, float* __restrict__ deg, float* __restrict__ dinv) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int dw = counts[r];
    deg[r] = (float)dw;
    dinv[r] = dw > 0 ? (float)(1.0 / sqrt((double)dw)) : 0.f;
}

// both directions of every edge whose ROW lies in [row_begin, row_end): key = (row - row_begin) << cb | col, appended at
// a device cursor (the order is irrelevant: the keys are sorted next)
__global__ void emit_block_keys_kernel(const long long* __restrict__ tu, const long long* __restrict__ ti, long long E,
                                       int nu, int ni, int cb, int row_begin, int row_end, u64* __restrict__ keys,
                                       long long cap, unsigned long long* cursor, int* status) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    u64 k0 = 0, k1 = 0; int n = 0;
    if (e < E) {
        const long long u = tu[e], i = ti[e];
        if (u < 0 || u >= nu || i < 0 || i >= ni) atomicExch(status, 1);
        else {
            const int r = (int)u, c = nu + (int)i;
            if (r >= row_begin && r < row_end) { k0 = ((u64)(r - row_begin) << cb) | (u64)c; n = 1; }
            if (c >= row_begin && c < row_end) { const u64 k = ((u64)(c - row_begin) << cb) | (u64)r; if (n) k1 = k; else k0 = k; ++n; }
        }
    }
    // warp-aggregated append
    const unsigned lane = threadIdx.x & 31;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    const long long at = (long long)base + incl - n;
    if (n >= 1) { if (at < cap) keys[at] = k0; else atomicExch(status, 2); }
    if (n == 2) { if (at + 1 < cap) keys[at + 1] = k1; else atomicExch(status, 2); }
}

__global__ void coo_to_csr_kernel(const long long* __restrict__ rows, const long long* __restrict__ cols,
                                  long long nnz, int n_rows, int* __restrict__ indptr, int* __restrict__ indices) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    indices[j] = (int)cols[j];
    const int row = (int)rows[j];
    const int prev_row = (j == 0) ? -1 : (int)rows[j - 1];
    for (int r = prev_row + 1; r <= row; ++r) indptr[r] = (int)j;
    if (j == nnz - 1) for (int r = row + 1; r <= n_rows; ++r) indptr[r] = (int)nnz;
}

__global__ void fill_i32_kernel(int* p, long long n, int v) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) p[j] = v;
}

static int bits_for(long long max_value) { int b = 1; while ((1LL << b) <= max_value) ++b; return b; }

struct Layout { size_t keysA, keysB, hist, scan_tmp, excl, headpos, erow, rawptr, total, end; long long nblk; };

static Layout make_layout_keys(long long n, int N) {
    Layout L;
    L.nblk = (n + kSortTile - 1) / kSortTile; if (L.nblk < 1) L.nblk = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    L.keysA = take(sizeof(u64) * (size_t)n);
    L.keysB = take(sizeof(u64) * (size_t)n);
    L.hist = take(sizeof(int) * (size_t)kRadix * L.nblk);
    const long long scan_len = (n > (long long)kRadix * L.nblk) ? n : (long long)kRadix * L.nblk;
    L.scan_tmp = take(sizeof(int) * scan_tmp_ints(scan_len));
    L.excl = take(sizeof(int) * (size_t)n);
    L.headpos = take(sizeof(int) * (size_t)(n + 1));
    L.erow = take(sizeof(int) * (size_t)n);
    L.rawptr = take(sizeof(int) * (size_t)(N + 1));
    L.total = take(16);
    L.end = o;
    return L;
}

static Layout make_layout(long long E, int N) { return make_layout_keys(2 * E, N); }

// sorted keys -> indptr/indices (+ erow/headpos/rawptr for the value pass); `bits` = significant key bits
static int sort_and_assemble(BuildArgs& a, const Layout& L, char* w, long long n, int bits, u64** sorted_out, cudaStream_t st) {
    u64* kA = reinterpret_cast<u64*>(w + L.keysA);
    u64* kB = reinterpret_cast<u64*>(w + L.keysB);
    int* hist = reinterpret_cast<int*>(w + L.hist);
    int* scan_tmp = reinterpret_cast<int*>(w + L.scan_tmp);
    const int passes = (bits + 7) / 8;
    const int nblk = (int)((n + kSortTile - 1) / kSortTile);
    for (int p = 0; p < passes; ++p) {
        radix_hist_kernel<<<nblk, kSortThreads, 0, st>>>(kA, n, 8 * p, hist, nblk);
        LGCN_CHECK_LAUNCH("radix_hist_kernel");
        if (int rc = scan_exclusive(hist, hist, (long long)kRadix * nblk, scan_tmp, a.total, st)) return rc;
        radix_scatter_kernel<<<nblk, kSortThreads, 0, st>>>(kA, kB, n, 8 * p, hist, nblk);
        LGCN_CHECK_LAUNCH("radix_scatter_kernel");
        u64* t = kA; kA = kB; kB = t;
    }
    const unsigned nb = (unsigned)((n + 255) / 256);
    head_flags_kernel<<<nb, 256, 0, st>>>(kA, n, a.excl);
    LGCN_CHECK_LAUNCH("head_flags_kernel");
    if (int rc = scan_exclusive(a.excl, a.excl, n, scan_tmp, a.total, st)) return rc;
    assemble_kernel<<<nb, 256, 0, st>>>(a, kA, n);
    LGCN_CHECK_LAUNCH("assemble_kernel");
    *sorted_out = kA;
    return 0;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" size_t lgcn_csr_build_workspace_bytes(int64_t E, int32_t n_users, int32_t m_items) {
    if (E < 0 || n_users < 0 || m_items < 0) return 0;
    return make_layout(E, n_users + m_items).end;
}

extern "C" int lgcn_csr_build(const int64_t* train_user, const int64_t* train_item, int64_t E,
                              int32_t n_users, int32_t m_items,
                              int32_t* indptr, int32_t* indices, float* vals, float* deg, float* dinv,
                              int64_t* nnz_out, int32_t* status_out,
                              void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(E >= 0 && n_users > 0 && m_items > 0, "csr_build: bad sizes E=%lld nu=%d ni=%d", (long long)E, n_users, m_items);
    LGCN_CHECK_ARG(2 * E < 0x7fffffffLL, "csr_build: 2E=%lld exceeds int32 CSR offsets", (long long)(2 * E));
    LGCN_CHECK_ARG((long long)n_users + m_items < 0x7fffffffLL, "csr_build: N exceeds int32");
    LGCN_CHECK_ARG(indptr && deg && dinv && nnz_out && status_out, "csr_build: null output");
    LGCN_CHECK_ARG(E == 0 || (train_user && train_item && indices && vals), "csr_build: null edge/entry arrays");
    const int N = n_users + m_items;
    const Layout L = make_layout(E, N);
    LGCN_CHECK_ARG(workspace && workspace_bytes >= L.end, "csr_build: workspace %zu < %zu bytes", workspace_bytes, L.end);
    LGCN_CHECK_ARG(((uintptr_t)workspace % 256) == 0, "csr_build: workspace must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char* w = static_cast<char*>(workspace);
    BuildArgs a;
    a.tu = reinterpret_cast<const long long*>(train_user); a.ti = reinterpret_cast<const long long*>(train_item);
    a.E = E; a.nu = n_users; a.ni = m_items; a.N = N; a.cb = bits_for(N - 1);
    a.indptr = indptr; a.indices = indices; a.vals = vals; a.deg = deg; a.dinv = dinv;
    a.nnz_out = reinterpret_cast<long long*>(nnz_out); a.status = status_out;
    a.excl = reinterpret_cast<int*>(w + L.excl); a.headpos = reinterpret_cast<int*>(w + L.headpos);
    a.erow = reinterpret_cast<int*>(w + L.erow); a.rawptr = reinterpret_cast<int*>(w + L.rawptr);
    a.total = reinterpret_cast<int*>(w + L.total);
    cudaMemsetAsync(status_out, 0, sizeof(int32_t), st);
    if (E == 0) {
        a.row_base = 0; a.n_rows = N;
        empty_graph_kernel<<<(N + 1 + 255) / 256, 256, 0, st>>>(a);
        LGCN_CHECK_LAUNCH("empty_graph_kernel");
        return 0;
    }
    const long long n = 2 * E;
    a.row_base = 0; a.n_rows = N;
    emit_keys_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(a, reinterpret_cast<u64*>(w + L.keysA));
    LGCN_CHECK_LAUNCH("emit_keys_kernel");
    u64* kA = nullptr;
    if (int rc = sort_and_assemble(a, L, w, n, 2 * a.cb, &kA, st)) return rc;
    const unsigned nb = (unsigned)((n + 255) / 256);
    degree_kernel<<<(N + 255) / 256, 256, 0, st>>>(a);
    LGCN_CHECK_LAUNCH("degree_kernel");
    values_kernel<<<nb, 256, 0, st>>>(a, n);
    LGCN_CHECK_LAUNCH("values_kernel");
    return 0;
}

extern "C" int lgcn_degree_accumulate(const int64_t* train_user, const int64_t* train_item, int64_t E, int32_t n_users, int32_t m_items,
                                      int32_t* deg_counts, int32_t* status_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(E >= 0 && n_users > 0 && m_items > 0 && deg_counts && status_out, "degree_accumulate: bad arguments");
    LGCN_CHECK_ARG(E == 0 || (train_user && train_item), "degree_accumulate: null edge arrays");
    if (E == 0) return 0;
    degree_count_kernel<<<(unsigned)((E + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const long long*>(train_user), reinterpret_cast<const long long*>(train_item), E, n_users, m_items, deg_counts, status_out);
    LGCN_CHECK_LAUNCH("degree_count_kernel");
    return 0;
}

extern "C" int lgcn_degree_finalize(const int32_t* deg_counts, int32_t n_nodes, float* deg, float* dinv, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(deg_counts && deg && dinv && n_nodes > 0, "degree_finalize: bad arguments");
    degree_finalize_kernel<<<(n_nodes + 255) / 256, 256, 0, as_stream(stream)>>>(deg_counts, n_nodes, deg, dinv);
    LGCN_CHECK_LAUNCH("degree_finalize_kernel");
    return 0;
}

extern "C" size_t lgcn_csr_rows_workspace_bytes(int64_t n_keys, int32_t n_rows_local) {
    if (n_keys < 0 || n_rows_local < 0) return 0;
    return make_layout_keys(n_keys > 0 ? n_keys : 1, n_rows_local).end;
}

extern "C" int lgcn_csr_rows_emit(const int64_t* train_user, const int64_t* train_item, int64_t E, int32_t n_users, int32_t m_items,
                                  int32_t row_begin, int32_t row_end, int64_t n_keys_cap, uint64_t* cursor_dev, int32_t* status_out,
                                  void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(E >= 0 && n_users > 0 && m_items > 0 && cursor_dev && status_out, "csr_rows_emit: bad arguments");
    const int N = n_users + m_items;
    LGCN_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "csr_rows_emit: bad row block [%d,%d)", row_begin, row_end);
    LGCN_CHECK_ARG(n_keys_cap >= 0 && n_keys_cap < 0x7fffffffLL, "csr_rows_emit: n_keys_cap=%lld exceeds int32 CSR offsets", (long long)n_keys_cap);
    const Layout L = make_layout_keys(n_keys_cap > 0 ? n_keys_cap : 1, row_end - row_begin);
    LGCN_CHECK_ARG(workspace && workspace_bytes >= L.end && ((uintptr_t)workspace % 256) == 0, "csr_rows_emit: workspace too small or misaligned");
    if (E == 0) return 0;
    LGCN_CHECK_ARG(train_user && train_item, "csr_rows_emit: null edge arrays");
    emit_block_keys_kernel<<<(unsigned)((E + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const long long*>(train_user), reinterpret_cast<const long long*>(train_item), E, n_users, m_items,
        bits_for(N - 1), row_begin, row_end, reinterpret_cast<u64*>(static_cast<char*>(workspace) + L.keysA), n_keys_cap,
        reinterpret_cast<unsigned long long*>(cursor_dev), status_out);
    LGCN_CHECK_LAUNCH("emit_block_keys_kernel");
    return 0;
}

extern "C" int lgcn_csr_rows_finish(int64_t n_keys, int64_t n_keys_cap, int32_t n_users, int32_t m_items, int32_t row_begin, int32_t row_end,
                                    const float* dinv, int32_t* indptr, int32_t* indices, float* vals, int64_t* nnz_out,
                                    void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    const int N = n_users + m_items;
    LGCN_CHECK_ARG(n_users > 0 && m_items > 0 && 0 <= row_begin && row_begin <= row_end && row_end <= N, "csr_rows_finish: bad row block");
    LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= n_keys_cap && n_keys_cap < 0x7fffffffLL, "csr_rows_finish: bad key count %lld (cap %lld)", (long long)n_keys, (long long)n_keys_cap);
    LGCN_CHECK_ARG(dinv && indptr && nnz_out && (n_keys == 0 || (indices && vals)), "csr_rows_finish: null output");
    const int n_rows = row_end - row_begin;
    const Layout L = make_layout_keys(n_keys_cap > 0 ? n_keys_cap : 1, n_rows);
    LGCN_CHECK_ARG(workspace && workspace_bytes >= L.end && ((uintptr_t)workspace % 256) == 0, "csr_rows_finish: workspace too small or misaligned");
    cudaStream_t st = as_stream(stream);
    char* w = static_cast<char*>(workspace);
    BuildArgs a;
    a.tu = nullptr; a.ti = nullptr; a.E = 0; a.nu = n_users; a.ni = m_items; a.N = N; a.cb = bits_for(N - 1);
    a.row_base = row_begin; a.n_rows = n_rows;
    a.indptr = indptr; a.indices = indices; a.vals = vals; a.deg = nullptr; a.dinv = const_cast<float*>(dinv);
    a.nnz_out = reinterpret_cast<long long*>(nnz_out); a.status = nullptr;
    a.excl = reinterpret_cast<int*>(w + L.excl); a.headpos = reinterpret_cast<int*>(w + L.headpos);
    a.erow = reinterpret_cast<int*>(w + L.erow); a.rawptr = reinterpret_cast<int*>(w + L.rawptr);
    a.total = reinterpret_cast<int*>(w + L.total);
    if (n_keys == 0) {
        fill_i32_kernel<<<(n_rows + 1 + 255) / 256, 256, 0, st>>>(indptr, n_rows + 1, 0);
        LGCN_CHECK_LAUNCH("fill_i32_kernel");
        cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), st);
        return 0;
    }
    u64* kA = nullptr;
    const int bits = a.cb + bits_for(n_rows > 1 ? n_rows - 1 : 1);
    if (int rc = sort_and_assemble(a, L, w, n_keys, bits, &kA, st)) return rc;
    values_kernel<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(a, n_keys);
    LGCN_CHECK_LAUNCH("values_kernel");
    return 0;
}

extern "C" int lgcn_coo_to_csr(const int64_t* rows, const int64_t* cols, int64_t nnz, int32_t n_rows,
                               int32_t* indptr, int32_t* indices, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && n_rows >= 0 && nnz >= 0 && nnz < 0x7fffffffLL, "coo_to_csr: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (nnz == 0) {
        fill_i32_kernel<<<(n_rows + 1 + 255) / 256, 256, 0, st>>>(indptr, n_rows + 1, 0);
        LGCN_CHECK_LAUNCH("fill_i32_kernel");
        return 0;
    }
    LGCN_CHECK_ARG(rows && cols && indices, "coo_to_csr: null arrays");
    coo_to_csr_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const long long*>(rows), reinterpret_cast<const long long*>(cols), nnz, n_rows, indptr, indices);
    LGCN_CHECK_LAUNCH("coo_to_csr_kernel");
    return 0;
}
