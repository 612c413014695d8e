// sampler.cu — device-side BPR negative sampler.
//
// Replaces utils.UniformSample_original / sources/sampling.cpp (reference code/utils.py:68-81,
// code/sources/sampling.cpp:27-56) AND the shuffle that follows it (code/Procedure.py:55, code/utils.py:142-151):
// every user gets exactly train_num/user_num triples (user, uniform positive from its CSR row, rejection-sampled
// negative not in the row), and the triples come out already permuted, directly in the (3, n) int64 layout the
// training step reads — no host sampler phase (0.2 s C++ / 16 s Python per gowalla epoch) and no per-epoch H2D.
//
// Randomness is counter-based (splitmix64 of seed, epoch, sample id, attempt), so the stream is reproducible and
// independent of the launch geometry.  It is NOT the glibc rand() stream of the reference: parity for this
// component is distributional (same per-user counts, same support), see tests/test_gpu_kernels.py.
// The permutation is a 4-round Feistel network on the next power of four with cycle walking: a bijection of
// [0, n) computed independently by every thread.
#include "common.cuh"

namespace lgcn {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ unsigned long long feistel_perm(unsigned long long p, unsigned long long n, int half_bits, unsigned long long key) {
    const unsigned long long mask = (1ull << half_bits) - 1ull;
    do {
        unsigned long long l = p >> half_bits, r = p & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const unsigned long long f = splitmix64(r ^ (key + 0x1234567ull * (round + 1))) & mask;
            const unsigned long long nl = r; r = l ^ f; l = nl;
        }
        p = (l << half_bits) | r;
    } while (p >= n);                                   // cycle walking keeps it a bijection on [0, n)
    return p;
}

__global__ void __launch_bounds__(256)
sample_bpr_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int n_users, int m_items,
                  long long per_user, long long n, int half_bits, unsigned long long key,
                  long long* __restrict__ users, long long* __restrict__ pos, long long* __restrict__ neg, int* __restrict__ status) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long s = (long long)feistel_perm((unsigned long long)p, (unsigned long long)n, half_bits, key);
    const int u = (int)(s / per_user);
    const int lo = __ldg(indptr + u), hi = __ldg(indptr + u + 1), deg = hi - lo;
    unsigned long long st = splitmix64(key ^ ((unsigned long long)s * 0xD1342543DE82EF95ull));
    long long pi = 0, ni = 0;
    bool found = false;
    if (deg > 0) pi = (long long)__ldg(indices + lo + (int)(st % (unsigned long long)deg)) - n_users;
    else if (status) atomicOr(status, 1);               // a user without train items: the triple is fabricated — the caller must fail
    if (deg < m_items) {
        for (int attempt = 0; attempt < 1 << 20; ++attempt) {
            st = splitmix64(st);
            const int cand = (int)(st % (unsigned long long)m_items);
            const int keyc = cand + n_users;
            int l = lo, h = hi;
            while (l < h) { const int mid = (l + h) >> 1; if (__ldg(indices + mid) < keyc) l = mid + 1; else h = mid; }
            if (!(l < hi && __ldg(indices + l) == keyc)) { ni = cand; found = true; break; }
        }
    }
    if (!found && status) atomicOr(status, 2);           // the user interacted with every item (or 2^20 rejections in a row)
    users[p] = u; pos[p] = pi; neg[p] = ni;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_sample_bpr(const int32_t* indptr, const int32_t* indices, int32_t n_users, int32_t m_items,
                               int64_t train_num, uint64_t seed, uint64_t epoch,
                               int64_t* users_out, int64_t* pos_out, int64_t* neg_out, int32_t* status_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(indptr && indices && users_out && pos_out && neg_out, "sample_bpr: null argument");
    LGCN_CHECK_ARG(n_users > 0 && m_items > 0 && train_num >= 0, "sample_bpr: bad sizes");
    const long long per_user = train_num / n_users;
    const long long n = per_user * n_users;
    if (n == 0) return 0;
    int bits = 2; while ((1ull << bits) < (unsigned long long)n) bits += 2;       // even number of bits
    const unsigned long long key = seed * 0x9E3779B97F4A7C15ull + epoch * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull;
    sample_bpr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        indptr, indices, n_users, m_items, per_user, n, bits / 2, key,
        reinterpret_cast<long long*>(users_out), reinterpret_cast<long long*>(pos_out), reinterpret_cast<long long*>(neg_out), status_out);
    LGCN_CHECK_LAUNCH("sample_bpr_kernel");
    return 0;
}
