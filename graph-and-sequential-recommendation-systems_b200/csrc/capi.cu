// capi.cu — runtime glue behind the C ABI: error channel, device info, and the host-side negative
// sampler that keeps the semantics of the reference's only native component
// (code/sources/sampling.cpp:27-56, 88-91) so utils.UniformSample_original works unchanged.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <cstdio>

namespace lgcn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return 1;
}

static int g_sm_count = 0, g_smem_optin = 0;
static void query_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_sm_count = 148; g_smem_optin = 227 * 1024; return; }
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
    if (g_smem_optin <= 0) g_smem_optin = 227 * 1024;
}
int sm_count() { if (!g_sm_count) query_device(); return g_sm_count; }
int max_smem_optin() { if (!g_smem_optin) query_device(); return g_smem_optin; }

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_abi_version(void) { return LGCN_ABI_VERSION; }
extern "C" const char* lgcn_last_error(void) { return g_err; }

extern "C" int lgcn_device_info(int32_t* out_host) {
    LGCN_CHECK_ARG(out_host, "device_info: null output");
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail("device_info: %s", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    out_host[0] = sm_count(); out_host[1] = max_smem_optin(); out_host[2] = major * 10 + minor;
    return 0;
}

// Peer access from the current device to `peer_device` (needed before a kernel may dereference memory of another GPU
// that was mapped into this process, e.g. through CUDA IPC).  Idempotent.
extern "C" int lgcn_enable_peer_access(int32_t peer_device) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail("enable_peer_access: %s", cudaGetErrorString(e));
    if (dev == peer_device) return 0;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dev, peer_device);
    if (!can) return fail("enable_peer_access: device %d cannot access device %d", dev, peer_device);
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return fail("enable_peer_access: %s", cudaGetErrorString(e));
    return 0;
}

__global__ void poke_kernel(float* p, float v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// Diagnostic: store `value` into n floats at `ptr` from a kernel on the current device; returns the CUDA error code after
// synchronising the stream (0 = ok).  out_host[0..3] = pointer attributes {type, device, current device, canAccessPeer}.
extern "C" int lgcn_debug_poke(float* ptr, float value, int32_t n, int32_t* out_host, lgcn_stream_t stream) {
    cudaPointerAttributes at; memset(&at, 0, sizeof(at));
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    int dev = -1; cudaGetDevice(&dev);
    int can = -1; if (e == cudaSuccess && at.device != dev) cudaDeviceCanAccessPeer(&can, dev, at.device);
    if (out_host) { out_host[0] = (int)at.type; out_host[1] = at.device; out_host[2] = dev; out_host[3] = can; }
    poke_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(ptr, value, n);
    e = cudaStreamSynchronize(as_stream(stream));
    if (e != cudaSuccess) return fail("debug_poke: %s", cudaGetErrorString(e));
    return 0;
}

// ---- host sampler: glibc rand() stream, same draw order as sampling.cpp ---------------------
// The reference's sampler draws from glibc's rand() (srand(seed), code/sources/sampling.cpp:88-91).  The generator is restated
// here — glibc's TYPE_3 additive feedback generator: degree 31, separation 3, seeded by the Lehmer LCG 16807 mod 2^31-1, first
// 310 outputs discarded, output = state word >> 1 — so that the stream (i) is the same on any C library, (ii) costs no lock per
// draw, and (iii) can be snapshotted and rewound (lgcn_sampler_get/set_state: the epoch prefetch in utils.py rewinds when the
// next call turns out not to be the epoch sample it ran ahead for).  tests/test_host.py checks it against this machine's libc
// rand() and against a recording of the reference's compiled module.
namespace lgcn {
struct GlibcRand {
    int32_t r[31]; int f, b;            // f = "front" index (starts at 3), b = "rear" index (starts at 0)
    void seed(uint32_t s) {
        if (s == 0) s = 1;
        r[0] = (int32_t)s;
        for (int i = 1; i < 31; ++i) {
            const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            long word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; ++i) next();
    }
    inline int next() {
        const uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
        r[f] = (int32_t)v;
        if (++f >= 31) { f = 0; ++b; } else if (++b >= 31) b = 0;
        return (int)(v >> 1);
    }
};
static GlibcRand g_rand = [] { GlibcRand g; g.seed(1); return g; }();     // an unseeded program behaves like srand(1)
}  // namespace lgcn

extern "C" void lgcn_sampler_seed(uint32_t seed) { g_rand.seed(seed); }
// state = 33 int32 words (31 ring words, front index, rear index)
extern "C" void lgcn_sampler_get_state(int32_t* state33_host) { if (state33_host) { memcpy(state33_host, g_rand.r, 31 * 4); state33_host[31] = g_rand.f; state33_host[32] = g_rand.b; } }
extern "C" int lgcn_sampler_set_state(const int32_t* state33_host) {
    if (!state33_host || state33_host[31] < 0 || state33_host[31] >= 31 || state33_host[32] < 0 || state33_host[32] >= 31) { set_error("sampler_set_state: bad state"); return 1; }
    memcpy(g_rand.r, state33_host, 31 * 4); g_rand.f = state33_host[31]; g_rand.b = state33_host[32];
    return 0;
}

extern "C" int64_t lgcn_sample_negative(int32_t user_num, int32_t item_num, int64_t train_num,
                                        const int64_t* allpos_indptr_host, const int32_t* allpos_items_host,
                                        int32_t neg_num, int32_t* out_host) {
    if (user_num <= 0 || item_num <= 0 || neg_num < 1 || !allpos_indptr_host || !allpos_items_host || !out_host) {
        set_error("sample_negative: bad arguments"); return -1;
    }
    const int64_t per_user = train_num / user_num;
    const int row = neg_num + 2;
    for (int32_t user = 0; user < user_num; ++user) {
        const int32_t* pos = allpos_items_host + allpos_indptr_host[user];
        const int64_t npos = allpos_indptr_host[user + 1] - allpos_indptr_host[user];
        if (npos <= 0 && per_user > 0) { set_error("sample_negative: user %d has no positive item", user); return -2; }
        if (npos >= item_num && per_user > 0) { set_error("sample_negative: user %d interacted with every item", user); return -3; }
        for (int64_t pair = 0; pair < per_user; ++pair) {
            int32_t* o = out_host + ((int64_t)user * per_user + pair) * row;
            o[0] = user;
            o[1] = pos[g_rand.next() % npos];
            for (int idx = 2; idx < row; ++idx) {
                int neg;
                do { neg = g_rand.next() % item_num; } while (std::find(pos, pos + npos, neg) != pos + npos);
                o[idx] = neg;
            }
        }
    }
    return (int64_t)user_num * per_user;
}

// sample_negative_ByUser (code/sources/sampling.cpp:58-86): one row per LISTED user, same draw order
extern "C" int64_t lgcn_sample_negative_by_user(const int32_t* users_host, int64_t n_listed, int32_t user_num, int32_t item_num,
                                                const int64_t* allpos_indptr_host, const int32_t* allpos_items_host,
                                                int32_t neg_num, int32_t* out_host) {
    if (n_listed < 0 || user_num <= 0 || item_num <= 0 || neg_num < 1 || !allpos_indptr_host || !allpos_items_host || (n_listed && (!users_host || !out_host))) {
        set_error("sample_negative_by_user: bad arguments"); return -1;
    }
    const int row = neg_num + 2;
    for (int64_t k = 0; k < n_listed; ++k) {
        const int32_t user = users_host[k];
        if (user < 0 || user >= user_num) { set_error("sample_negative_by_user: user %d outside [0,%d)", user, user_num); return -4; }
        const int32_t* pos = allpos_items_host + allpos_indptr_host[user];
        const int64_t npos = allpos_indptr_host[user + 1] - allpos_indptr_host[user];
        if (npos <= 0) { set_error("sample_negative_by_user: user %d has no positive item", user); return -2; }
        if (npos >= item_num) { set_error("sample_negative_by_user: user %d interacted with every item", user); return -3; }
        int32_t* o = out_host + k * row;
        o[0] = user;
        o[1] = pos[g_rand.next() % npos];
        for (int idx = 2; idx < row; ++idx) {
            int neg;
            do { neg = g_rand.next() % item_num; } while (std::find(pos, pos + npos, neg) != pos + npos);
            o[idx] = neg;
        }
    }
    return n_listed;
}

// randint (code/sources/sampling.cpp:22-25, exported to Python at :100): next value of the same rand() stream
extern "C" int32_t lgcn_randint(int32_t end) {
    if (end <= 0) { set_error("randint: end must be positive"); return -1; }
    return g_rand.next() % end;
}

// ---- ingest: the reference's interaction files --------------------------------------------------
// "uid item item ..." per line, whitespace separated; blank lines and lines without items are skipped
// (code/dataloader.py:82-115).  One pass over the file; pairs beyond `capacity` are counted, not stored.
extern "C" int64_t lgcn_parse_interactions(const char* path, int64_t* users_out_host, int64_t* items_out_host, int64_t capacity,
                                           int64_t* max_user_out_host, int64_t* max_item_out_host) {
    if (!path) { set_error("parse_interactions: null path"); return -1; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("parse_interactions: cannot open %s", path); return -1; }
    std::vector<char> buf;
    {
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
        fclose(f);
    }
    buf.push_back('\n');
    int64_t n = 0, max_u = -1, max_i = -1, line_no = 1;
    const char* p = buf.data();
    const char* end = p + buf.size();
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    while (p < end) {
        // one line
        int64_t uid = 0; bool have_uid = false; int64_t first_pair = n, line_max = -1;
        while (p < end && *p != '\n') {
            if (is_space(*p)) { ++p; continue; }
            bool neg = false;
            if (*p == '-' || *p == '+') { neg = (*p == '-'); ++p; }
            if (p >= end || *p < '0' || *p > '9') { set_error("parse_interactions: %s line %lld: not an integer", path, (long long)line_no); return -2; }
            int64_t v = 0;
            while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; }
            if (p < end && *p != '\n' && !is_space(*p)) { set_error("parse_interactions: %s line %lld: not an integer", path, (long long)line_no); return -2; }
            if (neg) v = -v;
            if (!have_uid) { uid = v; have_uid = true; continue; }
            if (n < capacity && users_out_host && items_out_host) { users_out_host[n] = uid; items_out_host[n] = v; }
            if (v > line_max) line_max = v;
            ++n;
        }
        if (n > first_pair) { if (uid > max_u) max_u = uid; if (line_max > max_i) max_i = line_max; }
        ++p; ++line_no;
    }
    if (max_user_out_host) *max_user_out_host = max_u;
    if (max_item_out_host) *max_item_out_host = max_i;
    return n;
}
