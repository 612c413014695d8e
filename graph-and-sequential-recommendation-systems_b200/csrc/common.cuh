// common.cuh — shared device/host helpers for liblgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/lgcn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblgcn_b200 is written for sm_100a (B200) only"
#endif

namespace lgcn {

// ---- error reporting (thread-local message behind lgcn_last_error) -------------------------
void set_error(const char* fmt, ...);
int fail(const char* fmt, ...);  // sets the message, returns 1

#define LGCN_CHECK_ARG(cond, ...) \
    do { if (!(cond)) return ::lgcn::fail(__VA_ARGS__); } while (0)

#define LGCN_CHECK_LAUNCH(what) \
    do { cudaError_t e__ = cudaGetLastError(); \
         if (e__ != cudaSuccess) return ::lgcn::fail("%s: %s", what, cudaGetErrorString(e__)); } while (0)

static inline cudaStream_t as_stream(lgcn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();            // cached cudaDevAttrMultiProcessorCount of the current device
int max_smem_optin();      // cached cudaDevAttrMaxSharedMemoryPerBlockOptin

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__

// streaming 4-byte loads that should not displace gathered embedding rows in L1
__device__ __forceinline__ int ld_stream_i32(const int* p) {
    int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
// 16-byte gather through the read-only path, keep in L1 (rows of popular nodes are re-read)
__device__ __forceinline__ float4 ld_gather_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// own-row operands are read exactly once
__device__ __forceinline__ float4 ld_once_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// L2-coherent load (bypasses the non-coherent L1) for data another CTA wrote in this launch
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// NVSwitch multicast store: `p` is a multimem address, the 16 bytes land in every replica of the group (sm_90+)
__device__ __forceinline__ void st_multicast_f4(float4* p, const float4& v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// vectorised reduction: one 16-byte atomic add, no return value (sm_90+)
__device__ __forceinline__ void red_add_f4(float4* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& x) {
    a.x = fmaf(s, x.x, a.x); a.y = fmaf(s, x.y, a.y); a.z = fmaf(s, x.z, a.z); a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void f4_add(float4& a, const float4& x) {
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// torch.optim.Adam (defaults) element update with the rounding points pinned: exp_avg.lerp_(g, 1-b1);
// exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2); denom = sqrt(v)/bc2_sqrt + eps; p.addcdiv_(m, denom, -step)
__device__ __forceinline__ void adam_update1(float& p, float& m, float& v, float g, float b2, float w1, float w2,
                                             float step_size, float bc2s, float eps) {
    m = __fmaf_rn(w1, __fsub_rn(g, m), m);
    v = __fmaf_rn(__fmul_rn(w2, g), g, __fmul_rn(v, b2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);
    p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);
}

#endif  // __CUDACC__

}  // namespace lgcn
