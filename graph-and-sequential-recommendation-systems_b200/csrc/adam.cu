// adam.cu — torch.optim.Adam (defaults) as device kernels; replaces self.opt.step()
// (reference code/utils.py:51,62).  The bias-correction scalars are computed on the device in
// double precision — the same arithmetic torch does with Python floats — so a whole training step
// can sit inside one CUDA graph with no host-side step counter.
#include "common.cuh"

namespace lgcn {

__device__ void adam_refresh(lgcn_adam_scalars_t* s) {
    const int t = s->step > 0 ? s->step : 1;
    s->step_size = (float)(s->lr_d / (1.0 - pow(s->beta1_d, (double)t)));
    s->bc2_sqrt = (float)sqrt(1.0 - pow(s->beta2_d, (double)t));
}

__global__ void adam_init_kernel(lgcn_adam_scalars_t* s, double lr, double b1, double b2, double eps, int step) {
    s->lr_d = lr; s->beta1_d = b1; s->beta2_d = b2;
    s->beta1 = (float)b1; s->beta2 = (float)b2; s->w1 = (float)(1.0 - b1); s->w2 = (float)(1.0 - b2);
    s->eps = (float)eps; s->pad0 = 0.f; s->step = step; s->pad1 = 0;
    adam_refresh(s);
}

__global__ void adam_tick_kernel(lgcn_adam_scalars_t* s) {
    s->step = s->step + 1;
    adam_refresh(s);
}

// The head of a captured training step in ONE launch: Adam step counter + bias corrections (thread 0), and either the
// move of the resident epoch's batch window (thread 0) or the pull of a host batch out of mapped pinned memory (all threads).
__global__ void __launch_bounds__(256)
step_begin_kernel(lgcn_adam_scalars_t* s, int* advance_ctl, int B_cap, int4* __restrict__ stage_dst, const int4* __restrict__ stage_src, long long n16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        s->step = s->step + 1;
        adam_refresh(s);
        if (advance_ctl != nullptr) {
            const int off = advance_ctl[0] + advance_ctl[1];
            int B = advance_ctl[2] - off; if (B > B_cap) B = B_cap; if (B < 0) B = 0;
            advance_ctl[0] = off; advance_ctl[1] = B;
        }
    }
    if (i < n16) stage_dst[i] = stage_src[i];
}

__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, const float4* __restrict__ G,
            long long n4, const lgcn_adam_scalars_t* __restrict__ sc) {
    const float b2 = sc->beta2, eps = sc->eps, step_size = sc->step_size, bc2s = sc->bc2_sqrt;
    const float w1 = sc->w1, w2 = sc->w2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p = P[i], m = M[i], v = V[i]; const float4 g = ld_once_f4(G + i);
        adam_update1(p.x, m.x, v.x, g.x, b2, w1, w2, step_size, bc2s, eps);
        adam_update1(p.y, m.y, v.y, g.y, b2, w1, w2, step_size, bc2s, eps);
        adam_update1(p.z, m.z, v.z, g.z, b2, w1, w2, step_size, bc2s, eps);
        adam_update1(p.w, m.w, v.w, g.w, b2, w1, w2, step_size, bc2s, eps);
        P[i] = p; M[i] = m; V[i] = v;
    }
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_adam_init(lgcn_adam_scalars_t* scalars_dev, double lr, double beta1, double beta2, double eps,
                              int32_t step, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(scalars_dev, "adam_init: null scalars");
    adam_init_kernel<<<1, 1, 0, as_stream(stream)>>>(scalars_dev, lr, beta1, beta2, eps, step);
    LGCN_CHECK_LAUNCH("adam_init_kernel");
    return 0;
}

extern "C" int lgcn_adam_tick(lgcn_adam_scalars_t* scalars_dev, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(scalars_dev, "adam_tick: null scalars");
    adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(scalars_dev);
    LGCN_CHECK_LAUNCH("adam_tick_kernel");
    return 0;
}

extern "C" int lgcn_adam_f32(float* P, float* M, float* V, const float* G, int64_t n,
                             const lgcn_adam_scalars_t* scalars_dev, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(P && M && V && G && scalars_dev, "adam: null argument");
    LGCN_CHECK_ARG(n >= 0 && n % 4 == 0, "adam: n=%lld must be a multiple of 4", (long long)n);
    LGCN_CHECK_ARG(((uintptr_t)P % 16) == 0 && ((uintptr_t)M % 16) == 0 && ((uintptr_t)V % 16) == 0 && ((uintptr_t)G % 16) == 0, "adam: 16-byte alignment required");
    if (n == 0) return 0;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(P), reinterpret_cast<float4*>(M),
        reinterpret_cast<float4*>(V), reinterpret_cast<const float4*>(G), n4, scalars_dev);
    LGCN_CHECK_LAUNCH("adam_kernel");
    return 0;
}

extern "C" int lgcn_step_begin(lgcn_adam_scalars_t* scalars_dev, int32_t* advance_ctl_dev, int32_t B_cap,
                               void* stage_dst, const void* stage_src, int64_t stage_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(scalars_dev && B_cap > 0, "step_begin: null scalars or bad B_cap");
    LGCN_CHECK_ARG(stage_bytes >= 0 && stage_bytes % 16 == 0 && (stage_bytes == 0 || (stage_dst && stage_src)), "step_begin: bad staging arguments");
    LGCN_CHECK_ARG(((uintptr_t)stage_dst % 16) == 0 && ((uintptr_t)stage_src % 16) == 0, "step_begin: staging blocks must be 16-byte aligned");
    LGCN_CHECK_ARG(!(advance_ctl_dev && stage_bytes), "step_begin: a step either advances the resident window or pulls a host batch");
    const long long n16 = stage_bytes / 16;
    const unsigned blocks = (unsigned)(n16 > 0 ? (n16 + 255) / 256 : 1);
    step_begin_kernel<<<blocks, 256, 0, as_stream(stream)>>>(scalars_dev, advance_ctl_dev, B_cap, static_cast<int4*>(stage_dst),
                                                             static_cast<const int4*>(stage_src), n16);
    LGCN_CHECK_LAUNCH("step_begin_kernel");
    return 0;
}
