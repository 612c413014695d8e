// popgate.cu — the reference's popularity gate as kernels: fusion of the propagated item embeddings with a
// popularity vector, and the BPR step on the fused embeddings with its closed-form backward.
//
// Replaces  _fuse_item_embeddings   code/model.py:139-157   (pop MLP 1 -> H1 -> d, gate MLP 2d -> H2 -> 1, sigmoid mix)
//           bpr_loss (pop-gate on)  code/model.py:162-183   (BPR + L2 on the FUSED item rows, minus coeff * gate entropy)
//           and their autograd backward (code/utils.py:61) down to the rows of `out` and the 8 MLP tensors.
//
// Per item i with propagated row x (d floats) and popularity scalar p:
//   a1 = W1 p + b1, h1 = relu(a1)            (H1)          v = W2 h1 + b2              (d)   "pop_vec"
//   a2 = G1 [x; v] + c1, h2 = relu(a2)       (H2)          g = sigmoid((G2 h2 + c2)/T)       "gate"
//   f  = g x + (1 - g) v                                                              "fused"
// Backward for upstream (df, dg_ext):  dg = <df, x - v> + dg_ext;  dlogit = dg g (1-g) / T;  dG2 = dlogit h2; dc2 = dlogit;
//   da2 = dlogit G2 [a2 > 0];  dG1 = da2 (x) [x; v];  dc1 = da2;  dz = G1^T da2;  dx = g df + dz[:d];  dv = (1-g) df + dz[d:];
//   dW2 = dv (x) h1;  db2 = dv;  da1 = (W2^T dv) [a1 > 0];  dW1 = da1 p;  db1 = da1.
// One WARP per item (forward-all kernel) or per triple (BPR kernel: pos and neg items back to back).  The 10.5 k MLP floats
// sit in shared memory with padded row strides (conflict-free for both orientations); weight gradients are summed in shared
// memory per CTA and flushed with one global atomicAdd per weight per CTA.  The arithmetic is tiny (~30 k MAC per item,
// <= 4096 items per step): this is a latency kernel, written for clarity; what matters is that the variant trains through
// the same fused step as the plain model (no autograd graph, no materialised M x d intermediates per step).
//
// Parameter block `pg` (flat float32, the layout lgcn_popgate_param_count documents):
//   W1[H1] b1[H1] W2[d][H1] b2[d] G1[H2][2d] c1[H2] G2[H2] c2[1]      (= nn.Linear weights/biases of pop_mlp.0/.2, gate_mlp.0/.2)
#include "common.cuh"

namespace lgcn {

constexpr int kPgThreads = 128;          // 4 warps per CTA
constexpr int kPgWarps = kPgThreads / 32;
constexpr int kPgMaxH1 = 64, kPgMaxH2 = 128;

struct PgDims { int d, H1, H2; };
__host__ __device__ inline int pg_off_b1(const PgDims& s) { return s.H1; }
__host__ __device__ inline int pg_off_W2(const PgDims& s) { return 2 * s.H1; }
__host__ __device__ inline int pg_off_b2(const PgDims& s) { return 2 * s.H1 + s.d * s.H1; }
__host__ __device__ inline int pg_off_G1(const PgDims& s) { return pg_off_b2(s) + s.d; }
__host__ __device__ inline int pg_off_c1(const PgDims& s) { return pg_off_G1(s) + s.H2 * 2 * s.d; }
__host__ __device__ inline int pg_off_G2(const PgDims& s) { return pg_off_c1(s) + s.H2; }
__host__ __device__ inline int pg_off_c2(const PgDims& s) { return pg_off_G2(s) + s.H2; }
__host__ __device__ inline int pg_count(const PgDims& s) { return pg_off_c2(s) + 1; }

// shared-memory image of the MLPs: rows of W2 padded to H1+1 and rows of G1 to 2d+1 floats
struct PgSmem {
    float *W1, *b1, *W2, *b2, *G1, *c1, *G2, *c2;
    __device__ void carve(float* base, const PgDims& s) {
        W1 = base; b1 = W1 + s.H1; W2 = b1 + s.H1; b2 = W2 + s.d * (s.H1 + 1);
        G1 = b2 + s.d; c1 = G1 + s.H2 * (2 * s.d + 1); G2 = c1 + s.H2; c2 = G2 + s.H2;
    }
};
__host__ __device__ inline int pg_smem_floats(const PgDims& s) { return 2 * s.H1 + s.d * (s.H1 + 1) + s.d + s.H2 * (2 * s.d + 1) + 2 * s.H2 + 1; }

__device__ void pg_load_weights(PgSmem& w, const float* __restrict__ pg, const PgDims& s) {
    for (int i = threadIdx.x; i < s.H1; i += blockDim.x) { w.W1[i] = pg[i]; w.b1[i] = pg[pg_off_b1(s) + i]; }
    for (int i = threadIdx.x; i < s.d * s.H1; i += blockDim.x) w.W2[(i / s.H1) * (s.H1 + 1) + i % s.H1] = pg[pg_off_W2(s) + i];
    for (int i = threadIdx.x; i < s.d; i += blockDim.x) w.b2[i] = pg[pg_off_b2(s) + i];
    for (int i = threadIdx.x; i < s.H2 * 2 * s.d; i += blockDim.x) w.G1[(i / (2 * s.d)) * (2 * s.d + 1) + i % (2 * s.d)] = pg[pg_off_G1(s) + i];
    for (int i = threadIdx.x; i < s.H2; i += blockDim.x) { w.c1[i] = pg[pg_off_c1(s) + i]; w.G2[i] = pg[pg_off_G2(s) + i]; }
    if (threadIdx.x == 0) w.c2[0] = pg[pg_off_c2(s)];
}

// per-warp scratch of ONE item: what the backward needs from the forward
template <int D> struct PgItem { float z[2 * D]; float a1[kPgMaxH1]; float a2[kPgMaxH2]; float da2[kPgMaxH2]; float dz[2 * D]; };

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// forward of one item by one warp.  x[q] = element lane + 32 q of the propagated row.  Returns the gate; f[] = fused row.
template <int D>
__device__ float pg_forward(const PgSmem& w, const PgDims& s, PgItem<D>& it, const float (&x)[D / 32], float p, float inv_temp,
                            float (&v)[D / 32], float (&f)[D / 32]) {
    constexpr int KPL = D / 32;
    const int lane = threadIdx.x & 31;
    for (int j = lane; j < s.H1; j += 32) it.a1[j] = fmaf(w.W1[j], p, w.b1[j]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int k = lane + 32 * q;
        float acc = w.b2[k];
        const float* row = w.W2 + k * (s.H1 + 1);
        for (int j = 0; j < s.H1; ++j) acc = fmaf(row[j], fmaxf(it.a1[j], 0.f), acc);
        v[q] = acc;
        it.z[k] = x[q]; it.z[D + k] = acc;
    }
    __syncwarp();
    float part = 0.f;
    for (int m = lane; m < s.H2; m += 32) {
        float acc = w.c1[m];
        const float* row = w.G1 + m * (2 * D + 1);
        for (int k = 0; k < 2 * D; ++k) acc = fmaf(row[k], it.z[k], acc);
        it.a2[m] = acc;
        part = fmaf(w.G2[m], fmaxf(acc, 0.f), part);
    }
    const float logit = (warp_sum(part) + w.c2[0]) * inv_temp;
    const float g = 1.f / (1.f + expf(-logit));
#pragma unroll
    for (int q = 0; q < KPL; ++q) f[q] = g * x[q] + (1.f - g) * v[q];
    __syncwarp();
    return g;
}

// backward of one item by one warp; weight gradients go to the CTA's shared accumulators `gw` (same padded layout as PgSmem).
// Returns dx[] (gradient of the propagated row).
template <int D>
__device__ void pg_backward(const PgSmem& w, const PgSmem& gw, const PgDims& s, PgItem<D>& it, const float (&x)[D / 32], const float (&v)[D / 32],
                            float p, float g, float inv_temp, const float (&df)[D / 32], float dg_ext, float (&dx)[D / 32]) {
    constexpr int KPL = D / 32;
    const int lane = threadIdx.x & 31;
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < KPL; ++q) part = fmaf(df[q], x[q] - v[q], part);
    const float dg = warp_sum(part) + dg_ext;
    const float dlogit = dg * g * (1.f - g) * inv_temp;
    if (lane == 0) atomicAdd(gw.c2, dlogit);
    for (int m = lane; m < s.H2; m += 32) {
        const float a2 = it.a2[m];
        atomicAdd(gw.G2 + m, dlogit * fmaxf(a2, 0.f));
        const float da2 = a2 > 0.f ? dlogit * w.G2[m] : 0.f;
        it.da2[m] = da2;
        atomicAdd(gw.c1 + m, da2);
    }
    __syncwarp();
    for (int kk = lane; kk < 2 * D; kk += 32) {
        const float zk = it.z[kk];
        float acc = 0.f;
        for (int m = 0; m < s.H2; ++m) {
            const float d2 = it.da2[m];
            if (d2 != 0.f) { acc = fmaf(w.G1[m * (2 * D + 1) + kk], d2, acc); atomicAdd(gw.G1 + m * (2 * D + 1) + kk, d2 * zk); }
        }
        it.dz[kk] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int k = lane + 32 * q;
        dx[q] = fmaf(g, df[q], it.dz[k]);
        const float dv = fmaf(1.f - g, df[q], it.dz[D + k]);
        it.dz[D + k] = dv;
        atomicAdd(gw.b2 + k, dv);
        float* grow = gw.W2 + k * (s.H1 + 1);
        for (int j = 0; j < s.H1; ++j) { const float h = it.a1[j]; if (h > 0.f) atomicAdd(grow + j, dv * h); }
    }
    __syncwarp();
    for (int j = lane; j < s.H1; j += 32) {
        float dh = 0.f;
        for (int k = 0; k < D; ++k) dh = fmaf(w.W2[k * (s.H1 + 1) + j], it.dz[D + k], dh);
        const float da1 = it.a1[j] > 0.f ? dh : 0.f;
        atomicAdd(gw.W1 + j, da1 * p);
        atomicAdd(gw.b1 + j, da1);
    }
    __syncwarp();
}

__device__ void pg_flush_grads(const PgSmem& gw, float* __restrict__ gpg, const PgDims& s) {
    for (int i = threadIdx.x; i < s.H1; i += blockDim.x) { atomicAdd(gpg + i, gw.W1[i]); atomicAdd(gpg + pg_off_b1(s) + i, gw.b1[i]); }
    for (int i = threadIdx.x; i < s.d * s.H1; i += blockDim.x) { const float g = gw.W2[(i / s.H1) * (s.H1 + 1) + i % s.H1]; if (g != 0.f) atomicAdd(gpg + pg_off_W2(s) + i, g); }
    for (int i = threadIdx.x; i < s.d; i += blockDim.x) atomicAdd(gpg + pg_off_b2(s) + i, gw.b2[i]);
    for (int i = threadIdx.x; i < s.H2 * 2 * s.d; i += blockDim.x) { const float g = gw.G1[(i / (2 * s.d)) * (2 * s.d + 1) + i % (2 * s.d)]; if (g != 0.f) atomicAdd(gpg + pg_off_G1(s) + i, g); }
    for (int i = threadIdx.x; i < s.H2; i += blockDim.x) { atomicAdd(gpg + pg_off_c1(s) + i, gw.c1[i]); atomicAdd(gpg + pg_off_G2(s) + i, gw.G2[i]); }
    if (threadIdx.x == 0) atomicAdd(gpg + pg_off_c2(s), gw.c2[0]);
}

struct PgArgs {
    const float* out; const float* pop; const float* pg; PgDims s; float inv_temp;
    int n_users, m_items;
    // forward-all
    float* fused; float* gate;
    // BPR step
    const long long* users; const long long* pos; const long long* neg; int B_cap; const int* ctl;
    float decay, entropy_coeff;
    float* loss_out; float* G; float* gpg;
    int* counter; float* partials;
};

// fused[i] = g x + (1-g) v for every item (evaluation / getUsersRating / getEmbedding outside the fused step)
template <int D>
__global__ void __launch_bounds__(kPgThreads)
popgate_fuse_kernel(const __grid_constant__ PgArgs a) {
    extern __shared__ float smem[];
    PgSmem w; w.carve(smem, a.s);
    PgItem<D>* items = reinterpret_cast<PgItem<D>*>(smem + ((pg_smem_floats(a.s) + 3) & ~3));
    pg_load_weights(w, a.pg, a.s);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x * kPgWarps + warp; i < a.m_items; i += gridDim.x * kPgWarps) {
        float x[D / 32], v[D / 32], f[D / 32];
        const float* row = a.out + (size_t)(a.n_users + i) * D;
#pragma unroll
        for (int q = 0; q < D / 32; ++q) x[q] = row[lane + 32 * q];
        const float g = pg_forward<D>(w, a.s, items[warp], x, a.pop[i], a.inv_temp, v, f);
#pragma unroll
        for (int q = 0; q < D / 32; ++q) a.fused[(size_t)i * D + lane + 32 * q] = f[q];
        if (a.gate && lane == 0) a.gate[i] = g;
    }
}

// BPR on (out[user], fused[pos], fused[neg]) with the gate-entropy term, and the whole backward: G += d total / d out,
// gpg += d total / d (MLP parameters), total = (bpr - coeff * entropy) + decay * reg   (code/model.py:162-183, code/utils.py:55-57)
template <int D>
__global__ void __launch_bounds__(kPgThreads)
popgate_bpr_kernel(const __grid_constant__ PgArgs a) {
    constexpr int KPL = D / 32;
    extern __shared__ float smem[];
    const int wfl = (pg_smem_floats(a.s) + 3) & ~3;
    PgSmem w, gw; w.carve(smem, a.s); gw.carve(smem + wfl, a.s);
    PgItem<D>* items = reinterpret_cast<PgItem<D>*>(smem + 2 * wfl);        // [warp][2]: pos and neg of the current triple
    __shared__ float s_part[kPgWarps][3];
    __shared__ int s_last;
    pg_load_weights(w, a.pg, a.s);
    for (int i = threadIdx.x; i < wfl; i += blockDim.x) smem[wfl + i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int off = a.ctl[0];
    const int B = min(a.ctl[1], a.B_cap);
    const float inv = 1.f / (float)(B > 0 ? B : 1);
    float bpr_w = 0.f, reg_w = 0.f, ent_w = 0.f;
    for (int t = blockIdx.x * kPgWarps + warp; t < B; t += gridDim.x * kPgWarps) {
        const long long u = a.users[off + t], pi = a.pos[off + t], ni = a.neg[off + t];
        float U[KPL], XP[KPL], XN[KPL], VP[KPL], VN[KPL], FP[KPL], FN[KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            U[q] = a.out[(size_t)u * D + lane + 32 * q];
            XP[q] = a.out[(size_t)(a.n_users + pi) * D + lane + 32 * q];
            XN[q] = a.out[(size_t)(a.n_users + ni) * D + lane + 32 * q];
        }
        const float pp = a.pop[pi], pn = a.pop[ni];
        PgItem<D>& ip = items[2 * warp]; PgItem<D>& in = items[2 * warp + 1];
        const float gp = pg_forward<D>(w, a.s, ip, XP, pp, a.inv_temp, VP, FP);
        const float gn = pg_forward<D>(w, a.s, in, XN, pn, a.inv_temp, VN, FN);
        float ps = 0.f, ns = 0.f, uu = 0.f, fpp = 0.f, fnn = 0.f;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            ps = fmaf(U[q], FP[q], ps); ns = fmaf(U[q], FN[q], ns);
            uu = fmaf(U[q], U[q], uu); fpp = fmaf(FP[q], FP[q], fpp); fnn = fmaf(FN[q], FN[q], fnn);
        }
        ps = warp_sum(ps); ns = warp_sum(ns); uu = warp_sum(uu); fpp = warp_sum(fpp); fnn = warp_sum(fnn);
        const float z = ps - ns;
        const float ez = expf(-fabsf(z));
        bpr_w += fmaxf(-z, 0.f) + log1pf(ez);                                  // softplus(-z) = -logsigmoid(z)
        const float sg = (z >= 0.f) ? ez / (1.f + ez) : 1.f / (1.f + ez);      // sigmoid(-z)
        reg_w += 0.5f * (uu + fpp + fnn);
        // gate entropy over the 2B gates of the batch (clamped like torch.clamp: no gradient outside the range)
        const float lo = 1e-6f, hi = 1.f - 1e-6f;
        const float gcp = fminf(fmaxf(gp, lo), hi), gcn = fminf(fmaxf(gn, lo), hi);
        ent_w += -(gcp * logf(gcp) + (1.f - gcp) * logf(1.f - gcp)) - (gcn * logf(gcn) + (1.f - gcn) * logf(1.f - gcn));
        const float ce = a.entropy_coeff * 0.5f * inv;
        const float dgp = (gp > lo && gp < hi) ? ce * (logf(gcp) - logf(1.f - gcp)) : 0.f;
        const float dgn = (gn > lo && gn < hi) ? ce * (logf(gcn) - logf(1.f - gcn)) : 0.f;
        if (a.G != nullptr) {
            const float ca = sg * inv, cr = a.decay * inv;
            float dfp[KPL], dfn[KPL], dxp[KPL], dxn[KPL];
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                atomicAdd(a.G + (size_t)u * D + lane + 32 * q, ca * (FN[q] - FP[q]) + cr * U[q]);
                dfp[q] = -ca * U[q] + cr * FP[q];
                dfn[q] = ca * U[q] + cr * FN[q];
            }
            pg_backward<D>(w, gw, a.s, ip, XP, VP, pp, gp, a.inv_temp, dfp, dgp, dxp);
            pg_backward<D>(w, gw, a.s, in, XN, VN, pn, gn, a.inv_temp, dfn, dgn, dxn);
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                atomicAdd(a.G + (size_t)(a.n_users + pi) * D + lane + 32 * q, dxp[q]);
                atomicAdd(a.G + (size_t)(a.n_users + ni) * D + lane + 32 * q, dxn[q]);
            }
        }
    }
    if (lane == 0) { s_part[warp][0] = bpr_w; s_part[warp][1] = reg_w; s_part[warp][2] = ent_w; }
    __syncthreads();
    if (a.gpg != nullptr) pg_flush_grads(gw, a.gpg, a.s);
    if (threadIdx.x == 0) {
        float l = 0.f, r = 0.f, e = 0.f;
        for (int q = 0; q < kPgWarps; ++q) { l += s_part[q][0]; r += s_part[q][1]; e += s_part[q][2]; }
        a.partials[3 * blockIdx.x] = l; a.partials[3 * blockIdx.x + 1] = r; a.partials[3 * blockIdx.x + 2] = e;
        __threadfence();
        s_last = (atomicAdd(a.counter, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 32) {                                  // fixed-order reduction of the CTA partials
        float l = 0.f, r = 0.f, e = 0.f;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) { l += __ldcg(a.partials + 3 * b); r += __ldcg(a.partials + 3 * b + 1); e += __ldcg(a.partials + 3 * b + 2); }
        l = warp_sum(l); r = warp_sum(r); e = warp_sum(e);
        if (threadIdx.x == 0) {
            const float loss = l * inv - a.entropy_coeff * (e * 0.5f * inv), reg = r * inv;
            const float total = loss + a.decay * reg;
            a.loss_out[0] = loss; a.loss_out[1] = reg; a.loss_out[2] = total; a.loss_out[3] += total;
            *a.counter = 0;
        }
    }
}

static int pg_grid(int work_items) {
    int g = (work_items + kPgWarps - 1) / kPgWarps;
    const int cap = 2 * sm_count();
    return g < 1 ? 1 : (g > cap ? cap : g);
}

static int pg_check(const PgDims& s) {
    LGCN_CHECK_ARG(s.d == 32 || s.d == 64 || s.d == 128, "popgate: d=%d unsupported (32,64,128)", s.d);
    LGCN_CHECK_ARG(s.H1 >= 1 && s.H1 <= kPgMaxH1 && s.H2 >= 1 && s.H2 <= kPgMaxH2, "popgate: hidden sizes (%d,%d) out of range (<= %d, <= %d)", s.H1, s.H2, kPgMaxH1, kPgMaxH2);
    return 0;
}

template <int D> static size_t pg_fuse_smem(const PgDims& s) { return sizeof(float) * ((pg_smem_floats(s) + 3) & ~3) + sizeof(PgItem<D>) * kPgWarps; }
template <int D> static size_t pg_bpr_smem(const PgDims& s) { return sizeof(float) * 2 * ((pg_smem_floats(s) + 3) & ~3) + sizeof(PgItem<D>) * 2 * kPgWarps; }

template <int D>
static int launch_fuse(const PgArgs& a, cudaStream_t st) {
    const size_t smem = pg_fuse_smem<D>(a.s);
    LGCN_CHECK_ARG(smem <= (size_t)max_smem_optin(), "popgate: %zu bytes of shared memory needed", smem);
    cudaFuncSetAttribute(popgate_fuse_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    popgate_fuse_kernel<D><<<pg_grid(a.m_items), kPgThreads, smem, st>>>(a);
    LGCN_CHECK_LAUNCH("popgate_fuse_kernel");
    return 0;
}

template <int D>
static int launch_pg_bpr(const PgArgs& a, cudaStream_t st, int grid) {
    const size_t smem = pg_bpr_smem<D>(a.s);
    LGCN_CHECK_ARG(smem <= (size_t)max_smem_optin(), "popgate: %zu bytes of shared memory needed", smem);
    cudaFuncSetAttribute(popgate_bpr_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    popgate_bpr_kernel<D><<<grid, kPgThreads, smem, st>>>(a);
    LGCN_CHECK_LAUNCH("popgate_bpr_kernel");
    return 0;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int32_t lgcn_popgate_param_count(int32_t d, int32_t pop_hidden, int32_t gate_hidden) {
    return pg_count(PgDims{d, pop_hidden, gate_hidden});
}

extern "C" int lgcn_popgate_fuse(const float* out, int32_t n_users, int32_t m_items, int32_t d, const float* item_pop,
                                 const float* params, int32_t pop_hidden, int32_t gate_hidden, float temperature,
                                 float* fused_out, float* gate_out, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(out && item_pop && params && fused_out && m_items > 0 && n_users >= 0 && temperature > 0.f, "popgate_fuse: bad arguments");
    PgArgs a{};
    a.out = out; a.pop = item_pop; a.pg = params; a.s = PgDims{d, pop_hidden, gate_hidden}; a.inv_temp = 1.f / temperature;
    a.n_users = n_users; a.m_items = m_items; a.fused = fused_out; a.gate = gate_out;
    if (int rc = pg_check(a.s)) return rc;
    cudaStream_t st = as_stream(stream);
    switch (d) {
        case 32:  return launch_fuse<32>(a, st);
        case 64:  return launch_fuse<64>(a, st);
        default:  return launch_fuse<128>(a, st);
    }
}

extern "C" size_t lgcn_popgate_bpr_workspace_bytes(int32_t B_cap) {
    if (B_cap <= 0) return 0;
    return 16 + sizeof(float) * 3 * (size_t)(2 * 1024);        // counter + 3 partials per CTA (grid <= 2 x SM count)
}

extern "C" int lgcn_popgate_bpr_fwd_bwd(const float* out, const int64_t* users, const int64_t* pos, const int64_t* neg,
                                        int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items, int32_t d,
                                        const float* item_pop, const float* params, int32_t pop_hidden, int32_t gate_hidden,
                                        float temperature, float entropy_coeff, float decay,
                                        float* loss_out, float* G, float* params_grad,
                                        void* workspace, size_t workspace_bytes, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(out && users && pos && neg && batch_ctl_dev && item_pop && params && loss_out, "popgate_bpr: null argument");
    LGCN_CHECK_ARG(B_cap > 0 && temperature > 0.f, "popgate_bpr: bad B_cap / temperature");
    LGCN_CHECK_ARG((G == nullptr) == (params_grad == nullptr), "popgate_bpr: G and params_grad go together (both NULL = forward only)");
    LGCN_CHECK_ARG(workspace && workspace_bytes >= lgcn_popgate_bpr_workspace_bytes(B_cap) && ((uintptr_t)workspace % 16) == 0, "popgate_bpr: workspace too small or misaligned");
    PgArgs a{};
    a.out = out; a.pop = item_pop; a.pg = params; a.s = PgDims{d, pop_hidden, gate_hidden}; a.inv_temp = 1.f / temperature;
    a.n_users = n_users; a.m_items = m_items;
    a.users = reinterpret_cast<const long long*>(users); a.pos = reinterpret_cast<const long long*>(pos); a.neg = reinterpret_cast<const long long*>(neg);
    a.B_cap = B_cap; a.ctl = batch_ctl_dev; a.decay = decay; a.entropy_coeff = entropy_coeff;
    a.loss_out = loss_out; a.G = G; a.gpg = params_grad;
    a.counter = static_cast<int*>(workspace); a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16);
    if (int rc = pg_check(a.s)) return rc;
    int grid = pg_grid(B_cap);
    if (grid > 2048) grid = 2048;
    cudaStream_t st = as_stream(stream);
    switch (d) {
        case 32:  return launch_pg_bpr<32>(a, st, grid);
        case 64:  return launch_pg_bpr<64>(a, st, grid);
        default:  return launch_pg_bpr<128>(a, st, grid);
    }
}
