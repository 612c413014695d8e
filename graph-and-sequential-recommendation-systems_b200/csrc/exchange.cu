// exchange.cu — device-side rendezvous of the ranks of the row partition.
//
// The reference has no multi-GPU path (code/dataloader.py:192-201 is the dead A_split); the row partition exchanges
// every layer's row block with stores issued by K1's epilogue (spmm.cu: multimem.st / peer st), so the only thing a
// layer still needs from "the collective" is ordering: all ranks' stores of layer k are visible before layer k+1 gathers.
// Round 1 used a 4-byte NCCL all-reduce for that (7 per step, ~35 us each, not graph-capturable together with the rest).
// This kernel is the replacement: one CTA, one thread per rank, flags in peer-mapped memory.
//
//   signal : fence.sys (orders this GPU's earlier stores, incl. the previous kernels' peer stores, before the flag)
//            st.release.sys  peer[p].flags[rank] = epoch          for every rank p (own slot included)
//   wait   : ld.acquire.sys  flags_local[p] until (int)(value - epoch) >= 0
//
// epoch lives in device memory and is incremented by the kernel, so a captured CUDA graph keeps counting across replays.
#include "common.cuh"

namespace lgcn {

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

struct BarrierArgs {
    unsigned* flags_local; unsigned* peer[LGCN_MAX_PEERS + 1];
    int rank, world; unsigned* epoch; int* err; unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(32) rank_barrier_kernel(const __grid_constant__ BarrierArgs a) {
    const int t = threadIdx.x;
    const unsigned e = *a.epoch + 1u;
    __syncwarp();
    if (t < a.world) {
        __threadfence_system();
        st_release_sys_u32(a.peer[t] + a.rank, e);
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while ((int)(ld_acquire_sys_u32(a.flags_local + t) - e) < 0) {
            if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > a.timeout_ns) {
                if (a.err) atomicExch(a.err, 1 + t);
                break;
            }
        }
        __threadfence_system();
    }
    __syncwarp();
    if (t == 0) *a.epoch = e;
}

// Copy between device memory and MAPPED PINNED HOST memory from a kernel (16-byte words).  Put at the two ends of the
// captured training step, it replaces the cudaMemcpyAsync H2D of the batch and the D2H of the loss: no copy-engine hop
// between the copies and the kernels, and stageOne becomes "write the pinned staging block, replay one graph, wait".
__global__ void __launch_bounds__(256) copy16_kernel(int4* __restrict__ dst, const int4* __restrict__ src, long long n16, int to_host) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src[i];
    if (to_host) __threadfence_system();
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_copy_words(void* dst, const void* src, int64_t n_bytes, int32_t dst_is_host, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(dst && src && n_bytes >= 0 && n_bytes % 16 == 0, "copy_words: bad arguments (n_bytes must be a multiple of 16)");
    LGCN_CHECK_ARG(((uintptr_t)dst % 16) == 0 && ((uintptr_t)src % 16) == 0, "copy_words: 16-byte alignment required");
    if (n_bytes == 0) return 0;
    const long long n16 = n_bytes / 16;
    copy16_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, as_stream(stream)>>>(static_cast<int4*>(dst), static_cast<const int4*>(src), n16, dst_is_host ? 1 : 0);
    LGCN_CHECK_LAUNCH("copy16_kernel");
    return 0;
}

extern "C" int lgcn_rank_barrier(uint32_t* flags_local, void* const* peer_flags_host, int32_t rank, int32_t world,
                                 uint32_t* epoch_dev, int32_t* err_dev, int32_t timeout_ms, lgcn_stream_t stream) {
    LGCN_CHECK_ARG(flags_local && peer_flags_host && epoch_dev, "rank_barrier: null argument");
    LGCN_CHECK_ARG(world >= 1 && world <= LGCN_MAX_PEERS + 1 && rank >= 0 && rank < world, "rank_barrier: rank %d of %d out of range", rank, world);
    BarrierArgs a;
    a.flags_local = flags_local; a.rank = rank; a.world = world; a.epoch = epoch_dev; a.err = err_dev;
    a.timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 20000) * 1000000ull;
    for (int p = 0; p <= LGCN_MAX_PEERS; ++p) a.peer[p] = nullptr;
    for (int p = 0; p < world; ++p) {
        LGCN_CHECK_ARG(peer_flags_host[p], "rank_barrier: flags of rank %d not mapped", p);
        a.peer[p] = static_cast<unsigned*>(peer_flags_host[p]);
    }
    rank_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(a);
    LGCN_CHECK_LAUNCH("rank_barrier_kernel");
    return 0;
}
