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

}  // namespace lgcn

using namespace lgcn;

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
