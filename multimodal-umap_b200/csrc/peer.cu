// Multi-GPU optimiser step over NVLink peer memory (one process per GPU).
//
// The reference is single-process; sharding its optimiser by edges leaves one exchange per epoch: every rank
// holds a partial gradient of the replicated embeddings (model.py:439-476 split over ranks).  Instead of an NCCL
// all-reduce of the whole gradient followed by the same Adam step on every rank, the reduction, the Adam step
// and the redistribution are ONE kernel over peer pointers: rank r owns a 1/W shard of the flat parameter
// buffer, sums that shard of all W gradient buffers with direct peer loads (fixed order: identical bits on
// every rank), applies torch.optim.Adam's arithmetic once, and stores the new parameters into all W replicas.
// Per rank and epoch (W-1)/W of the buffer crosses NVLink in each direction; two flag barriers in symmetric
// memory bracket the kernel ("gradients complete" before, "parameters delivered" after).
#include "common.cuh"

namespace mmu {

struct PeerPtrs {
    uint64_t p[MMU_PEER_MAX];
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// thread w tells rank w that this rank has arrived (sequence number `seq` in slot `slot`), then waits for rank w
__global__ void peer_barrier_kernel(PeerPtrs flags, int world, int rank, int slot, uint32_t seq) {
    const int w = threadIdx.x;
    if (w >= world) return;
    __threadfence_system();
    uint32_t *remote = reinterpret_cast<uint32_t *>(flags.p[w]) + slot * MMU_PEER_MAX + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(flags.p[rank]) + slot * MMU_PEER_MAX + w;
    const unsigned long long t0 = global_ns();
    uint32_t v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - seq) >= 0) break;
        if (global_ns() - t0 > 20000000000ull) __trap();      // a peer is gone: fail the launch instead of hanging
        __nanosleep(100);
    }
}

__device__ __forceinline__ float4 ld_peer(const float4 *p) {
    float4 r;    // peer memory is not cached in this GPU's L2; skip L1 as well (the line changes every epoch)
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(256)
adam_peer_kernel(PeerPtrs params, PeerPtrs grads, float *__restrict__ m, float *__restrict__ v, int64_t lo4, int64_t hi4,
                 int world, int rank, float beta2, float omb1, float omb2, float eps, const OptState *__restrict__ st) {
    const float neg_step = -st->step_size, bc2_sqrt = st->bc2_sqrt;
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {       // adam_kernel's arithmetic (layout_sgd.cu)
        mm = __fadd_rn(mm, __fmul_rn(omb1, __fsub_rn(gg, mm)));
        vv = __fadd_rn(__fmul_rn(vv, beta2), __fmul_rn(__fmul_rn(omb2, gg), gg));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        pp = __fadd_rn(pp, __fdiv_rn(__fmul_rn(neg_step, mm), denom));
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
        float4 gg = ld_peer(reinterpret_cast<const float4 *>(grads.p[0]) + i);
        for (int w = 1; w < world; ++w) {                                   // fixed order: the same sum on every rank
            const float4 t = ld_peer(reinterpret_cast<const float4 *>(grads.p[w]) + i);
            gg.x = __fadd_rn(gg.x, t.x); gg.y = __fadd_rn(gg.y, t.y); gg.z = __fadd_rn(gg.z, t.z); gg.w = __fadd_rn(gg.w, t.w);
        }
        float4 pp = reinterpret_cast<const float4 *>(params.p[rank])[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        for (int w = 0; w < world; ++w) reinterpret_cast<float4 *>(params.p[w])[i] = pp;
    }
}

static int fill_ptrs(PeerPtrs &dst, const uint64_t *src, int world) {
    for (int w = 0; w < MMU_PEER_MAX; ++w) dst.p[w] = w < world ? src[w] : 0;
    for (int w = 0; w < world; ++w)
        if (!src[w] || (src[w] & 15)) return 1;
    return 0;
}

}  // namespace mmu

extern "C" int mmu_peer_barrier(const uint64_t *peer_flags, int world, int rank, int slot, uint32_t seq,
                                mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(peer_flags && world >= 1 && world <= MMU_PEER_MAX && rank >= 0 && rank < world,
                  "mmu_peer_barrier: bad world/rank");
    MMU_CHECK_ARG(slot >= 0 && slot < 2, "mmu_peer_barrier: slot outside [0,2)");
    PeerPtrs f;
    MMU_CHECK_ARG(fill_ptrs(f, peer_flags, world) == 0, "mmu_peer_barrier: null or unaligned peer pointer");
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(f, world, rank, slot, seq);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_adam_step_peer(const uint64_t *peer_params, const uint64_t *peer_grads, float *m, float *v, int64_t n,
                                  int world, int rank, double beta1, double beta2, double eps, const uint32_t *state,
                                  mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(peer_params && peer_grads && m && v && state, "mmu_adam_step_peer: null pointer");
    MMU_CHECK_ARG(world >= 1 && world <= MMU_PEER_MAX && rank >= 0 && rank < world, "mmu_adam_step_peer: bad world/rank");
    MMU_CHECK_ARG(n >= 0 && (n & 3) == 0, "mmu_adam_step_peer: n must be a multiple of 4");
    MMU_CHECK_ARG(((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                  "mmu_adam_step_peer: m / v must be 16-byte aligned");
    PeerPtrs pp, gg;
    MMU_CHECK_ARG(fill_ptrs(pp, peer_params, world) == 0 && fill_ptrs(gg, peer_grads, world) == 0,
                  "mmu_adam_step_peer: null or unaligned peer pointer");
    const int64_t n4 = n >> 2;
    const int64_t lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;     // this rank's shard
    if (hi4 == lo4) return MMU_OK;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int64_t want = (hi4 - lo4 + 255) / 256;
    unsigned blocks = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    adam_peer_kernel<<<blocks, 256, 0, as_stream(stream)>>>(pp, gg, m, v, lo4, hi4, world, rank, (float)beta2,
                                                            (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                            reinterpret_cast<const OptState *>(state));
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
