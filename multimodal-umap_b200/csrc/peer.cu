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

// ------------------------------------------------------------------ fused epoch tail (one launch per epoch)
// Everything of a multi-GPU epoch after the force kernels in ONE kernel (round 1 used five launches: state advance,
// barrier, shard step, barrier, gradient clear -- 131 us per epoch at 8 GPUs, most of it launch gaps and two
// round trips of flags):
//   1. block 0 tells every peer "my gradients are complete" (flag slot 0); every block waits for all peers' flags;
//   2. each block takes a slice of this rank's 1/W shard: sum of the W partial gradients -- ONE
//      multimem.ld_reduce.add per 16 bytes when the buffers are mapped to an NVSwitch multicast address (the switch
//      adds the W replicas: one request per element instead of W), else W peer loads in rank order --, Adam once,
//      new parameters stored to all W replicas (one multimem.st, or W peer stores), and ZEROS stored over the shard
//      of all W gradient buffers: only this rank reads that shard, and no rank accumulates into it again before the
//      closing barrier, so the separate gradient clear disappears;
//   3. the last block to finish tells every peer "my shard is delivered" (flag slot 1); block 0 waits for all peers'
//      slot-1 flags (the next epoch's force kernels read parameters the peers wrote) and advances the device-resident
//      optimiser state (epoch, Adam step, step size, bias correction) -- the bias corrections of THIS step are
//      computed by every block from the old state at its start.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4 *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_peer(float4 *p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void flags_arrive(const PeerPtrs &flags, int world, int rank, int slot, uint32_t seq) {
    // thread w < world tells rank w; called by one warp
    const int w = threadIdx.x;
    if (w < world) {
        uint32_t *remote = reinterpret_cast<uint32_t *>(flags.p[w]) + slot * MMU_PEER_MAX + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
    }
}
__device__ __forceinline__ void flags_wait(const PeerPtrs &flags, int world, int rank, int slot, uint32_t seq) {
    const int w = threadIdx.x;
    if (w < world) {
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(flags.p[rank]) + slot * MMU_PEER_MAX + w;
        const unsigned long long t0 = global_ns();
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - seq) >= 0) break;
            if (global_ns() - t0 > 20000000000ull) __trap();  // a peer is gone: fail the launch instead of hanging
            __nanosleep(40);
        }
    }
}

template <bool MC, int WT>      // WT: compile-time world size (0 = run-time) so that the W peer loads are all in flight
__global__ void __launch_bounds__(256)
epoch_tail_peer_kernel(PeerPtrs params, PeerPtrs grads, PeerPtrs flags, const float4 *__restrict__ mc_params,
                       const float4 *__restrict__ mc_grads, float *__restrict__ m, float *__restrict__ v, int64_t lo4,
                       int64_t hi4, int world_rt, int rank, uint32_t seq_arg, double lr, double beta1d, double beta2d,
                       float beta2, float omb1, float omb2, float eps, OptState *__restrict__ st,
                       unsigned int *__restrict__ done_counter, uint32_t *__restrict__ seq_word) {
    const int world = WT ? WT : world_rt;
    // barrier sequence number: given by the host, or (seq_arg == 0) kept in device memory and advanced by this kernel,
    // so that the launch has no per-epoch argument and an epoch can be replayed from a CUDA graph
    const uint32_t seq = seq_arg ? seq_arg : *seq_word + 1u;
    __shared__ float s_step[2];
    if (threadIdx.x == 0) {
        const uint32_t step = st->step + 1;                       // the step this launch applies
        s_step[0] = (float)(lr / (1.0 - pow(beta1d, (double)step)));
        s_step[1] = (float)sqrt(1.0 - pow(beta2d, (double)step));
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        __threadfence_system();
        flags_arrive(flags, world, rank, 0, seq);
    }
    if (threadIdx.x < 32) flags_wait(flags, world, rank, 0, seq);
    __syncthreads();
    const float neg_step = -s_step[0], bc2_sqrt = s_step[1];
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {       // adam_kernel's arithmetic (layout_sgd.cu)
        mm = __fadd_rn(mm, __fmul_rn(omb1, __fsub_rn(gg, mm)));
        vv = __fadd_rn(__fmul_rn(vv, beta2), __fmul_rn(__fmul_rn(omb2, gg), gg));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        pp = __fadd_rn(pp, __fdiv_rn(__fmul_rn(neg_step, mm), denom));
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
        float4 gg;
        if (MC) {
            gg = multimem_ld_reduce_add(mc_grads + i);
        } else {
            float4 t[WT ? WT : 1];
            if (WT) {
#pragma unroll
                for (int w = 0; w < (WT ? WT : 1); ++w) t[w] = ld_peer(reinterpret_cast<const float4 *>(grads.p[w]) + i);
                gg = t[0];
#pragma unroll
                for (int w = 1; w < (WT ? WT : 1); ++w) {
                    gg.x = __fadd_rn(gg.x, t[w].x); gg.y = __fadd_rn(gg.y, t[w].y);
                    gg.z = __fadd_rn(gg.z, t[w].z); gg.w = __fadd_rn(gg.w, t[w].w);
                }
            } else {
                gg = ld_peer(reinterpret_cast<const float4 *>(grads.p[0]) + i);
                for (int w = 1; w < world; ++w) {                           // fixed order: the same sum on every rank
                    const float4 u = ld_peer(reinterpret_cast<const float4 *>(grads.p[w]) + i);
                    gg.x = __fadd_rn(gg.x, u.x); gg.y = __fadd_rn(gg.y, u.y); gg.z = __fadd_rn(gg.z, u.z); gg.w = __fadd_rn(gg.w, u.w);
                }
            }
        }
        float4 pp = reinterpret_cast<const float4 *>(params.p[rank])[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        if (MC) {
            multimem_st(const_cast<float4 *>(mc_params) + i, pp);
            multimem_st(const_cast<float4 *>(mc_grads) + i, zero);
        } else {
            for (int w = 0; w < world; ++w) {
                st_peer(reinterpret_cast<float4 *>(params.p[w]) + i, pp);
                st_peer(reinterpret_cast<float4 *>(grads.p[w]) + i, zero);
            }
        }
    }
    // grid-wide completion: the last block to arrive signals the peers
    __syncthreads();                     // the CTA's stores happen-before thread 0's fence through the barrier (cumulativity)
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence_system();
        flags_arrive(flags, world, rank, 1, seq);
        if (threadIdx.x == 0) *done_counter = 0u;                         // ready for the next epoch's launch
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) flags_wait(flags, world, rank, 1, seq);
        __syncthreads();
        if (threadIdx.x == 0) {
            // every block of THIS rank has read the old state before its slot-1 flag could be sent
            const uint32_t step = st->step + 1;
            st->step = step;
            st->epoch = st->epoch + 1;
            st->step_size = s_step[0];
            st->bc2_sqrt = s_step[1];
            if (!seq_arg) *seq_word = seq;
        }
    }
}

// ------------------------------------------------------------------ push form of the epoch tail (small tables)
// The pull form above is a chain of NVLink round trips (flag -> remote load / in-switch reduce -> store -> flag:
// measured 67 us per epoch on 12 MB tables at 2 AND at 8 GPUs, i.e. latency, not volume).  Here nothing is ever loaded
// from a peer -- remote STORES are posted, a rank only waits for flags:
//   A. every rank PUSHES its partial gradient to the owners' inboxes (shard s of my buffer -> slot `rank` of rank s's
//      inbox) and zeroes its own buffer as it goes (local); the last block to finish fences and raises flag slot 0;
//   B. every rank waits for the W flags, sums the W slots of its inbox (LOCAL loads, fixed order), applies Adam once
//      and pushes the new parameters into all W replicas; the last block fences and raises flag slot 1; block 0 waits
//      for the W slot-1 flags and advances the optimiser state.
// Two one-way hops instead of five.  Per rank and epoch 2 x (W-1)/W of the buffer leaves over NVLink (the pull form
// with multimem moves half of that), which is why this form is used for small tables only (where latency is the cost).
// The grid is one block per SM: blocks spin on peers' flags, so all of them have to become resident.
template <int WT>
__global__ void __launch_bounds__(512)
epoch_tail_push_kernel(PeerPtrs params, PeerPtrs inbox, PeerPtrs flags, float *__restrict__ grad, float *__restrict__ m,
                       float *__restrict__ v, int64_t n4, int64_t slot4, int world_rt, int rank, double lr, double beta1d,
                       double beta2d, float beta2, float omb1, float omb2, float eps, OptState *__restrict__ st,
                       unsigned int *__restrict__ done_counter, uint32_t *__restrict__ seq_word, int skip_mask) {
    // skip_mask (measurement only, option tail_skip_mask; results are then wrong): 1 = no gradient push, 2 = no shard step
    const int world = WT ? WT : world_rt;
    const uint32_t seq = *seq_word + 1u;
    __shared__ float s_step[2];
    __shared__ unsigned int s_last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- A: push my partial gradient to the owners, clear it.  Shard by shard (no per-element owner arithmetic), four
    // independent local loads in flight per thread before their remote stores
    float4 *g4 = reinterpret_cast<float4 *>(grad);
    for (int s = 0; s < ((skip_mask & 1) ? 0 : world); ++s) {
        const int64_t lo_s = n4 * s / world, hi_s = n4 * (s + 1) / world;
        float4 *dst = reinterpret_cast<float4 *>(inbox.p[s]) + (int64_t)rank * slot4 - lo_s;
        int64_t i = lo_s + tid;
        for (; i + 3 * stride < hi_s; i += 4 * stride) {
            const float4 a = g4[i], b = g4[i + stride], c = g4[i + 2 * stride], d = g4[i + 3 * stride];
            st_peer(dst + i, a); st_peer(dst + i + stride, b); st_peer(dst + i + 2 * stride, c); st_peer(dst + i + 3 * stride, d);
            g4[i] = zero; g4[i + stride] = zero; g4[i + 2 * stride] = zero; g4[i + 3 * stride] = zero;
        }
        for (; i < hi_s; i += stride) {
            const float4 gv = g4[i];
            st_peer(dst + i, gv);
            g4[i] = zero;
        }
    }
    __syncthreads();                 // the CTA's stores happen-before thread 0's fence through the barrier (cumulativity)
    if (threadIdx.x == 0) __threadfence_system();
    if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence_system();
        flags_arrive(flags, world, rank, 0, seq);
    }
    // ---- B: my shard is complete in my inbox once every rank has raised slot 0
    if (threadIdx.x == 64) {             // (another warp than the one polling: the double-precision pow overlaps the wait)
        const uint32_t step = st->step + 1;
        s_step[0] = (float)(lr / (1.0 - pow(beta1d, (double)step)));
        s_step[1] = (float)sqrt(1.0 - pow(beta2d, (double)step));
    }
    if (threadIdx.x < 32) flags_wait(flags, world, rank, 0, seq);
    __syncthreads();
    const float neg_step = -s_step[0], bc2_sqrt = s_step[1];
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {       // adam_kernel's arithmetic (layout_sgd.cu)
        mm = __fadd_rn(mm, __fmul_rn(omb1, __fsub_rn(gg, mm)));
        vv = __fadd_rn(__fmul_rn(vv, beta2), __fmul_rn(__fmul_rn(omb2, gg), gg));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        pp = __fadd_rn(pp, __fdiv_rn(__fmul_rn(neg_step, mm), denom));
    };
    const int64_t lo4 = n4 * rank / world, hi4 = (skip_mask & 2) ? n4 * rank / world : n4 * (rank + 1) / world;
    const float4 *in4 = reinterpret_cast<const float4 *>(inbox.p[rank]);
    for (int64_t i = lo4 + tid; i < hi4; i += stride) {
        float4 gg = ld_peer(in4 + (i - lo4));                                   // slot 0; L1 may hold last epoch's line
        for (int w = 1; w < world; ++w) {                                       // fixed order: a deterministic sum
            const float4 u = ld_peer(in4 + (int64_t)w * slot4 + (i - lo4));
            gg.x = __fadd_rn(gg.x, u.x); gg.y = __fadd_rn(gg.y, u.y); gg.z = __fadd_rn(gg.z, u.z); gg.w = __fadd_rn(gg.w, u.w);
        }
        float4 pp = reinterpret_cast<const float4 *>(params.p[rank])[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        for (int w = 0; w < world; ++w) st_peer(reinterpret_cast<float4 *>(params.p[w]) + i, pp);
    }
    __syncthreads();                 // the CTA's stores happen-before thread 0's fence through the barrier (cumulativity)
    if (threadIdx.x == 0) __threadfence_system();
    if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == 2u * gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence_system();
        flags_arrive(flags, world, rank, 1, seq);
        if (threadIdx.x == 0) *done_counter = 0u;
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) flags_wait(flags, world, rank, 1, seq);
        __syncthreads();
        if (threadIdx.x == 0) {
            st->step = st->step + 1;
            st->epoch = st->epoch + 1;
            st->step_size = s_step[0];
            st->bc2_sqrt = s_step[1];
            *seq_word = seq;
        }
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

extern "C" int mmu_epoch_tail_peer(const uint64_t *peer_params, const uint64_t *peer_grads, const uint64_t *peer_flags,
                                   uint64_t mc_params, uint64_t mc_grads, float *m, float *v, int64_t n, int world, int rank,
                                   uint32_t seq, double lr, double beta1, double beta2, double eps, uint32_t *state,
                                   uint32_t *done_counter, mmu_stream_t stream) {
    // done_counter[0]: grid-completion counter; done_counter[1]: device-resident barrier sequence (used when seq == 0)
    using namespace mmu;
    MMU_CHECK_ARG(peer_params && peer_grads && peer_flags && m && v && state && done_counter, "mmu_epoch_tail_peer: null pointer");
    MMU_CHECK_ARG(world >= 1 && world <= MMU_PEER_MAX && rank >= 0 && rank < world, "mmu_epoch_tail_peer: bad world/rank");
    MMU_CHECK_ARG(n >= 0 && (n & 3) == 0, "mmu_epoch_tail_peer: n must be a multiple of 4");
    MMU_CHECK_ARG(((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) | mc_params | mc_grads) & 15) == 0,
                  "mmu_epoch_tail_peer: m / v / multicast pointers must be 16-byte aligned");
    MMU_CHECK_ARG((mc_params == 0) == (mc_grads == 0), "mmu_epoch_tail_peer: give both multicast pointers or none");
    PeerPtrs pp, gg, ff;
    MMU_CHECK_ARG(fill_ptrs(pp, peer_params, world) == 0 && fill_ptrs(gg, peer_grads, world) == 0 &&
                  fill_ptrs(ff, peer_flags, world) == 0, "mmu_epoch_tail_peer: null or unaligned peer pointer");
    const int64_t n4 = n >> 2;
    const int64_t lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;     // this rank's shard
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int64_t want = (hi4 - lo4 + 255) / 256;
    if (want < 1) want = 1;
    // peer accesses have microseconds of latency: as many threads in flight as the shard has 16-byte elements, up to 8
    // blocks per SM (block 0 spins on the others' completion counter at the end; none of them waits for block 0)
    long long per_sm = option(OPT_TAIL_BLOCKS_PER_SM);
    if (per_sm < 1) per_sm = 1;
    const unsigned blocks = (unsigned)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    const float4 *mcp = reinterpret_cast<const float4 *>(mc_params), *mcg = reinterpret_cast<const float4 *>(mc_grads);
    OptState *os = reinterpret_cast<OptState *>(state);
    cudaStream_t st = as_stream(stream);
#define MMU_TAIL(MCV, WTV)                                                                                          \
    epoch_tail_peer_kernel<MCV, WTV><<<blocks, 256, 0, st>>>(pp, gg, ff, mcp, mcg, m, v, lo4, hi4, world, rank, seq, lr, beta1, \
                                                             beta2, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),  \
                                                             (float)eps, os, done_counter, done_counter + 1)
    if (mc_params) MMU_TAIL(true, 0);
    else if (world == 2) MMU_TAIL(false, 2);
    else if (world == 4) MMU_TAIL(false, 4);
    else if (world == 8) MMU_TAIL(false, 8);
    else MMU_TAIL(false, 0);
#undef MMU_TAIL
    note_kernel(SITE_EPOCH_TAIL, "epoch_tail_peer_kernel<%s,W=%d>", mc_params ? "multimem" : "peer-loads", world);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_epoch_tail_push(const uint64_t *peer_params, const uint64_t *peer_inbox, const uint64_t *peer_flags,
                                   float *grad, float *m, float *v, int64_t n, int64_t inbox_slot_floats, int world, int rank,
                                   double lr, double beta1, double beta2, double eps, uint32_t *state, uint32_t *done_counter,
                                   mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(peer_params && peer_inbox && peer_flags && grad && m && v && state && done_counter,
                  "mmu_epoch_tail_push: null pointer");
    MMU_CHECK_ARG(world >= 1 && world <= MMU_PEER_MAX && rank >= 0 && rank < world, "mmu_epoch_tail_push: bad world/rank");
    MMU_CHECK_ARG(n >= 0 && (n & 3) == 0 && (inbox_slot_floats & 3) == 0, "mmu_epoch_tail_push: sizes must be multiples of 4");
    const int64_t n4 = n >> 2, slot4 = inbox_slot_floats >> 2;
    MMU_CHECK_ARG(slot4 >= (n4 + world - 1) / world, "mmu_epoch_tail_push: inbox slot smaller than a shard");
    MMU_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                  "mmu_epoch_tail_push: grad / m / v must be 16-byte aligned");
    PeerPtrs pp, ib, ff;
    MMU_CHECK_ARG(fill_ptrs(pp, peer_params, world) == 0 && fill_ptrs(ib, peer_inbox, world) == 0 &&
                  fill_ptrs(ff, peer_flags, world) == 0, "mmu_epoch_tail_push: null or unaligned peer pointer");
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const unsigned blocks = (unsigned)sms;                 // every block spins on peers' flags: all must become resident
    OptState *os = reinterpret_cast<OptState *>(state);
    cudaStream_t st = as_stream(stream);
#define MMU_PUSH(WTV)                                                                                                       \
    epoch_tail_push_kernel<WTV><<<blocks, 512, 0, st>>>(pp, ib, ff, grad, m, v, n4, slot4, world, rank, lr, beta1, beta2,   \
                                                        (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, \
                                                        os, done_counter, done_counter + 1, (int)option(OPT_TAIL_SKIP_MASK))
    if (world == 2) MMU_PUSH(2);
    else if (world == 4) MMU_PUSH(4);
    else if (world == 8) MMU_PUSH(8);
    else MMU_PUSH(0);
#undef MMU_PUSH
    note_kernel(SITE_EPOCH_TAIL, "epoch_tail_push_kernel<W=%d>", world);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
