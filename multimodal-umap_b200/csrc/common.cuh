// common.cuh -- shared helpers for the sm_100a kernels behind include/mmumap.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmumap.h"

namespace mmu {

void set_error(const char *fmt, ...);

#define MMU_CHECK_ARG(cond, ...)              \
    do {                                      \
        if (!(cond)) {                        \
            mmu::set_error(__VA_ARGS__);      \
            return MMU_ERR_ARG;               \
        }                                     \
    } while (0)

#define MMU_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            mmu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                           __FILE__, __LINE__);                                          \
            return MMU_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

void count_launch(int n);   // n of OUR kernels were launched (mmu_launch_count)

#define MMU_LAUNCH_CHECK() MMU_LAUNCH_CHECK_N(1)
#define MMU_LAUNCH_CHECK_N(n_launched)                                                   \
    do {                                                                                 \
        mmu::count_launch(n_launched);                                                   \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            mmu::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                          \
            return MMU_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

int sm_count();   // cached; 0 when no device

// launch sites with per-device one-time setup and/or a reported kernel variant (mmu_last_kernel)
enum : int { SITE_KNN_CANDIDATES = 0, SITE_KNN_EXACT, SITE_EIGH_SMALL, SITE_EDGE_FORCES, SITE_EPOCH_TAIL, SITE_BLOCK_OPS,
             MMU_SITE_COUNT };
bool first_use_on_device(int site);                  // true exactly once per (site, current device)
void note_kernel(int site, const char *fmt, ...);    // name of the kernel variant the site launched last

// A/B switches: read from the environment once at load, changed with mmu_set_option (never getenv per launch)
enum : int { OPT_FORCE_STAGED = 0, OPT_KNN_CTA_PAIRS, OPT_KNN_WINDOW_MB, OPT_SGD_WINDOW_MB, OPT_KNN_FOLD_NORMS, OPT_TAIL_BLOCKS_PER_SM, OPT_TAIL_SKIP_MASK, OPT_COUNT };
long long option(int id);

static inline cudaStream_t as_stream(mmu_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------- optimiser state words
struct OptState {
    uint32_t epoch;      // RNG counter (advances once per epoch)
    uint32_t step;       // Adam step
    float step_size;     // lr / (1 - beta1^step)
    float bc2_sqrt;      // sqrt(1 - beta2^step)
    uint32_t reserved[4];
};
static_assert(sizeof(OptState) == MMU_OPT_STATE_WORDS * 4, "OptState layout");

// ---------------------------------------------------------------- Philox4x32-10
struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// uniform in [0,1) with 24 random bits, the resolution torch.rand gives for float32
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// unbiased-enough integer in [0, n): multiply-shift on 32 random bits
__device__ __forceinline__ uint32_t urange(uint32_t x, uint32_t n) {
    return (uint32_t)(((uint64_t)x * (uint64_t)n) >> 32);
}

// RNG stream ids (4th counter word)
enum : uint32_t { STREAM_KEEP = 0, STREAM_NEG = 1, STREAM_INFONCE = 16 };

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint64_t dist_key(float d, int32_t idx) {
    return ((uint64_t)__float_as_uint(d) << 32) | (uint32_t)idx;   // d >= 0
}
__device__ __forceinline__ float key_dist(uint64_t key) { return __uint_as_float((uint32_t)(key >> 32)); }
__device__ __forceinline__ int32_t key_idx(uint64_t key) { return (int32_t)(uint32_t)(key & 0xffffffffu); }
#define MMU_KEY_EMPTY 0x7f800000ffffffffull   // (+inf, idx = -1)

__device__ __forceinline__ void red_add_f32(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mmu
