// knn_tc.cu -- K1/K2: exact kNN through the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the reference's candidate search and per-row top-k
// (/root/reference/impl/model.py:81-195; distance :109,:163; selection :181-193; self exclusion
// :88,:166).  Three launches, all on the caller's stream:
//
//   prep       X (fp32) -> centred, power-of-two scaled fp16 copy X16 [rows_pad x D_pad] plus the
//              fp32 squared norms of the centred/scaled rows (translation and scaling leave the
//              ranking unchanged; centring minimises |x||y|, the quantity the error bound uses).
//   candidates dense contraction S~[q][j] = |Y_j|^2 - 2 X_q.Y_j on tcgen05.mma (kind::f16, fp32
//              accumulators in TMEM, operands staged by TMA with the 128-byte swizzle, a 3/4-stage
//              mbarrier pipeline, two TMEM accumulator buffers so that the epilogue of tile i
//              overlaps the MMAs of tile i+1).  The epilogue is a fused per-row top-K' selection:
//              one thread per TMEM lane (= query row) keeps its K' best approximate scores in a
//              max-heap in shared memory (slot-major, so every access is bank-conflict free) and
//              filters each accumulator value with ONE compare against the heap root.
//   rescore    per query row (one warp): certify, from the approximate scores alone, that no point
//              outside the candidate list can be among the k nearest -- every non-candidate has
//              S~ >= tau (the list's largest score) and |S~ - S| <= eps, a rigorous bound on the fp16
//              rounding + accumulation error -- then evaluate the reference's own fp32 distance
//              expression in the canonical order of oracle/knn_oracle.c for the few candidates
//              that can still reach the top k and rank them by (distance, index).  Rows that
//              cannot be certified (ties, degenerate data) are listed for the exhaustive fp32
//              kernel (mmu_knn_exact_f32), so the result is bit-exact by construction.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mmu {

constexpr int TC_BM = 128;        // query rows per CTA (= TMEM lanes)
constexpr int TC_BN = 256;        // db rows per tile (= UMMA N, TMEM columns per accumulator)
constexpr int TC_BK = 64;         // fp16 elements per k-block: 128 bytes = one swizzle span
constexpr int TC_STAGES = 3;
constexpr int TC_EPI_GROUPS = 4;  // epilogue warps per TMEM lane quarter; group g scans a 256/GROUPS-column slice of a tile
constexpr int TC_THREADS = 128 + 128 * TC_EPI_GROUPS;   // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue
constexpr int TC_KP = 64;         // candidates kept per row and db split (TC_EPI_GROUPS lists of TC_KPG)
constexpr int TC_KPG = TC_KP / TC_EPI_GROUPS;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_CH = 32;         // accumulator columns scanned per epilogue step
constexpr int TC_LIST_BYTES = TC_KP * TC_BM * 8 + (TC_KP / 8) * TC_BM * 4;   // slots + group maxima
constexpr int TC_SMEM_BYTES = 1024 + TC_STAGES * TC_STAGE_BYTES + TC_LIST_BYTES + 256;
static_assert(TC_SMEM_BYTES <= 232448, "shared memory budget");
// NCTA = 2: a CTA pair (one TPC) runs ONE tcgen05.mma.cta_group::2 of M = 256 per step.  Each CTA stages its own
// 128 query rows and HALF of the database tile (128 of the 256 rows), so the operand bytes an SM moves through
// its shared memory per MMA drop from 24 KB (12 KB TMA fill + 12 KB operand read) to 16 KB -- at cta_group::1
// that traffic, not the tensor pipe, is what bounds the kernel (ncu: tensor pipe 68 % = 128/192 B/clk).
template <int NCTA>
struct TcCfg {
    static constexpr int B_ROWS = TC_BN / NCTA;
    static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int STAGES = NCTA == 2 ? 4 : TC_STAGES;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + TC_LIST_BYTES + 256;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area");
};
constexpr int RS_PMAX = 128;      // most candidates rescored per row
#define F_INF __int_as_float(0x7f800000)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// both CTAs of a pair load into their own shared memory; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 8 consecutive fp32 columns: thread i of the warp receives lane (base+i)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// v[i] for a run-time i in [0,32): a five-level select tree (31 SEL), registers cannot be indexed
__device__ __forceinline__ float pick32(const float (&v)[32], int i) {
    float a[16], b[8], c[4], d[2];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (i & 16) ? v[j + 16] : v[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (i & 8) ? a[j + 8] : a[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (i & 4) ? b[j + 4] : b[j];
#pragma unroll
    for (int j = 0; j < 2; ++j) d[j] = (i & 2) ? c[j + 2] : c[j];
    return (i & 1) ? d[1] : d[0];
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: 8-row groups of 1024 bytes (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D fp32 (bits 4-5 = 1), A/B fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
// (the instruction descriptor is built where the MMAs are issued: M = 128 per CTA, 256 for a CTA pair)

// ------------------------------------------------------------------ prep
// column sums (deterministic two-level reduction) and the largest |x|
__global__ void __launch_bounds__(256)
tc_colsum_kernel(const float *__restrict__ x, int64_t n, int dim, int rows_per_block, float *__restrict__ partial,
                 uint32_t *__restrict__ maxabs_bits) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    float mx = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float s = 0.f;
        for (int64_t r = r0; r < r1; ++r) {
            float v = x[r * dim + c];
            s += v;
            mx = fmaxf(mx, fabsf(v));
        }
        partial[(int64_t)blockIdx.x * dim + c] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxabs_bits, __float_as_uint(mx));
}

// prm[0] = scale (power of two), prm[1] = largest |Y|^2 (bits, atomicMax), prm[2] = largest |x| bits
__global__ void tc_finish_stats_kernel(const float *__restrict__ partial, int n_blocks, int dim, int64_t n,
                                       float *__restrict__ mean, uint32_t *__restrict__ prm, int center) {
    __shared__ float s_mx[32];
    float mu_max = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < n_blocks; ++b) s += partial[(int64_t)b * dim + c];
        float mu = center ? s / (float)n : 0.f;
        mean[c] = mu;
        mu_max = fmaxf(mu_max, fabsf(mu));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mu_max = fmaxf(mu_max, __shfl_xor_sync(0xffffffffu, mu_max, o));
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mu_max;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_mx[w]);
        float bound = __uint_as_float(prm[2]) + m;            // >= max |x - mean|
        int e = 0;
        if (bound > 0.f && bound < F_INF) {
            (void)frexpf(bound, &e);                          // bound = f * 2^e, f in [0.5, 1)
            e = 14 - e;                                       // bound * 2^e in [2^13, 2^14): fp16-safe
        }
        e = max(-100, min(100, e));
        reinterpret_cast<float *>(prm)[0] = ldexpf(1.0f, e);
        prm[1] = 0u;
    }
}

// one warp per row: fp16 copy of (x - mean) * scale, zero padded to dim_pad; squared norm of the
// unrounded values; rows >= n are zero with norm = pad_norm
// split == 0: one fp16 value per component.  split == 1 / 2: error-compensated operands for the
// second certification level -- every value v is written as hi = fp16(v) and lo = fp16(v - hi), the row
// is three dim_pad-wide sections [hi | hi | lo] (split 1, query side) or [hi | lo | hi] (split 2, database
// side), so that the unchanged contraction kernel evaluates hi.hi + hi.lo + lo.hi (error ~2^-21 |x||y|
// instead of 2^-10 |x||y|) with a three times longer k loop.
__global__ void __launch_bounds__(256)
tc_convert_kernel(const float *__restrict__ x, int64_t n, int64_t n_pad, int dim, int dim_pad, int split,
                  const float *__restrict__ mean, const uint32_t *__restrict__ prm, __half *__restrict__ x16,
                  float *__restrict__ norm2, float pad_norm, uint32_t *__restrict__ max_norm_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_pad) return;
    const float scale = __uint_as_float(prm[0]);
    const int width = split ? 3 * dim_pad : dim_pad;
    __half *dst = x16 + row * width;
    float acc = 0.f;
    if (row < n) {
        const float *src = x + row * dim;
        for (int c = lane * 2; c < dim_pad; c += 64) {
            float v0 = (c < dim) ? (src[c] - mean[c]) * scale : 0.f;
            float v1 = (c + 1 < dim) ? (src[c + 1] - mean[c + 1]) * scale : 0.f;
            acc = fmaf(v0, v0, acc);
            acc = fmaf(v1, v1, acc);
            const __half2 hi = __floats2half2_rn(v0, v1);
            *reinterpret_cast<__half2 *>(dst + c) = hi;
            if (split) {
                const float2 hf = __half22float2(hi);
                const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                *reinterpret_cast<__half2 *>(dst + dim_pad + c) = split == 1 ? hi : lo;
                *reinterpret_cast<__half2 *>(dst + 2 * dim_pad + c) = split == 1 ? lo : hi;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            norm2[row] = acc;
            if (max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(acc));
        }
    } else {
        for (int c = lane * 2; c < width; c += 64) *reinterpret_cast<__half2 *>(dst + c) = __floats2half2_rn(0.f, 0.f);
        if (lane == 0) norm2[row] = pad_norm;
    }
}

// ------------------------------------------------------------------ candidates (tcgen05)
// Per-row candidate list: TC_KP slots in 8-slot groups, slot-major in shared memory (address =
// slot * 128 + row, so the 32 rows of a warp always hit 32 different banks), with the maximum of
// every group cached in `gmax`.  The owner thread keeps the list maximum `tau` and its group
// `gstar` in registers.  Replacing the maximum costs two rounds of independent loads (the 8 slots
// of one group, then the 8 group maxima) instead of a pointer-chasing heap walk.
constexpr int TC_GROUPS = TC_KPG / 8;                // 8-slot groups per list
__device__ __forceinline__ void list_replace_max(float *__restrict__ sc, int32_t *__restrict__ id,
                                                 float *__restrict__ gmax, float s, int32_t j, float &tau, int &gstar) {
    float e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = sc[(gstar * 8 + u) * TC_BM];
    int pos = 0;
    float m = e[0];
#pragma unroll
    for (int u = 1; u < 8; ++u)
        if (e[u] > m) { m = e[u]; pos = u; }
    float ng = s;
#pragma unroll
    for (int u = 0; u < 8; ++u) ng = fmaxf(ng, (u == pos) ? s : e[u]);
    sc[(gstar * 8 + pos) * TC_BM] = s;
    id[(gstar * 8 + pos) * TC_BM] = j;
    gmax[gstar * TC_BM] = ng;
    float best = ng;
    int bg = gstar;
#pragma unroll
    for (int g = 0; g < TC_GROUPS; ++g) {
        float gm = gmax[g * TC_BM];
        if (g != gstar && gm > best) { best = gm; bg = g; }
    }
    tau = best;
    gstar = bg;
}

struct TcParams {
    const float *ynorm;      // [n_db_pad]  |Y_j|^2 (+inf on padding rows)
    int n_kblocks;           // dim_pad / 64
    int n_tiles;             // n_db_pad / 256
    int tiles_per_split;
    int n_splits;
    int split_major;         // grid = (n_splits, n_qblocks): co-resident CTAs share query blocks instead of db tiles
    int n_qblocks;           // real query blocks (a pair launch rounds the grid up to an even number)
    int window_begin;        // this launch covers tiles [window_begin, window_begin + window_tiles) of every split
    int window_tiles;
    int resume;              // continue the lists an earlier window left in cand_idx / cand_score
    int32_t *cand_idx;       // [n_qblocks][n_splits][TC_KP][128]
    float *cand_score;       // same layout
    float *tau;              // [n_qblocks][n_splits][TC_EPI_GROUPS][128]
    // pruned search (single-CTA form, one split): every query block walks its OWN set of database tiles --
    // a contiguous range [qb_tile_begin[qb], qb_tile_end[qb]) or the list tile_list[tile_ptr[qb] .. tile_ptr[qb+1])
    const int32_t *qb_tile_begin;
    const int32_t *qb_tile_end;
    const int32_t *tile_ptr;
    const int32_t *tile_list;
};

// NCTA = 1: one CTA per query block.  NCTA = 2: clusters of two CTAs along the query-block axis of the grid; the pair
// shares every database tile (rank r stages rows [128 r, 128 r + 128) of it) and the even-rank CTA issues the MMAs
// for both.  tm_db's box is TC_BN / NCTA rows.
template <int NCTA>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_candidates_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                         const TcParams p) {
    using Cfg = TcCfg<NCTA>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    // 1024-byte alignment (128-byte swizzle atoms) comes from the declaration, so that every pointer
    // below keeps its shared-memory provenance and compiles to LDS/STS
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *stage_base = smem;
    float *list_sc = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
    int32_t *list_id = reinterpret_cast<int32_t *>(list_sc + TC_KP * TC_BM);
    float *list_gmax = reinterpret_cast<float *>(list_id + TC_KP * TC_BM);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + TC_LIST_BYTES);
    uint64_t *full_bar = bars;                    // [STAGES]  TMA -> MMA        (pair: the leader's counts both CTAs' bytes)
    uint64_t *empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA        (pair: commit multicast to both CTAs)
    uint64_t *acc_full = bars + 2 * STAGES;       // [2]       MMA -> epilogue   (pair: commit multicast to both CTAs)
    uint64_t *acc_empty = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA   (pair: both CTAs' warps arrive on the leader's)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rasterisation: x is the fast index of the block scheduler.  Query-block-major (default) keeps ~148
    // different query blocks and ONE database region in flight -- right while 148 query tiles fit in L2
    // beside it; for long rows (148 x 128 x D x 2 B of query tiles alone overflow the L2) split-major keeps
    // few query blocks and all database regions in flight instead.
    // A pair is always two x-neighbours of the grid: split-major pair launches use grid = (2 n_splits, n_qblocks / 2)
    // with x = 2 split + rank.
    const int qblock = !p.split_major ? blockIdx.x : NCTA == 2 ? 2 * blockIdx.y + (blockIdx.x & 1) : blockIdx.y;
    const int split = !p.split_major ? blockIdx.y : NCTA == 2 ? (blockIdx.x >> 1) : blockIdx.x;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) __trap();   // swizzle atoms need 1024-byte alignment
    int t0 = split * p.tiles_per_split + p.window_begin;
    int t1 = min(min(p.n_tiles, (split + 1) * p.tiles_per_split), t0 + p.window_tiles);
    const int32_t *my_list = nullptr;
    int tstep = 1;
    // per-block tile sets (pruned search): the splits of a query block take its tiles in turn (split s: tiles s, s + S, ...),
    // each with its own lists -- a row's neighbours are spread over S x 4 lists by tile parity and column slice
    if (p.tile_list) {
        const int a = qblock < p.n_qblocks ? p.tile_ptr[qblock] : 0, b = qblock < p.n_qblocks ? p.tile_ptr[qblock + 1] : 0;
        my_list = p.tile_list + a;
        t0 = split; t1 = b - a; tstep = p.n_splits;
    } else if (p.qb_tile_begin) {
        t0 = (qblock < p.n_qblocks ? p.qb_tile_begin[qblock] : 0) + split;
        t1 = qblock < p.n_qblocks ? p.qb_tile_end[qblock] : 0;
        tstep = p.n_splits;
    }
    const int n_my_tiles = max(0, (t1 - t0 + tstep - 1) / tstep);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_db)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], NCTA * 4 * TC_EPI_GROUPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (NCTA == 2) {      // the same warp of both CTAs allocates the same columns in both SMs
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * TC_BN) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * TC_BN) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();    // the peer's barriers are initialised before anything is signalled across
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int ti = t0; ti < t1; ti += tstep) {
                const int t = my_list ? my_list[ti] : ti;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *a_dst = stage_base + stage * STAGE_BYTES;
                    uint8_t *b_dst = a_dst + TC_A_BYTES;
                    if (NCTA == 2) {
                        // the leader expects the bytes of both CTAs; the peer's bytes may land first (the
                        // transaction count goes negative for a moment, the phase cannot complete before the
                        // leader's arrive)
                        const uint32_t leader_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        tma_load_2d_pair(a_dst, &tm_q, leader_bar, kb * TC_BK, qblock * TC_BM);
                        tma_load_2d_pair(b_dst, &tm_db, leader_bar, kb * TC_BK, t * TC_BN + (int)cta_rank * Cfg::B_ROWS);
                    } else {
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(a_dst, &tm_q, &full_bar[stage], kb * TC_BK, qblock * TC_BM);
                        tma_load_2d(b_dst, &tm_db, &full_bar[stage], kb * TC_BK, t * TC_BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((NCTA * TC_BM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < n_my_tiles; ++it) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_BN;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base + stage * STAGE_BYTES);
                    const uint64_t adesc = umma_smem_desc(a_addr);
                    const uint64_t bdesc = umma_smem_desc(a_addr + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {    // +32 bytes (2 x 16 B) per UMMA_K = 16 halfs
                        if (NCTA == 2) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
                        else umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
                    }
                    // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                    if (NCTA == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    if (kb == p.n_kblocks - 1) {
                        if (NCTA == 2) umma_commit_pair(&acc_full[acc]); else umma_commit(&acc_full[acc]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: fused per-row top-K' =====
        // Two warps per TMEM lane quarter: group g of a quarter scans columns [128g, 128g+128) of every tile
        // into its own TC_KPG-slot list (slots [g*TC_KPG, (g+1)*TC_KPG) of the row).  A row's candidates are
        // the union of its lists and tau = min over them is still a bound for every non-candidate.
        const int quarter = warp & 3;                  // TMEM lanes 32*quarter .. +31
        const int grp = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;
        float *sc = list_sc + grp * TC_KPG * TC_BM + row;
        int32_t *id = list_id + grp * TC_KPG * TC_BM + row;
        float *gmax = list_gmax + grp * TC_GROUPS * TC_BM + row;
        const int64_t list_base = (((int64_t)qblock * p.n_splits + split) * TC_KP + grp * TC_KPG) * TC_BM;
        if (p.resume && qblock < p.n_qblocks) {
            for (int s = 0; s < TC_KPG; ++s) {
                sc[s * TC_BM] = p.cand_score[list_base + (int64_t)s * TC_BM + row];
                id[s * TC_BM] = p.cand_idx[list_base + (int64_t)s * TC_BM + row];
            }
        } else {
            for (int s = 0; s < TC_KPG; ++s) { sc[s * TC_BM] = F_INF; id[s * TC_BM] = -1; }
        }
        float tau = -F_INF;
        int gstar = 0;
        for (int g = 0; g < TC_GROUPS; ++g) {
            float gm = sc[g * 8 * TC_BM];
            for (int u = 1; u < 8; ++u) gm = fmaxf(gm, sc[(g * 8 + u) * TC_BM]);
            gmax[g * TC_BM] = gm;
            if (gm > tau) { tau = gm; gstar = g; }
        }
        constexpr int COLS = TC_BN / TC_EPI_GROUPS;
        for (int it = 0; it < n_my_tiles; ++it) {
            const int acc = it & 1;
            const int n0 = (my_list ? my_list[t0 + it * tstep] : t0 + it * tstep) * TC_BN + grp * COLS;
            mbar_wait(&acc_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TC_BN + grp * COLS;
#pragma unroll 1
            for (int c = 0; c < COLS / TC_CH; ++c) {
                uint32_t v[TC_CH];
                tmem_ld_32x32(taddr + c * TC_CH, v);
                float4 yn[TC_CH / 4];
                const float4 *yp = reinterpret_cast<const float4 *>(p.ynorm + n0 + c * TC_CH);
#pragma unroll
                for (int i = 0; i < TC_CH / 4; ++i) yn[i] = __ldg(yp + i);
                tmem_ld_wait();
                // branch-free scan: scores stay in registers, hits go to a bit mask
                float sv[TC_CH];
                uint32_t mask = 0;
#pragma unroll
                for (int i = 0; i < TC_CH; ++i) {
                    const float y = (i & 3) == 0 ? yn[i >> 2].x : (i & 3) == 1 ? yn[i >> 2].y : (i & 3) == 2 ? yn[i >> 2].z : yn[i >> 2].w;
                    sv[i] = fmaf(-2.0f, __uint_as_float(v[i]), y);
                    mask |= (sv[i] < tau) ? (1u << i) : 0u;
                }
                // drain: each lane inserts its own hits (rare after the first tiles); the score of a hit is
                // fetched from the register array with a select tree
                while (__any_sync(0xffffffffu, mask != 0u)) {
                    const int i = mask ? __ffs(mask) - 1 : 0;
                    const float s = pick32(sv, i);
                    if (mask) {
                        mask &= mask - 1;
                        if (s < tau) list_replace_max(sc, id, gmax, s, n0 + c * TC_CH + i, tau, gstar);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NCTA == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[acc]), 0));
                else mbar_arrive(&acc_empty[acc]);
            }
        }
        // flush: [qblock][split][slot][row] keeps the stores coalesced
        if (qblock < p.n_qblocks) {                    // the odd block out of a pair launch computes padding only
            for (int s = 0; s < TC_KPG; ++s) {
                p.cand_idx[list_base + (int64_t)s * TC_BM + row] = id[s * TC_BM];
                p.cand_score[list_base + (int64_t)s * TC_BM + row] = sc[s * TC_BM];
            }
            p.tau[(((int64_t)qblock * p.n_splits + split) * TC_EPI_GROUPS + grp) * TC_BM + row] = tau;
        }
    }

    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();    // neither CTA leaves (or frees TMEM) while the other still reads its memory
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (NCTA == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TC_BN) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TC_BN) : "memory");
    }
}

// ------------------------------------------------------------------ certify + canonical fp32 rescoring
struct RsParams {
    const float *query;      // fp32 originals
    const float *db;
    int64_t n_query, n_db;
    int dim, k, exclude_self;
    int64_t query_index_base;
    const int32_t *query_gid;   // nullable: db index of query row q (self exclusion for gathered query rows)
    const int32_t *db_gid;      // nullable: index REPORTED (and used for ties / self exclusion) for db row j -- the database was
                                // re-ordered (cluster sorted) and db_gid maps a row of the re-ordered copy to its original index
    const float *xnorm;      // [n_query_pad] |X_q|^2 (centred, scaled)
    const uint32_t *prm;     // prm[1] = largest |Y|^2 bits
    int n_splits;
    const int32_t *cand_idx;
    const float *cand_score;
    const float *tau;
    float c_rel;             // error bound of -2 X.Y relative to |X||Y|
    float c_norm;            // relative error bound of the fp32 norm |Y|^2
    float c_abs;             // absolute error (scaled units) from fp16 underflow of tiny components
    float gamma;             // relative error bound of the canonical fp32 squared distance
    int32_t *out_idx;
    float *out_dist;
    int32_t *stats;          // [0] rows needing the exhaustive kernel, [1] candidates rescored, [2] rows certified
    int32_t *fallback_rows;
};

constexpr int RS_WARPS = 4;
constexpr int RS_TS = 33;    // tile row stride (floats): conflict-free transposed access (scalar path)

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool VEC4>
__global__ void __launch_bounds__(RS_WARPS * 32)
knn_tc_rescore_kernel(const RsParams p) {
    // two tiles per warp: [32 candidates][32 floats], 16-byte chunks XOR-swizzled by the candidate
    // number (VEC4 path, cp.async double buffering) -- or one padded transposed tile (scalar path)
    __shared__ __align__(16) float s_tile[RS_WARPS][2 * 32 * 32 + 64];
    __shared__ int32_t s_p[RS_WARPS][RS_PMAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (q >= p.n_query) return;
    const int qblock = (int)(q / TC_BM), r = (int)(q % TC_BM);
    const int n_cand = p.n_splits * TC_KP;                     // <= 512
    const int64_t qglob = p.query_gid ? (int64_t)p.query_gid[q] : q + p.query_index_base;
    const int64_t base = (int64_t)qblock * p.n_splits * TC_KP * TC_BM;

    // gather this row's candidates: entry e = split * KP + slot
    constexpr int PER_LANE = 16;
    float cs[PER_LANE];
    int32_t ci[PER_LANE];
    float tau = F_INF;
#pragma unroll
    for (int u = 0; u < PER_LANE; ++u) {
        int e = u * 32 + lane;
        cs[u] = F_INF;
        ci[u] = -1;
        if (e < n_cand) {
            int32_t j = p.cand_idx[base + (int64_t)e * TC_BM + r];
            float s = p.cand_score[base + (int64_t)e * TC_BM + r];
            const bool in = j >= 0 && j < p.n_db;
            bool self = p.exclude_self && in && (int64_t)(p.db_gid ? p.db_gid[j] : j) == qglob;
            if (in && !self) { cs[u] = s; ci[u] = j; }
        }
    }
    for (int s = lane; s < p.n_splits * TC_EPI_GROUPS; s += 32)
        tau = fminf(tau, p.tau[((int64_t)qblock * p.n_splits * TC_EPI_GROUPS + s) * TC_BM + r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tau = fminf(tau, __shfl_xor_sync(0xffffffffu, tau, o));

    // k-th smallest approximate score (k rounds of extract-min over (score, entry) keys)
    float kth = -F_INF;
    {
        uint64_t last = 0ull;
        bool first = true;
        for (int round = 0; round < p.k; ++round) {
            uint64_t best = ~0ull;
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {
                if (ci[u] < 0) continue;
                // order-preserving key: scores may be negative
                uint32_t b = __float_as_uint(cs[u]);
                b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                uint64_t key = ((uint64_t)b << 32) | (uint32_t)(u * 32 + lane);
                if ((first || key > last) && key < best) best = key;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other < best ? other : best;
            }
            if (best == ~0ull) { kth = F_INF; break; }             // fewer than k candidates
            last = best;
            first = false;
            uint32_t b = (uint32_t)(best >> 32);
            b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
            kth = __uint_as_float(b);
        }
    }
    const float x2 = p.xnorm[q];
    const float ymax2 = __uint_as_float(p.prm[1]);
    // |S~ - S| <= eps for every db point; the few ulps of the epilogue's own fp32 ops are in c_rel
    const float eps = p.c_rel * sqrtf(x2) * sqrtf(ymax2) + p.c_norm * ymax2 + p.c_abs;
    const float upper = kth + eps;                              // >= true k-th smallest score
    const float margin = 2.2f * p.gamma * fmaxf(upper + x2, 0.f) + 1e-30f;
    const float cut = upper + margin;                           // true score above this: cannot be in the top k
    bool certified = (kth < F_INF) && (tau - eps > cut);
    // candidates that can still be in the top k
    int np = 0;
    if (certified) {
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) {
            bool keep = ci[u] >= 0 && (cs[u] - eps <= cut);
            unsigned m = __ballot_sync(0xffffffffu, keep);
            int pos = np + __popc(m & ((1u << lane) - 1u));
            if (keep && pos < RS_PMAX) s_p[warp][pos] = ci[u];
            np += __popc(m);
        }
        if (np > RS_PMAX || np < p.k) certified = false;
    }
    if (!certified) {
        if (lane == 0) {
            int slot = atomicAdd(&p.stats[0], 1);
            p.fallback_rows[slot] = (int32_t)q;
        }
        return;
    }
    __syncwarp();
    if (lane == 0) { atomicAdd(&p.stats[1], np); atomicAdd(&p.stats[2], 1); }

    // canonical fp32 distances (oracle/knn_oracle.c): acc = fmaf(x[t]-y[t], x[t]-y[t], acc), t ascending
    const float *xq = p.query + q * (int64_t)p.dim;
    float *tile = s_tile[warp];
    uint64_t keys[RS_PMAX / 32];
#pragma unroll
    for (int g = 0; g < RS_PMAX / 32; ++g) {
        keys[g] = ~0ull;
        if (g * 32 >= np) continue;                             // warp-uniform
        const int my = g * 32 + lane;
        const int32_t my_j = my < np ? s_p[warp][my] : -1;
        float acc = 0.f;
        if (VEC4) {
            // Each candidate row is streamed in 128-byte pieces (8 lanes x 16 B, 4 candidates per
            // instruction) with cp.async into tile[buf][cand][chunk ^ (cand & 7)]; chunk n+1 is in flight
            // while chunk n is consumed; lane c then walks candidate c's 32 floats in order with
            // conflict-free 16-byte shared loads.  The query chunk rides in the same group.
            const int part = lane & 7;
            int32_t cjs[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int c = rr * 4 + (lane >> 3);
                cjs[rr] = g * 32 + c < np ? s_p[warp][g * 32 + c] : -1;
            }
            float *xs = tile + 2 * 32 * 32;                                    // [2][32] query chunk
            auto issue = [&](int buf, int tb) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int c = rr * 4 + (lane >> 3);
                    float *dst = tile + buf * 1024 + c * 32 + ((part ^ (c & 7)) << 2);
                    if (cjs[rr] >= 0 && tb + part * 4 < p.dim)
                        cp_async16(dst, p.db + (int64_t)cjs[rr] * p.dim + tb + part * 4);
                    else
                        *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (tb + lane < p.dim) cp_async4(xs + buf * 32 + lane, xq + tb + lane);
                else xs[buf * 32 + lane] = 0.f;
                cp_async_commit();
            };
            __syncwarp();
            issue(0, 0);
            int buf = 0;
            for (int tb = 0; tb < p.dim; tb += 32, buf ^= 1) {
                if (tb + 32 < p.dim) { issue(buf ^ 1, tb + 32); cp_async_wait<1>(); }
                else cp_async_wait<0>();
                __syncwarp();
                const float *row = tile + buf * 1024 + lane * 32;
#pragma unroll
                for (int pc = 0; pc < 8; ++pc) {
                    const float4 y = *reinterpret_cast<const float4 *>(row + ((pc ^ (lane & 7)) << 2));
                    const float4 x = *reinterpret_cast<const float4 *>(xs + buf * 32 + pc * 4);
                    float d0 = x.x - y.x; acc = fmaf(d0, d0, acc);     // zero padding beyond dim adds exactly 0
                    float d1 = x.y - y.y; acc = fmaf(d1, d1, acc);
                    float d2 = x.z - y.z; acc = fmaf(d2, d2, acc);
                    float d3 = x.w - y.w; acc = fmaf(d3, d3, acc);
                }
                __syncwarp();
            }
        } else {
            for (int tb = 0; tb < p.dim; tb += 32) {
                const float xv = (tb + lane < p.dim) ? xq[tb + lane] : 0.f;
                __syncwarp();
                for (int c = 0; c < 32; ++c) {
                    const int cj = g * 32 + c < np ? s_p[warp][g * 32 + c] : -1;
                    float v = 0.f;
                    if (cj >= 0 && tb + lane < p.dim) v = __ldg(p.db + (int64_t)cj * p.dim + tb + lane);
                    tile[lane * RS_TS + c] = v;
                }
                __syncwarp();
                const int tn = min(32, p.dim - tb);
                for (int t = 0; t < tn; ++t) {
                    const float x = __shfl_sync(0xffffffffu, xv, t);
                    const float diff = x - tile[t * RS_TS + lane];
                    acc = fmaf(diff, diff, acc);
                }
            }
        }
        if (my_j >= 0) keys[g] = dist_key(sqrtf(acc), p.db_gid ? p.db_gid[my_j] : my_j);
    }
    // rank by counting; keys are distinct (distinct indices)
#pragma unroll
    for (int g = 0; g < RS_PMAX / 32; ++g) {
        if (g * 32 >= np) continue;
        int rank = 0;
#pragma unroll
        for (int h = 0; h < RS_PMAX / 32; ++h) {
            if (h * 32 >= np) continue;
            for (int l = 0; l < 32; ++l) {
                uint64_t other = __shfl_sync(0xffffffffu, keys[h], l);
                rank += other < keys[g] ? 1 : 0;
            }
        }
        if (keys[g] != ~0ull && rank < p.k) {
            p.out_idx[q * p.k + rank] = key_idx(keys[g]);
            p.out_dist[q * p.k + rank] = key_dist(keys[g]);
        }
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static int make_map(CUtensorMap *map, const __half *base, int64_t rows, int dim_pad, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MMU_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)dim_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dim_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc); return MMU_ERR_CUDA; }
    return MMU_OK;
}

static inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct TcLayout {
    int64_t q_pad, n_pad;
    int dim_pad, width, n_qblocks, n_tiles, n_splits, tiles_per_split, stat_blocks, rows_per_stat_block;
    bool shared_operand;
    size_t off_prm, off_mean, off_partial, off_db16, off_ynorm, off_q16, off_xnorm, off_cidx, off_cscore, off_tau, total;
};

static TcLayout tc_layout(int64_t n_query, int64_t n_db, int dim, bool shared_operand, int min_splits, int split) {
    TcLayout L;
    if (split) shared_operand = false;               // the two operands differ: [hi|hi|lo] vs [hi|lo|hi]
    L.shared_operand = shared_operand;
    L.dim_pad = (int)round_up(dim, TC_BK);
    L.width = split ? 3 * L.dim_pad : L.dim_pad;
    L.n_pad = round_up(n_db, TC_BN);                 // also a multiple of TC_BM
    L.q_pad = shared_operand ? L.n_pad : round_up(n_query, 2 * TC_BM);   // whole CTA pairs
    L.n_qblocks = (int)(round_up(n_query, TC_BM) / TC_BM);
    L.n_tiles = (int)(L.n_pad / TC_BN);
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // split the db range so that the grid fills whole waves (one CTA per SM)
    int best_s = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 8 && s <= L.n_tiles; ++s) {
        if (s > 1 && L.n_tiles / s < 4) break;
        double waves = (double)L.n_qblocks * s / sms;
        double eff = waves / (double)((int64_t)waves + (waves > (int64_t)waves ? 1 : 0));
        if (eff > best_eff + 0.02) { best_eff = eff; best_s = s; }
    }
    // more splits = deeper candidate pool (64 per split): the caller raises min_splits for rows whose
    // neighbourhood gaps are too small for one list to certify
    if (min_splits > best_s) best_s = min_splits;
    if (min_splits < 0) best_s = -min_splits;        // pinned to exactly that many splits (per-block tile sets of the pruned search)
    if (best_s > 8) best_s = 8;
    if (best_s > L.n_tiles) best_s = L.n_tiles;
    L.n_splits = best_s;
    L.tiles_per_split = (L.n_tiles + L.n_splits - 1) / L.n_splits;
    L.stat_blocks = (int)((n_db + 31) / 32 < 296 ? (n_db + 31) / 32 : 296);
    if (L.stat_blocks < 1) L.stat_blocks = 1;
    L.rows_per_stat_block = (int)((n_db + L.stat_blocks - 1) / L.stat_blocks);
    size_t o = 0;
    L.off_prm = o; o += 256;
    L.off_mean = o; o += align256(sizeof(float) * L.dim_pad);
    L.off_partial = o; o += align256(sizeof(float) * (size_t)296 * dim) * (shared_operand ? 1 : 2);
    L.off_db16 = o; o += align256(sizeof(__half) * (size_t)L.n_pad * L.width);
    L.off_ynorm = o; o += align256(sizeof(float) * (size_t)L.n_pad);
    L.off_q16 = o; if (!shared_operand) o += align256(sizeof(__half) * (size_t)L.q_pad * L.width);
    L.off_xnorm = o; o += align256(sizeof(float) * (size_t)L.q_pad);
    size_t cand = (size_t)L.n_qblocks * L.n_splits * TC_KP * TC_BM;
    L.off_cidx = o; o += align256(sizeof(int32_t) * cand);
    L.off_cscore = o; o += align256(sizeof(float) * cand);
    L.off_tau = o; o += align256(sizeof(float) * (size_t)L.n_qblocks * L.n_splits * TC_EPI_GROUPS * TC_BM);
    L.total = o;
    return L;
}

}  // namespace mmu

extern "C" size_t mmu_knn_tc_workspace_bytes(int64_t n_query, int64_t n_db, int dim, int query_is_db, int min_splits,
                                             int precision) {
    if (n_query <= 0 || n_db <= 0 || dim <= 0) return 0;
    return mmu::tc_layout(n_query, n_db, dim, query_is_db != 0, min_splits, precision != 0).total;
}

static void tc_error_constants(int dim, int dim_pad, int width, int split, float *c_rel, float *c_norm, float *c_abs, float *gamma) {
    // fp16 rounding of both operands (2 * 2^-11 on each product, x2 for the -2 factor, 1% headroom) plus
    // the fp32 accumulation across dim_pad/16 MMAs and the final fma
    *c_rel = 1.01f * 0x1p-9f + ((float)(dim_pad / 16) + 16.f) * 0x1p-22f;
    if (split)   // hi+lo of each operand is exact to 2^-22 relative, the dropped lo.lo term is <= 2^-22 |x||y|; x2 for -2
        *c_rel = 7.0f * 0x1p-22f + ((float)(width / 16) + 16.f) * 0x1p-22f;
    // components below the fp16 subnormal spacing: <= 2^-25 absolute per element and operand, scaled values <= 2^14
    *c_abs = (float)dim * 0x1p-9f;
    *c_norm = ((float)(dim_pad / 32) + 8.f) * 0x1p-23f;      // lane-strided fma chain + warp tree + final fma
    *gamma = ((float)dim + 4.f) * 0x1p-24f;
}

extern "C" int mmu_knn_tc_layout(int64_t n_query, int64_t n_db, int dim, int query_is_db, int min_splits, int precision,
                                 int64_t *out_words, float *out_consts) {
    using namespace mmu;
    MMU_CHECK_ARG(out_words && out_consts && n_query > 0 && n_db > 0 && dim > 0, "mmu_knn_tc_layout: bad arguments");
    const TcLayout L = tc_layout(n_query, n_db, dim, query_is_db != 0, min_splits, precision != 0);
    out_words[0] = (int64_t)L.off_prm;     out_words[1] = (int64_t)(L.shared_operand ? L.off_ynorm : L.off_xnorm);
    out_words[2] = (int64_t)L.off_ynorm;   out_words[3] = (int64_t)L.off_cidx;
    out_words[4] = (int64_t)L.off_cscore;  out_words[5] = (int64_t)L.off_tau;
    out_words[6] = L.n_qblocks;            out_words[7] = L.n_splits;
    out_words[8] = L.n_tiles;              out_words[9] = TC_KP;
    out_words[10] = TC_BM;                 out_words[11] = TC_BN;
    tc_error_constants(dim, L.dim_pad, L.width, precision != 0, &out_consts[0], &out_consts[1], &out_consts[2], &out_consts[3]);
    return MMU_OK;
}

extern "C" int mmu_knn_tc(const float *query, int64_t n_query, const float *db, int64_t n_db, int dim, int k,
                          int exclude_self, int64_t query_index_base, const int32_t *query_gid, int query_is_db,
                          int min_splits, int precision, void *workspace, size_t workspace_bytes, int32_t *out_idx,
                          float *out_dist, int32_t *stats, int32_t *fallback_rows, mmu_stream_t stream) {
    return mmu_knn_tc_ex(query, n_query, db, n_db, dim, k, exclude_self, query_index_base, query_gid, query_is_db, min_splits,
                         precision, workspace, workspace_bytes, out_idx, out_dist, stats, fallback_rows, 7, nullptr, nullptr,
                         nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int mmu_knn_tc_ex(const float *query, int64_t n_query, const float *db, int64_t n_db, int dim, int k,
                             int exclude_self, int64_t query_index_base, const int32_t *query_gid, int query_is_db,
                             int min_splits, int precision, void *workspace, size_t workspace_bytes, int32_t *out_idx,
                             float *out_dist, int32_t *stats, int32_t *fallback_rows, int stages,
                             const int32_t *qb_tile_begin, const int32_t *qb_tile_end, const int32_t *tile_ptr,
                             const int32_t *tile_list, int resume, const int32_t *db_gid, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(query && db && workspace && out_idx && out_dist && stats && fallback_rows, "mmu_knn_tc: null pointer");
    MMU_CHECK_ARG(stages >= 1 && stages <= 7, "mmu_knn_tc_ex: stages must be a mask of 1 (prep) | 2 (candidates) | 4 (rescore)");
    MMU_CHECK_ARG((qb_tile_begin == nullptr) == (qb_tile_end == nullptr) && (tile_ptr == nullptr) == (tile_list == nullptr),
                  "mmu_knn_tc_ex: tile ranges / tile lists come in pairs");
    const bool pruned = qb_tile_begin || tile_list;
    MMU_CHECK_ARG(k >= 1 && k <= MMU_KNN_TC_MAX_K, "mmu_knn_tc: k=%d outside [1,%d]", k, MMU_KNN_TC_MAX_K);
    MMU_CHECK_ARG(dim >= 1 && n_db >= 1 && n_query >= 0, "mmu_knn_tc: bad sizes");
    MMU_CHECK_ARG(n_db < (int64_t)2147483647 - TC_BN, "mmu_knn_tc: db index exceeds int32");
    MMU_CHECK_ARG(!query_is_db || (query == db && n_query == n_db && query_index_base == 0),
                  "mmu_knn_tc: query_is_db requires identical query and db");
    cudaStream_t st = as_stream(stream);
    if (stages & 4) MMU_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * 4, st));
    if (n_query == 0) return MMU_OK;
    MMU_CHECK_ARG(min_splits >= -8 && min_splits <= 8, "mmu_knn_tc: min_splits outside [-8,8]");
    MMU_CHECK_ARG(precision == 0 || precision == 1, "mmu_knn_tc: precision must be 0 (fp16) or 1 (split fp16)");
    const TcLayout L = tc_layout(n_query, n_db, dim, query_is_db != 0, min_splits, precision != 0);
    const int split = precision != 0;
    MMU_CHECK_ARG(workspace_bytes >= L.total, "mmu_knn_tc: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    MMU_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "mmu_knn_tc: workspace must be 256-byte aligned");
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    uint32_t *prm = reinterpret_cast<uint32_t *>(ws + L.off_prm);
    float *mean = reinterpret_cast<float *>(ws + L.off_mean);
    float *partial = reinterpret_cast<float *>(ws + L.off_partial);
    __half *db16 = reinterpret_cast<__half *>(ws + L.off_db16);
    float *ynorm = reinterpret_cast<float *>(ws + L.off_ynorm);
    __half *q16 = L.shared_operand ? db16 : reinterpret_cast<__half *>(ws + L.off_q16);
    float *xnorm = L.shared_operand ? ynorm : reinterpret_cast<float *>(ws + L.off_xnorm);
    int32_t *cidx = reinterpret_cast<int32_t *>(ws + L.off_cidx);
    float *cscore = reinterpret_cast<float *>(ws + L.off_cscore);
    float *tau = reinterpret_cast<float *>(ws + L.off_tau);

    MMU_CHECK_ARG(!pruned || min_splits < 0, "mmu_knn_tc_ex: per-block tile sets need a pinned split count (min_splits < 0)");
    if (stages & 1) {
    // ---- prep
    MMU_CUDA(cudaMemsetAsync(prm, 0, 256, st));
    tc_colsum_kernel<<<L.stat_blocks, 256, 0, st>>>(db, n_db, dim, L.rows_per_stat_block, partial, prm + 2);
    int launches = 1;
    if (!L.shared_operand) {
        // the queries only contribute to the range of the common scale
        int qb = (int)((n_query + 31) / 32 < 296 ? (n_query + 31) / 32 : 296);
        int rpb = (int)((n_query + qb - 1) / qb);
        // the queries' partial sums are not used: they go to the second half of the scratch area
        float *scratch = partial + (size_t)296 * dim;
        tc_colsum_kernel<<<qb, 256, 0, st>>>(query, n_query, dim, rpb, scratch, prm + 2);
        ++launches;
    }
    tc_finish_stats_kernel<<<1, 1024, 0, st>>>(partial, L.stat_blocks, dim, n_db, mean, prm, 1);
    tc_convert_kernel<<<(unsigned)((L.n_pad * 32 + 255) / 256), 256, 0, st>>>(db, n_db, L.n_pad, dim, L.dim_pad,
                                                                              split ? 2 : 0, mean, prm, db16, ynorm,
                                                                              HUGE_VALF, prm + 1);
    launches += 2;
    if (!L.shared_operand) {
        tc_convert_kernel<<<(unsigned)((L.q_pad * 32 + 255) / 256), 256, 0, st>>>(query, n_query, L.q_pad, dim, L.dim_pad,
                                                                                  split ? 1 : 0, mean, prm, q16, xnorm, 0.f,
                                                                                  nullptr);
        ++launches;
    }
    MMU_LAUNCH_CHECK_N(launches);
    }

    if (stages & 2) {
    // ---- candidates
    // CTA pairs (cta_group::2) for long rows; option knn_cta_pairs = 0 keeps one CTA per query block (A/B measurements)
    const int pairs_allowed = option(OPT_KNN_CTA_PAIRS) != 0;
    // Short rows (width < 512: at most 7 k-blocks per tile) are bound by the per-tile epilogue, not by operand
    // traffic; there a pair only couples the two epilogues (1M x 128, k = 30: 795 ms paired, 695 ms unpaired).
    // (per-block tile sets: the two CTAs of a pair would have to walk the same tiles, so the pruned search is single-CTA)
    const int ncta = (pairs_allowed && !pruned && L.n_qblocks >= 2 && L.width >= 512) ? 2 : 1;
    CUtensorMap tm_q, tm_db;
    int rc = make_map(&tm_q, q16, L.q_pad, L.width, TC_BM);
    if (rc) return rc;
    rc = make_map(&tm_db, db16, L.n_pad, L.width, TC_BN / ncta);
    if (rc) return rc;
    if (first_use_on_device(SITE_KNN_CANDIDATES)) {     // the attribute is per device
        MMU_CUDA(cudaFuncSetAttribute(knn_tc_candidates_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<1>::SMEM_BYTES));
        MMU_CUDA(cudaFuncSetAttribute(knn_tc_candidates_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<2>::SMEM_BYTES));
    }
    TcParams tp;
    tp.ynorm = ynorm;
    tp.n_kblocks = L.width / TC_BK;
    tp.n_tiles = L.n_tiles;
    tp.tiles_per_split = L.tiles_per_split;
    tp.n_splits = L.n_splits;
    tp.n_qblocks = L.n_qblocks;
    tp.cand_idx = cidx;
    tp.cand_score = cscore;
    tp.tau = tau;
    {
        int sms = sm_count();
        if (sms <= 0) sms = 148;
        const size_t q_tiles_in_flight = (size_t)sms * TC_BM * L.width * 2;
        tp.split_major = (L.n_splits > 1 && q_tiles_in_flight > ((size_t)48 << 20) && L.n_qblocks <= 65535) ? 1 : 0;   // grid.y limit
    }
    tp.window_begin = 0;
    tp.window_tiles = L.tiles_per_split;
    tp.resume = resume ? 1 : 0;
    tp.qb_tile_begin = qb_tile_begin; tp.qb_tile_end = qb_tile_end; tp.tile_ptr = tile_ptr; tp.tile_list = tile_list;
    if (pruned) tp.split_major = 0;
    if (ncta == 2) {
        // Pairs do not stay in lock-step over a long database pass the way single CTAs do (traced on 1M x 768: the
        // spread of the CTAs' positions grows by ~0.2 ms per wave until every pair streams its tiles from DRAM,
        // 3.9 TB instead of 82 GB).  So a database region that does not fit in L2 is walked in windows of ~48 MB,
        // one launch per window over ALL query blocks; the lists carry over through cand_idx / cand_score.
        const size_t tile_bytes = (size_t)TC_BN * L.width * 2;
        const size_t region = (size_t)L.n_splits * L.tiles_per_split * tile_bytes;
        // option knn_window_mb: window size; set explicitly (>= 0) it applies to every pair launch (tests), 0 = one launch
        const long long wopt = option(OPT_KNN_WINDOW_MB);
        const bool forced = wopt >= 0;
        const long long window_mb = forced ? wopt : 48;
        // (split-major launches keep few query blocks and every split in flight; measured fine as one launch)
        if (window_mb > 0 && (forced || (region > ((size_t)160 << 20) && !tp.split_major))) {
            size_t w = ((size_t)window_mb << 20) / ((size_t)L.n_splits * tile_bytes);
            tp.window_tiles = (int)(w < 8 ? 8 : w);
            if (tp.window_tiles > L.tiles_per_split) tp.window_tiles = L.tiles_per_split;
        }
        const unsigned qb = (unsigned)((L.n_qblocks + 1) & ~1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = tp.split_major ? dim3(2 * L.n_splits, qb / 2) : dim3(qb, L.n_splits);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TcCfg<2>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;    // two query blocks, x-neighbours in the grid
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        for (int w0 = 0; w0 < L.tiles_per_split; w0 += tp.window_tiles) {
            tp.window_begin = w0;
            tp.resume = (w0 > 0 || resume) ? 1 : 0;
            MMU_CUDA(cudaLaunchKernelEx(&cfg, knn_tc_candidates_kernel<2>, tm_q, tm_db, tp));
            if (w0 > 0) mmu_launch_count_add(1);
        }
    } else {
        const dim3 grid = tp.split_major ? dim3(L.n_splits, L.n_qblocks) : dim3(L.n_qblocks, L.n_splits);
        knn_tc_candidates_kernel<1><<<grid, TC_THREADS, TcCfg<1>::SMEM_BYTES, st>>>(tm_q, tm_db, tp);
    }
    note_kernel(SITE_KNN_CANDIDATES, "knn_tc_candidates_kernel<%d>(%s%s, width %d, %d split%s)", ncta,
                ncta == 2 ? "cta_group::2 pairs" : "cta_group::1", pruned ? ", per-block tile sets" : "", L.width, L.n_splits,
                L.n_splits > 1 ? "s" : "");
    MMU_LAUNCH_CHECK();
    }

    if (!(stages & 4)) return MMU_OK;
    // ---- certify + rescore
    RsParams rp;
    rp.query = query; rp.db = db; rp.n_query = n_query; rp.n_db = n_db; rp.dim = dim; rp.k = k;
    rp.exclude_self = exclude_self; rp.query_index_base = query_index_base; rp.query_gid = query_gid; rp.db_gid = db_gid;
    rp.xnorm = xnorm; rp.prm = prm; rp.n_splits = L.n_splits;
    rp.cand_idx = cidx; rp.cand_score = cscore; rp.tau = tau;
    tc_error_constants(dim, L.dim_pad, L.width, split, &rp.c_rel, &rp.c_norm, &rp.c_abs, &rp.gamma);
    rp.out_idx = out_idx; rp.out_dist = out_dist; rp.stats = stats; rp.fallback_rows = fallback_rows;
    unsigned rblocks = (unsigned)((n_query + RS_WARPS - 1) / RS_WARPS);
    bool vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    if (vec4) knn_tc_rescore_kernel<true><<<rblocks, RS_WARPS * 32, 0, st>>>(rp);
    else knn_tc_rescore_kernel<false><<<rblocks, RS_WARPS * 32, 0, st>>>(rp);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
