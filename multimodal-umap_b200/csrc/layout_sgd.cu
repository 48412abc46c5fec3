// layout_sgd.cu -- K7 (edge sampling + attractive/repulsive forces), K8 (InfoNCE), K9 (Adam).
//
// ref: /root/reference/impl/model.py:396-481 (_train), :312-334 (force terms),
//      :364-394 (_infonce_loss), :403,:474-476 (Adam).
//
// The reference builds an autograd graph per epoch and lets index_put_(accumulate=True) scatter
// the gradient; here the closed-form gradient of the same loss is accumulated straight into a
// dense gradient table with vector red.global.add (Hogwild-style: no ordering between edges),
// and a fused Adam pass consumes and clears it.  Edges are COO (row, col, w) sorted by
// (row, col) as .coalesce() leaves them; a kept edge list (positions into that COO) is either
// uploaded from the host-replayed sample stream (parity mode) or produced on the device by
// mmu_edge_sample (Philox4x32-10 counter stream, throughput mode).
#include "common.cuh"

namespace mmu {

// ------------------------------------------------------------------ optimiser state
__global__ void opt_state_init_kernel(OptState *s) {
    s->epoch = 0; s->step = 0; s->step_size = 0.f; s->bc2_sqrt = 1.f;
    for (int i = 0; i < 4; ++i) s->reserved[i] = 0;
}
__global__ void opt_state_advance_kernel(OptState *s, double lr, double beta1, double beta2) {
    uint32_t step = s->step + 1;
    s->step = step;
    s->epoch = s->epoch + 1;
    double bc1 = 1.0 - pow(beta1, (double)step);
    double bc2 = 1.0 - pow(beta2, (double)step);
    s->step_size = (float)(lr / bc1);
    s->bc2_sqrt = (float)sqrt(bc2);
}


// ------------------------------------------------------------------ K7a: edge sampling
// Kept-edge list header (device, 4 x int32): [0] number of kept edges, [1] capacity of the record array (written
// by the host once), [2] overflow flag (set when an epoch kept more than the capacity; the host checks it).
// One 16-byte record per kept edge: {edge position, row, col, row-batch}.  The force kernels start every edge
// from this single sequential load (staged through shared memory) instead of a kept_pos -> row/col chase.
//
// thread handles 4 consecutive edges (one Philox call -> 4 uniforms)
__global__ void __launch_bounds__(256)
edge_sample_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ col, const float *__restrict__ w,
                   int64_t edge_lo, int64_t nnz, int batch_size, uint64_t seed, const OptState *__restrict__ st,
                   int4 *__restrict__ kept_rec, int32_t *__restrict__ kept_hdr, int32_t *__restrict__ batch_kept,
                   int64_t epoch_override) {
    // edges [edge_lo, nnz) of the global COO; the Philox counter is the GLOBAL quad index, so a
    // shard draws exactly what the single-GPU run draws for the same edges
    // epoch_override >= 0: the host's epoch number; -1: the device-resident counter; -2: the counter + 1 (the NEXT epoch's
    // sample, drawn while this epoch's forces run, inside a captured CUDA graph that has no per-epoch arguments)
    const uint32_t epoch = epoch_override >= 0 ? (uint32_t)epoch_override : st->epoch + (epoch_override == -2 ? 1u : 0u);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n4 = (nnz + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int cap = kept_hdr[1];
    __shared__ int s_wtot[8];
    __shared__ int s_base;
    // every thread of a block iterates the same number of times (block-level compaction below)
    for (int64_t bbase = (edge_lo >> 2) + (int64_t)blockIdx.x * blockDim.x; bbase < n4; bbase += stride) {
        const int64_t q = bbase + threadIdx.x;
        int32_t pos[4];
        int cnt = 0;
        int32_t brow[4], erow[4];
        if (q < n4) {
            const Philox4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), epoch, STREAM_KEEP, k0, k1);
            const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
            const int64_t e0 = q * 4;
            float wv[4];
            if (e0 >= edge_lo && e0 + 3 < nnz) {
                const float4 t = *reinterpret_cast<const float4 *>(w + e0);      // quads are 16-byte aligned
                wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) wv[i] = (e0 + i >= edge_lo && e0 + i < nnz) ? w[e0 + i] : -1.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (u01(rv[i]) < wv[i]) {                                         // ref: model.py:432  rand < w
                    pos[cnt] = (int32_t)(e0 + i);
                    erow[cnt] = row[e0 + i];
                    brow[cnt] = erow[cnt] / batch_size;
                    ++cnt;
                }
            }
        }
        // compaction: warp scan, block scan over the 8 warp totals, ONE atomic per block and iteration
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { int t = s_wtot[u]; s_wtot[u] = tot; tot += t; }
            s_base = tot ? atomicAdd(&kept_hdr[0], tot) : 0;
            if (tot && s_base + tot > cap) kept_hdr[2] = 1;                       // more kept edges than the list holds
        }
        __syncthreads();
        const int off = s_base + s_wtot[warp] + incl - cnt;
        for (int i = 0; i < cnt; ++i)
            if (off + i < cap) kept_rec[off + i] = make_int4(pos[i], erow[i], col[pos[i]], brow[i]);
        // per-batch counts, aggregated on the warp's most common batch (edges are row sorted)
        unsigned has = __ballot_sync(0xffffffffu, cnt > 0);
        int src_lane = has ? __ffs(has) - 1 : 0;
        int b_ref = __shfl_sync(0xffffffffu, cnt ? brow[0] : -1, src_lane);
        int same = 0;
        for (int i = 0; i < cnt; ++i) {
            if (brow[i] == b_ref) ++same;
            else atomicAdd(&batch_kept[brow[i]], 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) same += __shfl_xor_sync(0xffffffffu, same, o);
        if (lane == 0 && same) atomicAdd(&batch_kept[b_ref], same);
        __syncthreads();                                                          // s_wtot / s_base reused next iteration
    }
}

// host sample stream: the kept positions come from the replayed CPU draws; build their records
__global__ void __launch_bounds__(256)
edge_records_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                    const int32_t *__restrict__ kept_pos, int64_t n_kept, int batch_size, int4 *__restrict__ kept_rec,
                    int32_t *__restrict__ kept_hdr) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) { kept_hdr[0] = (int32_t)n_kept; kept_hdr[2] = n_kept > kept_hdr[1] ? 1 : 0; }
    if (e >= n_kept || e >= kept_hdr[1]) return;
    const int32_t p = kept_pos[e];
    const int32_t r = row[p];
    kept_rec[e] = make_int4(p, r, col[p], r / batch_size);
}

// ------------------------------------------------------------------ K7b: forces
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int VEC>
struct Vec {
    float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec(const float *p) {
    Vec<VEC> r;
    if (VEC == 4) { float4 t = *reinterpret_cast<const float4 *>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else if (VEC == 2) { float2 t = *reinterpret_cast<const float2 *>(p); r.v[0] = t.x; r.v[1] = t.y; }
    else { r.v[0] = *p; }
    return r;
}
// L2-only load (ld.global.cg): random tail rows of a table larger than L1 only pollute it
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec_cg(const float *p) {
    Vec<VEC> r;
    if (VEC == 4) { float4 t = __ldcg(reinterpret_cast<const float4 *>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else if (VEC == 2) { float2 t = __ldcg(reinterpret_cast<const float2 *>(p)); r.v[0] = t.x; r.v[1] = t.y; }
    else { r.v[0] = __ldcg(p); }
    return r;
}
template <int VEC>
__device__ __forceinline__ void red_vec(float *p, const Vec<VEC> &g, float sign) {
    if (VEC == 4) red_add_v4(p, sign * g.v[0], sign * g.v[1], sign * g.v[2], sign * g.v[3]);
    else if (VEC == 2) red_add_v2(p, sign * g.v[0], sign * g.v[1]);
    else red_add_f32(p, sign * g.v[0]);
}

// attractive: d/ds log(1+a s^b);  repulsive: d/ds -log(a s^b/(1+a s^b)+1e-6)   (x2 for ds/dy)
__device__ __forceinline__ void attr_terms(float s_raw, float a, float b, float &coef, float &loss) {
    float s = fmaxf(s_raw, 1e-6f);
    float sb = powf(s, b);
    loss = logf(1.0f + a * sb);
    coef = (s_raw >= 1e-6f) ? (2.0f * a * b * sb / s) / (1.0f + a * sb) : 0.0f;
}
__device__ __forceinline__ void rep_terms(float s_raw, float a, float b, float &coef, float &loss) {
    float s = fmaxf(s_raw, 1e-6f);
    float sb = powf(s, b);
    float q = a * sb;
    float f = q / (1.0f + q) + 1e-6f;
    loss = -logf(f);
    coef = (s_raw >= 1e-6f) ? -(2.0f * a * b * sb / s) / (f * (1.0f + q) * (1.0f + q)) : 0.0f;
}

__device__ __forceinline__ int kept_total(const int32_t *__restrict__ hdr) { return min(hdr[0], hdr[1]); }

// Plain loop form: one group of LANES threads per kept edge (grid-stride), each thread owns VEC consecutive
// components (dim == LANES*VEC), any num_rep.  The A/B partner of the staged kernel below (option force_staged = 0)
// and the path of num_rep values it has no instantiation for.
template <int VEC, int LANES>
__global__ void __launch_bounds__(256)
edge_forces_kernel(const int4 *__restrict__ kept_rec, const int32_t *__restrict__ kept_hdr,
                   const int32_t *__restrict__ neg, const int32_t *__restrict__ batch_kept, int n_batches,
                   int num_rep, uint32_t rep_count, const float *__restrict__ head,
                   const float *__restrict__ tail, float *__restrict__ grad_head, float *__restrict__ grad_tail,
                   int dim, float a, float b, uint64_t seed, const OptState *__restrict__ st,
                   float *__restrict__ loss_out) {
    const int n_kept = kept_total(kept_hdr);
    const uint32_t epoch = st->epoch;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int gl = threadIdx.x % LANES;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
    const float inv_nb = 1.0f / (float)n_batches;
    float loss_acc = 0.f;
    // groups of the same warp must iterate together (shuffles): round the trip count per warp
    const int64_t gpw = 32 / LANES;                                  // groups per warp
    const int64_t wfirst = (gid / gpw) * gpw;
    for (int64_t e0 = wfirst; e0 < n_kept; e0 += n_groups) {
        const int64_t e = e0 + (gid - wfirst);
        const bool active = e < n_kept;
        int32_t p = 0, i = 0, j = 0;
        float sc_a = 0.f, sc_r = 0.f;
        if (active) {
            const int4 rec = kept_rec[e];
            p = rec.x; i = rec.y; j = rec.z;
            float kb = (float)batch_kept[rec.w];
            sc_a = inv_nb / kb;                                      // mean over kept, mean over batches
            sc_r = inv_nb / (kb * (float)num_rep);
        }
        Vec<VEC> yi = load_vec<VEC>(head + (int64_t)i * dim + gl * VEC);
        Vec<VEC> gi;
#pragma unroll
        for (int c = 0; c < VEC; ++c) gi.v[c] = 0.f;
        // attractive term (ref: model.py:312-322)
        {
            Vec<VEC> yj = load_vec<VEC>(tail + (int64_t)j * dim + gl * VEC);
            Vec<VEC> df;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) { df.v[c] = yi.v[c] - yj.v[c]; s = fmaf(df.v[c], df.v[c], s); }
            s = group_sum<LANES>(s);
            float coef, l;
            attr_terms(s, a, b, coef, l);
            coef *= sc_a;
            if (gl == 0) loss_acc += l * sc_a;
            Vec<VEC> g;
#pragma unroll
            for (int c = 0; c < VEC; ++c) { g.v[c] = coef * df.v[c]; gi.v[c] += g.v[c]; }
            if (active && grad_tail) red_vec<VEC>(grad_tail + (int64_t)j * dim + gl * VEC, g, -1.0f);
        }
        // repulsive terms (ref: model.py:324-334, 441-449)
        Philox4 rnd = {0, 0, 0, 0};
        for (int r = 0; r < num_rep; ++r) {
            uint32_t l_idx;
            if (neg) {
                l_idx = active ? (uint32_t)neg[e * num_rep + r] : 0u;
            } else {
                if ((r & 3) == 0) rnd = philox4x32_10((uint32_t)p, (uint32_t)(r >> 2), epoch, STREAM_NEG, k0, k1);
                uint32_t x = (r & 3) == 0 ? rnd.x : (r & 3) == 1 ? rnd.y : (r & 3) == 2 ? rnd.z : rnd.w;
                l_idx = urange(x, rep_count);
            }
            Vec<VEC> yl = load_vec<VEC>(tail + (int64_t)l_idx * dim + gl * VEC);
            Vec<VEC> df;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) { df.v[c] = yi.v[c] - yl.v[c]; s = fmaf(df.v[c], df.v[c], s); }
            s = group_sum<LANES>(s);
            float coef, l;
            rep_terms(s, a, b, coef, l);
            coef *= sc_r;
            if (gl == 0) loss_acc += l * sc_r;
            Vec<VEC> g;
#pragma unroll
            for (int c = 0; c < VEC; ++c) { g.v[c] = coef * df.v[c]; gi.v[c] += g.v[c]; }
            if (active && grad_tail) red_vec<VEC>(grad_tail + (int64_t)l_idx * dim + gl * VEC, g, -1.0f);
        }
        if (active) red_vec<VEC>(grad_head + (int64_t)i * dim + gl * VEC, gi, 1.0f);
    }
    if (loss_out) {
        loss_acc = warp_sum(loss_acc);
        if ((threadIdx.x & 31) == 0 && loss_acc != 0.f) atomicAdd(loss_out, loss_acc);
    }
}

#ifndef MMU_STAGED_WIN_BPS
#define MMU_STAGED_WIN_BPS 3      // 4 (<= 64 registers, no spills) measured 10.66 against 10.51 ms/epoch on 10M x 2-D: no gain
#endif
// ------------------------------------------------------------------ K7b, staged run form (default)
// FAST selects s^b = ex2(b*lg2(s)) and an approximate reciprocal (device sample stream); !FAST keeps
// powf/div exactly as the loop version (host-replayed stream, parity tests).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool FAST>
__device__ __forceinline__ float pair_coef(float s_raw, bool attractive, float a, float b, float sc_a, float sc_r,
                                           bool want_loss, float &loss) {
    if (FAST) {
        const float s = fmaxf(s_raw, 1e-6f);
        const float q = a * ex2_approx(b * __log2f(s));          // a s^b
        const float opq = 1.0f + q;
        const float num = (attractive ? sc_a : -sc_r) * 2.0f * b * q;
        const float den = attractive ? s * opq : s * opq * fmaf(1e-6f, opq, q);
        if (want_loss) loss = attractive ? sc_a * logf(opq) : -sc_r * logf(q / opq + 1e-6f);
        return (s_raw >= 1e-6f) ? __fdividef(num, den) : 0.0f;
    } else {
        float coef, l;
        if (attractive) { attr_terms(s_raw, a, b, coef, l); coef *= sc_a; l *= sc_a; }
        else { rep_terms(s_raw, a, b, coef, l); coef *= sc_r; l *= sc_r; }
        loss = l;
        return coef;
    }
}

__device__ __forceinline__ void cp_async16_cg(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// records one group handles per round (runs of equal head rows are amortised over them) and the threads' share of
// the shared-memory staging buffers: 2 x GROUPS x (T+1) records of 16 bytes, at most 40 KB per block
template <int LANES>
struct StageCfg {
    static constexpr int T = LANES == 1 ? 4 : LANES == 2 ? 8 : 16;
    static constexpr int GROUPS = 256 / LANES;
    static constexpr int STRIDE = T + 1;           // one pad record per group: the groups' 16-byte reads hit distinct banks
};

// A block owns a contiguous, equal share of the kept list (row sorted inside every sampler block's chunk, so
// consecutive records mostly share their head row) and walks it in rounds: all 256 threads copy the round's records
// global -> shared with coalesced 16-byte cp.async (double buffered: round r+1 is in flight while round r is
// processed), then every group of LANES lanes takes a contiguous run of `tr` <= T records out of shared memory.  The
// head gradient of a run is accumulated in registers and leaves with ONE vector red per (run, row); the R+1 tail rows
// of an edge are gathered first (R+1 independent 16-byte loads in flight per lane), the R+1 force coefficients are
// evaluated by different lanes of the group and exchanged by shuffle, the negatives' Philox call is split the same way.
//
// WIN (tables larger than the L2: C4's 10M x 2-D p and g are 80 MB each, and a random 8-byte access moves a 32-byte
// DRAM sector): the launch handles only the (edge, tail) pairs whose TAIL row lies in [win_lo, win_hi); the host
// launches once per window, so that every random gather / red of a pass hits a slice of p and g that stays L2
// resident, DRAM sees each table once per pass, and the kept records stream through.  The negatives are counter
// based (Philox keyed on the edge position), so every pass regenerates the same draws: the same pairs, the same
// arithmetic as the single pass -- only the order of the atomics changes.
// blocks per SM.  The windowed small-row form (LANES == 1: d = 2, 4) is issue bound with a third of the warp slots in use
// (ncu r02: issue active 69 %, warps active 33 %); a fourth block per SM was tried for it and did not pay (see above)
constexpr int staged_blocks_per_sm(int lanes, bool win) { return (lanes == 1 && win) ? MMU_STAGED_WIN_BPS : 3; }

template <int VEC, int LANES, int R, bool FAST, bool WIN>
__global__ void __launch_bounds__(256, staged_blocks_per_sm(LANES, WIN))
edge_forces_staged_kernel(const int4 *__restrict__ kept_rec, const int32_t *__restrict__ kept_hdr,
                          const int32_t *__restrict__ neg, const int32_t *__restrict__ batch_kept, int n_batches,
                          uint32_t rep_count, const float *__restrict__ head, const float *__restrict__ tail,
                          float *__restrict__ grad_head, float *__restrict__ grad_tail, float a, float b, uint64_t seed,
                          const OptState *__restrict__ st, float *__restrict__ loss_out, uint32_t win_lo, uint32_t win_len) {
    using SC = StageCfg<LANES>;
    constexpr int DIM = VEC * LANES;
    constexpr int NP = R + 1;                                 // pairs per kept edge: 1 attractive + R repulsive
    constexpr int NCALL = (R + 3) / 4;                        // Philox calls per edge
    constexpr int ROUNDS = (NP + LANES - 1) / LANES;
    constexpr int T = SC::T, GROUPS = SC::GROUPS, STRIDE = SC::STRIDE;
    __shared__ int4 s_rec[2][GROUPS * STRIDE];
    const int n_kept = kept_total(kept_hdr);
    const uint32_t epoch = st->epoch;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int gl = threadIdx.x % LANES;
    const int grp = threadIdx.x / LANES;
    const float inv_nb = 1.0f / (float)n_batches;
    const bool want_loss = loss_out != nullptr;
    float loss_acc = 0.f;
    // this block's share, split into rounds of GROUPS x tr records with tr <= T as equal as possible
    const int per_block = (int)(((int64_t)n_kept + gridDim.x - 1) / gridDim.x);
    const int64_t blk_lo = (int64_t)blockIdx.x * per_block;
    const int64_t blk_hi = min((int64_t)n_kept, blk_lo + per_block);
    const int mine = blk_hi > blk_lo ? (int)(blk_hi - blk_lo) : 0;
    const int n_rounds = (mine + GROUPS * T - 1) / (GROUPS * T);
    const int tr = n_rounds ? (mine + n_rounds * GROUPS - 1) / (n_rounds * GROUPS) : 0;
    const int per_round = tr * GROUPS;

    auto issue = [&](int buf, int round) {
        const int64_t base = blk_lo + (int64_t)round * per_round;
        for (int r = threadIdx.x; r < per_round; r += 256) {
            if (base + r < blk_hi) cp_async16_cg(&s_rec[buf][(r / tr) * STRIDE + (r % tr)], kept_rec + base + r);
        }
        cp_async_commit_group();
    };
    if (n_rounds) issue(0, 0);
    int32_t cur_i = -1;
    Vec<VEC> gi;
#pragma unroll
    for (int c = 0; c < VEC; ++c) gi.v[c] = 0.f;
    for (int round = 0, buf = 0; round < n_rounds; ++round, buf ^= 1) {
        if (round + 1 < n_rounds) { issue(buf ^ 1, round + 1); cp_async_wait_group<1>(); }
        else cp_async_wait_group<0>();
        __syncthreads();
        const int64_t e_begin = blk_lo + (int64_t)round * per_round + (int64_t)grp * tr;
        for (int t = 0; t < tr; ++t) {
            const int64_t e = e_begin + t;
            const bool active = e < blk_hi;
            int32_t p = 0, i = 0;
            uint32_t t_idx[NP];
            float sc_a = 0.f, sc_r = 0.f;
            t_idx[0] = 0;
            if (active) {
                const int4 rec = s_rec[buf][grp * STRIDE + t];
                p = rec.x;
                i = rec.y;
                t_idx[0] = (uint32_t)rec.z;
                const float kb = (float)batch_kept[rec.w];
                sc_a = inv_nb / kb;
                sc_r = inv_nb / (kb * (float)R);
                if (i != cur_i) {                                     // group-uniform: the lanes of a group share e
                    if (cur_i >= 0) red_vec<VEC>(grad_head + (int64_t)cur_i * DIM + gl * VEC, gi, 1.0f);
#pragma unroll
                    for (int c = 0; c < VEC; ++c) gi.v[c] = 0.f;
                    cur_i = i;
                }
            }
            if (neg) {
#pragma unroll
                for (int r = 0; r < R; ++r) t_idx[r + 1] = active ? (uint32_t)neg[e * R + r] : 0u;
            } else if (LANES >= NCALL) {
                // lane c of the group evaluates call c (same counters as the loop version: identical draws)
                const Philox4 w = philox4x32_10((uint32_t)p, (uint32_t)(gl % NCALL), epoch, STREAM_NEG, k0, k1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t mine_w = (r & 3) == 0 ? w.x : (r & 3) == 1 ? w.y : (r & 3) == 2 ? w.z : w.w;
                    t_idx[r + 1] = urange(__shfl_sync(0xffffffffu, mine_w, r >> 2, LANES), rep_count);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NCALL; ++c) {
                    const Philox4 w = philox4x32_10((uint32_t)p, (uint32_t)c, epoch, STREAM_NEG, k0, k1);
                    if (4 * c + 0 < R) t_idx[4 * c + 1] = urange(w.x, rep_count);
                    if (4 * c + 1 < R) t_idx[4 * c + 2] = urange(w.y, rep_count);
                    if (4 * c + 2 < R) t_idx[4 * c + 3] = urange(w.z, rep_count);
                    if (4 * c + 3 < R) t_idx[4 * c + 4] = urange(w.w, rep_count);
                }
            }
            // which pairs this launch handles (all of them without windows)
            bool inw[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) inw[q] = active && (!WIN || (t_idx[q] - win_lo) < win_len);
            // gather: head row + the tail rows, then differences and squared distances
            const Vec<VEC> yi = load_vec<VEC>(head + (int64_t)i * DIM + gl * VEC);
            Vec<VEC> df[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (WIN) {
                    if (inw[q]) df[q] = load_vec_cg<VEC>(tail + (int64_t)t_idx[q] * DIM + gl * VEC);
                    else {
#pragma unroll
                        for (int c = 0; c < VEC; ++c) df[q].v[c] = 0.f;
                    }
                } else {
                    df[q] = load_vec<VEC>(tail + (int64_t)t_idx[q] * DIM + gl * VEC);
                }
            }
            float s[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < VEC; ++c) { df[q].v[c] = yi.v[c] - df[q].v[c]; acc = fmaf(df[q].v[c], df[q].v[c], acc); }
                s[q] = group_sum<LANES>(acc);
            }
            // coefficients: pair q is evaluated by lane q % LANES
            float cv[ROUNDS];
#pragma unroll
            for (int u0 = 0; u0 < ROUNDS; ++u0) {
                float sv = 1.0f;
                bool valid = false;
#pragma unroll
                for (int u = 0; u < LANES; ++u)
                    if (u0 * LANES + u < NP && gl == u) { sv = s[u0 * LANES + u]; valid = inw[u0 * LANES + u]; }
                float l = 0.f;
                const float cf = pair_coef<FAST>(sv, u0 == 0 && gl == 0, a, b, sc_a, sc_r, want_loss, l);
                cv[u0] = valid ? cf : 0.f;
                if (want_loss && valid) loss_acc += l;
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const float coef = LANES == 1 ? cv[q] : __shfl_sync(0xffffffffu, cv[q / LANES], q % LANES, LANES);
                Vec<VEC> g;
#pragma unroll
                for (int c = 0; c < VEC; ++c) { g.v[c] = coef * df[q].v[c]; gi.v[c] += g.v[c]; }
                if (inw[q] && grad_tail) red_vec<VEC>(grad_tail + (int64_t)t_idx[q] * DIM + gl * VEC, g, -1.0f);
            }
        }
        __syncthreads();                                              // the buffer is refilled two rounds later
    }
    if (cur_i >= 0) red_vec<VEC>(grad_head + (int64_t)cur_i * DIM + gl * VEC, gi, 1.0f);
    if (want_loss) {
        loss_acc = warp_sum(loss_acc);
        if ((threadIdx.x & 31) == 0 && loss_acc != 0.f) atomicAdd(loss_out, loss_acc);
    }
}

// generic dimension: one warp per kept edge, lane owns components lane, lane+32, ... (dim <= 128)
__global__ void __launch_bounds__(256)
edge_forces_generic_kernel(const int4 *__restrict__ kept_rec, const int32_t *__restrict__ kept_hdr,
                           const int32_t *__restrict__ neg, const int32_t *__restrict__ batch_kept,
                           int n_batches, int num_rep, uint32_t rep_count,
                           const float *__restrict__ head, const float *__restrict__ tail,
                           float *__restrict__ grad_head, float *__restrict__ grad_tail, int dim, float a,
                           float b, uint64_t seed, const OptState *__restrict__ st, float *__restrict__ loss_out) {
    const int n_kept = kept_total(kept_hdr);
    const uint32_t epoch = st->epoch;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_nb = 1.0f / (float)n_batches;
    float loss_acc = 0.f;
    for (int64_t e = wid; e < n_kept; e += n_warps) {
        const int4 rec = kept_rec[e];
        const int32_t p = rec.x, i = rec.y, j = rec.z;
        float kb = (float)batch_kept[rec.w];
        float sc_a = inv_nb / kb, sc_r = inv_nb / (kb * (float)num_rep);
        float yi[4], gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) { int cc = lane + 32 * c; yi[c] = cc < dim ? head[(int64_t)i * dim + cc] : 0.f; }
        Philox4 rnd = {0, 0, 0, 0};
        for (int r = -1; r < num_rep; ++r) {
            uint32_t t_idx;
            if (r < 0) t_idx = (uint32_t)j;
            else if (neg) t_idx = (uint32_t)neg[e * num_rep + r];
            else {
                if ((r & 3) == 0) rnd = philox4x32_10((uint32_t)p, (uint32_t)(r >> 2), epoch, STREAM_NEG, k0, k1);
                uint32_t x = (r & 3) == 0 ? rnd.x : (r & 3) == 1 ? rnd.y : (r & 3) == 2 ? rnd.z : rnd.w;
                t_idx = urange(x, rep_count);
            }
            float df[4], s = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int cc = lane + 32 * c;
                float yt = cc < dim ? tail[(int64_t)t_idx * dim + cc] : 0.f;
                df[c] = yi[c] - yt;
                s = fmaf(df[c], df[c], s);
            }
            s = warp_sum(s);
            float coef, l;
            if (r < 0) { attr_terms(s, a, b, coef, l); coef *= sc_a; l *= sc_a; }
            else { rep_terms(s, a, b, coef, l); coef *= sc_r; l *= sc_r; }
            if (lane == 0) loss_acc += l;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int cc = lane + 32 * c;
                float g = coef * df[c];
                gi[c] += g;
                if (cc < dim && grad_tail) red_add_f32(grad_tail + (int64_t)t_idx * dim + cc, -g);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { int cc = lane + 32 * c; if (cc < dim) red_add_f32(grad_head + (int64_t)i * dim + cc, gi[c]); }
    }
    if (loss_out && lane == 0 && loss_acc != 0.f) atomicAdd(loss_out, loss_acc);
}

// ------------------------------------------------------------------ K7c: invert-mode forces
// ref: model.py:336-362 (_inv_attr_loss, _inv_rep_loss) as used by _train in mode "invert"
// (model.py:437,447).  The variable is a Q x D table in DATA space (D up to thousands), the tails
// are rows of the fitted modality's data with their fit-time sigma / rho.  One warp per kept edge:
// phase 1 computes the 1+R squared distances with all 1+R tail rows streamed together (float4 per lane when
// D % 4 == 0: 1+R independent 16-byte loads in flight per lane), phase 2 forms sum_p coef_p (x - y_p) per
// component -- the rows are L1/L2 hits now -- and issues ONE red per component (vector red when D % 4 == 0).
constexpr int INV_MAXP = 17;      // 1 attractive + up to 16 repulsive pairs

template <int NPT>                // compile-time pair count (1 + num_rep) or 0 = run-time (<= INV_MAXP)
__global__ void __launch_bounds__(256)
invert_forces_kernel(const int4 *__restrict__ kept_rec, const int32_t *__restrict__ kept_hdr,
                     const int32_t *__restrict__ neg, const int32_t *__restrict__ batch_kept, int n_batches,
                     int num_rep, uint32_t rep_count, const float *__restrict__ head,
                     const float *__restrict__ data, const float *__restrict__ sigma, const float *__restrict__ rho,
                     float *__restrict__ grad_head, int dim, float a, float b, uint64_t seed,
                     const OptState *__restrict__ st, float *__restrict__ loss_out) {
    constexpr int MAXP = NPT ? NPT : INV_MAXP;
    const int n_kept = kept_total(kept_hdr);
    const uint32_t epoch = st->epoch;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_nb = 1.0f / (float)n_batches;
    const int np = NPT ? NPT : 1 + num_rep;
    const bool vec4 = (dim & 3) == 0;
    float loss_acc = 0.f;
    for (int64_t e = wid; e < n_kept; e += n_warps) {
        const int4 rec = kept_rec[e];
        const int32_t p = rec.x, i = rec.y;
        const float kb = (float)batch_kept[rec.w];
        const float sc_a = inv_nb / kb, sc_r = inv_nb / (kb * (float)num_rep);
        const float *xi = head + (int64_t)i * dim;
        uint32_t t_idx[MAXP];
        float coef[MAXP], s_raw[MAXP];
        Philox4 rnd = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {
            t_idx[q] = 0; coef[q] = 0.f; s_raw[q] = 0.f;
            if (q >= np) continue;
            if (q == 0) t_idx[q] = (uint32_t)rec.z;
            else if (neg) t_idx[q] = (uint32_t)neg[e * num_rep + (q - 1)];
            else {
                const int r = q - 1;
                if ((r & 3) == 0) rnd = philox4x32_10((uint32_t)p, (uint32_t)(r >> 2), epoch, STREAM_NEG, k0, k1);
                const uint32_t x = (r & 3) == 0 ? rnd.x : (r & 3) == 1 ? rnd.y : (r & 3) == 2 ? rnd.z : rnd.w;
                t_idx[q] = urange(x, rep_count);
            }
        }
        // phase 1: squared distances; per-pair accumulation order = components ascending within a lane, then the
        // warp tree -- the same for the scalar and the float4 layout of the lane's components is NOT required
        // (the test tolerance is relative 1e-4), so each path uses its natural order
        if (vec4) {
            for (int c = lane * 4; c < dim; c += 128) {
                const float4 x = *reinterpret_cast<const float4 *>(xi + c);
#pragma unroll
                for (int q = 0; q < MAXP; ++q) {
                    if (q >= np) continue;
                    const float4 y = __ldg(reinterpret_cast<const float4 *>(data + (int64_t)t_idx[q] * dim + c));
                    float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
                    s_raw[q] = fmaf(d0, d0, s_raw[q]); s_raw[q] = fmaf(d1, d1, s_raw[q]);
                    s_raw[q] = fmaf(d2, d2, s_raw[q]); s_raw[q] = fmaf(d3, d3, s_raw[q]);
                }
            }
        } else {
            for (int c = lane; c < dim; c += 32) {
                const float x = xi[c];
#pragma unroll
                for (int q = 0; q < MAXP; ++q) {
                    if (q >= np) continue;
                    const float df = x - data[(int64_t)t_idx[q] * dim + c];
                    s_raw[q] = fmaf(df, df, s_raw[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {
            if (q >= np) continue;
            const float sr = warp_sum(s_raw[q]);
            const float s = fmaxf(sr, 1e-6f);
            const float dist = sqrtf(s);
            const float sg = sigma[t_idx[q]];
            float dls, l;
            if (q == 0) {
                const float sb = powf(s, b);
                const float w = 1.0f / (1.0f + a * sb);
                const float u = w * sg + 1e-6f;
                l = sc_a * dist / u;
                dls = sc_a * (1.0f / (2.0f * dist * u) + dist * sg * a * b * (sb / s) * w * w / (u * u));
            } else {
                const float c_raw = dist - rho[t_idx[q]];
                const float ex = expf(-fmaxf(c_raw, 1e-6f) / (sg + 1e-6f));
                const float om = 1.0f - ex + 1e-6f;
                l = -sc_r * logf(om);
                dls = (c_raw >= 1e-6f) ? sc_r * (-ex / (sg + 1e-6f)) / (om * 2.0f * dist) : 0.f;
            }
            coef[q] = (sr >= 1e-6f) ? 2.0f * dls : 0.f;
            if (lane == 0) loss_acc += l;
        }
        // phase 2
        if (vec4) {
            for (int c = lane * 4; c < dim; c += 128) {
                const float4 x = *reinterpret_cast<const float4 *>(xi + c);
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
                for (int q = 0; q < MAXP; ++q) {
                    if (q >= np) continue;
                    const float4 y = __ldg(reinterpret_cast<const float4 *>(data + (int64_t)t_idx[q] * dim + c));
                    g0 = fmaf(coef[q], x.x - y.x, g0); g1 = fmaf(coef[q], x.y - y.y, g1);
                    g2 = fmaf(coef[q], x.z - y.z, g2); g3 = fmaf(coef[q], x.w - y.w, g3);
                }
                red_add_v4(grad_head + (int64_t)i * dim + c, g0, g1, g2, g3);
            }
        } else {
            for (int c = lane; c < dim; c += 32) {
                const float x = xi[c];
                float g = 0.f;
#pragma unroll
                for (int q = 0; q < MAXP; ++q)
                    if (q < np) g = fmaf(coef[q], x - data[(int64_t)t_idx[q] * dim + c], g);
                red_add_f32(grad_head + (int64_t)i * dim + c, g);
            }
        }
    }
    if (loss_out && lane == 0 && loss_acc != 0.f) atomicAdd(loss_out, loss_acc);
}

// ------------------------------------------------------------------ K8: InfoNCE
constexpr int NCE_MAX = 16;   // 1 positive + up to 15 negatives

// Device sample stream: the reference re-draws randperm(num) every epoch (model.py:373), so which anchors fall into
// the short last chunk (and get the larger per-anchor weight 1/(clen*n_chunks), model.py:392-394) changes from epoch
// to epoch.  The device stream keeps position order but rotates it by a per-epoch, per-direction random offset:
// anchor position t holds row (t + offset) mod num, so the short chunk is a different row range every epoch and all
// rows are weighted equally in expectation.  (The host stream uploads the reference's own permutation instead.)
__device__ __forceinline__ int32_t nce_rotated_anchor(int64_t t, int64_t num, uint32_t epoch, uint32_t stream_id, uint64_t seed) {
    const Philox4 r = philox4x32_10(0xffffffffu, 0xffffffffu, epoch, STREAM_INFONCE + stream_id, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    const int64_t off = (int64_t)urange(r.x, (uint32_t)num);
    const int64_t v = t + off;
    return (int32_t)(v >= num ? v - num : v);
}

__global__ void __launch_bounds__(128)
infonce_kernel(const float *__restrict__ e0_, const float *__restrict__ e1_, int64_t num, int64_t a_lo, int64_t a_hi, int dim,
               const int32_t *__restrict__ perm_, const int32_t *__restrict__ neg_, const int32_t *__restrict__ perm_rev,
               const int32_t *__restrict__ neg_rev, int n_neg, int chunk,
               float weight, float temperature, float *__restrict__ grad0_, float *__restrict__ grad1_,
               uint64_t seed, uint32_t stream_id_, const OptState *__restrict__ st, float *__restrict__ loss_out) {
    const bool rev = blockIdx.y == 1;
    const float *__restrict__ e0 = rev ? e1_ : e0_;
    const float *__restrict__ e1 = rev ? e0_ : e1_;
    float *__restrict__ grad0 = rev ? grad1_ : grad0_;
    float *__restrict__ grad1 = rev ? grad0_ : grad1_;
    const int32_t *__restrict__ perm = rev ? perm_rev : perm_;
    const int32_t *__restrict__ neg = rev ? neg_rev : neg_;
    const uint32_t stream_id = stream_id_ + (rev ? 1u : 0u);
    const int64_t t = a_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float loss_local = 0.f;
    if (t < a_hi) {
        const int64_t n_chunks = (num + chunk - 1) / chunk;
        const int64_t cidx = t / chunk;
        const int64_t clen = min((int64_t)chunk, num - cidx * chunk);
        const float wgt = weight / ((float)clen * (float)n_chunks);       // ref: model.py:392,394
        const int32_t i = perm ? perm[t] : nce_rotated_anchor(t, num, st->epoch, stream_id, seed);
        const int M = 1 + n_neg;
        int32_t ids[NCE_MAX];
        bool ok[NCE_MAX];
        float nrm[NCE_MAX], cs[NCE_MAX];
        ids[0] = i; ok[0] = true;
        if (neg) {
            for (int m = 1; m < M; ++m) ids[m] = neg[t * n_neg + (m - 1)];
        } else {
            const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
            Philox4 rnd = {0, 0, 0, 0};
            for (int m = 1; m < M; ++m) {
                int q = m - 1;
                if ((q & 3) == 0) rnd = philox4x32_10((uint32_t)t, (uint32_t)(q >> 2), st->epoch, STREAM_INFONCE + stream_id, k0, k1);
                uint32_t x = (q & 3) == 0 ? rnd.x : (q & 3) == 1 ? rnd.y : (q & 3) == 2 ? rnd.z : rnd.w;
                ids[m] = (int32_t)urange(x, (uint32_t)num);
            }
        }
        for (int m = 1; m < M; ++m) ok[m] = ids[m] != i;                     // ref: model.py:386
        const float *ap = e0 + (int64_t)i * dim;
        float na2 = 0.f;
        for (int c = 0; c < dim; ++c) na2 = fmaf(ap[c], ap[c], na2);
        const float na = fmaxf(sqrtf(na2), 1e-12f);                          // F.normalize eps
        float mx = -__int_as_float(0x7f800000);
        for (int m = 0; m < M; ++m) {
            nrm[m] = 1.f; cs[m] = 0.f;
            if (!ok[m]) continue;
            const float *ep = e1 + (int64_t)ids[m] * dim;
            float n2 = 0.f, dt = 0.f;
            for (int c = 0; c < dim; ++c) { float v = ep[c]; n2 = fmaf(v, v, n2); dt = fmaf(ap[c], v, dt); }
            nrm[m] = fmaxf(sqrtf(n2), 1e-12f);
            cs[m] = dt / (na * nrm[m]);                                      // cosine similarity
            mx = fmaxf(mx, cs[m] / temperature);
        }
        float den = 0.f;
        for (int m = 0; m < M; ++m) if (ok[m]) den += expf(cs[m] / temperature - mx);
        loss_local = wgt * -(cs[0] / temperature - mx - logf(den));          // -log_softmax[:,0]
        // c_m = softmax_m - delta_m0 ; sum_m c_m * cos_m
        float cm[NCE_MAX];
        float ccs = 0.f;
        for (int m = 0; m < M; ++m) {
            cm[m] = ok[m] ? expf(cs[m] / temperature - mx) / den - (m == 0 ? 1.f : 0.f) : 0.f;
            ccs = fmaf(cm[m], cs[m], ccs);
        }
        const float sa = wgt / (temperature * na);
        for (int c = 0; c < dim; ++c) {
            const float ac = ap[c];
            const float uc = ac / na;
            float acc = 0.f;
            for (int m = 0; m < M; ++m) {
                if (!ok[m]) continue;
                float vc = e1[(int64_t)ids[m] * dim + c] / nrm[m];
                acc = fmaf(cm[m], vc, acc);
                // d/d e_m = w c_m (u - v_m (v_m.u)) / (tau |e_m|)
                float gm = wgt * cm[m] * (uc - vc * cs[m]) / (temperature * nrm[m]);
                red_add_f32(grad1 + (int64_t)ids[m] * dim + c, gm);
            }
            // d/d a = w (sum_m c_m v_m - u sum_m c_m cos_m) / (tau |a|)
            red_add_f32(grad0 + (int64_t)i * dim + c, sa * (acc - uc * ccs));
        }
    }
    if (loss_out) {
        loss_local = warp_sum(loss_local);
        if ((threadIdx.x & 31) == 0 && loss_local != 0.f) atomicAdd(loss_out, loss_local);
    }
}

// K8, vectorised variant: a group of LANES lanes per anchor, each lane owning VEC consecutive
// components (dim == VEC*LANES): the anchor row and its 1+n_neg candidate rows are gathered with
// 16-byte loads, norms and dot products are group reductions, and the gradients leave as vector
// red.global.add -- the same closed form as infonce_kernel (ref: model.py:364-394).
template <int VEC, int LANES, int MT>
__global__ void __launch_bounds__(256)
infonce_vec_kernel(const float *__restrict__ e0_, const float *__restrict__ e1_, int64_t num, int64_t a_lo, int64_t a_hi,
                   const int32_t *__restrict__ perm_, const int32_t *__restrict__ neg_, const int32_t *__restrict__ perm_rev,
                   const int32_t *__restrict__ neg_rev, int n_neg, int chunk,
                   float weight, float temperature, float *__restrict__ grad0_, float *__restrict__ grad1_,
                   uint64_t seed, uint32_t stream_id_, const OptState *__restrict__ st, float *__restrict__ loss_out) {
    // blockIdx.y == 1: the reverse direction (anchors in e1, candidates in e0) of the same pair; both
    // read the same embedding state, so one grid evaluates L_ij + L_ji (model.py:467-472)
    const bool rev = blockIdx.y == 1;
    const float *__restrict__ e0 = rev ? e1_ : e0_;
    const float *__restrict__ e1 = rev ? e0_ : e1_;
    float *__restrict__ grad0 = rev ? grad1_ : grad0_;
    float *__restrict__ grad1 = rev ? grad0_ : grad1_;
    const int32_t *__restrict__ perm = rev ? perm_rev : perm_;
    const int32_t *__restrict__ neg = rev ? neg_rev : neg_;
    const uint32_t stream_id = stream_id_ + (rev ? 1u : 0u);
    constexpr int DIM = VEC * LANES;
    const int gl = threadIdx.x % LANES;
    const int64_t t = a_lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const bool active = t < a_hi;
    const int64_t tt = active ? t : a_lo;
    const int64_t n_chunks = (num + chunk - 1) / chunk;
    const int64_t cidx = tt / chunk;
    const int64_t clen = min((int64_t)chunk, num - cidx * chunk);
    const float wgt = weight / ((float)clen * (float)n_chunks);           // ref: model.py:392,394
    const int32_t i = perm ? perm[tt] : nce_rotated_anchor(tt, num, st->epoch, stream_id, seed);
    const int M = MT ? MT : 1 + n_neg;
    constexpr int MCAP = MT ? MT : NCE_MAX;
    int32_t ids[MCAP];
    ids[0] = i;
    if (neg) {
#pragma unroll
        for (int m = 1; m < MCAP; ++m)
            if (m < M) ids[m] = neg[tt * n_neg + (m - 1)];
    } else {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        constexpr int NCALL = (MCAP - 1 + 3) / 4;
        if (MT && LANES >= NCALL) {
            // lane c of the group evaluates Philox call c (same counters, same draws as the serial form)
            const Philox4 w = philox4x32_10((uint32_t)tt, (uint32_t)(gl % NCALL), st->epoch, STREAM_INFONCE + stream_id, k0, k1);
#pragma unroll
            for (int m = 1; m < MCAP; ++m) {
                const int q = m - 1;
                const uint32_t mine = (q & 3) == 0 ? w.x : (q & 3) == 1 ? w.y : (q & 3) == 2 ? w.z : w.w;
                ids[m] = (int32_t)urange(__shfl_sync(0xffffffffu, mine, q >> 2, LANES), (uint32_t)num);
            }
        } else {
            Philox4 rnd = {0, 0, 0, 0};
#pragma unroll
            for (int m = 1; m < MCAP; ++m) {
                if (m >= M) continue;
                int q = m - 1;
                if ((q & 3) == 0) rnd = philox4x32_10((uint32_t)tt, (uint32_t)(q >> 2), st->epoch, STREAM_INFONCE + stream_id, k0, k1);
                uint32_t x = (q & 3) == 0 ? rnd.x : (q & 3) == 1 ? rnd.y : (q & 3) == 2 ? rnd.z : rnd.w;
                ids[m] = (int32_t)urange(x, (uint32_t)num);
            }
        }
    }
    const Vec<VEC> av = load_vec<VEC>(e0 + (int64_t)i * DIM + gl * VEC);
    // with a compile-time candidate count all rows are gathered first (independent loads in flight)
    Vec<VEC> evr[MT ? MT : 1];
    if (MT) {
#pragma unroll
        for (int m = 0; m < MT; ++m) evr[m] = load_vec<VEC>(e1 + (int64_t)ids[m] * DIM + gl * VEC);
    }
    float na2 = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) na2 = fmaf(av.v[c], av.v[c], na2);
    const float na = fmaxf(sqrtf(group_sum<LANES>(na2)), 1e-12f);       // F.normalize eps
    // one reciprocal per quantity instead of a division per use (the kernel is instruction bound)
    const float inv_t = 1.0f / temperature;
    const float inv_na = 1.0f / na;
    float inv_nrm[MCAP], cs[MCAP], ex[MCAP];
    float mx = -__int_as_float(0x7f800000);
#pragma unroll
    for (int m = 0; m < MCAP; ++m) {
        if (m >= M) continue;
        const bool ok = m == 0 || ids[m] != i;                           // ref: model.py:386
        const Vec<VEC> ev = MT ? evr[MT ? m : 0] : load_vec<VEC>(e1 + (int64_t)ids[m] * DIM + gl * VEC);
        float n2 = 0.f, dt = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) { n2 = fmaf(ev.v[c], ev.v[c], n2); dt = fmaf(av.v[c], ev.v[c], dt); }
        n2 = group_sum<LANES>(n2);
        dt = group_sum<LANES>(dt);
        inv_nrm[m] = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
        cs[m] = dt * inv_na * inv_nrm[m];
        if (ok) mx = fmaxf(mx, cs[m] * inv_t);
    }
    float den = 0.f;
#pragma unroll
    for (int m = 0; m < MCAP; ++m)
        if (m < M) {
            ex[m] = (m == 0 || ids[m] != i) ? expf(fmaf(cs[m], inv_t, -mx)) : 0.f;
            den += ex[m];
        }
    const float inv_den = 1.0f / den;
    float loss_local = (active && gl == 0) ? wgt * -(fmaf(cs[0], inv_t, -mx) - logf(den)) : 0.f;
    float ccs = 0.f;
    Vec<VEC> acc;
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc.v[c] = 0.f;
#pragma unroll
    for (int m = 0; m < MCAP; ++m) {
        if (m >= M) continue;
        const bool ok = m == 0 || ids[m] != i;
        if (!ok) continue;                                               // group-uniform
        const float cm = ex[m] * inv_den - (m == 0 ? 1.f : 0.f);
        ccs = fmaf(cm, cs[m], ccs);
        const Vec<VEC> ev = MT ? evr[MT ? m : 0] : load_vec<VEC>(e1 + (int64_t)ids[m] * DIM + gl * VEC);
        const float inv_n = inv_nrm[m];
        const float sm = wgt * cm * inv_t * inv_n;
        Vec<VEC> g;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const float vc = ev.v[c] * inv_n, uc = av.v[c] * inv_na;
            acc.v[c] = fmaf(cm, vc, acc.v[c]);
            g.v[c] = sm * (uc - vc * cs[m]);                             // d/d e_m = w c_m (u - v_m cos_m)/(tau |e_m|)
        }
        if (active) red_vec<VEC>(grad1 + (int64_t)ids[m] * DIM + gl * VEC, g, 1.0f);
    }
    const float sa = wgt * inv_t * inv_na;
    Vec<VEC> ga;
#pragma unroll
    for (int c = 0; c < VEC; ++c) ga.v[c] = sa * (acc.v[c] - av.v[c] * inv_na * ccs);   // d/d a
    if (active) red_vec<VEC>(grad0 + (int64_t)i * DIM + gl * VEC, ga, 1.0f);
    if (loss_out) {
        loss_local = warp_sum(loss_local);
        if ((threadIdx.x & 31) == 0 && loss_local != 0.f) atomicAdd(loss_out, loss_local);
    }
}

// ------------------------------------------------------------------ K9: Adam
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
            int64_t n, float beta2, float omb1, float omb2, float eps, const OptState *__restrict__ st,
            int zero_grad) {
    // torch.optim.Adam (single-tensor path) operation by operation, one rounding each: no fma
    // contraction, so that the result equals the CPU reference bit for bit.
    const float neg_step = -st->step_size, bc2_sqrt = st->bc2_sqrt;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {
        mm = __fadd_rn(mm, __fmul_rn(omb1, __fsub_rn(gg, mm)));                        // exp_avg.lerp_(grad, 1-beta1)
        vv = __fadd_rn(__fmul_rn(vv, beta2), __fmul_rn(__fmul_rn(omb2, gg), gg));      // mul_(beta2).addcmul_(g, g, 1-beta2)
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        pp = __fadd_rn(pp, __fdiv_rn(__fmul_rn(neg_step, mm), denom));                 // addcdiv_(exp_avg, denom, -step_size)
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                           reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const int64_t n4 = aligned ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], gg = reinterpret_cast<float4 *>(g)[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        if (zero_grad) reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        upd(pp, g[i], mm, vv);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (zero_grad) g[i] = 0.f;
    }
}

static inline unsigned persistent_blocks(int threads, int per_sm) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    (void)threads;
    return (unsigned)(sms * per_sm);
}

}  // namespace mmu

extern "C" int mmu_opt_state_init(uint32_t *state, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(state, "mmu_opt_state_init: null pointer");
    opt_state_init_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<OptState *>(state));
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_opt_state_advance(uint32_t *state, double lr, double beta1, double beta2, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(state, "mmu_opt_state_advance: null pointer");
    opt_state_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<OptState *>(state), lr, beta1, beta2);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_edge_sample_at(const int32_t *row, const int32_t *col, const float *w, int64_t edge_lo,
                                  int64_t edge_hi, int batch_size, int n_batches, uint64_t seed, int64_t epoch,
                                  const uint32_t *state, int32_t *kept_rec, int32_t *kept_hdr, int32_t *batch_kept,
                                  mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(row && col && w && state && kept_rec && kept_hdr && batch_kept, "mmu_edge_sample: null pointer");
    MMU_CHECK_ARG((reinterpret_cast<uintptr_t>(kept_rec) & 15) == 0, "mmu_edge_sample: kept_rec must be 16-byte aligned");
    MMU_CHECK_ARG(batch_size >= 1 && n_batches >= 1, "mmu_edge_sample: bad batch geometry");
    MMU_CHECK_ARG(edge_lo >= 0 && edge_hi >= edge_lo && edge_hi < ((int64_t)1 << 31), "mmu_edge_sample: bad edge range");
    cudaStream_t st = as_stream(stream);
    MMU_CUDA(cudaMemsetAsync(kept_hdr, 0, sizeof(int32_t), st));            // the count; capacity and flag stay
    MMU_CUDA(cudaMemsetAsync(batch_kept, 0, sizeof(int32_t) * (size_t)n_batches, st));
    if (edge_hi == edge_lo) return MMU_OK;
    int64_t n4 = (edge_hi + 3) / 4 - edge_lo / 4;
    int64_t want = (n4 + 255) / 256;
    unsigned cap = persistent_blocks(256, 8);
    unsigned blocks = (unsigned)(want < (int64_t)cap ? want : cap);
    edge_sample_kernel<<<blocks, 256, 0, st>>>(row, col, w, edge_lo, edge_hi, batch_size, seed,
                                               reinterpret_cast<const OptState *>(state),
                                               reinterpret_cast<int4 *>(kept_rec), kept_hdr, batch_kept, epoch);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_edge_sample_range(const int32_t *row, const int32_t *col, const float *w, int64_t edge_lo,
                                     int64_t edge_hi, int batch_size, int n_batches, uint64_t seed,
                                     const uint32_t *state, int32_t *kept_rec, int32_t *kept_hdr, int32_t *batch_kept,
                                     mmu_stream_t stream) {
    return mmu_edge_sample_at(row, col, w, edge_lo, edge_hi, batch_size, n_batches, seed, -1, state, kept_rec, kept_hdr,
                              batch_kept, stream);
}

extern "C" int mmu_edge_records(const int32_t *row, const int32_t *col, const int32_t *kept_pos, int64_t n_kept,
                                int batch_size, int32_t *kept_rec, int32_t *kept_hdr, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(row && col && kept_pos && kept_rec && kept_hdr, "mmu_edge_records: null pointer");
    MMU_CHECK_ARG((reinterpret_cast<uintptr_t>(kept_rec) & 15) == 0, "mmu_edge_records: kept_rec must be 16-byte aligned");
    MMU_CHECK_ARG(n_kept >= 0 && n_kept < ((int64_t)1 << 31) && batch_size >= 1, "mmu_edge_records: bad sizes");
    const unsigned blocks = (unsigned)((n_kept + 255) / 256 > 0 ? (n_kept + 255) / 256 : 1);
    edge_records_kernel<<<blocks, 256, 0, as_stream(stream)>>>(row, col, kept_pos, n_kept, batch_size,
                                                               reinterpret_cast<int4 *>(kept_rec), kept_hdr);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_edge_forces(const int32_t *kept_rec, const int32_t *kept_hdr, const int32_t *neg,
                               const int32_t *batch_kept, int n_batches, int num_rep, int64_t rep_count,
                               const float *head, const float *tail, float *grad_head, float *grad_tail, int dim,
                               float a, float b, uint64_t seed, const uint32_t *state, float *loss, int fast_math,
                               int64_t window_rows, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(kept_rec && kept_hdr && batch_kept && head && tail && grad_head && state, "mmu_edge_forces: null pointer");
    MMU_CHECK_ARG((reinterpret_cast<uintptr_t>(kept_rec) & 15) == 0, "mmu_edge_forces: kept_rec must be 16-byte aligned");
    MMU_CHECK_ARG(dim >= 1 && dim <= 128, "mmu_edge_forces: dim=%d outside [1,128]", dim);
    MMU_CHECK_ARG(num_rep >= 0 && rep_count >= 1 && rep_count < ((int64_t)1 << 31), "mmu_edge_forces: bad negatives");
    MMU_CHECK_ARG(n_batches >= 1 && window_rows >= 0, "mmu_edge_forces: bad batch geometry / window");
    cudaStream_t st = as_stream(stream);
    const OptState *os = reinterpret_cast<const OptState *>(state);
    const int4 *rec = reinterpret_cast<const int4 *>(kept_rec);
    const bool vec_dim = dim == 2 || dim == 4 || dim == 8 || dim == 16 || dim == 32 || dim == 64 || dim == 128;
    const bool staged = vec_dim && (num_rep == 8 || num_rep == 4) && option(OPT_FORCE_STAGED) != 0;
    if (!staged) {
        // loop form (any num_rep; option force_staged = 0) and the generic-dimension kernel: one pass, no windows
        const unsigned blocks = persistent_blocks(256, 8);
#define MMU_FORCES(V, L)                                                                                          \
    edge_forces_kernel<V, L><<<blocks, 256, 0, st>>>(rec, kept_hdr, neg, batch_kept, n_batches, num_rep,          \
                                                     (uint32_t)rep_count, head, tail, grad_head, grad_tail, dim,  \
                                                     a, b, seed, os, loss)
        switch (dim) {
            case 2: MMU_FORCES(2, 1); break;
            case 4: MMU_FORCES(4, 1); break;
            case 8: MMU_FORCES(4, 2); break;
            case 16: MMU_FORCES(4, 4); break;
            case 32: MMU_FORCES(4, 8); break;
            case 64: MMU_FORCES(4, 16); break;
            case 128: MMU_FORCES(4, 32); break;
            default:
                edge_forces_generic_kernel<<<blocks, 256, 0, st>>>(rec, kept_hdr, neg, batch_kept, n_batches, num_rep,
                                                                   (uint32_t)rep_count, head, tail, grad_head, grad_tail,
                                                                   dim, a, b, seed, os, loss);
        }
#undef MMU_FORCES
        note_kernel(SITE_EDGE_FORCES, vec_dim ? "edge_forces_kernel<dim=%d>(num_rep=%d)" : "edge_forces_generic_kernel(dim=%d,num_rep=%d)",
                    dim, num_rep);
        MMU_LAUNCH_CHECK();
        return MMU_OK;
    }
    // staged run form: a grid of whole waves at the kernel's occupancy (__launch_bounds__)
    const bool windows = window_rows > 0 && window_rows < rep_count;
    const unsigned blocks = persistent_blocks(256, staged_blocks_per_sm(dim <= 4 ? 1 : 2, windows));
    const int n_windows = windows ? (int)((rep_count + window_rows - 1) / window_rows) : 1;
#define MMU_STAGED(V, L, RR, FASTV, WINV)                                                                        \
    edge_forces_staged_kernel<V, L, RR, FASTV, WINV><<<blocks, 256, 0, st>>>(                                     \
        rec, kept_hdr, neg, batch_kept, n_batches, (uint32_t)rep_count, head, tail, grad_head, grad_tail, a, b,  \
        seed, os, loss, win_lo, win_len)
#define MMU_STAGED_R(V, L, RR)                                         \
    do {                                                               \
        if (fast_math && windows) MMU_STAGED(V, L, RR, true, true);    \
        else if (fast_math) MMU_STAGED(V, L, RR, true, false);         \
        else if (windows) MMU_STAGED(V, L, RR, false, true);           \
        else MMU_STAGED(V, L, RR, false, false);                       \
    } while (0)
#define MMU_STAGED_DIM(V, L)                          \
    do {                                              \
        if (num_rep == 8) MMU_STAGED_R(V, L, 8);      \
        else MMU_STAGED_R(V, L, 4);                   \
    } while (0)
    for (int wdx = 0; wdx < n_windows; ++wdx) {
        const uint32_t win_lo = windows ? (uint32_t)(wdx * window_rows) : 0u;
        const uint32_t win_len = windows ? (uint32_t)(rep_count - win_lo < window_rows ? rep_count - win_lo : window_rows)
                                         : (uint32_t)rep_count;
        switch (dim) {
            case 2: MMU_STAGED_DIM(2, 1); break;
            case 4: MMU_STAGED_DIM(4, 1); break;
            case 8: MMU_STAGED_DIM(4, 2); break;
            case 16: MMU_STAGED_DIM(4, 4); break;
            case 32: MMU_STAGED_DIM(4, 8); break;
            case 64: MMU_STAGED_DIM(4, 16); break;
            default: MMU_STAGED_DIM(4, 32); break;
        }
    }
#undef MMU_STAGED_DIM
#undef MMU_STAGED_R
#undef MMU_STAGED
    note_kernel(SITE_EDGE_FORCES, "edge_forces_staged_kernel<dim=%d,R=%d,%s,%s>%s", dim, num_rep, fast_math ? "fast" : "ieee",
                windows ? "windowed" : "one-pass", windows ? " x tail windows" : "");
    MMU_LAUNCH_CHECK_N(n_windows);
    return MMU_OK;
}

extern "C" int mmu_invert_forces(const int32_t *kept_rec, const int32_t *kept_hdr, const int32_t *neg,
                                 const int32_t *batch_kept, int n_batches, int num_rep, int64_t rep_count,
                                 const float *head, const float *data, const float *sigma, const float *rho,
                                 float *grad_head, int dim, float a, float b, uint64_t seed, const uint32_t *state,
                                 float *loss, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(kept_rec && kept_hdr && batch_kept && head && data && sigma && rho && grad_head && state,
                  "mmu_invert_forces: null pointer");
    MMU_CHECK_ARG(dim >= 1, "mmu_invert_forces: bad dim");
    MMU_CHECK_ARG(num_rep >= 0 && num_rep < INV_MAXP, "mmu_invert_forces: num_rep=%d outside [0,%d)", num_rep, INV_MAXP);
    MMU_CHECK_ARG(rep_count >= 1 && rep_count < ((int64_t)1 << 31), "mmu_invert_forces: bad rep_count");
    MMU_CHECK_ARG(n_batches >= 1, "mmu_invert_forces: bad batch geometry");
    MMU_CHECK_ARG((dim & 3) != 0 || ((reinterpret_cast<uintptr_t>(head) | reinterpret_cast<uintptr_t>(data) |
                                      reinterpret_cast<uintptr_t>(grad_head)) & 15) == 0,
                  "mmu_invert_forces: tables must be 16-byte aligned when dim is a multiple of 4");
    const int4 *rec = reinterpret_cast<const int4 *>(kept_rec);
    const OptState *os = reinterpret_cast<const OptState *>(state);
    const unsigned blocks = persistent_blocks(256, 8);
    if (num_rep == 8)
        invert_forces_kernel<9><<<blocks, 256, 0, as_stream(stream)>>>(rec, kept_hdr, neg, batch_kept, n_batches, num_rep,
                                                                       (uint32_t)rep_count, head, data, sigma, rho,
                                                                       grad_head, dim, a, b, seed, os, loss);
    else
        invert_forces_kernel<0><<<blocks, 256, 0, as_stream(stream)>>>(rec, kept_hdr, neg, batch_kept, n_batches, num_rep,
                                                                       (uint32_t)rep_count, head, data, sigma, rho,
                                                                       grad_head, dim, a, b, seed, os, loss);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_infonce_range(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                                 int dim, const int32_t *perm, const int32_t *neg, int n_neg, int chunk, float weight,
                                 float temperature, float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                                 const uint32_t *state, float *loss, mmu_stream_t stream);

extern "C" int mmu_infonce(const float *e0, const float *e1, int64_t num, int dim, const int32_t *perm,
                           const int32_t *neg, int n_neg, int chunk, float weight, float temperature, float *grad0,
                           float *grad1, uint64_t seed, uint32_t stream_id, const uint32_t *state, float *loss,
                           mmu_stream_t stream) {
    return mmu_infonce_range(e0, e1, num, 0, num, dim, perm, neg, n_neg, chunk, weight, temperature, grad0, grad1, seed,
                             stream_id, state, loss, stream);
}

static int infonce_launch(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                          int dim, const int32_t *perm, const int32_t *neg, const int32_t *perm_rev,
                          const int32_t *neg_rev, int directions, int n_neg, int chunk, float weight,
                          float temperature, float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                          const uint32_t *state, float *loss, mmu_stream_t stream);

extern "C" int mmu_infonce_range(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                                 int dim, const int32_t *perm, const int32_t *neg, int n_neg, int chunk, float weight,
                                 float temperature, float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                                 const uint32_t *state, float *loss, mmu_stream_t stream) {
    return infonce_launch(e0, e1, num, anchor_lo, anchor_hi, dim, perm, neg, nullptr, nullptr, 1, n_neg, chunk, weight,
                          temperature, grad0, grad1, seed, stream_id, state, loss, stream);
}

extern "C" int mmu_infonce_bidir(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                                 int dim, const int32_t *perm_fwd, const int32_t *neg_fwd, const int32_t *perm_rev,
                                 const int32_t *neg_rev, int n_neg, int chunk, float weight, float temperature,
                                 float *grad0, float *grad1, uint64_t seed, uint32_t stream_id, const uint32_t *state,
                                 float *loss, mmu_stream_t stream) {
    return infonce_launch(e0, e1, num, anchor_lo, anchor_hi, dim, perm_fwd, neg_fwd, perm_rev, neg_rev, 2, n_neg, chunk,
                          weight, temperature, grad0, grad1, seed, stream_id, state, loss, stream);
}

static int infonce_launch(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                          int dim, const int32_t *perm, const int32_t *neg, const int32_t *perm_rev,
                          const int32_t *neg_rev, int directions, int n_neg, int chunk, float weight,
                          float temperature, float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                          const uint32_t *state, float *loss, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(anchor_lo >= 0 && anchor_lo <= anchor_hi && anchor_hi <= num, "mmu_infonce: bad anchor range");
    MMU_CHECK_ARG(e0 && e1 && grad0 && grad1 && state, "mmu_infonce: null pointer");
    MMU_CHECK_ARG(n_neg >= 0 && n_neg < NCE_MAX, "mmu_infonce: n_neg=%d outside [0,%d)", n_neg, NCE_MAX);
    MMU_CHECK_ARG(dim >= 1 && chunk >= 1 && temperature > 0.f, "mmu_infonce: bad dim/chunk/temperature");
    MMU_CHECK_ARG(num >= 0 && num < ((int64_t)1 << 31), "mmu_infonce: num must be < 2^31");
    if (anchor_hi == anchor_lo) return MMU_OK;
    const OptState *os = reinterpret_cast<const OptState *>(state);
    const int64_t cnt = anchor_hi - anchor_lo;
#define MMU_NCE_M(V, L, MTV)                                                                                       \
    infonce_vec_kernel<V, L, MTV><<<dim3((unsigned)((cnt * L + 255) / 256), directions), 256, 0, as_stream(stream)>>>( \
        e0, e1, num, anchor_lo, anchor_hi, perm, neg, perm_rev, neg_rev, n_neg, chunk, weight, temperature, grad0, \
        grad1, seed, stream_id, os, loss)
#define MMU_NCE(V, L)                                  \
    do {                                               \
        if (n_neg == 9) MMU_NCE_M(V, L, 10);           \
        else MMU_NCE_M(V, L, 0);                       \
    } while (0)
    switch (dim) {
        case 4: MMU_NCE(4, 1); break;
        case 8: MMU_NCE(4, 2); break;
        case 16: MMU_NCE(4, 4); break;
        case 32: MMU_NCE(4, 8); break;
        case 64: MMU_NCE(4, 16); break;
        case 128: MMU_NCE(4, 32); break;
        default:
            infonce_kernel<<<dim3((unsigned)((cnt + 127) / 128), directions), 128, 0, as_stream(stream)>>>(
                e0, e1, num, anchor_lo, anchor_hi, dim, perm, neg, perm_rev, neg_rev, n_neg, chunk, weight, temperature,
                grad0, grad1, seed, stream_id, os, loss);
    }
#undef MMU_NCE
#undef MMU_NCE_M
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_adam_step(float *p, float *g, float *m, float *v, int64_t n, double beta1, double beta2, double eps,
                             const uint32_t *state, int zero_grad, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(p && g && m && v && state, "mmu_adam_step: null pointer");
    if (n == 0) return MMU_OK;
    int64_t want = (n / 4 + 255) / 256 + 1;
    unsigned cap = persistent_blocks(256, 16);
    unsigned blocks = (unsigned)(want < (int64_t)cap ? want : cap);
    // 1-beta is formed in double and rounded once, as torch does with its Python-float betas
    adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, (float)beta2, (float)(1.0 - beta1),
                                                       (float)(1.0 - beta2), (float)eps,
                                                       reinterpret_cast<const OptState *>(state), zero_grad);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
