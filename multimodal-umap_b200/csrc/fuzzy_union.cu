// fuzzy_union.cu -- K5: S = G + G^T - G*G^T as a radix-sort COO merge.
//
// ref: /root/reference/impl/model.py:271  (graph + graph.T - graph*graph.T).coalesce()
//
// G arrives as the fixed-degree graph of K4: n rows x k entries, columns ascending per row, so
// enumerating its entries in memory order gives them sorted by (src,dst).  A STABLE LSD radix
// sort of those entries on the dst key alone therefore yields G^T in CSR order (dst-major,
// src ascending).  Every output row r is then the sorted union of two sorted lists
//   A = G[r,:] (k entries)    and    B = G^T[r,:] (indeg(r) entries),
// merged by rank arithmetic (no per-row sort), with fl(fl(a+b)-fl(a*b)) on mutual edges.
// All passes stream 4-byte keys/payloads with coalesced reads; indices are bit-exact.
#include "common.cuh"

namespace mmu {

// ------------------------------------------------------------------ generic exclusive scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                         // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

template <typename TI, typename TO>
__global__ void scan_tile_sums(const TI *__restrict__ in, int64_t n, TO *__restrict__ partial) {
    __shared__ TO red[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    TO s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t e = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (e < n) s += (TO)in[e];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        TO t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

template <typename TO>
__global__ void scan_partials(TO *partial, int64_t m) {   // single block, in-place exclusive
    __shared__ TO buf[1024];
    __shared__ TO carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += 1024) {
        int64_t e = base + threadIdx.x;
        TO v = (e < m) ? partial[e] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            TO t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        TO incl = buf[threadIdx.x];
        if (e < m) partial[e] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

// out[e] = exclusive prefix; if total_slot >= 0, out[total_slot] = grand total (n+1 arrays)
template <typename TI, typename TO>
__global__ void scan_apply(const TI *__restrict__ in, int64_t n, const TO *__restrict__ partial,
                           TO *__restrict__ out, int write_total) {
    __shared__ TO warp_tot[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    // blocked arrangement: thread t owns items base + t*ITEMS .. +ITEMS-1
    TO v[SCAN_ITEMS];
    TO s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t e = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        v[i] = (e < n) ? (TO)in[e] : 0;
        s += v[i];
    }
    TO incl = s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        TO t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    TO woff = 0;
    for (int i = 0; i < w; ++i) woff += warp_tot[i];
    TO run = partial[blockIdx.x] + woff + incl - s;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t e = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (e < n) out[e] = run;
        run += v[i];
        if (write_total && e == n - 1) out[n] = run;
    }
}

template <typename TI, typename TO>
static int exclusive_scan(const TI *in, int64_t n, TO *out, TO *partial, int write_total, cudaStream_t st) {
    if (n == 0) return MMU_OK;
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<TI, TO><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, partial);
    scan_partials<TO><<<1, 1024, 0, st>>>(partial, tiles);
    scan_apply<TI, TO><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, partial, out, write_total);
    MMU_LAUNCH_CHECK_N(3);
    return MMU_OK;
}

// ------------------------------------------------------------------ stable LSD radix pass
constexpr int RS_WARPS = 4;          // warps per CTA, one tile per warp
constexpr int RS_TILE = 2048;        // items per warp tile
constexpr int RS_MAX_BINS = 2048;

// hist[bin * n_tiles + tile]
__global__ void __launch_bounds__(RS_WARPS * 32)
radix_hist_kernel(const int32_t *__restrict__ key, int64_t n_items, int shift, int bins, int64_t n_tiles,
                  uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t cnt[];     // [RS_WARPS][bins]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * RS_WARPS + w;
    uint32_t *c = cnt + w * bins;
    for (int b = lane; b < bins; b += 32) c[b] = 0;
    __syncwarp();
    if (tile < n_tiles) {
        int64_t base = tile * RS_TILE;
        for (int i = 0; i < RS_TILE; i += 32) {
            int64_t e = base + i + lane;
            if (e < n_items) atomicAdd(&c[((uint32_t)key[e] >> shift) & (bins - 1)], 1u);
        }
        __syncwarp();
        for (int b = lane; b < bins; b += 32) hist[(int64_t)b * n_tiles + tile] = c[b];
    }
}

// FIRST: payload src is implicit (e / k) and keys/weights come from the graph arrays
template <bool FIRST>
__global__ void __launch_bounds__(RS_WARPS * 32)
radix_scatter_kernel(const int32_t *__restrict__ key_in, const int32_t *__restrict__ src_in,
                     const float *__restrict__ w_in, int64_t n_items, int k, int shift, int bins,
                     int64_t n_tiles, const uint32_t *__restrict__ offs, int32_t *__restrict__ key_out,
                     int32_t *__restrict__ src_out, float *__restrict__ w_out) {
    extern __shared__ uint32_t cnt[];     // [RS_WARPS][bins] running output cursors
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * RS_WARPS + w;
    if (tile >= n_tiles) return;
    uint32_t *c = cnt + w * bins;
    for (int b = lane; b < bins; b += 32) c[b] = offs[(int64_t)b * n_tiles + tile];
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    int64_t base = tile * RS_TILE;
    for (int i = 0; i < RS_TILE; i += 32) {
        int64_t e = base + i + lane;
        bool valid = e < n_items;
        unsigned act = __ballot_sync(0xffffffffu, valid);
        if (!act) break;
        int32_t kk = valid ? key_in[e] : 0;
        uint32_t d = valid ? (((uint32_t)kk >> shift) & (bins - 1)) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t pos = 0;
        if (valid) pos = c[d] + __popc(peers & lt);
        __syncwarp();
        if (valid && (peers & lt) == 0) c[d] += __popc(peers);     // group leader advances the cursor
        __syncwarp();
        if (valid) {
            key_out[pos] = kk;
            src_out[pos] = FIRST ? (int32_t)(e / k) : src_in[e];
            w_out[pos] = w_in[e];
        }
    }
}

// ------------------------------------------------------------------ in-degree histogram
__global__ void indeg_kernel(const int32_t *__restrict__ col, int64_t n_items, uint32_t *__restrict__ indeg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_items) atomicAdd(&indeg[col[e]], 1u);
}

__device__ __forceinline__ int lower_bound_g(const int32_t *__restrict__ a, int n, int32_t x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per row: out_cnt[r] = k + indeg - |A n B|
__global__ void __launch_bounds__(256)
union_count_kernel(const int32_t *__restrict__ col, int64_t n, int k, const uint32_t *__restrict__ tptr,
                   const int32_t *__restrict__ tsrc, uint32_t *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t b0 = tptr[r], b1 = tptr[r + 1];
    const int nb = (int)(b1 - b0);
    const int32_t *B = tsrc + b0;
    int hit = 0;
    for (int h = 0; h < 2; ++h) {
        int i = lane + 32 * h;
        if (i < k) {
            int32_t a = col[r * k + i];
            int p = lower_bound_g(B, nb, a);
            hit += (p < nb && B[p] == a);
        }
    }
    for (int o = 16; o > 0; o >>= 1) hit += __shfl_xor_sync(0xffffffffu, hit, o);
    if (lane == 0) out_cnt[r] = (uint32_t)(k + nb - hit);
}

__global__ void __launch_bounds__(256)
union_write_kernel(const int32_t *__restrict__ col, const float *__restrict__ w, int64_t n, int k,
                   const uint32_t *__restrict__ tptr, const int32_t *__restrict__ tsrc,
                   const float *__restrict__ tw, const int64_t *__restrict__ out_rowptr,
                   int32_t *__restrict__ out_row, int32_t *__restrict__ out_col, float *__restrict__ out_val,
                   int32_t row_base) {
    __shared__ int32_t As[8][MMU_MAX_K];
    __shared__ int32_t Pf[8][MMU_MAX_K + 1];     // Pf[i] = #flagged A elements with index < i
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t b0 = tptr[r], b1 = tptr[r + 1];
    const int nb = (int)(b1 - b0);
    const int32_t *B = tsrc + b0;
    const float *Bw = tw + b0;
    const int64_t obase = out_rowptr[r];
    const unsigned lt = (1u << lane) - 1u;

    int32_t a[2];
    float wa[2];
    int lb[2];
    bool fl[2];
    for (int h = 0; h < 2; ++h) {
        int i = lane + 32 * h;
        a[h] = 0x7fffffff; wa[h] = 0.f; lb[h] = 0; fl[h] = false;
        if (i < k) {
            a[h] = col[r * k + i];
            wa[h] = w[r * k + i];
            lb[h] = lower_bound_g(B, nb, a[h]);
            fl[h] = (lb[h] < nb && B[lb[h]] == a[h]);
            As[wl][i] = a[h];
        }
    }
    unsigned m0 = __ballot_sync(0xffffffffu, fl[0]);
    unsigned m1 = __ballot_sync(0xffffffffu, fl[1]);
    int before[2] = {__popc(m0 & lt), __popc(m0) + __popc(m1 & lt)};
    for (int h = 0; h < 2; ++h) {
        int i = lane + 32 * h;
        if (i < k) Pf[wl][i] = before[h];
    }
    if (lane == 0) Pf[wl][k] = __popc(m0) + __popc(m1);
    __syncwarp();
    // A side
    for (int h = 0; h < 2; ++h) {
        int i = lane + 32 * h;
        if (i < k) {
            int64_t pos = obase + i + lb[h] - before[h];
            float v = wa[h];
            if (fl[h]) {
                float wb = Bw[lb[h]];
                v = __fsub_rn(__fadd_rn(wa[h], wb), __fmul_rn(wa[h], wb));   // ref: model.py:271
            }
            out_row[pos] = (int32_t)r + row_base;
            out_col[pos] = a[h];
            out_val[pos] = v;
        }
    }
    // B side: entries not in A
    for (int j = lane; j < nb; j += 32) {
        int32_t b = B[j];
        int lo = 0, hi = k;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (As[wl][mid] < b) lo = mid + 1; else hi = mid;
        }
        if (lo < k && As[wl][lo] == b) continue;
        int64_t pos = obase + lo + j - Pf[wl][lo];
        out_row[pos] = (int32_t)r + row_base;
        out_col[pos] = b;
        out_val[pos] = Bw[j];
    }
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct UnionPlan {
    int64_t items, n_tiles;
    int bits, passes, width, bins;
    size_t off_key[2], off_src[2], off_w[2], off_hist, off_partial, off_indeg, off_tptr, off_cnt, total;
};

static UnionPlan make_plan_items(int64_t n, int64_t items);
static UnionPlan make_plan(int64_t n, int k) { return make_plan_items(n, n * (int64_t)k); }
static UnionPlan make_plan_items(int64_t n, int64_t items) {
    UnionPlan p;
    p.items = items > 0 ? items : 1;
    p.n_tiles = (p.items + RS_TILE - 1) / RS_TILE;
    p.bits = 1;
    while (((int64_t)1 << p.bits) < n) ++p.bits;
    p.passes = (p.bits + 10) / 11;
    p.width = (p.bits + p.passes - 1) / p.passes;
    p.bins = 1 << p.width;
    size_t o = 0;
    for (int i = 0; i < 2; ++i) {
        p.off_key[i] = o; o += align_up(sizeof(int32_t) * p.items);
        p.off_src[i] = o; o += align_up(sizeof(int32_t) * p.items);
        p.off_w[i] = o; o += align_up(sizeof(float) * p.items);
    }
    p.off_hist = o; o += align_up(sizeof(uint32_t) * (size_t)p.bins * p.n_tiles);
    size_t scan_len = (size_t)p.bins * p.n_tiles;
    if ((size_t)n + 1 > scan_len) scan_len = (size_t)n + 1;
    p.off_partial = o; o += align_up(sizeof(int64_t) * ((scan_len + SCAN_TILE - 1) / SCAN_TILE + 1));
    p.off_indeg = o; o += align_up(sizeof(uint32_t) * (n + 1));
    p.off_tptr = o; o += align_up(sizeof(uint32_t) * (n + 1));
    p.off_cnt = o; o += align_up(sizeof(uint32_t) * (n + 1));
    p.total = o;
    return p;
}

}  // namespace mmu

extern "C" size_t mmu_union_workspace_bytes(int64_t n, int k) {
    if (n <= 0 || k <= 0) return 0;
    return mmu::make_plan(n, k).total;
}

extern "C" int mmu_fuzzy_union(const int32_t *col, const float *w, int64_t n, int k, void *workspace,
                               size_t workspace_bytes, int64_t *out_rowptr, int32_t *out_row,
                               int32_t *out_col, float *out_val, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(col && w && workspace && out_rowptr && out_row && out_col && out_val, "mmu_fuzzy_union: null pointer");
    MMU_CHECK_ARG(k >= 1 && k <= MMU_MAX_K, "mmu_fuzzy_union: k=%d outside [1,%d]", k, MMU_MAX_K);
    MMU_CHECK_ARG(n >= 1 && n * (int64_t)k < (int64_t)1 << 31, "mmu_fuzzy_union: n*k must be < 2^31");
    UnionPlan p = make_plan(n, k);
    MMU_CHECK_ARG(workspace_bytes >= p.total, "mmu_fuzzy_union: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    cudaStream_t st = as_stream(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *key[2] = {reinterpret_cast<int32_t *>(ws + p.off_key[0]), reinterpret_cast<int32_t *>(ws + p.off_key[1])};
    int32_t *src[2] = {reinterpret_cast<int32_t *>(ws + p.off_src[0]), reinterpret_cast<int32_t *>(ws + p.off_src[1])};
    float *wv[2] = {reinterpret_cast<float *>(ws + p.off_w[0]), reinterpret_cast<float *>(ws + p.off_w[1])};
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + p.off_hist);
    uint32_t *partial32 = reinterpret_cast<uint32_t *>(ws + p.off_partial);
    int64_t *partial64 = reinterpret_cast<int64_t *>(ws + p.off_partial);
    uint32_t *indeg = reinterpret_cast<uint32_t *>(ws + p.off_indeg);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(ws + p.off_tptr);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(ws + p.off_cnt);

    // 1. G^T by stable LSD radix sort on the dst key
    unsigned rs_blocks = (unsigned)((p.n_tiles + RS_WARPS - 1) / RS_WARPS);
    size_t rs_smem = sizeof(uint32_t) * RS_WARPS * p.bins;
    const int32_t *kin = col;
    const int32_t *sin = nullptr;
    const float *win = w;
    int cur = 0;
    for (int pass = 0; pass < p.passes; ++pass) {
        int shift = pass * p.width;
        radix_hist_kernel<<<rs_blocks, RS_WARPS * 32, rs_smem, st>>>(kin, p.items, shift, p.bins, p.n_tiles, hist);
        int rc = exclusive_scan<uint32_t, uint32_t>(hist, (int64_t)p.bins * p.n_tiles, hist, partial32, 0, st);
        if (rc) return rc;
        if (pass == 0)
            radix_scatter_kernel<true><<<rs_blocks, RS_WARPS * 32, rs_smem, st>>>(
                kin, sin, win, p.items, k, shift, p.bins, p.n_tiles, hist, key[cur], src[cur], wv[cur]);
        else
            radix_scatter_kernel<false><<<rs_blocks, RS_WARPS * 32, rs_smem, st>>>(
                kin, sin, win, p.items, k, shift, p.bins, p.n_tiles, hist, key[cur], src[cur], wv[cur]);
        MMU_LAUNCH_CHECK_N(2);
        kin = key[cur]; sin = src[cur]; win = wv[cur];
        cur ^= 1;
    }
    const int32_t *tsrc = sin;
    const float *tw = win;

    // 2. CSR offsets of G^T from the in-degree histogram
    MMU_CUDA(cudaMemsetAsync(indeg, 0, sizeof(uint32_t) * (n + 1), st));
    indeg_kernel<<<(unsigned)((p.items + 255) / 256), 256, 0, st>>>(col, p.items, indeg);
    int rc = exclusive_scan<uint32_t, uint32_t>(indeg, n, tptr, partial32, 1, st);
    if (rc) return rc;

    // 3. output row sizes, 4. offsets, 5. merge
    unsigned wblocks = (unsigned)((n * 32 + 255) / 256);
    union_count_kernel<<<wblocks, 256, 0, st>>>(col, n, k, tptr, tsrc, cnt);
    rc = exclusive_scan<uint32_t, int64_t>(cnt, n, out_rowptr, partial64, 1, st);
    if (rc) return rc;
    union_write_kernel<<<wblocks, 256, 0, st>>>(col, w, n, k, tptr, tsrc, tw, out_rowptr, out_row, out_col, out_val, 0);
    MMU_LAUNCH_CHECK_N(3);
    return MMU_OK;
}

// ------------------------------------------------------------------ row block of the union (multi-GPU)
// Rows [row_base, row_base + n_rows) of S = G + G^T - G*G^T from (a) the block's own rows of G (col_block / w_block,
// fixed degree k) and (b) the in-edges of the block: the entries (src, dst, w) of the WHOLE graph with dst inside the block,
// in src-major order, given as in_key = dst - row_base, in_src, in_w (n_in of them; the host filters them out of the
// replicated kNN result).  Same sort + merge as mmu_fuzzy_union on 1/W of the entries; the ranks' blocks are all-gathered.
extern "C" size_t mmu_union_rows_workspace_bytes(int64_t n_rows, int64_t n_in) {
    if (n_rows <= 0) return 0;
    return mmu::make_plan_items(n_rows, n_in).total;
}

extern "C" int mmu_fuzzy_union_rows(const int32_t *col_block, const float *w_block, int64_t n_rows, int k, const int32_t *in_key,
                                    const int32_t *in_src, const float *in_w, int64_t n_in, int64_t row_base, void *workspace,
                                    size_t workspace_bytes, int64_t *out_rowptr, int32_t *out_row, int32_t *out_col,
                                    float *out_val, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(col_block && w_block && workspace && out_rowptr && out_row && out_col && out_val, "mmu_fuzzy_union_rows: null pointer");
    MMU_CHECK_ARG(n_in == 0 || (in_key && in_src && in_w), "mmu_fuzzy_union_rows: null in-edge arrays");
    MMU_CHECK_ARG(k >= 1 && k <= MMU_MAX_K, "mmu_fuzzy_union_rows: k=%d outside [1,%d]", k, MMU_MAX_K);
    MMU_CHECK_ARG(n_rows >= 1 && n_in >= 0 && n_in < (int64_t)1 << 31 && row_base >= 0 && row_base + n_rows < (int64_t)1 << 31,
                  "mmu_fuzzy_union_rows: bad sizes");
    UnionPlan p = make_plan_items(n_rows, n_in);
    MMU_CHECK_ARG(workspace_bytes >= p.total, "mmu_fuzzy_union_rows: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    cudaStream_t st = as_stream(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *key[2] = {reinterpret_cast<int32_t *>(ws + p.off_key[0]), reinterpret_cast<int32_t *>(ws + p.off_key[1])};
    int32_t *src[2] = {reinterpret_cast<int32_t *>(ws + p.off_src[0]), reinterpret_cast<int32_t *>(ws + p.off_src[1])};
    float *wv[2] = {reinterpret_cast<float *>(ws + p.off_w[0]), reinterpret_cast<float *>(ws + p.off_w[1])};
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + p.off_hist);
    uint32_t *partial32 = reinterpret_cast<uint32_t *>(ws + p.off_partial);
    int64_t *partial64 = reinterpret_cast<int64_t *>(ws + p.off_partial);
    uint32_t *indeg = reinterpret_cast<uint32_t *>(ws + p.off_indeg);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(ws + p.off_tptr);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(ws + p.off_cnt);
    const int32_t *kin = in_key;
    const int32_t *sin = in_src;
    const float *win = in_w;
    if (n_in > 0) {
        unsigned rs_blocks = (unsigned)((p.n_tiles + RS_WARPS - 1) / RS_WARPS);
        size_t rs_smem = sizeof(uint32_t) * RS_WARPS * p.bins;
        int cur = 0;
        for (int pass = 0; pass < p.passes; ++pass) {
            int shift = pass * p.width;
            radix_hist_kernel<<<rs_blocks, RS_WARPS * 32, rs_smem, st>>>(kin, n_in, shift, p.bins, p.n_tiles, hist);
            int rc = exclusive_scan<uint32_t, uint32_t>(hist, (int64_t)p.bins * p.n_tiles, hist, partial32, 0, st);
            if (rc) return rc;
            radix_scatter_kernel<false><<<rs_blocks, RS_WARPS * 32, rs_smem, st>>>(kin, sin, win, n_in, k, shift, p.bins, p.n_tiles,
                                                                                  hist, key[cur], src[cur], wv[cur]);
            MMU_LAUNCH_CHECK_N(2);
            kin = key[cur]; sin = src[cur]; win = wv[cur];
            cur ^= 1;
        }
    }
    MMU_CUDA(cudaMemsetAsync(indeg, 0, sizeof(uint32_t) * (n_rows + 1), st));
    if (n_in > 0) indeg_kernel<<<(unsigned)((n_in + 255) / 256), 256, 0, st>>>(in_key, n_in, indeg);
    int rc = exclusive_scan<uint32_t, uint32_t>(indeg, n_rows, tptr, partial32, 1, st);
    if (rc) return rc;
    unsigned wblocks = (unsigned)((n_rows * 32 + 255) / 256);
    union_count_kernel<<<wblocks, 256, 0, st>>>(col_block, n_rows, k, tptr, sin ? sin : col_block, cnt);
    rc = exclusive_scan<uint32_t, int64_t>(cnt, n_rows, out_rowptr, partial64, 1, st);
    if (rc) return rc;
    union_write_kernel<<<wblocks, 256, 0, st>>>(col_block, w_block, n_rows, k, tptr, sin ? sin : col_block, win ? win : w_block,
                                                out_rowptr, out_row, out_col, out_val, (int32_t)row_base);
    MMU_LAUNCH_CHECK_N(3);
    return MMU_OK;
}
