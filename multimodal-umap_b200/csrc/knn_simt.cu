// knn_simt.cu -- exhaustive fp32 kNN on CUDA cores (exact by construction) + list merge.
//
// Replaces the reference's NN-descent candidate search and per-row top-k
// (/root/reference/impl/model.py:81-195) with an exhaustive evaluation of its own distance
// expression (model.py:109,163) under the canonical accumulation order of
// oracle/knn_oracle.c: acc = fmaf(x[t]-y[t], x[t]-y[t], acc), t ascending; dist = sqrtf(acc);
// rank by (dist, index).
//
// Role on B200: (a) the fallback for rows the tcgen05 candidate path cannot certify,
// (b) the on-device reference the tensor-core path is checked against.  A 128x64 tile of
// pairs per CTA, 8x4 pairs per thread, operands staged through shared memory in 32-wide
// K slabs; per-row sorted top-k lists live in shared memory and tile results are filtered
// by one compare against the row's current k-th best before a warp-cooperative insert.
#include "common.cuh"

namespace mmu {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BN = 64;    // db rows per tile
constexpr int BK = 32;    // K slab
constexpr int QS = BQ + 4;
constexpr int DS = BN + 4;
constexpr int KNN_THREADS = 256;

// shared memory: [ staging (Qs, Ds)  |aliased with|  dist tile BQ x BN ]  +  lists BQ x k u64
constexpr int STAGE_FLOATS = BK * QS + BK * DS;
constexpr int TILE_FLOATS = BQ * (BN + 1);
constexpr int REGION_A_FLOATS = (STAGE_FLOATS > TILE_FLOATS) ? STAGE_FLOATS : TILE_FLOATS;

__device__ __forceinline__ void warp_insert(uint64_t *list, int k, uint64_t key, int lane) {
    // list sorted ascending, key < list[k-1].  Lanes own positions lane and lane+32.
    uint64_t a = (lane < k) ? list[lane] : ~0ull;
    uint64_t b = (lane + 32 < k) ? list[lane + 32] : ~0ull;
    int ins = __popc(__ballot_sync(0xffffffffu, a < key)) + __popc(__ballot_sync(0xffffffffu, b < key));
    // shift right by one from ins
    uint64_t pa = (lane > 0 && lane - 1 < k) ? list[lane - 1] : 0ull;             // predecessor of pos lane
    uint64_t pb = (lane + 31 < k) ? list[lane + 31] : 0ull;                        // predecessor of pos lane+32
    __syncwarp();
    if (lane < k && lane > ins) list[lane] = pa;
    if (lane + 32 < k && lane + 32 > ins) list[lane + 32] = pb;
    if (lane == 0) list[ins] = key;
    __syncwarp();
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_exact_f32_kernel(const float *__restrict__ query, int64_t n_query, const int32_t *__restrict__ query_ids,
                     const float *__restrict__ db, int64_t n_db, int dim, int k, int exclude_self,
                     int64_t q_base, int64_t db_base, int merge_existing,
                     int32_t *__restrict__ out_idx, float *__restrict__ out_dist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *regionA = reinterpret_cast<float *>(smem_raw);
    float *Qs = regionA;                 // [BK][QS]
    float *Ds = regionA + BK * QS;       // [BK][DS]
    float *tile = regionA;               // [BQ][BN+1]  (aliases the staging area)
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem_raw + sizeof(float) * REGION_A_FLOATS);   // [BQ][k]
    __shared__ int64_t rowid[BQ];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4;             // 0..15 -> query rows ty*8 .. +7
    const int tx = tid & 15;             // 0..15 -> db cols   tx*4 .. +3
    const int64_t q0 = (int64_t)blockIdx.x * BQ;

    for (int r = tid; r < BQ; r += KNN_THREADS) {
        int64_t t = q0 + r;
        rowid[r] = (t < n_query) ? (query_ids ? (int64_t)query_ids[t] : t) : -1;
    }
    __syncthreads();
    for (int e = tid; e < BQ * k; e += KNN_THREADS) {
        int r = e / k, p = e - r * k;
        uint64_t key = MMU_KEY_EMPTY;
        if (merge_existing && rowid[r] >= 0) {
            int32_t id = out_idx[rowid[r] * k + p];
            if (id >= 0) key = dist_key(out_dist[rowid[r] * k + p], id);
        }
        lists[e] = key;
    }

    for (int64_t n0 = 0; n0 < n_db; n0 += BN) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

        for (int k0 = 0; k0 < dim; k0 += BK) {
            __syncthreads();   // previous slab / previous tile's selection done with region A
            // stage Q slab: BQ rows x BK dims, transposed into Qs[t][row]
            for (int e = tid; e < BQ * BK; e += KNN_THREADS) {
                int r = e / BK, t = e - r * BK;
                float v = 0.0f;
                if (rowid[r] >= 0 && k0 + t < dim) v = query[rowid[r] * (int64_t)dim + k0 + t];
                Qs[t * QS + r] = v;
            }
            for (int e = tid; e < BN * BK; e += KNN_THREADS) {
                int r = e / BK, t = e - r * BK;
                float v = 0.0f;
                if (n0 + r < n_db && k0 + t < dim) v = db[(n0 + r) * (int64_t)dim + k0 + t];
                Ds[t * DS + r] = v;
            }
            __syncthreads();
#pragma unroll 4
            for (int t = 0; t < BK; ++t) {
                float4 qa = *reinterpret_cast<const float4 *>(&Qs[t * QS + ty * 8]);
                float4 qb = *reinterpret_cast<const float4 *>(&Qs[t * QS + ty * 8 + 4]);
                float4 dv = *reinterpret_cast<const float4 *>(&Ds[t * DS + tx * 4]);
                float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float diff = qv[i] - dd[j];
                        acc[i][j] = fmaf(diff, diff, acc[i][j]);
                    }
            }
        }
        __syncthreads();   // all threads done reading staging before the tile overwrites it
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tile[(ty * 8 + i) * (BN + 1) + tx * 4 + j] = sqrtf(acc[i][j]);
        __syncthreads();

        // selection: warp w owns rows w*16 .. +15
        for (int rr = 0; rr < BQ / 8; ++rr) {
            int r = warp * (BQ / 8) + rr;
            if (rowid[r] < 0) continue;                       // warp-uniform
            uint64_t *list = lists + r * k;
            int64_t qglob = rowid[r] + q_base;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int c = lane + 32 * h;
                int64_t j = n0 + c;
                uint64_t key = MMU_KEY_EMPTY;
                if (j < n_db && !(exclude_self && (j + db_base) == qglob))
                    key = dist_key(tile[r * (BN + 1) + c], (int32_t)(j + db_base));
                uint64_t worst = list[k - 1];
                unsigned pass = __ballot_sync(0xffffffffu, key < worst);
                while (pass) {
                    int src = __ffs(pass) - 1;
                    pass &= pass - 1;
                    uint64_t cand = __shfl_sync(0xffffffffu, key, src);
                    if (cand < list[k - 1]) warp_insert(list, k, cand, lane);
                }
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < BQ * k; e += KNN_THREADS) {
        int r = e / k, p = e - r * k;
        if (rowid[r] < 0) continue;
        uint64_t key = lists[e];
        bool empty = (key == MMU_KEY_EMPTY);
        out_idx[rowid[r] * k + p] = empty ? -1 : key_idx(key);
        out_dist[rowid[r] * k + p] = empty ? __int_as_float(0x7f800000) : key_dist(key);
    }
}

// K3: two-pointer merge of two sorted lists per row
__global__ void knn_merge_kernel(const int32_t *__restrict__ ia, const float *__restrict__ da,
                                 const int32_t *__restrict__ ib, const float *__restrict__ db_,
                                 int64_t n_rows, int k, int32_t *__restrict__ oi, float *__restrict__ od) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t *pa = ia + r * k, *pb = ib + r * k;
    const float *qa = da + r * k, *qb = db_ + r * k;
    int a = 0, b = 0;
    for (int o = 0; o < k; ++o) {
        uint64_t ka = (a < k && pa[a] >= 0) ? dist_key(qa[a], pa[a]) : MMU_KEY_EMPTY;
        uint64_t kb = (b < k && pb[b] >= 0) ? dist_key(qb[b], pb[b]) : MMU_KEY_EMPTY;
        uint64_t key;
        if (ka <= kb) { key = ka; ++a; if (ka == kb && ka != MMU_KEY_EMPTY) ++b; }   // same point in both lists
        else { key = kb; ++b; }
        bool empty = (key == MMU_KEY_EMPTY);
        oi[r * k + o] = empty ? -1 : key_idx(key);
        od[r * k + o] = empty ? __int_as_float(0x7f800000) : key_dist(key);
    }
}

}  // namespace mmu

extern "C" int mmu_knn_exact_f32(const float *query, int64_t n_query, const int32_t *query_ids,
                                 const float *db, int64_t n_db, int dim, int k, int exclude_self,
                                 int64_t query_index_base, int64_t db_index_base, int merge_existing,
                                 int32_t *out_idx, float *out_dist, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(query && db && out_idx && out_dist, "mmu_knn_exact_f32: null pointer");
    MMU_CHECK_ARG(k >= 1 && k <= MMU_MAX_K, "mmu_knn_exact_f32: k=%d outside [1,%d]", k, MMU_MAX_K);
    MMU_CHECK_ARG(dim >= 1 && n_db >= 0 && n_query >= 0, "mmu_knn_exact_f32: bad sizes");
    MMU_CHECK_ARG(n_db + db_index_base < (int64_t)2147483647, "mmu_knn_exact_f32: db index exceeds int32");
    if (n_query == 0) return MMU_OK;
    size_t smem = sizeof(float) * REGION_A_FLOATS + sizeof(uint64_t) * (size_t)BQ * k;
    if (first_use_on_device(SITE_KNN_EXACT))      // the attribute is per device
        MMU_CUDA(cudaFuncSetAttribute(knn_exact_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    int64_t grid = (n_query + BQ - 1) / BQ;
    knn_exact_f32_kernel<<<(unsigned)grid, KNN_THREADS, smem, as_stream(stream)>>>(
        query, n_query, query_ids, db, n_db, dim, k, exclude_self, query_index_base, db_index_base,
        merge_existing, out_idx, out_dist);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_knn_merge(const int32_t *idx_a, const float *dist_a, const int32_t *idx_b,
                             const float *dist_b, int64_t n_rows, int k, int32_t *out_idx,
                             float *out_dist, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(idx_a && dist_a && idx_b && dist_b && out_idx && out_dist, "mmu_knn_merge: null pointer");
    MMU_CHECK_ARG(k >= 1 && k <= 1024, "mmu_knn_merge: bad k");
    MMU_CHECK_ARG(out_idx != idx_a && out_idx != idx_b, "mmu_knn_merge: output must not alias an input");
    if (n_rows == 0) return MMU_OK;
    int threads = 128;
    knn_merge_kernel<<<(unsigned)((n_rows + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
        idx_a, dist_a, idx_b, dist_b, n_rows, k, out_idx, out_dist);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
