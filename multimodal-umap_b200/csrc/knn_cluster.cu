// knn_cluster.cu -- cluster structure for the pruned exact kNN search (umap_b200/knn_pruned.py).
//
// Role in the path: /root/reference/impl/model.py:81-195 searches neighbours by 2-hop expansion of a random graph; the
// engine searches exactly, and on large clustered low-dimensional inputs restricts the exact search to the clusters that
// can hold a neighbour.  The clusters come from farthest-point (greedy k-centre) sampling: every cluster of the data, however
// small, gets a centroid, which is what makes the ball bounds tight.  The greedy loop is inherently sequential -- n_centroids
// rounds of {distance of every sample row to the newest centroid, running minimum, arg-max} -- and as ~4 ATen launches per
// round it cost 80 us a round (330 ms for 4,096 centroids); here it is ONE persistent kernel, one CTA per SM, a grid-wide
// barrier per round (the sample, 65,536 x D floats, stays L2 resident): ~8 us a round.
#include "common.cuh"

namespace mmu {

__device__ __forceinline__ void fps_grid_sync(unsigned int *count, volatile unsigned int *gen, unsigned int n_blocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(count, 1u) == n_blocks - 1u) {
            *count = 0u;
            __threadfence();
            atomicAdd(const_cast<unsigned int *>(gen), 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// bar: [0] arrival counter, [1] generation; best: 3 rotating 64-bit slots {distance bits : ~row} (atomicMax)
__global__ void __launch_bounds__(1024)
fps_kernel(const float *__restrict__ sub, int n_sub, int dim, int n_centroids, float *__restrict__ d2,
           unsigned long long *__restrict__ best, unsigned int *__restrict__ bar, float *__restrict__ cent,
           int32_t *__restrict__ picked) {
    extern __shared__ float s_c[];
    __shared__ unsigned long long s_best;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sub; i += gridDim.x * blockDim.x) d2[i] = __int_as_float(0x7f800000);
    if (blockIdx.x == 0 && threadIdx.x < 3) best[threadIdx.x] = 0ull;
    fps_grid_sync(&bar[0], &bar[1], gridDim.x);
    int cur = 0;
    for (int c = 0; c < n_centroids; ++c) {
        for (int t = threadIdx.x; t < dim; t += blockDim.x) {
            const float v = sub[(int64_t)cur * dim + t];
            s_c[t] = v;
            if (blockIdx.x == 0) cent[(int64_t)c * dim + t] = v;
        }
        if (threadIdx.x == 0) {
            s_best = 0ull;
            // clear the slot of round c+1 now: it was last read right after the barrier of round c-2, and no block can be
            // in round c before every block has passed the barrier of round c-1, i.e. finished that read; the clear is
            // visible to every block after this round's barrier.  (The slot of round c-1 may still be being read.)
            if (blockIdx.x == 0) { picked[c] = cur; best[(c + 1) % 3] = 0ull; }
        }
        __syncthreads();
        unsigned long long mine = 0ull;
        for (int row = gwarp; row < n_sub; row += n_warps) {
            const float *x = sub + (int64_t)row * dim;
            float acc = 0.f;
            for (int t = lane; t < dim; t += 32) { const float df = x[t] - s_c[t]; acc = fmaf(df, df, acc); }
            acc = warp_sum(acc);
            const float nd = fminf(d2[row], acc);
            if (lane == 0) d2[row] = nd;
            const unsigned long long key = ((unsigned long long)__float_as_uint(nd) << 32) | (unsigned int)(0xffffffffu - (unsigned int)row);
            mine = key > mine ? key : mine;
        }
        if (lane == 0 && mine) atomicMax(&s_best, mine);
        __syncthreads();
        if (threadIdx.x == 0 && s_best) atomicMax(&best[c % 3], s_best);
        fps_grid_sync(&bar[0], &bar[1], gridDim.x);
        cur = (int)(0xffffffffu - (unsigned int)(*reinterpret_cast<volatile unsigned long long *>(&best[c % 3]) & 0xffffffffull));
        if (cur < 0 || cur >= n_sub) cur = 0;
    }
}

}  // namespace mmu

extern "C" size_t mmu_fps_workspace_bytes(int64_t n_sub) { return sizeof(float) * (size_t)n_sub + 64; }

extern "C" int mmu_fps_centroids(const float *sub, int64_t n_sub, int dim, int n_centroids, void *workspace,
                                 size_t workspace_bytes, float *out_centroids, int32_t *out_rows, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(sub && workspace && out_centroids && out_rows, "mmu_fps_centroids: null pointer");
    MMU_CHECK_ARG(n_sub >= 1 && n_sub < ((int64_t)1 << 31) && dim >= 1 && dim <= 8192 && n_centroids >= 1,
                  "mmu_fps_centroids: bad sizes");
    MMU_CHECK_ARG(workspace_bytes >= mmu_fps_workspace_bytes(n_sub) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "mmu_fps_centroids: workspace too small or unaligned");
    cudaStream_t st = as_stream(stream);
    // [best: 3 x u64 | bar: 2 x u32 | pad | d2: n_sub floats]
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    unsigned long long *best = reinterpret_cast<unsigned long long *>(ws);
    unsigned int *bar = reinterpret_cast<unsigned int *>(ws + 32);
    float *d2 = reinterpret_cast<float *>(ws + 64);
    MMU_CUDA(cudaMemsetAsync(ws, 0, 64, st));
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // one block per SM: the grid barrier needs every block resident
    // 32 warps per SM: a round is a chain of L2-latency-bound row loads per warp, so more warps = fewer rows per warp
    fps_kernel<<<sms, 1024, sizeof(float) * (size_t)dim, st>>>(sub, (int)n_sub, dim, n_centroids, d2, best, bar, out_centroids, out_rows);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
