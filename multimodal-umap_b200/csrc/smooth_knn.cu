// smooth_knn.cu -- K4: rho / sigma / membership weights, one warp per row.
//
// ref: /root/reference/impl/model.py:33-61 (get_sigmas: Newton through autograd),
//      :197-209 (rho = row min, w = exp(-(d-rho)/sigma) or 1/(1+a d^2b), coalesced COO).
// Bandwidth bound: reads k*(4+4) B, writes k*(4+4)+8 B per row; everything else stays in
// registers (k <= 64: lane owns entries `lane` and `lane+32`).
#include "common.cuh"

namespace mmu {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <bool INVERT>
__global__ void __launch_bounds__(256)
smooth_knn_kernel(const int32_t *__restrict__ idx, const float *__restrict__ dist, int64_t n_rows, int k,
                  int solver, int n_iter, float *__restrict__ sigma_out, float *__restrict__ rho_out,
                  int32_t *__restrict__ col_sorted, float *__restrict__ w_sorted, float a, float b) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const bool v0 = lane < k, v1 = lane + 32 < k;
    const float inf = __int_as_float(0x7f800000);
    float d0 = v0 ? dist[row * k + lane] : inf;
    float d1 = v1 ? dist[row * k + lane + 32] : inf;
    int32_t c0 = v0 ? idx[row * k + lane] : -1;
    int32_t c1 = v1 ? idx[row * k + lane + 32] : -1;
    float w0, w1;

    if (INVERT) {
        // ref: model.py:206
        w0 = 1.0f / (1.0f + a * powf(d0, 2.0f * b));
        w1 = 1.0f / (1.0f + a * powf(d1, 2.0f * b));
    } else {
        const float rho = warp_min(fminf(d0, d1));
        const float e0 = d0 - rho, e1 = d1 - rho;
        const float target = log2f((float)k);
        float sigma;
        if (solver == MMU_SIGMA_NEWTON) {
            // ref: model.py:52-59, derivative in closed form: d/ds sum exp(-e/s) = sum p * e/(s*s)
            sigma = 1.0f;
            for (int it = 0; it < n_iter; ++it) {
                float p0 = v0 ? expf(-e0 / sigma) : 0.0f;
                float p1 = v1 ? expf(-e1 / sigma) : 0.0f;
                float s2 = sigma * sigma;
                float g0 = v0 ? p0 * (e0 / s2) : 0.0f;
                float g1 = v1 ? p1 * (e1 / s2) : 0.0f;
                float val = warp_sum(p0 + p1) - target;
                float grad = warp_sum(g0 + g1);
                sigma = fmaxf(sigma - val / (grad + 1e-6f), 1e-6f);
            }
        } else {
            // bisection of the same equation (model.py:46-50); doubling until bracketed
            float lo = 0.0f, hi = inf, mid = 1.0f;
            for (int it = 0; it < n_iter; ++it) {
                float p0 = v0 ? expf(-e0 / mid) : 0.0f;
                float p1 = v1 ? expf(-e1 / mid) : 0.0f;
                float s = warp_sum(p0 + p1);
                if (fabsf(s - target) < 1e-5f) break;        // warp-uniform
                if (s > target) { hi = mid; mid = (lo + hi) * 0.5f; }
                else { lo = mid; mid = isinf(hi) ? mid * 2.0f : (lo + hi) * 0.5f; }
            }
            sigma = fmaxf(mid, 1e-6f);
        }
        w0 = v0 ? expf(-e0 / sigma) : 0.0f;
        w1 = v1 ? expf(-e1 / sigma) : 0.0f;
        if (lane == 0) {
            if (sigma_out) sigma_out[row] = sigma;
            if (rho_out) rho_out[row] = rho;
        }
    }

    // order the row by column (what .coalesce() yields, model.py:208): rank by counting
    const uint32_t u0 = (uint32_t)c0, u1 = (uint32_t)c1;     // idx=-1 padding sorts last
    int r0 = 0, r1 = 0;
    for (int j = 0; j < k; ++j) {
        uint32_t cj = __shfl_sync(0xffffffffu, (j < 32) ? u0 : u1, j & 31);
        r0 += (cj < u0) || (cj == u0 && j < lane);
        r1 += (cj < u1) || (cj == u1 && j < lane + 32);
    }
    if (v0) { col_sorted[row * k + r0] = c0; w_sorted[row * k + r0] = w0; }
    if (v1) { col_sorted[row * k + r1] = c1; w_sorted[row * k + r1] = w1; }
}

}  // namespace mmu

extern "C" int mmu_smooth_knn(const int32_t *idx, const float *dist, int64_t n_rows, int k, int solver,
                              int n_iter, float *sigma, float *rho, int32_t *col_sorted, float *w_sorted,
                              mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(idx && dist && col_sorted && w_sorted, "mmu_smooth_knn: null pointer");
    MMU_CHECK_ARG(k >= 1 && k <= MMU_MAX_K, "mmu_smooth_knn: k=%d outside [1,%d]", k, MMU_MAX_K);
    MMU_CHECK_ARG(solver == MMU_SIGMA_BISECT || solver == MMU_SIGMA_NEWTON, "mmu_smooth_knn: bad solver %d", solver);
    MMU_CHECK_ARG(n_iter >= 1, "mmu_smooth_knn: n_iter must be >= 1");
    MMU_CHECK_ARG((const void *)idx != (const void *)col_sorted && (const void *)dist != (const void *)w_sorted,
                  "mmu_smooth_knn: outputs must not alias inputs");
    if (n_rows == 0) return MMU_OK;
    int threads = 256;
    int64_t blocks = (n_rows * 32 + threads - 1) / threads;
    smooth_knn_kernel<false><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(
        idx, dist, n_rows, k, solver, n_iter, sigma, rho, col_sorted, w_sorted, 0.f, 0.f);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_invert_weights(const int32_t *idx, const float *dist, int64_t n_rows, int k, float a,
                                  float b, int32_t *col_sorted, float *w_sorted, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(idx && dist && col_sorted && w_sorted, "mmu_invert_weights: null pointer");
    MMU_CHECK_ARG(k >= 1 && k <= MMU_MAX_K, "mmu_invert_weights: k=%d outside [1,%d]", k, MMU_MAX_K);
    if (n_rows == 0) return MMU_OK;
    int threads = 256;
    int64_t blocks = (n_rows * 32 + threads - 1) / threads;
    smooth_knn_kernel<true><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(
        idx, dist, n_rows, k, 0, 0, nullptr, nullptr, col_sorted, w_sorted, a, b);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
