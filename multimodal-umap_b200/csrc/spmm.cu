// spmm.cu -- K6: transform initialisation (row-normalised fixed-degree SpMM) and a CSR SpMM.
//
// ref: /root/reference/impl/model.py:236-252 (embed_query: row-normalise, sparse @ dense)
//      /root/reference/impl/model.py:227,232 (operator applications inside the spectral init)
#include "common.cuh"

namespace mmu {

// one thread per output element; the k (col, w) pairs of a row are read by `dim` adjacent
// threads (broadcast), ref rows are read as contiguous dim-float segments.
__global__ void __launch_bounds__(256)
embed_query_kernel(const int32_t *__restrict__ col, const float *__restrict__ w, int64_t n_rows, int k,
                   const float *__restrict__ ref, int dim, float *__restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * dim) return;
    int64_t r = e / dim;
    int c = (int)(e - r * dim);
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += w[r * k + j];
    s = fmaxf(s, 1e-6f);                                  // ref: model.py:247 clamp(min=1e-6)
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
        int32_t cj = col[r * k + j];
        if (cj >= 0) acc = fmaf(w[r * k + j] / s, ref[(int64_t)cj * dim + c], acc);
    }
    out[e] = acc;
}

// Y = A X, one warp per row, lanes over the m columns (m <= 32 per sweep), edges sequential.
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                const float *__restrict__ val, int64_t n, const float *__restrict__ x, int m,
                float alpha, float beta, const float *__restrict__ z, float gamma, float *__restrict__ y) {
    // y = alpha * (A x) + beta * x + gamma * z   (z nullable): one Chebyshev step per launch
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int64_t e0 = rowptr[r], e1 = rowptr[r + 1];
    for (int c0 = 0; c0 < m; c0 += 32) {
        int c = c0 + lane;
        float acc = 0.f;
        for (int64_t eb = e0; eb < e1; eb += 32) {
            int64_t e = eb + lane;
            int32_t cj = (e < e1) ? col[e] : 0;
            float vj = (e < e1) ? val[e] : 0.f;
            int cnt = (int)min((int64_t)32, e1 - eb);
            int j = 0;
            // four independent gathers in flight per lane (the loop is latency bound otherwise)
            for (; j + 4 <= cnt; j += 4) {
                int32_t c0_ = __shfl_sync(0xffffffffu, cj, j), c1_ = __shfl_sync(0xffffffffu, cj, j + 1);
                int32_t c2_ = __shfl_sync(0xffffffffu, cj, j + 2), c3_ = __shfl_sync(0xffffffffu, cj, j + 3);
                float v0 = __shfl_sync(0xffffffffu, vj, j), v1 = __shfl_sync(0xffffffffu, vj, j + 1);
                float v2 = __shfl_sync(0xffffffffu, vj, j + 2), v3 = __shfl_sync(0xffffffffu, vj, j + 3);
                if (c < m) {
                    float x0 = x[(int64_t)c0_ * m + c], x1 = x[(int64_t)c1_ * m + c];
                    float x2 = x[(int64_t)c2_ * m + c], x3 = x[(int64_t)c3_ * m + c];
                    acc = fmaf(v0, x0, acc);
                    acc = fmaf(v1, x1, acc);
                    acc = fmaf(v2, x2, acc);
                    acc = fmaf(v3, x3, acc);
                }
            }
            for (; j < cnt; ++j) {
                int32_t cc = __shfl_sync(0xffffffffu, cj, j);
                float vv = __shfl_sync(0xffffffffu, vj, j);
                if (c < m) acc = fmaf(vv, x[(int64_t)cc * m + c], acc);
            }
        }
        if (c < m) {
            float out = alpha * acc;
            if (beta != 0.f) out = fmaf(beta, x[r * m + c], out);
            if (z) out = fmaf(gamma, z[r * m + c], out);
            y[r * m + c] = out;
        }
    }
}

// m == 32 (the spectral solver's block width): 8 lanes per row, each lane owning 4 consecutive
// columns (16-byte loads, one 128-byte segment per edge and row group), 4 rows per warp and two edges
// per step in flight -- four times the memory-level parallelism of the warp-per-row form.
__global__ void __launch_bounds__(256)
spmm_csr_m32_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const float *__restrict__ val, int64_t n, const float *__restrict__ x,
                    float alpha, float beta, const float *__restrict__ z, float gamma, float *__restrict__ y) {
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const bool live = r < n;
    const int64_t e0 = live ? rowptr[r] : 0, e1 = live ? rowptr[r + 1] : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // all four groups of a warp iterate together (shuffles): trip count = the longest row of the warp
    int64_t len = e1 - e0;
    for (int o = 8; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    for (int64_t eb = 0; eb < len; eb += 8) {
        const int64_t e = e0 + eb + sub;
        const int32_t cj = (e < e1) ? col[e] : 0;
        const float vj = (e < e1) ? val[e] : 0.f;           // padding edges contribute 0 * x[0]
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const int32_t ca = __shfl_sync(0xffffffffu, cj, j, 8), cb = __shfl_sync(0xffffffffu, cj, j + 1, 8);
            const float va = __shfl_sync(0xffffffffu, vj, j, 8), vb = __shfl_sync(0xffffffffu, vj, j + 1, 8);
            const float4 xa = *reinterpret_cast<const float4 *>(x + (int64_t)ca * 32 + sub * 4);
            const float4 xb = *reinterpret_cast<const float4 *>(x + (int64_t)cb * 32 + sub * 4);
            acc.x = fmaf(va, xa.x, acc.x); acc.y = fmaf(va, xa.y, acc.y); acc.z = fmaf(va, xa.z, acc.z); acc.w = fmaf(va, xa.w, acc.w);
            acc.x = fmaf(vb, xb.x, acc.x); acc.y = fmaf(vb, xb.y, acc.y); acc.z = fmaf(vb, xb.z, acc.z); acc.w = fmaf(vb, xb.w, acc.w);
        }
    }
    if (!live) return;
    float4 out = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
    const int64_t o = r * 32 + sub * 4;
    if (beta != 0.f) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + o);
        out.x = fmaf(beta, xv.x, out.x); out.y = fmaf(beta, xv.y, out.y); out.z = fmaf(beta, xv.z, out.z); out.w = fmaf(beta, xv.w, out.w);
    }
    if (z) {
        const float4 zv = *reinterpret_cast<const float4 *>(z + o);
        out.x = fmaf(gamma, zv.x, out.x); out.y = fmaf(gamma, zv.y, out.y); out.z = fmaf(gamma, zv.z, out.z); out.w = fmaf(gamma, zv.w, out.w);
    }
    *reinterpret_cast<float4 *>(y + o) = out;
}

}  // namespace mmu

extern "C" int mmu_embed_query(const int32_t *col, const float *w, int64_t n_rows, int k, const float *ref,
                               int dim, float *out, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(col && w && ref && out, "mmu_embed_query: null pointer");
    MMU_CHECK_ARG(k >= 1 && dim >= 1, "mmu_embed_query: bad k/dim");
    if (n_rows == 0) return MMU_OK;
    int64_t total = n_rows * dim;
    embed_query_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(col, w, n_rows, k, ref, dim, out);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_spmm_csr(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                            const float *x, int m, float *y, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(rowptr && col && val && x && y, "mmu_spmm_csr: null pointer");
    MMU_CHECK_ARG(m >= 1, "mmu_spmm_csr: bad m");
    MMU_CHECK_ARG(x != y, "mmu_spmm_csr: x and y must not alias");
    if (n == 0) return MMU_OK;
    spmm_csr_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(rowptr, col, val, n, x, m, 1.f, 0.f,
                                                                                     nullptr, 0.f, y);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_spmm_csr_axpby(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                                  const float *x, int m, float alpha, float beta, const float *z, float gamma,
                                  float *y, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(rowptr && col && val && x && y, "mmu_spmm_csr_axpby: null pointer");
    MMU_CHECK_ARG(m >= 1, "mmu_spmm_csr_axpby: bad m");
    MMU_CHECK_ARG(x != y, "mmu_spmm_csr_axpby: x and y must not alias");
    if (n == 0) return MMU_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & 15) == 0;
    if (m == 32 && aligned) {
        const int64_t warps = (n + 3) / 4;
        spmm_csr_m32_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(rowptr, col, val, n, x, alpha,
                                                                                                 beta, z, gamma, y);
        MMU_LAUNCH_CHECK();
        return MMU_OK;
    }
    spmm_csr_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(rowptr, col, val, n, x, m, alpha,
                                                                                     beta, z, gamma, y);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
