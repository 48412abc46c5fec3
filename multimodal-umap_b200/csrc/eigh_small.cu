// Symmetric eigendecomposition of one small dense matrix (n <= 64) in a single CTA: parallel cyclic
// Jacobi.  Role in the path: the Rayleigh-Ritz and SVQB steps of the spectral initialisation
// (/root/reference/impl/model.py:232 hands the whole eigenproblem to torch.lobpcg, whose inner dense
// solves are of this kind).  The engine's Chebyshev-filtered subspace iteration needs three of these
// per iteration on 32 x 32 matrices; on the device they cost no host round trip and no stream sync.
//
// Round-robin ordering: n/2 disjoint rotations per step, n-1 steps per sweep; each step is
// A <- J^T A J, V <- V J with J the product of the step's rotations, done as a column pass and a
// row pass over double-buffered shared-memory copies.  Jacobi's eigenvalues are accurate to
// eps * |A| in the absolute and, for positive definite Gram matrices, in the relative sense.
#include "common.cuh"

namespace mmu {

constexpr int EIGH_MAX_N = 64;
constexpr int EIGH_MAX_SWEEPS = 16;

constexpr int EIGH_THREADS = 512;

__global__ void __launch_bounds__(EIGH_THREADS)
eigh_small_kernel(const float *__restrict__ a_in, int n, float *__restrict__ lam_out, float *__restrict__ v_out,
                  const int *__restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;         // the enclosing iteration has converged (block_eig.cu control block)
    extern __shared__ float sm[];
    const int ne = (n + 1) & ~1;                 // even number of players; index n (if any) is a dummy
    const int np = ne / 2;                       // rotations per step
    const int nn = n * n;
    float *a0 = sm, *a1 = sm + nn, *v0 = sm + 2 * nn, *v1 = sm + 3 * nn;
    float *pc = sm + 4 * nn;                     // [np] cosine of pair k
    float *ps = pc + np;                         // [np] sine
    int *pp = reinterpret_cast<int *>(ps + np);  // [np] smaller index of pair k
    int *pq = pp + np;                           // [np] larger index (>= n: the pair is idle)
    float *lam = reinterpret_cast<float *>(pq + np);   // [n]
    __shared__ unsigned s_off, s_dia;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < nn; e += nt) {
        const int i = e / n, j = e - i * n;
        a0[e] = 0.5f * (a_in[e] + a_in[j * n + i]);
        v0[e] = i == j ? 1.0f : 0.0f;
    }
    if (tid == 0) { s_off = 0u; s_dia = 0u; }
    __syncthreads();
    for (int sweep = 0; sweep < EIGH_MAX_SWEEPS; ++sweep) {
        // convergence: largest off-diagonal against the largest diagonal magnitude (block-wide maxima through
        // two shared words; the values are non-negative, so the integer compare orders them)
        float off = 0.f, dia = 0.f;
        for (int e = tid; e < nn; e += nt) {
            const int i = e / n, j = e - i * n;
            const float x = fabsf(a0[e]);
            if (i == j) dia = fmaxf(dia, x); else off = fmaxf(off, x);
        }
        atomicMax(&s_off, __float_as_uint(off));
        atomicMax(&s_dia, __float_as_uint(dia));
        __syncthreads();
        const float offm = __uint_as_float(s_off), diam = __uint_as_float(s_dia);
        __syncthreads();
        if (tid == 0) { s_off = 0u; s_dia = 0u; }
        if (offm <= 1e-7f * diam) break;
        for (int step = 0; step < ne - 1; ++step) {
            // pairing of this step (chess tournament): (ne-1, step) and ((step+k) mod m, (step-k) mod m), m = ne-1
            if (tid < np) {
                const int m = ne - 1;
                int p = tid == 0 ? m : (step + tid) % m;
                int q = tid == 0 ? step : (step - tid + m) % m;
                if (p > q) { const int t = p; p = q; q = t; }
                float c = 1.0f, s = 0.0f;
                if (q < n) {
                    const float apq = a0[p * n + q];
                    if (apq != 0.0f) {
                        const float tau = __fdividef(a0[q * n + q] - a0[p * n + p], 2.0f * apq);
                        const float at = fabsf(tau);
                        // t = sign(tau) / (|tau| + sqrt(1 + tau^2)); huge |tau| (tiny a_pq): t -> 1 / (2 tau)
                        const float t = at > 1e18f ? __fdividef(0.5f, tau)
                                                   : __fdividef(tau >= 0.f ? 1.0f : -1.0f, at + sqrtf(fmaf(tau, tau, 1.0f)));
                        // correctly rounded: a biased c (rsqrt.approx) makes every rotation shrink its plane by
                        // ~1e-7, which adds up to 2e-5 relative over the ~450 rotations an index takes part in
                        c = __fdiv_rn(1.0f, __fsqrt_rn(fmaf(t, t, 1.0f)));
                        s = t * c;
                    }
                }
                pc[tid] = c; ps[tid] = s; pp[tid] = p; pq[tid] = q;
            }
            __syncthreads();
            // column pass: A1 = A0 J, V1 = V0 J; work item = (row i, pair k), k fastest
            for (int w = tid; w < n * np; w += nt) {
                const int i = w / np, k = w - i * np;
                const int p = pp[k], q = pq[k];
                if (q < n) {
                    const float c = pc[k], s = ps[k];
                    const float x = a0[i * n + p], y = a0[i * n + q];
                    a1[i * n + p] = c * x - s * y;            // col p' = c col p - s col q
                    a1[i * n + q] = s * x + c * y;            // col q' = s col p + c col q
                    const float vx = v0[i * n + p], vy = v0[i * n + q];
                    v1[i * n + p] = c * vx - s * vy;
                    v1[i * n + q] = s * vx + c * vy;
                } else {
                    a1[i * n + p] = a0[i * n + p];
                    v1[i * n + p] = v0[i * n + p];
                }
            }
            __syncthreads();
            // row pass: A0 = J^T A1; work item = (pair k, column j), j fastest
            for (int w = tid; w < n * np; w += nt) {
                const int k = w / n, j = w - k * n;
                const int p = pp[k], q = pq[k];
                if (q < n) {
                    const float c = pc[k], s = ps[k];
                    const float x = a1[p * n + j], y = a1[q * n + j];
                    // the rotated pair's own off-diagonal element is zero by construction: set it, not compute it
                    a0[p * n + j] = j == q ? 0.0f : c * x - s * y;
                    a0[q * n + j] = j == p ? 0.0f : s * x + c * y;
                } else {
                    a0[p * n + j] = a1[p * n + j];
                }
            }
            { float *t = v0; v0 = v1; v1 = t; }
            __syncthreads();
        }
    }
    // ascending order (torch.linalg.eigh's convention), ties by index
    for (int i = tid; i < n; i += nt) lam[i] = a0[i * n + i];
    __syncthreads();
    for (int e = tid; e < nn; e += nt) {
        const int i = e / n, j = e - i * n;     // element (i, j) of V: eigenvector j
        const float lj = lam[j];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (lam[u] < lj || (lam[u] == lj && u < j)) ? 1 : 0;
        v_out[i * n + rank] = v0[e];
        if (i == 0) lam_out[rank] = lj;
    }
}

}  // namespace mmu

extern "C" int mmu_eigh_small_flag(const float *a, int n, float *lam, float *v, const int *skip_flag, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(a && lam && v, "mmu_eigh_small: null pointer");
    MMU_CHECK_ARG(n >= 1 && n <= EIGH_MAX_N, "mmu_eigh_small: n=%d outside [1,%d]", n, EIGH_MAX_N);
    const int ne = (n + 1) & ~1;
    const size_t smem = sizeof(float) * ((size_t)4 * n * n + 2 * ne + n);
    if (first_use_on_device(SITE_EIGH_SMALL))     // the attribute is per device
        MMU_CUDA(cudaFuncSetAttribute(eigh_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(float) * (4 * EIGH_MAX_N * EIGH_MAX_N + 4 * EIGH_MAX_N))));
    eigh_small_kernel<<<1, EIGH_THREADS, smem, as_stream(stream)>>>(a, n, lam, v, skip_flag);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_eigh_small(const float *a, int n, float *lam, float *v, mmu_stream_t stream) {
    return mmu_eigh_small_flag(a, n, lam, v, nullptr, stream);
}
