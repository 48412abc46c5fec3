// block_eig.cu -- dense block algebra of the spectral initialisation, device resident.
//
// ref: /root/reference/impl/model.py:221-234 (embed_all: smallest eigenvectors of I - D^-1/2 S D^-1/2 + 1e-6 I through
// torch.lobpcg, whose Rayleigh-Ritz / orthonormalisation steps are dense block operations of this kind).
//
// The engine's Chebyshev-filtered subspace iteration (umap_b200/spectral.py) works on an n x B block (B = 8, 16 or 32
// columns).  Round 1 ran its dense steps as torch.bmm / cuBLAS / ~30 ATen launches per outer iteration with one host
// synchronisation per iteration (residual and filter edge came back to the host).  Here every step is a kernel of this
// library and every scalar the iteration needs -- Ritz values, residual, convergence flag, filter edge, Chebyshev
// coefficients -- lives in a small device-resident control block, so the host only enqueues:
//
//   gram      G = X^T Y          tall-skinny reduction, deterministic two-stage (per-CTA partials, then one CTA)
//   rotate    X <- X T           (and AX <- AX T with the column residuals |AX t_j - theta_j X t_j|^2 accumulated)
//   ritz      theta, residual, done flag, filter edge and Chebyshev coefficients from the Rayleigh-Ritz eigenvalues
//   svqb      T = D^-1 V Lambda^-1/2 from the eigen-decomposition of the (diagonally scaled) Gram matrix
//   spmm      Y = alpha A X + beta X + gamma Z with (alpha, beta, gamma) read from the control block
//
// All of them return immediately once the control block says "converged", so the host can enqueue a few outer
// iterations ahead and look at the flag once.
#include "common.cuh"

namespace mmu {

// control block (floats unless noted)
enum : int {
    BC_DONE = 0,        // int: 1 once every wanted Ritz pair has residual < tol
    BC_ITERS = 1,       // int: Rayleigh-Ritz steps taken
    BC_RES = 2,         // largest residual of the wanted pairs at the last Rayleigh-Ritz step
    BC_CUT = 3,         // lower edge of the damped interval [-1, cut]
    BC_COEF1 = 4,       // 3 floats: first Chebyshev step  (1/e, -c/e, 0)
    BC_COEFK = 8,       // 3 floats: later steps           (2/e, -2c/e, -1)
    BC_COEF_ID = 12,    // 3 floats: (1, 0, 0): plain operator application
    BC_THETA = 16,      // B floats: Ritz values, descending
    BC_RES2 = 16 + 64,  // B floats: column residuals squared (accumulated by rotate, consumed by ritz)
    BC_WORDS = 16 + 128
};

__device__ __forceinline__ bool bc_done(const float *ctl) { return ctl && reinterpret_cast<const int *>(ctl)[BC_DONE] != 0; }

// ------------------------------------------------------------------ Y = alpha A X + beta X + gamma Z, B columns
// B/4 lanes per row, each lane owning 4 consecutive columns (16-byte gathers), 32/(B/4) rows per warp, two edges per
// step in flight.  coef (device, 3 floats) or the by-value triple.
template <int B>
__global__ void __launch_bounds__(256)
spmm_block_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ val,
                  int64_t row_lo, int64_t n, const float *__restrict__ x, const float *__restrict__ coef,
                  const float *__restrict__ z, float *__restrict__ y, const float *__restrict__ ctl) {
    // rows [row_lo, n) of the operator (a rank's row block when the SpMM is sharded over GPUs; x is always the full block)
    if (bc_done(ctl)) return;
    constexpr int LPR = B / 4, RPW = 32 / LPR;
    const float alpha = coef[0], beta = coef[1], gamma = coef[2];
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const int64_t r = row_lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW + grp;
    const bool live = r < n;
    const int64_t e0 = live ? rowptr[r] : 0, e1 = live ? rowptr[r + 1] : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // all groups of a warp iterate together (shuffles): trip count = the longest row of the warp
    int64_t len = e1 - e0;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    for (int64_t eb = 0; eb < len; eb += LPR) {
        const int64_t e = e0 + eb + sub;
        const int32_t cj = (e < e1) ? col[e] : 0;
        const float vj = (e < e1) ? val[e] : 0.f;           // padding edges contribute 0 * x[0]
#pragma unroll
        for (int j = 0; j < LPR; j += 2) {
            const int32_t ca = __shfl_sync(0xffffffffu, cj, j, LPR), cb = __shfl_sync(0xffffffffu, cj, j + 1, LPR);
            const float va = __shfl_sync(0xffffffffu, vj, j, LPR), vb = __shfl_sync(0xffffffffu, vj, j + 1, LPR);
            const float4 xa = *reinterpret_cast<const float4 *>(x + (int64_t)ca * B + sub * 4);
            const float4 xb = *reinterpret_cast<const float4 *>(x + (int64_t)cb * B + sub * 4);
            acc.x = fmaf(va, xa.x, acc.x); acc.y = fmaf(va, xa.y, acc.y); acc.z = fmaf(va, xa.z, acc.z); acc.w = fmaf(va, xa.w, acc.w);
            acc.x = fmaf(vb, xb.x, acc.x); acc.y = fmaf(vb, xb.y, acc.y); acc.z = fmaf(vb, xb.z, acc.z); acc.w = fmaf(vb, xb.w, acc.w);
        }
    }
    if (!live) return;
    float4 out = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
    const int64_t o = r * B + sub * 4;
    if (beta != 0.f) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + o);
        out.x = fmaf(beta, xv.x, out.x); out.y = fmaf(beta, xv.y, out.y); out.z = fmaf(beta, xv.z, out.z); out.w = fmaf(beta, xv.w, out.w);
    }
    if (gamma != 0.f) {
        const float4 zv = *reinterpret_cast<const float4 *>(z + o);
        out.x = fmaf(gamma, zv.x, out.x); out.y = fmaf(gamma, zv.y, out.y); out.z = fmaf(gamma, zv.z, out.z); out.w = fmaf(gamma, zv.w, out.w);
    }
    *reinterpret_cast<float4 *>(y + o) = out;
}

// ------------------------------------------------------------------ G = X^T Y  (n x B blocks, B x B result)
constexpr int GRAM_ROWS = 64;          // rows staged per step
constexpr int GRAM_MAX_CTAS = 592;

template <int B>
__global__ void __launch_bounds__(256)
gram_partial_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t n, float *__restrict__ partial,
                    const float *__restrict__ ctl) {
    if (bc_done(ctl)) return;
    // thread t owns OUT = B*B/256 outputs of row i = t / TPR, columns j0 .. j0+OUT-1 (B = 8: one output of 64 threads)
    constexpr int OUT = (B * B >= 256) ? B * B / 256 : 1;
    constexpr int TPR = B / OUT;                       // threads per output row
    __shared__ float xs[GRAM_ROWS][B + 1], ys[GRAM_ROWS][B];
    const int t = threadIdx.x;
    const bool owner = t < B * TPR;
    const int i = owner ? t / TPR : 0, j0 = owner ? (t % TPR) * OUT : 0;
    float acc[OUT];
#pragma unroll
    for (int u = 0; u < OUT; ++u) acc[u] = 0.f;
    const int64_t n_steps = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    for (int64_t s = blockIdx.x; s < n_steps; s += gridDim.x) {
        const int64_t r0 = s * GRAM_ROWS;
        for (int e = t; e < GRAM_ROWS * B; e += 256) {
            const int rr = e / B, cc = e % B;
            const bool in = r0 + rr < n;
            xs[rr][cc] = in ? x[(r0 + rr) * B + cc] : 0.f;
            ys[rr][cc] = in ? y[(r0 + rr) * B + cc] : 0.f;
        }
        __syncthreads();
        if (owner) {
#pragma unroll 8
            for (int rr = 0; rr < GRAM_ROWS; ++rr) {
                const float xv = xs[rr][i];
#pragma unroll
                for (int u = 0; u < OUT; ++u) acc[u] = fmaf(xv, ys[rr][j0 + u], acc[u]);
            }
        }
        __syncthreads();
    }
    if (owner) {
#pragma unroll
        for (int u = 0; u < OUT; ++u) partial[(int64_t)blockIdx.x * B * B + i * B + j0 + u] = acc[u];
    }
}

// sums the per-CTA partials in a fixed order; mode 0: G as is; 1: symmetrised (G + G^T)/2;
// 2: symmetrised and scaled to unit diagonal, Gs_ij = G_ij / sqrt(G_ii G_jj), with diag[i] = 1/sqrt(G_ii) kept for svqb
template <int B>
__global__ void __launch_bounds__(B *B >= 256 ? 1024 : 64)
gram_reduce_kernel(const float *__restrict__ partial, int n_parts, int mode, float *__restrict__ g, float *__restrict__ dinv,
                   const float *__restrict__ ctl) {
    if (bc_done(ctl)) return;
    __shared__ float s[B][B + 1];
    __shared__ float sd[B];
    for (int e = threadIdx.x; e < B * B; e += blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < n_parts; ++p) a += partial[(int64_t)p * B * B + e];
        s[e / B][e % B] = a;
    }
    __syncthreads();
    if (threadIdx.x < B) sd[threadIdx.x] = mode == 2 ? rsqrtf(fmaxf(s[threadIdx.x][threadIdx.x], 1e-37f)) : 1.0f;
    __syncthreads();
    for (int e = threadIdx.x; e < B * B; e += blockDim.x) {
        const int i = e / B, j = e % B;
        float v = mode == 0 ? s[i][j] : 0.5f * (s[i][j] + s[j][i]);
        g[e] = v * sd[i] * sd[j];
    }
    if (dinv && threadIdx.x < B) dinv[threadIdx.x] = sd[threadIdx.x];
}

// ------------------------------------------------------------------ X <- X T (and AX <- AX T, residuals)
// T (B x B, row-major: out column j = sum_k X[:,k] T[k][j]); flip: use T's columns in reverse order (eigh returns
// ascending eigenvalues, the iteration wants descending).  If ax is given: the same rotation is applied to it and
// res2[j] += sum_rows (AX t_j - theta_j X t_j)^2 with theta_j = lam[flip ? B-1-j : j].  In place (rows are independent).
template <int B>
__global__ void __launch_bounds__(256)
rotate_kernel(const float *__restrict__ xin, float *__restrict__ xout, const float *__restrict__ axin, float *__restrict__ axout,
              int64_t n, const float *__restrict__ tmat, int flip, const float *__restrict__ lam, float *__restrict__ res2,
              const float *__restrict__ ctl) {
    if (bc_done(ctl)) return;
    constexpr int LPR = B / 4, RPW = 32 / LPR;
    __shared__ float4 ts[B][LPR];             // ts[k][s] = T[k][4s .. 4s+3] (after the optional column flip)
    __shared__ float s_res[B];
    for (int e = threadIdx.x; e < B * B; e += 256) {
        const int k = e / B, j = e % B;
        reinterpret_cast<float *>(&ts[k][0])[j] = tmat[k * B + (flip ? B - 1 - j : j)];
    }
    if (threadIdx.x < B) s_res[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    float th[4] = {0.f, 0.f, 0.f, 0.f};
    if (axin && lam) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int j = sub * 4 + u; th[u] = lam[flip ? B - 1 - j : j]; }
    }
    float racc[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_groups = (n + RPW - 1) / RPW;
    for (int64_t wg = warp0; wg < n_groups; wg += n_warps) {
        const int64_t r = wg * RPW + grp;
        const bool live = r < n;
        const float4 xv = live ? *reinterpret_cast<const float4 *>(xin + r * B + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 av = (live && axin) ? *reinterpret_cast<const float4 *>(axin + r * B + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ox = make_float4(0.f, 0.f, 0.f, 0.f), oa = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int src = k >> 2;
            const float xs_ = (k & 3) == 0 ? xv.x : (k & 3) == 1 ? xv.y : (k & 3) == 2 ? xv.z : xv.w;
            const float xk = __shfl_sync(0xffffffffu, xs_, src, LPR);
            const float4 tk = ts[k][sub];
            ox.x = fmaf(xk, tk.x, ox.x); ox.y = fmaf(xk, tk.y, ox.y); ox.z = fmaf(xk, tk.z, ox.z); ox.w = fmaf(xk, tk.w, ox.w);
            if (axin) {
                const float as_ = (k & 3) == 0 ? av.x : (k & 3) == 1 ? av.y : (k & 3) == 2 ? av.z : av.w;
                const float ak = __shfl_sync(0xffffffffu, as_, src, LPR);
                oa.x = fmaf(ak, tk.x, oa.x); oa.y = fmaf(ak, tk.y, oa.y); oa.z = fmaf(ak, tk.z, oa.z); oa.w = fmaf(ak, tk.w, oa.w);
            }
        }
        if (live) {
            *reinterpret_cast<float4 *>(xout + r * B + sub * 4) = ox;
            if (axin) {
                *reinterpret_cast<float4 *>(axout + r * B + sub * 4) = oa;
                float d;
                d = oa.x - th[0] * ox.x; racc[0] = fmaf(d, d, racc[0]);
                d = oa.y - th[1] * ox.y; racc[1] = fmaf(d, d, racc[1]);
                d = oa.z - th[2] * ox.z; racc[2] = fmaf(d, d, racc[2]);
                d = oa.w - th[3] * ox.w; racc[3] = fmaf(d, d, racc[3]);
            }
        }
    }
    if (axin && res2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(&s_res[sub * 4 + u], racc[u]);
        __syncthreads();
        if (threadIdx.x < B && s_res[threadIdx.x] != 0.f) atomicAdd(&res2[threadIdx.x], s_res[threadIdx.x]);
    }
}

// ------------------------------------------------------------------ Rayleigh-Ritz bookkeeping (one warp)
// lam: eigenvalues of X^T A X ascending (mmu_eigh_small).  Writes theta (descending), the largest residual of the m
// wanted pairs, the convergence flag, and the Chebyshev filter of the next sweep: damp [-1, cut] with
// cut = max(min(theta_min_of_block, theta_m - 0.05), -0.5) -- UMAP graphs of well separated clusters have one
// eigenvalue ~1 per cluster, a cluster wider than the block, so the edge stays a fixed distance below the wanted ones.
__global__ void ritz_kernel(const float *__restrict__ lam, int b, int m, float tol, int max_iters, float *__restrict__ ctl) {
    int *ictl = reinterpret_cast<int *>(ctl);
    if (ictl[BC_DONE]) return;
    if (threadIdx.x == 0) {
        float res = 0.f;
        for (int j = 0; j < m; ++j) res = fmaxf(res, sqrtf(fmaxf(ctl[BC_RES2 + j], 0.f)));
        for (int j = 0; j < b; ++j) { ctl[BC_THETA + j] = lam[b - 1 - j]; ctl[BC_RES2 + j] = 0.f; }
        const int it = ictl[BC_ITERS] + 1;
        ictl[BC_ITERS] = it;
        ctl[BC_RES] = res;
        if (res < tol || it >= max_iters) { ictl[BC_DONE] = 1; return; }
        const float lo = lam[0], theta_m = lam[b - m];
        const float cut = fmaxf(fminf(lo, theta_m - 0.05f), -0.5f);
        const float e = (cut + 1.0f) * 0.5f, c = (cut - 1.0f) * 0.5f;
        ctl[BC_CUT] = cut;
        ctl[BC_COEF1 + 0] = 1.0f / e; ctl[BC_COEF1 + 1] = -c / e; ctl[BC_COEF1 + 2] = 0.f;
        ctl[BC_COEFK + 0] = 2.0f / e; ctl[BC_COEFK + 1] = -2.0f * c / e; ctl[BC_COEFK + 2] = -1.0f;
    }
}

__global__ void block_ctl_init_kernel(float *__restrict__ ctl) {
    for (int i = threadIdx.x; i < BC_WORDS; i += blockDim.x) ctl[i] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) { ctl[BC_COEF_ID] = 1.0f; ctl[BC_COEF_ID + 1] = 0.f; ctl[BC_COEF_ID + 2] = 0.f; }
}

// ------------------------------------------------------------------ SVQB: T = D^-1 V Lambda^-1/2
// (lam, v) = eigen-decomposition of the unit-diagonal Gram matrix Gs = D^-1 G D^-1 (dinv = diagonal of D^-1, or null for
// D = I): X T has orthonormal columns.  Eigenvalues are clamped at 1e-10 of the largest (nearly dependent columns, which
// the Chebyshev filter produces by design).
__global__ void svqb_kernel(const float *__restrict__ lam, const float *__restrict__ v, const float *__restrict__ dinv, int b,
                            float *__restrict__ tmat, const float *__restrict__ ctl) {
    if (bc_done(ctl)) return;
    __shared__ float s_scale[64];
    if (threadIdx.x < b) {
        float mx = 0.f;
        for (int j = 0; j < b; ++j) mx = fmaxf(mx, lam[j]);
        s_scale[threadIdx.x] = rsqrtf(fmaxf(lam[threadIdx.x], fmaxf(mx * 1e-10f, 1e-37f)));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < b * b; e += blockDim.x) {
        const int k = e / b, j = e % b;
        tmat[e] = (dinv ? dinv[k] : 1.0f) * v[e] * s_scale[j];
    }
}

// ------------------------------------------------------------------ Cholesky-QR: T = L^-T with G = L L^T
// The second orthonormalisation of a sweep sees a Gram matrix that is already close to the identity (SVQB ran just
// before), where a Cholesky factorisation is stable and costs a few microseconds in one small CTA -- against ~100 us
// for another Jacobi eigendecomposition.  Thread i owns row i of L and, afterwards, column i of L^-1.
__global__ void __launch_bounds__(64)
chol_inv_kernel(const float *__restrict__ g, int b, float *__restrict__ tmat, const float *__restrict__ ctl) {
    if (bc_done(ctl)) return;
    __shared__ float l[64][65], li[64][65];
    const int i = threadIdx.x;
    for (int e = i; e < 64 * 65; e += 64) { (&l[0][0])[e] = 0.f; (&li[0][0])[e] = 0.f; }
    __syncthreads();
    for (int j = 0; j < b; ++j) {
        if (i >= j && i < b) {
            float s = 0.5f * (g[i * b + j] + g[j * b + i]);
            for (int k = 0; k < j; ++k) s = fmaf(-l[i][k], l[j][k], s);
            if (i == j) l[j][j] = sqrtf(fmaxf(s, 1e-30f));
            else l[i][j] = s;                                  // divided by l[j][j] below
        }
        __syncthreads();
        if (i > j && i < b) l[i][j] = l[i][j] / l[j][j];
        __syncthreads();
    }
    // column c = i of L^-1 by forward substitution
    if (i < b) {
        li[i][i] = 1.0f / l[i][i];
        for (int r = i + 1; r < b; ++r) {
            float s = 0.f;
            for (int k = i; k < r; ++k) s = fmaf(l[r][k], li[k][i], s);
            li[r][i] = -s / l[r][r];
        }
    }
    __syncthreads();
    for (int e = i; e < b * b; e += 64) {
        const int k = e / b, j = e % b;
        tmat[e] = li[j][k];                                    // T = (L^-1)^T, upper triangular
    }
}

static inline unsigned row_group_blocks(int64_t n, int b) {
    const int rpw = 32 / (b / 4);
    const int64_t warps = (n + rpw - 1) / rpw;
    return (unsigned)((warps * 32 + 255) / 256);
}

}  // namespace mmu

#define MMU_BLOCK_DISPATCH(B_, CALL8, CALL16, CALL32)       \
    do {                                                    \
        if ((B_) == 8) { CALL8; }                           \
        else if ((B_) == 16) { CALL16; }                    \
        else { CALL32; }                                    \
    } while (0)

extern "C" int mmu_block_ctl_words(void) { return mmu::BC_WORDS; }

extern "C" int mmu_block_ctl_init(float *ctl, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(ctl, "mmu_block_ctl_init: null pointer");
    block_ctl_init_kernel<<<1, 128, 0, as_stream(stream)>>>(ctl);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_block_spmm(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n, const float *x, int b,
                              const float *ctl, int coef_slot, const float *z, float *y, mmu_stream_t stream) {
    return mmu_block_spmm_rows(rowptr, col, val, 0, n, x, b, ctl, coef_slot, z, y, stream);
}

extern "C" int mmu_block_spmm_rows(const int64_t *rowptr, const int32_t *col, const float *val, int64_t row_lo, int64_t row_hi,
                                   const float *x, int b, const float *ctl, int coef_slot, const float *z, float *y,
                                   mmu_stream_t stream) {
    using namespace mmu;
    const int64_t n = row_hi;
    MMU_CHECK_ARG(row_lo >= 0 && row_lo <= row_hi, "mmu_block_spmm_rows: bad row range");
    MMU_CHECK_ARG(rowptr && col && val && x && y && ctl, "mmu_block_spmm: null pointer");
    MMU_CHECK_ARG(b == 8 || b == 16 || b == 32, "mmu_block_spmm: block width %d not in {8,16,32}", b);
    MMU_CHECK_ARG(coef_slot >= 0 && coef_slot <= 2, "mmu_block_spmm: coef_slot outside [0,2]");
    MMU_CHECK_ARG(x != y, "mmu_block_spmm: x and y must not alias");
    MMU_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & 15) == 0,
                  "mmu_block_spmm: blocks must be 16-byte aligned");
    if (n == row_lo) return MMU_OK;
    const float *coef = ctl + (coef_slot == 0 ? BC_COEF_ID : coef_slot == 1 ? BC_COEF1 : BC_COEFK);
    MMU_CHECK_ARG(coef_slot != 2 || z, "mmu_block_spmm: the three-term step needs z");
    const float *zz = z ? z : x;
    const unsigned blocks = row_group_blocks(n - row_lo, b);
    cudaStream_t st = as_stream(stream);
    MMU_BLOCK_DISPATCH(b, (spmm_block_kernel<8><<<blocks, 256, 0, st>>>(rowptr, col, val, row_lo, n, x, coef, zz, y, ctl)),
                       (spmm_block_kernel<16><<<blocks, 256, 0, st>>>(rowptr, col, val, row_lo, n, x, coef, zz, y, ctl)),
                       (spmm_block_kernel<32><<<blocks, 256, 0, st>>>(rowptr, col, val, row_lo, n, x, coef, zz, y, ctl)));
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" size_t mmu_block_gram_workspace_bytes(int b) { return sizeof(float) * (size_t)mmu::GRAM_MAX_CTAS * b * b; }

extern "C" int mmu_block_gram(const float *x, const float *y, int64_t n, int b, int mode, void *workspace, float *g,
                              float *dinv, const float *ctl, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(x && y && workspace && g, "mmu_block_gram: null pointer");
    MMU_CHECK_ARG(b == 8 || b == 16 || b == 32, "mmu_block_gram: block width %d not in {8,16,32}", b);
    MMU_CHECK_ARG(mode >= 0 && mode <= 2 && n >= 1, "mmu_block_gram: bad mode / n");
    MMU_CHECK_ARG(mode != 2 || dinv, "mmu_block_gram: mode 2 needs dinv");
    const int64_t steps = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    const int parts = (int)(steps < GRAM_MAX_CTAS ? steps : GRAM_MAX_CTAS);
    float *partial = static_cast<float *>(workspace);
    cudaStream_t st = as_stream(stream);
    MMU_BLOCK_DISPATCH(b, (gram_partial_kernel<8><<<parts, 256, 0, st>>>(x, y, n, partial, ctl)),
                       (gram_partial_kernel<16><<<parts, 256, 0, st>>>(x, y, n, partial, ctl)),
                       (gram_partial_kernel<32><<<parts, 256, 0, st>>>(x, y, n, partial, ctl)));
    MMU_BLOCK_DISPATCH(b, (gram_reduce_kernel<8><<<1, 64, 0, st>>>(partial, parts, mode, g, dinv, ctl)),
                       (gram_reduce_kernel<16><<<1, 1024, 0, st>>>(partial, parts, mode, g, dinv, ctl)),
                       (gram_reduce_kernel<32><<<1, 1024, 0, st>>>(partial, parts, mode, g, dinv, ctl)));
    MMU_LAUNCH_CHECK_N(2);
    return MMU_OK;
}

extern "C" int mmu_block_rotate(const float *x_in, float *x_out, const float *ax_in, float *ax_out, int64_t n, int b,
                                const float *tmat, int flip, const float *lam, float *ctl, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(x_in && x_out && tmat, "mmu_block_rotate: null pointer");
    MMU_CHECK_ARG(b == 8 || b == 16 || b == 32, "mmu_block_rotate: block width %d not in {8,16,32}", b);
    MMU_CHECK_ARG((ax_in == nullptr) == (ax_out == nullptr), "mmu_block_rotate: ax_in / ax_out go together");
    MMU_CHECK_ARG(!ax_in || (lam && ctl), "mmu_block_rotate: residuals need lam and the control block");
    if (n == 0) return MMU_OK;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    unsigned blocks = row_group_blocks(n, b);
    if (blocks > (unsigned)sms * 8) blocks = (unsigned)sms * 8;
    float *res2 = ctl ? ctl + BC_RES2 : nullptr;
    cudaStream_t st = as_stream(stream);
    MMU_BLOCK_DISPATCH(b, (rotate_kernel<8><<<blocks, 256, 0, st>>>(x_in, x_out, ax_in, ax_out, n, tmat, flip, lam, res2, ctl)),
                       (rotate_kernel<16><<<blocks, 256, 0, st>>>(x_in, x_out, ax_in, ax_out, n, tmat, flip, lam, res2, ctl)),
                       (rotate_kernel<32><<<blocks, 256, 0, st>>>(x_in, x_out, ax_in, ax_out, n, tmat, flip, lam, res2, ctl)));
    note_kernel(SITE_BLOCK_OPS, "spmm_block/gram_partial/rotate_kernel<B=%d>", b);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_block_ritz(const float *lam, int b, int m, float tol, int max_iters, float *ctl, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(lam && ctl, "mmu_block_ritz: null pointer");
    MMU_CHECK_ARG(b >= 1 && b <= 64 && m >= 1 && m <= b, "mmu_block_ritz: bad block geometry");
    ritz_kernel<<<1, 32, 0, as_stream(stream)>>>(lam, b, m, tol, max_iters, ctl);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_block_svqb(const float *lam, const float *v, const float *dinv, int b, float *tmat, const float *ctl,
                              mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(lam && v && tmat, "mmu_block_svqb: null pointer");
    MMU_CHECK_ARG(b >= 1 && b <= 64, "mmu_block_svqb: bad block width");
    svqb_kernel<<<1, 256, 0, as_stream(stream)>>>(lam, v, dinv, b, tmat, ctl);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}

extern "C" int mmu_block_cholqr(const float *g, int b, float *tmat, const float *ctl, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(g && tmat, "mmu_block_cholqr: null pointer");
    MMU_CHECK_ARG(b >= 1 && b <= 64, "mmu_block_cholqr: bad block width");
    chol_inv_kernel<<<1, 64, 0, as_stream(stream)>>>(g, b, tmat, ctl);
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
