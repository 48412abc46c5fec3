// roofs.cu -- measured roofs for the random-access kernels (bench.py reports them beside the HBM copy peak).
//
// The force kernel (layout_sgd.cu; ref: /root/reference/impl/model.py:312-334,439-449) is a stream of random
// row gathers and random row reds.  When its tables fit the L2 (BASELINE.json configs[1]: 12 MB) HBM bandwidth is
// the wrong roof: the ceiling is how many random row-sized L2 accesses the chip sustains.  This kernel measures
// exactly that, with the force kernel's own access shape -- groups of `row_floats / 4` lanes, one 16-byte vector
// load and one 16-byte vector red per lane and row (8-byte accesses for 2-float rows), ROWS_PER_STEP independent
// rows in flight per group -- and no arithmetic to speak of, so that bench.py can state the force kernel's
// achieved GB/s as a fraction of a roof measured on the same box, the same clocks, the same table size.
#include "common.cuh"

namespace mmu {

constexpr int ROOF_ROWS_PER_STEP = 8;      // two Philox calls per step, 8 rows in flight per group

template <int VEC, int LANES>
__global__ void __launch_bounds__(256, 3)
roof_random_rows_kernel(const float *__restrict__ table, float *__restrict__ accum, uint32_t n_rows, int64_t n_steps,
                        uint64_t seed, int do_gather, int do_red, float *__restrict__ sink) {
    constexpr int DIM = VEC * LANES;
    const int gl = threadIdx.x % LANES;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    float acc = 0.f;
    for (int64_t s = gid; s < n_steps; s += n_groups) {
        uint32_t idx[ROOF_ROWS_PER_STEP];
#pragma unroll
        for (int c = 0; c < ROOF_ROWS_PER_STEP / 4; ++c) {
            const Philox4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)c, 77u, k0, k1);
            idx[4 * c + 0] = urange(r.x, n_rows); idx[4 * c + 1] = urange(r.y, n_rows);
            idx[4 * c + 2] = urange(r.z, n_rows); idx[4 * c + 3] = urange(r.w, n_rows);
        }
        float v[ROOF_ROWS_PER_STEP][VEC];
        if (do_gather) {
#pragma unroll
            for (int q = 0; q < ROOF_ROWS_PER_STEP; ++q) {
                const float *p = table + (int64_t)idx[q] * DIM + gl * VEC;
                if (VEC == 4) { const float4 t = __ldcg(reinterpret_cast<const float4 *>(p)); v[q][0] = t.x; v[q][1 % VEC] = t.y; v[q][2 % VEC] = t.z; v[q][3 % VEC] = t.w; }
                else { const float2 t = __ldcg(reinterpret_cast<const float2 *>(p)); v[q][0] = t.x; v[q][1 % VEC] = t.y; }
            }
#pragma unroll
            for (int q = 0; q < ROOF_ROWS_PER_STEP; ++q)
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc += v[q][c];
        }
        if (do_red) {
#pragma unroll
            for (int q = 0; q < ROOF_ROWS_PER_STEP; ++q) {
                float *p = accum + (int64_t)idx[q] * DIM + gl * VEC;
                if (VEC == 4) red_add_v4(p, 1.f, 1.f, 1.f, 1.f);
                else red_add_v2(p, 1.f, 1.f);
            }
        }
    }
    if (acc == 123.456f) *sink = acc;       // keeps the gathers alive
}

}  // namespace mmu

extern "C" int mmu_roof_random_rows(const float *table, float *accum, int64_t n_rows, int row_floats, int64_t n_rows_touched,
                                    uint64_t seed, int do_gather, int do_red, float *sink, mmu_stream_t stream) {
    using namespace mmu;
    MMU_CHECK_ARG(table && accum && sink, "mmu_roof_random_rows: null pointer");
    MMU_CHECK_ARG(n_rows >= 1 && n_rows < ((int64_t)1 << 31) && n_rows_touched >= 0, "mmu_roof_random_rows: bad sizes");
    MMU_CHECK_ARG(row_floats == 2 || row_floats == 4 || row_floats == 16 || row_floats == 64,
                  "mmu_roof_random_rows: row_floats must be 2, 4, 16 or 64");
    const int64_t n_steps = n_rows_touched / ROOF_ROWS_PER_STEP;
    if (n_steps == 0) return MMU_OK;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const unsigned blocks = (unsigned)(sms * 3);
    cudaStream_t st = as_stream(stream);
    switch (row_floats) {
        case 2: roof_random_rows_kernel<2, 1><<<blocks, 256, 0, st>>>(table, accum, (uint32_t)n_rows, n_steps, seed, do_gather, do_red, sink); break;
        case 4: roof_random_rows_kernel<4, 1><<<blocks, 256, 0, st>>>(table, accum, (uint32_t)n_rows, n_steps, seed, do_gather, do_red, sink); break;
        case 16: roof_random_rows_kernel<4, 4><<<blocks, 256, 0, st>>>(table, accum, (uint32_t)n_rows, n_steps, seed, do_gather, do_red, sink); break;
        default: roof_random_rows_kernel<4, 16><<<blocks, 256, 0, st>>>(table, accum, (uint32_t)n_rows, n_steps, seed, do_gather, do_red, sink); break;
    }
    MMU_LAUNCH_CHECK();
    return MMU_OK;
}
