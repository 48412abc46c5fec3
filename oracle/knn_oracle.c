/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's kNN *semantics* evaluated exhaustively:
 *   - distance expression: LA.vector_norm(x_i - x_j) in fp32
 *       (/root/reference/impl/model.py:109 and :163)
 *   - selection rule: ascending distance, stable, self excluded in fit mode
 *       (/root/reference/impl/model.py:88,166 self-exclusion; :181-193 per-row top-k)
 *
 * The reference's own candidate search is randomised NN-descent (model.py:84-195)
 * whose output is seed dependent; the pin for "bit-exact neighbour indices" is
 * therefore the exhaustive evaluation of the same fp32 expression with a FIXED
 * accumulation order, which the CUDA rescoring kernel reproduces bit for bit:
 *
 *     acc = 0;  for t = 0..D-1 (ascending):  diff = x[t] - y[t] (fp32, rounded)
 *                                             acc  = fmaf(diff, diff, acc)
 *     dist = sqrtf(acc)          (IEEE correctly rounded)
 *     rank by (dist, index) ascending -- ties on dist go to the smaller index.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off -mfma -pthread)
 */
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* canonical fp32 distance, see header */
static inline float canon_dist(const float *x, const float *y, int D) {
    float acc = 0.0f;
    for (int t = 0; t < D; ++t) {
        float diff = x[t] - y[t];
        acc = fmaf(diff, diff, acc);
    }
    return sqrtf(acc);
}

static inline uint64_t make_key(float d, int64_t idx) {
    uint32_t bits;
    memcpy(&bits, &d, 4);          /* d >= 0 so the bit pattern orders like the value */
    return ((uint64_t)bits << 32) | (uint32_t)idx;
}

typedef struct {
    const float *query, *db;
    int64_t Q, N, self_offset;
    int D, k, exclude_self, tid, nthreads;
    int32_t *out_idx;
    float *out_dist;
} knn_job;

static void *knn_worker(void *arg) {
    knn_job *jb = (knn_job *)arg;
    const int k = jb->k, D = jb->D;
    uint64_t *best = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k);
    /* interleaved rows: thread t takes q = t, t+T, ... */
    for (int64_t q = jb->tid; q < jb->Q; q += jb->nthreads) {
        int cnt = 0;
        const float *x = jb->query + q * (int64_t)D;
        for (int64_t j = 0; j < jb->N; ++j) {
            if (jb->exclude_self && j == q + jb->self_offset) continue;
            float d = canon_dist(x, jb->db + j * (int64_t)D, D);
            uint64_t key = make_key(d, j);
            if (cnt < k) {
                int p = cnt++;
                while (p > 0 && best[p - 1] > key) { best[p] = best[p - 1]; --p; }
                best[p] = key;
            } else if (key < best[k - 1]) {
                int p = k - 1;
                while (p > 0 && best[p - 1] > key) { best[p] = best[p - 1]; --p; }
                best[p] = key;
            }
        }
        for (int i = 0; i < k; ++i) {
            if (i < cnt) {
                uint32_t bits = (uint32_t)(best[i] >> 32);
                float d;
                memcpy(&d, &bits, 4);
                jb->out_idx[q * k + i] = (int32_t)(best[i] & 0xffffffffu);
                jb->out_dist[q * k + i] = d;
            } else {
                jb->out_idx[q * k + i] = -1;
                jb->out_dist[q * k + i] = INFINITY;
            }
        }
    }
    free(best);
    return NULL;
}

/*
 * Exhaustive exact kNN.  query: Q x D, db: N x D (row-major fp32).
 * If exclude_self != 0 the pair (q, j) with j == q + self_offset is skipped
 * (fit mode, model.py:88,166).  Outputs are Q x k, sorted by (dist, idx).
 * Rows with fewer than k admissible candidates are padded with idx=-1, dist=+inf.
 * nthreads <= 0 selects one thread per online core (capped at 64).
 */
int oracle_knn_exact(const float *query, int64_t Q, const float *db, int64_t N, int D, int k,
                     int exclude_self, int64_t self_offset, int32_t *out_idx, float *out_dist,
                     int nthreads) {
    if (k <= 0 || k > 1024 || D <= 0) return 1;
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > 64) nthreads = 64;
    if (nthreads < 1) nthreads = 1;
    pthread_t th[64];
    knn_job jobs[64];
    for (int t = 0; t < nthreads; ++t) {
        knn_job jb = {query, db, Q, N, self_offset, D, k, exclude_self, t, nthreads, out_idx, out_dist};
        jobs[t] = jb;
        if (pthread_create(&th[t], NULL, knn_worker, &jobs[t]) != 0) return 2;
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

/* canonical distances for explicit pairs (used to audit near-ties against torch) */
int oracle_pair_dist(const float *query, const float *db, int D, const int64_t *qi,
                     const int64_t *dj, int64_t npairs, float *out) {
    for (int64_t p = 0; p < npairs; ++p)
        out[p] = canon_dist(query + qi[p] * (int64_t)D, db + dj[p] * (int64_t)D, D);
    return 0;
}
