/*
 * mmumap.h -- C ABI of the B200-native UMAP engine (libmmumap_b200.so).
 *
 * This is the lower drop-in boundary: the host-side mirror of the reference's Python API
 * (multimodal-umap_b200/impl/model.py, util.py) calls these entry points through ctypes.
 * The reference (aletheiaaaaa/Multimodal-UMAP) has no native layer of its own; each entry
 * point below replaces a group of torch calls in /root/reference/impl/model.py, cited per
 * function as "ref: model.py:<lines>".
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation crosses the ABI: the caller allocates, the kernels fill;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all calls are asynchronous on `stream` unless stated; none synchronises the device;
 *   - return value 0 = ok, non-zero = error; mmu_last_error() returns the message
 *     (thread-local, valid until the next failing call on the same thread);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     MMU_ERR_CUDA.
 *   - indices are int32 on the device (the reference uses int64 COO; the host mirror
 *     converts at the Python boundary), row offsets are int64, all reals are fp32.
 */
#ifndef MMUMAP_H_
#define MMUMAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMU_ABI_VERSION 1

#define MMU_OK 0
#define MMU_ERR_ARG 1      /* invalid argument (message says which) */
#define MMU_ERR_CUDA 2     /* CUDA runtime / launch error */
#define MMU_ERR_UNSUPPORTED 3

#define MMU_MAX_K 64       /* neighbours per row supported by the per-row kernels */

typedef void *mmu_stream_t;

int mmu_abi_version(void);
const char *mmu_last_error(void);
/* Number of kernels of this library launched by the process so far (every launch site counts
 * itself); bench.py reports the difference over its timed region as "gpu_launches". */
uint64_t mmu_launch_count(void);
/* Host code that replays a captured CUDA graph of this library's kernels adds the replayed
 * kernel launches here (a graph replay does not pass through the launch sites). */
void mmu_launch_count_add(uint64_t n);
/* Device properties the host uses to size persistent grids. Synchronous. */
int mmu_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes);
/* A/B switches of the kernels (measurement and tests only; results never depend on them).  They are read from
 * the environment ONCE when the library is loaded and can be changed afterwards with mmu_set_option:
 *   "force_staged"  (env MMUMAP_FORCE_RUNS,     default 1)  0 = plain loop force kernel instead of the staged run form
 *   "knn_cta_pairs" (env MMUMAP_KNN_CTA_PAIRS,  default 1)  0 = one CTA per query block instead of cta_group::2 pairs
 *   "knn_window_mb" (env MMUMAP_KNN_WINDOW_MB,  default -1) -1 = automatic, 0 = one launch, >0 = database window size
 *   "sgd_window_mb" (env MMUMAP_SGD_WINDOW_MB,  default -1) host hint for mmu_edge_forces' window_rows (-1 = automatic)
 *   "knn_fold_norms" (env MMUMAP_KNN_FOLD_NORMS, default 1) 0 = add |Y|^2 in the epilogue instead of inside the contraction
 *   "tail_blocks_per_sm" (env MMUMAP_TAIL_BLOCKS_PER_SM, default 1) grid of mmu_epoch_tail_peer in blocks per SM
 *   "tail_skip_mask" (env MMUMAP_TAIL_SKIP_MASK, default 0) MEASUREMENT ONLY (results are wrong when set): bit 0 drops the
 *                   gradient push of mmu_epoch_tail_push, bit 1 its shard step -- what is left is the synchronisation cost
 * Unknown names return MMU_ERR_ARG. */
int mmu_set_option(const char *name, int64_t value);
int mmu_get_option(const char *name, int64_t *value);
/* Name of the kernel variant the given launch site chose last ("knn_candidates", "edge_forces", "epoch_tail",
 * "block_ops", ...; "" if it has not launched): bench.py reports what actually ran instead of a constant. */
const char *mmu_last_kernel(const char *site);

/* ------------------------------------------------------------------------------------
 * K1/K2/K3  exact kNN graph            ref: model.py:81-195 (candidate search + per-row
 *           top-k), distance model.py:109,163, selection model.py:181-193,
 *           self-exclusion model.py:88,166.
 *
 * Result rows are sorted by (dist, index) ascending; dist = sqrtf(sum_t fmaf(diff,diff,.))
 * accumulated over t ascending (the canonical order of oracle/knn_oracle.c).
 * Rows with fewer than k admissible points are padded with idx=-1, dist=+inf.
 * ---------------------------------------------------------------------------------- */

/* Exhaustive fp32 kNN on CUDA cores (exact by construction; also the fallback for rows the
 * tensor-core path cannot certify).  query_ids (nullable): if given, only the n_query rows
 * query[query_ids[t]] are processed and results are written to row query_ids[t].
 * query_index_base / db_index_base are the global indices of query row 0 / db row 0
 * (multi-GPU shards); out_idx holds GLOBAL db indices.  If exclude_self, the pair whose
 * global indices coincide is skipped.  If merge_existing, out_idx/out_dist are read as the
 * running top-k (streaming over db shards, K3 fused). */
int mmu_knn_exact_f32(const float *query, int64_t n_query, const int32_t *query_ids,
                      const float *db, int64_t n_db, int dim, int k, int exclude_self,
                      int64_t query_index_base, int64_t db_index_base, int merge_existing,
                      int32_t *out_idx, float *out_dist, mmu_stream_t stream);

/* Tensor-core path (tcgen05.mma kind::f16, TMEM accumulators, TMA operand staging): candidates
 * are generated from a centred, power-of-two scaled fp16 copy of the data with a fused per-row
 * top-64 selection in the epilogue; each row is then CERTIFIED from a rigorous bound on the fp16
 * error (no point outside its candidate list can be among the k nearest) and the surviving
 * candidates are rescored with the canonical fp32 distance above, so certified rows equal
 * mmu_knn_exact_f32 bit for bit.  Rows that cannot be certified (exact ties, degenerate data) are
 * NOT written; their indices are appended to fallback_rows and counted in stats[0] -- the caller
 * runs mmu_knn_exact_f32 with query_ids = fallback_rows on them.
 *   stats[0] rows left for the exhaustive kernel, stats[1] candidates rescored, stats[2] rows
 *   certified, stats[3] reserved (4 x int32, zeroed by the call); fallback_rows: [n_query].
 * query_is_db != 0 (fit mode): query and db are the same array and share one fp16 copy.
 * out_idx holds db row numbers; exclude_self drops the pair j == query_index_base + q, or
 * j == query_gid[q] when query_gid (nullable, [n_query]) is given (gathered query rows).
 * min_splits (0..8, 0 = automatic; -S = exactly S): the database range is searched in at least that many splits,
 * each keeping its own 64 candidates per row -- a deeper candidate pool for rows whose
 * neighbourhood gaps are too small for one list to be certified (the host retries the rows of
 * fallback_rows with min_splits = 8 and precision = 1 before resorting to the exhaustive kernel).
 * precision 0: fp16 operands (error bound ~2^-9 |x||y| on a score).  precision 1: error-compensated
 * split operands hi + lo (three times the MMA work, error bound ~2^-20 |x||y|) for data whose
 * neighbour gaps are small against |x||y| (low dimension, norms large against local distances).
 * Kernel form (chosen by the call, same results): rows of padded width >= 512 run as CTA pairs
 * (tcgen05.mma.cta_group::2, M = 256: two query blocks share each database tile, each SM stages half
 * of it), and a database that does not fit in L2 is then walked in ~48 MB windows, one launch per
 * window, the per-row candidate lists carried between launches in the workspace; shorter rows use
 * one CTA per query block (cta_group::1, M = 128). */
#define MMU_KNN_TC_MAX_K 32
size_t mmu_knn_tc_workspace_bytes(int64_t n_query, int64_t n_db, int dim, int query_is_db, int min_splits,
                                  int precision);
int mmu_knn_tc(const float *query, int64_t n_query, const float *db, int64_t n_db, int dim, int k,
               int exclude_self, int64_t query_index_base, const int32_t *query_gid, int query_is_db,
               int min_splits, int precision, void *workspace, size_t workspace_bytes, int32_t *out_idx,
               float *out_dist, int32_t *stats, int32_t *fallback_rows, mmu_stream_t stream);

/* Staged / pruned form of mmu_knn_tc (same arguments, plus):
 *   stages         bit mask 1 = prep (fp16 operand copies, norms), 2 = candidates (tcgen05 kernel), 4 = certify + rescore;
 *                  mmu_knn_tc is stages = 7.  The stages of one search share `workspace`.
 *   qb_tile_begin / qb_tile_end [n_qblocks]   every 128-row query block searches only the 256-row database tiles
 *                  [begin, end) -- or
 *   tile_ptr [n_qblocks + 1] / tile_list      the tiles tile_list[tile_ptr[qb] .. tile_ptr[qb+1]) (no tile twice).
 *   resume         1 = continue the per-row candidate lists an earlier candidates stage left in the workspace.
 *   db_gid [n_db]  the database is a RE-ORDERED copy (cluster sorted): db_gid[j] is the original index of its row j; it is
 *                  what out_idx reports and what ties and self exclusion (against query_gid) are decided on.
 * This is the exact pruned search of umap_b200/knn_pruned.py: rows sorted by cluster, pass 1 over the home clusters'
 * tiles, ball bounds |c_a - c_b| - r_block - R_b against the pass-1 k-th distance select the tiles of pass 2, and the
 * usual certification + canonical fp32 rescoring finishes; a row is exact because every tile NOT visited is farther
 * than its k-th neighbour by the triangle inequality.  Needs a pinned split count (min_splits = -S): the S splits of a query
 * block take its tiles in turn, each with its own candidate lists.
 * mmu_knn_tc_layout: byte offsets inside the workspace {prm (scale at float 0, max |Y|^2 bits at word 1), |X|^2, |Y|^2,
 * cand_idx, cand_score, tau} and {n_qblocks, n_splits, n_tiles, candidates per split, query-block rows, tile rows} in
 * out_words[12]; the error-bound constants {c_rel, c_norm, c_abs, gamma} of the certification in out_consts[4]. */
int mmu_knn_tc_ex(const float *query, int64_t n_query, const float *db, int64_t n_db, int dim, int k,
                  int exclude_self, int64_t query_index_base, const int32_t *query_gid, int query_is_db,
                  int min_splits, int precision, void *workspace, size_t workspace_bytes, int32_t *out_idx,
                  float *out_dist, int32_t *stats, int32_t *fallback_rows, int stages,
                  const int32_t *qb_tile_begin, const int32_t *qb_tile_end, const int32_t *tile_ptr,
                  const int32_t *tile_list, int resume, const int32_t *db_gid, mmu_stream_t stream);
int mmu_knn_tc_layout(int64_t n_query, int64_t n_db, int dim, int query_is_db, int min_splits, int precision,
                      int64_t *out_words, float *out_consts);

/* Farthest-point (greedy k-centre) sampling of n_centroids rows of `sub` [n_sub x dim], for the pruned search: round c
 * picks the row farthest from the rows picked so far (round 0: row 0; ties: the smaller row number).  One persistent kernel,
 * one CTA per SM with a grid-wide barrier per round.  out_centroids [n_centroids x dim], out_rows [n_centroids] (row
 * numbers in `sub`).  workspace: mmu_fps_workspace_bytes(n_sub), 16-byte aligned. */
size_t mmu_fps_workspace_bytes(int64_t n_sub);
int mmu_fps_centroids(const float *sub, int64_t n_sub, int dim, int n_centroids, void *workspace, size_t workspace_bytes,
                      float *out_centroids, int32_t *out_rows, mmu_stream_t stream);

/* K3: merge two sorted per-row lists (e.g. from two db shards) into one sorted top-k. */
int mmu_knn_merge(const int32_t *idx_a, const float *dist_a, const int32_t *idx_b,
                  const float *dist_b, int64_t n_rows, int k, int32_t *out_idx, float *out_dist,
                  mmu_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K4  rho / sigma / membership weights   ref: model.py:33-61 (get_sigmas), :197-209
 * One warp per row.  rho = row minimum.  solver 0 = bisection (64 steps) of
 * sum_j exp(-(d_j-rho)/sigma) = log2(k); solver 1 = the reference's Newton iteration
 * (n_iter steps from sigma=1, closed-form derivative, clamp >= 1e-6) reproducing its
 * results including its divergent rows.  Outputs the row's k entries re-ordered by column
 * (the order .coalesce() gives, model.py:208): col_sorted / w_sorted [n_rows x k].
 * sigma / rho may be NULL.
 * ---------------------------------------------------------------------------------- */
#define MMU_SIGMA_BISECT 0
#define MMU_SIGMA_NEWTON 1
int mmu_smooth_knn(const int32_t *idx, const float *dist, int64_t n_rows, int k, int solver,
                   int n_iter, float *sigma, float *rho, int32_t *col_sorted, float *w_sorted,
                   mmu_stream_t stream);
/* invert-mode weights 1/(1 + a d^(2b))       ref: model.py:206 */
int mmu_invert_weights(const int32_t *idx, const float *dist, int64_t n_rows, int k, float a,
                       float b, int32_t *col_sorted, float *w_sorted, mmu_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K5  fuzzy union S = G + G^T - G*G^T     ref: model.py:271
 * G is the fixed-degree graph from K4 (n x k, columns ascending AND DISTINCT per row, as a kNN
 * result is; rows with repeated columns are outside the contract).  G^T is built by
 * a stable LSD radix sort on the column key; each row then merges its two sorted lists.
 * Output is coalesced COO/CSR: out_rowptr [n+1], out_row/out_col/out_val with capacity
 * 2*n*k entries, sorted by (row, col); values fl(fl(a+b)-fl(a*b)) on mutual edges.
 * out_rowptr[n] is the nnz.
 * ---------------------------------------------------------------------------------- */
size_t mmu_union_workspace_bytes(int64_t n, int k);
int mmu_fuzzy_union(const int32_t *col, const float *w, int64_t n, int k, void *workspace,
                    size_t workspace_bytes, int64_t *out_rowptr, int32_t *out_row,
                    int32_t *out_col, float *out_val, mmu_stream_t stream);

/* Rows [row_base, row_base + n_rows) of the same union (multi-GPU: SURVEY 8(e), "transpose = exchange of edges keyed by
 * destination row block, then local sort-merge"): col_block / w_block are those rows of G; in_key / in_src / in_w [n_in]
 * are the entries of the WHOLE graph whose destination lies in the block, in source-major order, with
 * in_key = dst - row_base.  The kNN result is replicated after its all-gather, so the exchange step is the host's filter
 * of that replica; each rank sorts and merges 1/W of the entries and the CSR blocks are all-gathered.  out_rowptr
 * [n_rows + 1] is LOCAL (starts at 0), out_row holds global row numbers; capacity k * n_rows + n_in entries. */
size_t mmu_union_rows_workspace_bytes(int64_t n_rows, int64_t n_in);
int mmu_fuzzy_union_rows(const int32_t *col_block, const float *w_block, int64_t n_rows, int k, const int32_t *in_key,
                         const int32_t *in_src, const float *in_w, int64_t n_in, int64_t row_base, void *workspace,
                         size_t workspace_bytes, int64_t *out_rowptr, int32_t *out_row, int32_t *out_col,
                         float *out_val, mmu_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K6  transform / invert initialisation    ref: model.py:236-252 (embed_query)
 * out[q] = sum_j (w_qj / max(sum_j w_qj, 1e-6)) * ref[col_qj]
 * ---------------------------------------------------------------------------------- */
int mmu_embed_query(const int32_t *col, const float *w, int64_t n_rows, int k, const float *ref,
                    int dim, float *out, mmu_stream_t stream);
/* CSR SpMM Y = A X (A n x n fp32 CSR, X n x m row-major), used by the spectral init
 * ref: model.py:227,232 (sp.mm / lobpcg operator application) */
int mmu_spmm_csr(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                 const float *x, int m, float *y, mmu_stream_t stream);

/* Eigendecomposition of ONE small symmetric matrix (n <= 64, row-major; (a + a^T)/2 is used) by parallel
 * cyclic Jacobi in a single CTA: lam[n] ascending, v[n x n] row-major with eigenvector j in column j
 * (torch.linalg.eigh's convention).  The dense Rayleigh-Ritz / orthonormalisation steps of the spectral
 * initialisation without a host round trip.  ref: model.py:232 (inside torch.lobpcg) */
int mmu_eigh_small(const float *a, int n, float *lam, float *v, mmu_stream_t stream);
/* the same, returning at once when *skip_flag != 0 (device int; the control block of the block eigensolver below) */
int mmu_eigh_small_flag(const float *a, int n, float *lam, float *v, const int *skip_flag, mmu_stream_t stream);

/* Block eigensolver operations: the dense n x B block algebra (B in {8,16,32}) of the spectral initialisation's
 * Chebyshev-filtered subspace iteration, with every scalar of the iteration in a device-resident control block so
 * that the host enqueues whole outer iterations without a synchronisation.        ref: model.py:221-234 (embed_all /
 * torch.lobpcg's Rayleigh-Ritz and orthonormalisation steps).
 * Control block: mmu_block_ctl_words() floats; word 0 (int) = converged flag, word 1 (int) = Rayleigh-Ritz steps
 * taken, word 2 = largest residual of the wanted pairs, word 3 = filter edge.  Every call below returns at once when
 * the flag is set.
 *   mmu_block_spmm    y = alpha A x + beta x + gamma z with (alpha, beta, gamma) from the control block:
 *                     coef_slot 0 = (1,0,0), 1 = first Chebyshev step (1/e, -c/e, 0), 2 = later steps (2/e, -2c/e, -1);
 *                     z may alias y.
 *   mmu_block_gram    g = x^T y (B x B, row-major), deterministic two-stage reduction (workspace of
 *                     mmu_block_gram_workspace_bytes(b)); mode 0 as is, 1 symmetrised, 2 symmetrised and scaled to unit
 *                     diagonal with dinv[i] = 1/sqrt(g_ii) returned.
 *   mmu_block_rotate  x_out = x_in T (T = tmat, columns reversed if flip); with ax_in also ax_out = ax_in T and the
 *                     control block's column residuals += |ax_out_j - lam'_j x_out_j|^2.  In place allowed.
 *   mmu_block_ritz    after mmu_eigh_small(x^T A x): Ritz values, residual, converged flag (residual < tol or
 *                     max_iters reached), next filter edge and Chebyshev coefficients.
 *   mmu_block_svqb    tmat = D^-1 V Lambda^-1/2 from the eigen-decomposition (lam, v) of the Gram matrix (dinv from
 *                     mode 2, or NULL): x tmat has orthonormal columns. */
int mmu_block_ctl_words(void);
int mmu_block_ctl_init(float *ctl, mmu_stream_t stream);
int mmu_block_spmm(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n, const float *x, int b,
                   const float *ctl, int coef_slot, const float *z, float *y, mmu_stream_t stream);
/* rows [row_lo, row_hi) of the same product only (x, z, y are still whole n x B blocks): a rank's share when the operator
 * applications of a large solve are sharded over GPUs and the row blocks of y are all-gathered afterwards. */
int mmu_block_spmm_rows(const int64_t *rowptr, const int32_t *col, const float *val, int64_t row_lo, int64_t row_hi,
                        const float *x, int b, const float *ctl, int coef_slot, const float *z, float *y,
                        mmu_stream_t stream);
size_t mmu_block_gram_workspace_bytes(int b);
int mmu_block_gram(const float *x, const float *y, int64_t n, int b, int mode, void *workspace, float *g, float *dinv,
                   const float *ctl, mmu_stream_t stream);
int mmu_block_rotate(const float *x_in, float *x_out, const float *ax_in, float *ax_out, int64_t n, int b,
                     const float *tmat, int flip, const float *lam, float *ctl, mmu_stream_t stream);
int mmu_block_ritz(const float *lam, int b, int m, float tol, int max_iters, float *ctl, mmu_stream_t stream);
int mmu_block_svqb(const float *lam, const float *v, const float *dinv, int b, float *tmat, const float *ctl,
                   mmu_stream_t stream);
/* Cholesky-QR step: tmat = L^-T with g = L L^T (g a B x B Gram matrix close to the identity, as after an SVQB pass):
 * x tmat has orthonormal columns.  One small CTA, a few microseconds. */
int mmu_block_cholqr(const float *g, int b, float *tmat, const float *ctl, mmu_stream_t stream);

/* y = alpha * (A x) + beta * x + gamma * z  (z nullable; z may alias y): one three-term
 * Chebyshev recurrence step of the spectral initialisation per launch.  ref: model.py:221-234 */
int mmu_spmm_csr_axpby(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                       const float *x, int m, float alpha, float beta, const float *z, float gamma,
                       float *y, mmu_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K7/K8/K9  layout optimiser               ref: model.py:396-481 (_train),
 *           :312-334 (attractive / repulsive terms), :364-394 (InfoNCE), :403,:474-476 (Adam)
 * ---------------------------------------------------------------------------------- */

/* Per-run optimiser state kept on the device so that epochs can be replayed from a CUDA
 * graph: 8 x uint32 words {epoch, step, step_size(f32), bc2_sqrt(f32), reserved...}. */
#define MMU_OPT_STATE_WORDS 8
int mmu_opt_state_init(uint32_t *state, mmu_stream_t stream);
/* epoch += 1, step += 1, recompute Adam's step_size = lr/(1-beta1^step) and
 * bc2_sqrt = sqrt(1-beta2^step) in double precision. */
int mmu_opt_state_advance(uint32_t *state, double lr, double beta1, double beta2, mmu_stream_t stream);

/* Kept-edge list of one epoch: `kept_rec` holds one 16-byte record {edge position, row, col, row-batch} per kept
 * edge (4 x int32, 16-byte aligned), `kept_hdr` (4 x int32 on the device) is its header: [0] number of kept
 * edges (written by the sampler / mmu_edge_records), [1] capacity of kept_rec in records (written by the HOST once,
 * before the first call), [2] overflow flag (set when an epoch kept more edges than the capacity: the surplus is
 * dropped and the host must treat the run as failed), [3] reserved.
 *
 * K7a (device sample stream): Bernoulli(w) keep per edge (ref: model.py:432) with a counter-based Philox4x32-10
 * stream keyed by (seed, state->epoch, edge position), for the edges [edge_lo, edge_hi) of the COO (the whole graph,
 * or one rank's shard of an edge-sharded multi-GPU optimisation: positions are GLOBAL and the random stream is keyed
 * on them, so the union over shards equals what one GPU draws).  Writes the records of the kept edges (row sorted
 * inside every 1024-edge chunk, chunks in any order) and the number kept in each row-batch (ref: model.py:423-424,
 * batch = row / batch_size) to batch_kept.  kept_hdr[0] and batch_kept are zeroed by the call. */
int mmu_edge_sample_range(const int32_t *row, const int32_t *col, const float *w, int64_t edge_lo,
                          int64_t edge_hi, int batch_size, int n_batches, uint64_t seed,
                          const uint32_t *state, int32_t *kept_rec, int32_t *kept_hdr, int32_t *batch_kept,
                          mmu_stream_t stream);

/* The same with the epoch number given by the host (epoch >= 0) instead of read from `state`:
 * lets the sampling of epoch e+1 run on a second stream while the forces of epoch e execute (it
 * does not depend on the embeddings).  epoch = -1 reads state->epoch; epoch = -2 reads state->epoch + 1 (the next
 * epoch's sample inside a captured CUDA graph, which cannot carry a per-epoch argument). */
int mmu_edge_sample_at(const int32_t *row, const int32_t *col, const float *w, int64_t edge_lo,
                       int64_t edge_hi, int batch_size, int n_batches, uint64_t seed, int64_t epoch,
                       const uint32_t *state, int32_t *kept_rec, int32_t *kept_hdr, int32_t *batch_kept,
                       mmu_stream_t stream);

/* Host sample stream: kept_pos [n_kept] are the edge positions the replayed CPU draws kept (ref: model.py:432,
 * in the reference's order); builds their records and sets kept_hdr[0] = n_kept. */
int mmu_edge_records(const int32_t *row, const int32_t *col, const int32_t *kept_pos, int64_t n_kept,
                     int batch_size, int32_t *kept_rec, int32_t *kept_hdr, mmu_stream_t stream);

/* K7b: forces of every kept edge and its num_rep negatives, accumulated with red.global.add
 * into the gradient table(s).  Gradient of
 *   mean_batches[ mean_kept log(1+a s^b) + mean_{kept*R} -log(a s^b/(1+a s^b)+1e-6) ],
 *   s = max(|y_i-y_j|^2, 1e-6)                         (ref: model.py:312-334,439-453).
 * head/grad_head: the table being optimised [n_head x dim]; tail: table the column indices
 * and negatives address (same pointer as head in fit mode, the frozen fitted table in
 * transform mode); grad_tail: NULL in transform mode (ref: model.py:399-401,416).
 * neg: [n_kept x num_rep] host-generated negative ids (ref: model.py:444) in kept-list order, or NULL to
 * draw them on the device (Philox keyed on the edge position, uniform in [0, rep_count)).
 * loss (nullable): device float accumulating the modality's loss.
 * fast_math != 0: s^b through ex2/lg2 and an approximate reciprocal (relative error of a force
 * coefficient <= ~1e-5; meant for the device sample stream); 0 = powf and IEEE division, the
 * arithmetic the parity tests pin to the reference.
 * window_rows > 0 (and < rep_count): the tail table is processed in windows of that many rows, one launch per
 * window over the whole kept list, each launch handling only the (edge, tail) pairs whose tail row lies in its
 * window -- for tables that do not fit the L2 (10M x 2-D: every random 8-byte gather / red would otherwise move
 * a 32-byte DRAM sector), so that the random traffic of a pass stays inside an L2-resident slice.  Same pairs,
 * same arithmetic; only the order of the atomics differs.  0 = one pass.
 * Kernel form: dim in {2,4,8,16,32,64,128} with num_rep in {4,8} runs the staged run-form kernel (records
 * staged through shared memory with cp.async, head gradient accumulated per run of equal rows); other
 * (dim, num_rep), or option force_staged = 0, the plain loop kernels.  mmu_last_kernel("edge_forces") names it. */
int mmu_edge_forces(const int32_t *kept_rec, const int32_t *kept_hdr, const int32_t *neg,
                    const int32_t *batch_kept, int n_batches, int num_rep, int64_t rep_count,
                    const float *head, const float *tail, float *grad_head, float *grad_tail, int dim,
                    float a, float b, uint64_t seed, const uint32_t *state, float *loss, int fast_math,
                    int64_t window_rows, mmu_stream_t stream);

/* K7c: invert-mode forces (inverse_transform).            ref: model.py:336-362, :437, :447
 * head / grad_head: the Q x dim table being reconstructed in DATA space; data [n x dim]: the
 * target modality's fitted data rows (constants); sigma / rho [n]: its fit-time sigma and rho.
 * Gradient of mean_batches[ mean_kept dist/(w sigma_j + 1e-6)
 *                          + mean_{kept*R} -log(1 - exp(-max(dist - rho_l, 1e-6)/(sigma_l+1e-6)) + 1e-6) ],
 * dist = sqrt(max(|x_i - y|^2, 1e-6)), w = 1/(1 + a dist^(2b)).  Other arguments as mmu_edge_forces. */
int mmu_invert_forces(const int32_t *kept_rec, const int32_t *kept_hdr, const int32_t *neg,
                      const int32_t *batch_kept, int n_batches, int num_rep, int64_t rep_count,
                      const float *head, const float *data, const float *sigma, const float *rho,
                      float *grad_head, int dim, float a, float b, uint64_t seed, const uint32_t *state,
                      float *loss, mmu_stream_t stream);

/* K8: InfoNCE gradient for one direction (anchors e0 -> positives/negatives e1).
 * ref: model.py:364-394.  perm [num] (nullable = identity) and neg [num x n_neg] (nullable =
 * device Philox) are the host draws of model.py:373,383 in anchor order; anchors are
 * weighted weight/(chunk_len*n_chunks) with chunks of `chunk` anchors (model.py:369,392-394);
 * `weight` carries alpha (model.py:467-472). */
int mmu_infonce(const float *e0, const float *e1, int64_t num, int dim, const int32_t *perm,
                const int32_t *neg, int n_neg, int chunk, float weight, float temperature,
                float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                const uint32_t *state, float *loss, mmu_stream_t stream);

/* The same for the anchors [anchor_lo, anchor_hi) of the `num` anchors only (anchor-sharded
 * multi-GPU optimisation); weights, chunks and the random stream are those of the full range. */
int mmu_infonce_range(const float *e0, const float *e1, int64_t num, int64_t anchor_lo,
                      int64_t anchor_hi, int dim, const int32_t *perm, const int32_t *neg, int n_neg,
                      int chunk, float weight, float temperature, float *grad0, float *grad1,
                      uint64_t seed, uint32_t stream_id, const uint32_t *state, float *loss,
                      mmu_stream_t stream);

/* Both directions of one modality pair in a single grid: forward = anchors e0 -> candidates e1 with
 * (perm_fwd, neg_fwd, stream_id), reverse = anchors e1 -> candidates e0 with (perm_rev, neg_rev,
 * stream_id + 1).  Equal to two mmu_infonce_range calls (the reference evaluates L_ij and L_ji on
 * the same embeddings, model.py:467-472).  `weight` carries alpha. */
int mmu_infonce_bidir(const float *e0, const float *e1, int64_t num, int64_t anchor_lo, int64_t anchor_hi,
                      int dim, const int32_t *perm_fwd, const int32_t *neg_fwd, const int32_t *perm_rev,
                      const int32_t *neg_rev, int n_neg, int chunk, float weight, float temperature,
                      float *grad0, float *grad1, uint64_t seed, uint32_t stream_id,
                      const uint32_t *state, float *loss, mmu_stream_t stream);

/* K9: fused Adam update over a dense table (torch.optim.Adam single-tensor semantics,
 * eps=1e-8 style denominator sqrt(v)/bc2_sqrt + eps), reading step_size / bc2_sqrt from
 * `state`; zero_grad != 0 also clears g.  Hyper-parameters are doubles because torch forms
 * 1-beta in double before rounding to fp32; every fp32 operation is rounded separately (no
 * fma contraction) so the update equals torch's CPU result bit for bit.
 *                                                  ref: model.py:403,474-476 */
int mmu_adam_step(float *p, float *g, float *m, float *v, int64_t n, double beta1, double beta2,
                  double eps, const uint32_t *state, int zero_grad, mmu_stream_t stream);

/* ----------------------------------------------------------------------------------
 * Multi-GPU optimiser step over NVLink peer memory (one process per GPU; the reference is single-process,
 * this is the exchange SURVEY.md 8(e) derives for model.py:439-476 sharded by edges).
 * The W ranks hold replicas of the flat parameter buffer and partial gradients in buffers that are mapped
 * into every rank's address space (peer_*[w] = device pointer, valid on THIS rank, to rank w's buffer).
 * mmu_adam_step_peer: rank r sums the r-th 1/W shard of all W gradient buffers in rank order, applies
 *   mmu_adam_step's arithmetic to it and writes the new parameters into all W parameter replicas; m / v are
 *   local, full-size (only the shard is touched).  It neither waits for peers nor clears gradients:
 * mmu_peer_barrier: flag barrier between the W ranks in symmetric memory (peer_flags[w] -> 2 x MMU_PEER_MAX
 *   uint32 of rank w, zero-initialised; slot 0 / 1; seq must increase by one per use of a slot).  Call it with
 *   slot 0 after the gradient kernels and before mmu_adam_step_peer, with slot 1 after it; the local gradient
 *   buffer may be cleared after the slot-1 barrier.  A peer that does not arrive within 20 s fails the launch.
 * ---------------------------------------------------------------------------------- */
#define MMU_PEER_MAX 16
int mmu_peer_barrier(const uint64_t *peer_flags, int world, int rank, int slot, uint32_t seq, mmu_stream_t stream);
int mmu_adam_step_peer(const uint64_t *peer_params, const uint64_t *peer_grads, float *m, float *v, int64_t n,
                       int world, int rank, double beta1, double beta2, double eps, const uint32_t *state,
                       mmu_stream_t stream);

/* The whole multi-GPU epoch tail in ONE launch (the default exchange; the two calls above remain as its A/B partner):
 * flag barrier "gradients complete" -> reduce this rank's 1/W shard of the W partial gradients + Adam + store the new
 * parameters to all W replicas + store zeros over that shard of all W gradient buffers (the gradient clear) -> flag
 * barrier "shard delivered" -> advance the optimiser state words (what mmu_opt_state_advance does; the bias
 * corrections of this step are formed from the old state inside the kernel).  mc_params / mc_grads: NVSwitch multicast
 * addresses of the same symmetric buffers (0 = not available): with them the reduction is one
 * multimem.ld_reduce.add and each broadcast one multimem.st per 16 bytes instead of W peer accesses.
 * done_counter: TWO zero-initialised uint32 in LOCAL device memory: [0] grid-completion counter (left at zero), [1] the
 * barrier sequence number kept on the device.  seq > 0: the host's sequence number (must increase by one per call);
 * seq == 0: the kernel uses and advances done_counter[1] -- no per-epoch argument, so the epoch can be captured into a
 * CUDA graph and replayed.  Flag slots as for mmu_peer_barrier. */
int mmu_epoch_tail_peer(const uint64_t *peer_params, const uint64_t *peer_grads, const uint64_t *peer_flags,
                        uint64_t mc_params, uint64_t mc_grads, float *m, float *v, int64_t n, int world, int rank,
                        uint32_t seq, double lr, double beta1, double beta2, double eps, uint32_t *state,
                        uint32_t *done_counter, mmu_stream_t stream);

/* Push form of the epoch tail, for small tables where NVLink LATENCY is the cost (measured: the pull form above takes
 * 67 us per epoch on 12 MB at 2 and at 8 GPUs alike): nothing is loaded from a peer.  Every rank pushes shard s of its
 * partial gradient `grad` (LOCAL buffer, cleared on the way) into slot `rank` of rank s's inbox, raises flag slot 0,
 * waits for the W flags, sums the W slots of its own inbox (local loads, fixed order), applies Adam once and pushes the
 * new parameters into all W replicas, raises flag slot 1 and waits for the W flags; then advances the optimiser state.
 * peer_inbox[w]: rank w's inbox, W slots of inbox_slot_floats floats each (>= ceil(n/W), multiple of 4).
 * done_counter: as for mmu_epoch_tail_peer (the barrier sequence is always device resident here). */
int mmu_epoch_tail_push(const uint64_t *peer_params, const uint64_t *peer_inbox, const uint64_t *peer_flags,
                        float *grad, float *m, float *v, int64_t n, int64_t inbox_slot_floats, int world, int rank,
                        double lr, double beta1, double beta2, double eps, uint32_t *state, uint32_t *done_counter,
                        mmu_stream_t stream);

/* ----------------------------------------------------------------------------------
 * Measured roof of the random-access kernels (reported by bench.py beside the HBM copy peak; not on the fit path).
 * Touches n_rows_touched uniformly random rows of a [n_rows x row_floats] table the way mmu_edge_forces does: one
 * 16-byte vector load per lane and row from `table` (do_gather) and/or one 16-byte vector red per lane and row into
 * `accum` (do_red), 8 rows in flight per group of row_floats/4 lanes, no arithmetic.  row_floats in {2,4,16,64}.
 * Bytes moved = n_rows_touched * row_floats * 4 * (do_gather + do_red).  ref: the gathers / index_put_ of
 * model.py:316-321,328-333 and their backward. */
int mmu_roof_random_rows(const float *table, float *accum, int64_t n_rows, int row_floats, int64_t n_rows_touched,
                         uint64_t seed, int do_gather, int do_red, float *sink, mmu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMUMAP_H_ */
