/*
 * pmm.h — C ABI of libpmm_b200.so: the B200-native replacement for polars-matmul's native layer.
 *
 * This is the drop-in boundary for the one hot path of NivekNey/polars-matmul v0.1.4:
 *     pl.col(..).pmm.topk(corpus, k, metric)   and   pl.col(..).pmm.matmul(corpus, flatten)
 * The reference crosses Python -> Rust at src/lib.rs:15-55 (`_matmul`, `_topk`, PyO3) and does all
 * numeric work in src/matmul.rs, src/metrics.rs, src/topk.rs on the CPU (faer GEMM + serial
 * quickselect).  A maintainer of the reference keeps src/lib.rs and python/polars_matmul/__init__.py
 * and replaces the bodies of `topk_impl` (src/matmul.rs:473) and `matmul_impl` (src/matmul.rs:295)
 * with calls to `pmm_topk` / `pmm_matmul` below (binding stubs: INTEGRATION.md).
 *
 * Conventions
 *   - plain C, pointers and sizes only; no CUDA, torch or Arrow types in signatures
 *     (a `void* stream` is a cudaStream_t passed opaquely; NULL = the library's own stream);
 *   - every function returns PMM_OK (0) or a PMM_ERR_* code; the message is in pmm_last_error()
 *     (thread-local) and contains the same substrings the reference's errors carry
 *     ("Unknown metric", "Empty series", "Dimension mismatch", "Zero-dimensional vectors",
 *     "First element is null"), so the Python shim can raise RuntimeError(msg) exactly as
 *     src/lib.rs:28,53 does;
 *   - the caller owns every input and output buffer; the library owns all device memory;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PMM_ERR_CUDA.
 */
#ifndef PMM_H_
#define PMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PMM_API __attribute__((visibility("default")))
#else
#define PMM_API
#endif

#define PMM_OK 0
#define PMM_ERR_INVALID 1     /* bad argument / reference-visible compute error (message matters) */
#define PMM_ERR_CUDA 2        /* CUDA runtime / driver failure, or no device */
#define PMM_ERR_UNSUPPORTED 3 /* valid request this build does not cover (says which) */

/* Storage dtype of a values buffer.  Working precision follows src/matmul.rs:308,427:
 * f32 iff BOTH sides are f32 (f16 storage is upcast exactly and counts as f32, README.md:154-156). */
#define PMM_DTYPE_F16 0
#define PMM_DTYPE_F32 1
#define PMM_DTYPE_F64 2

/* src/metrics.rs:10-17 */
#define PMM_METRIC_COSINE 0
#define PMM_METRIC_DOT 1
#define PMM_METRIC_EUCLIDEAN 2

/*
 * One embedding column in Arrow layout — what `series_to_matrix[_f32]` (src/matmul.rs:131-164) and
 * `try_extract_contiguous_*` (src/matmul.rs:39-95) read from a Polars Series.
 *   offsets == NULL : fixed-size rows (pl.Array / FixedSizeList): row i = values[i*dim .. (i+1)*dim)
 *   offsets != NULL : pl.List as 64-bit offsets (LargeList): row i = values[offsets[i] .. offsets[i+1]);
 *                     `dim` = length of row 0 (src/matmul.rs:237-239); shorter rows are zero padded;
 *                     a longer row is an error here (the reference panics, ndarray out-of-bounds).
 *   validity        : Arrow LSB validity bitmap over the child values (bit p for values[p]) or NULL;
 *                     a null element reads as 0.0 (src/matmul.rs:192,224,251,280).
 *   row_validity    : validity bitmap over rows or NULL; a null row reads as zeros (src/matmul.rs:248).
 * For the host entry points all pointers are host pointers; for the pmm_dev_* entry points all
 * pointers are device pointers.
 */
/* pmm_matrix_t::reserved flag: all buffers of the descriptor are device pointers already (fixed-size rows, no
 * bitmaps). Accepted for the QUERIES of pmm_topk_shard: a multi-GPU driver uploads the replicated query batch once
 * and broadcasts it over NVLink instead of sending it through the host link of every rank. */
#define PMM_MATRIX_ON_DEVICE 1
/* pmm_matrix_t::reserved flag: the column's rows live in SEVERAL host buffers (a multi-chunk Polars Series; the
 * reference's zero-copy path gives up there, `cont_slice` at src/matmul.rs:53, and concatenates on the host).
 * `values` then points to a pmm_chunks_t; fixed-size rows only (offsets / validity / row_validity NULL); n_rows = total.
 * Accepted by pmm_topk, pmm_matmul, pmm_corpus_create and pmm_topk_corpus (queries): the chunks are uploaded one after
 * the other into one device buffer. */
#define PMM_MATRIX_CHUNKED 2
typedef struct pmm_chunk {
    const void *values; /* n_rows * dim elements of the column's dtype */
    int64_t n_rows;
} pmm_chunk_t;
typedef struct pmm_chunks {
    int64_t n_chunks;
    const pmm_chunk_t *chunks;
} pmm_chunks_t;

typedef struct pmm_matrix {
    const void *values;
    const int64_t *offsets;
    const uint8_t *validity;
    const uint8_t *row_validity;
    int64_t n_rows;
    int64_t dim;
    int32_t dtype; /* PMM_DTYPE_* */
    int32_t reserved; /* flags: 0, or PMM_MATRIX_ON_DEVICE where an entry point says it accepts it */
} pmm_matrix_t;

/* ------------------------------------------------------------------ small helpers */

/* Metric::from_str, src/metrics.rs:19-27: case-insensitive, "l2" == euclidean.
 * Error text: "Unknown metric: '<name>'. Supported: cosine, dot, euclidean". */
PMM_API int pmm_metric_from_str(const char *name, int32_t *metric);

/* Metric::higher_is_better, src/metrics.rs:30-35. Returns 1/0. */
PMM_API int pmm_higher_is_better(int32_t metric);

/* is_f32_series x2, src/matmul.rs:13-19,308,427. Returns PMM_DTYPE_F32 or PMM_DTYPE_F64. */
PMM_API int pmm_working_dtype(int32_t left_dtype, int32_t right_dtype);

/* ------------------------------------------------------------------ host entry points (the boundary) */

/*
 * Replaces `_topk` (src/lib.rs:33-55) -> `topk_impl` (src/matmul.rs:473-519) ->
 * `compute_topk_indices_scores` (src/matmul.rs:420-469) -> `compute_similarity_matrix[_f32]`
 * (src/metrics.rs:258,314) -> `select_topk_with_scores[_f32]` (src/topk.rs:6,42).
 *
 *   k_eff = min(k, corpus->n_rows)  (src/matmul.rs:443), written to *k_actual.
 *   out_index [queries->n_rows * k_eff]  u32 corpus row numbers (`idx as u32`, src/matmul.rs:506)
 *   out_score [queries->n_rows * k_eff]  f64 (f32 working scores widened exactly, src/matmul.rs:447)
 *   Row i of both = the i-th query's matches, best first; ties: lower corpus index first; NaN last.
 * These are exactly the two child buffers of the List[Struct{index:u32, score:f64}] the reference
 * builds row by row (src/matmul.rs:497-518); list offsets are i*k_eff.
 *
 * Order of checks follows the reference: zero queries -> PMM_OK with *k_actual = min(k, N) and nothing
 * written, BEFORE the metric is parsed (src/matmul.rs:480-490); then metric; then "Empty series";
 * then "Dimension mismatch: left has {} dimensional vectors, right has {} dimensional vectors".
 * k < 0 -> PMM_ERR_INVALID (PyO3 rejects negative usize with OverflowError before the call).
 */
PMM_API int pmm_topk(const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t k, const char *metric,
             uint32_t *out_index, double *out_score, int64_t *k_actual);

/*
 * Replaces `_matmul` (src/lib.rs:15-30) -> `matmul_impl` (src/matmul.rs:295-417) ->
 * `matmul_slice_f32/_f64` (src/metrics.rs:160,111) / `matmul_f32/_f64` (src/metrics.rs:204,40).
 * out [left->n_rows * right->n_rows], row-major, element type = pmm_working_dtype(left, right):
 * the flat buffer the reference reshapes to Array[T, N] (src/matmul.rs:100-125); `flatten=True`
 * is the same buffer read as 1-D (python/polars_matmul/__init__.py:173-187).
 * Zero left rows -> PMM_OK, nothing written (src/matmul.rs:297-305).
 */
PMM_API int pmm_matmul(const pmm_matrix_t *left, const pmm_matrix_t *right, void *out);

/* ------------------------------------------------------------------ resident corpus
 * The reference re-marshals the corpus on every map_batches call (src/matmul.rs:430-431).  A handle
 * keeps the prepared corpus (tensor-core operand planes + norms) in HBM across calls. */
typedef struct pmm_corpus pmm_corpus_t;

/* `corpus` holds host pointers. `query_dtype` fixes the working precision (pmm_working_dtype). */
PMM_API int pmm_corpus_create(const pmm_matrix_t *corpus, int32_t query_dtype, pmm_corpus_t **out);
PMM_API int pmm_corpus_destroy(pmm_corpus_t *corpus);
PMM_API int64_t pmm_corpus_rows(const pmm_corpus_t *corpus);
/* Same contract as pmm_topk, corpus taken from the handle (host query buffers, host outputs). */
PMM_API int pmm_topk_corpus(const pmm_matrix_t *queries, const pmm_corpus_t *corpus, int64_t k,
                    const char *metric, uint32_t *out_index, double *out_score, int64_t *k_actual);

/* ------------------------------------------------------------------ device entry points
 * Same computations with every pointer in device memory (HBM-resident inputs and outputs), enqueued
 * on `stream`.  Used by the multi-GPU driver (one process per GPU) and for kernel-only measurement. */

/*
 * Local top-k of `queries` against `corpus` (a shard). Corpus row j is reported as index
 * index_base + j (global index of a sharded corpus).
 *   d_index / d_score : [Q * k_eff] final outputs, or both NULL
 *   d_candidates      : [Q * k_eff] packed candidates or NULL.  A candidate is one u64:
 *                       (ordered score key << 32) | ~index ; larger = better under the total order
 *                       (score best-first, lower index first), so shards merge by plain u64 max.
 *                       f32 working precision only.
 * Path: k_eff <= 248 runs the fused tcgen05 filter (operands rounded to f16 / TF32 planes) + exact re-scoring in the
 * working precision (f32 or f64; sequential FMA like the reference) + a per-query losslessness proof; f32 with
 * 248 < k_eff <= 2000 runs several passes of it; anything larger (or f64 beyond 248) takes the SIMT score-slab path.
 * Either way scores and indices equal the reference arithmetic's.
 */
PMM_API int pmm_dev_topk(const pmm_matrix_t *d_queries, const pmm_matrix_t *d_corpus, int64_t k, int32_t metric,
                 int64_t index_base, uint32_t *d_index, double *d_score, uint64_t *d_candidates,
                 void *stream);

/* Host inputs (this rank's corpus shard, replicated queries) -> this rank's exact local top-k as packed
 * candidates [Q * min(k, shard rows)] LEFT IN DEVICE MEMORY at d_candidates, ready for the all-gather.
 * The shard upload overlaps the compute (chunked H2D). Synchronous: the candidates are complete on return. */
PMM_API int pmm_topk_shard(const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard, int64_t k, int32_t metric,
                           int64_t index_base, uint64_t *d_candidates);

/* K-way merge of `n_lists` candidate lists laid out [n_lists][n_queries][k_in] (each sorted best
 * first, as pmm_dev_topk writes them; e.g. the all-gathered shards) into the final
 * [n_queries * k_out] index/score buffers.  k_out <= k_in <= 256. */
PMM_API int pmm_dev_merge_candidates(const uint64_t *d_lists, int64_t n_lists, int64_t n_queries, int64_t k_in,
                             int64_t k_out, int32_t metric, uint32_t *d_index, double *d_score,
                             void *stream);

/* Raw left * right^T into d_out [Q*N] of the working dtype. */
PMM_API int pmm_dev_matmul(const pmm_matrix_t *d_left, const pmm_matrix_t *d_right, void *d_out, void *stream);

/* Row norms (squared != 0: squared norms) in the storage-derived working dtype (f16 -> f32),
 * compute_norms_* / compute_squared_norms_*, src/metrics.rs:367-393. d_out [n_rows]. */
PMM_API int pmm_dev_norms(const pmm_matrix_t *d_x, int32_t squared, void *d_out, void *stream);

/* ------------------------------------------------------------------ groups of GPUs (north_star item 6)
 * The corpus partitions by rows: every GPU of a group scans its shard for ALL queries with the fused kernel and emits
 * its exact local top-k as packed candidates with global row numbers; the GPUs exchange candidates with NCCL over
 * NVLink (an all-to-all: GPU g receives only the rows of the queries it merges, [g Q/G, (g+1) Q/G)) and merge by
 * u64 max.  The total order of packed candidates makes the result independent of the number of shards.
 * libnccl is opened at run time (dlopen); without it every group function returns PMM_ERR_UNSUPPORTED.
 *
 * Two process models:
 *   single process  pmm_group_init_local: ncclCommInitAll over the first n visible devices (0 = all), one persistent
 *                   host thread per GPU.  pmm_topk / pmm_matmul use such a group on their own for large calls (options
 *                   "multi_gpu", "multi_gpu_min_gflop"); pmm_group_topk runs one call on a given group.
 *   process per GPU pmm_group_unique_id on one rank, distributed by the host application (MPI, torch.distributed, a
 *                   file ...), then pmm_group_init_rank on every rank with its device current; pmm_group_topk_shard is
 *                   collective: every rank passes ITS corpus shard and the same queries, k and metric.
 * f32 working precision, k <= 248. */
typedef struct pmm_group pmm_group_t;
#define PMM_GROUP_ID_BYTES 128
PMM_API int pmm_group_unique_id(void *id /* PMM_GROUP_ID_BYTES */);
PMM_API int pmm_group_init_rank(const void *id, int32_t rank, int32_t world, pmm_group_t **out);
PMM_API int pmm_group_init_local(int32_t n_devices, pmm_group_t **out);
PMM_API int pmm_group_destroy(pmm_group_t *group);
PMM_API int pmm_group_size(const pmm_group_t *group);

/* Same contract as pmm_topk (host buffers in and out), the corpus rows sharded over the GPUs of a single-process group;
 * the query batch crosses one host link and is broadcast over NVLink; every GPU writes its slice of the result. */
PMM_API int pmm_group_topk(pmm_group_t *group, const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t k,
                           const char *metric, uint32_t *out_index, double *out_score, int64_t *k_actual);

/* flags of pmm_group_topk_shard: one output mode ... */
#define PMM_GROUP_OUT_HOST_SLICE 1   /* out_* = host buffers [Q x k_eff]; this rank fills only the rows of ITS query slice */
#define PMM_GROUP_OUT_DEVICE_FULL 2  /* out_* = device buffers [Q x k_eff]; every rank receives the full result */
#define PMM_GROUP_OUT_DEVICE_SLICE 3 /* out_* = device buffers [Q x k_eff]; this rank fills only its query slice */
#define PMM_GROUP_OUT_HOST_FULL 4    /* out_* = host buffers [Q x k_eff]; every rank reads the full result back */
/* ... optionally OR-ed with: */
#define PMM_GROUP_QUERIES_FROM_ROOT 256 /* host queries (dense rows): only rank 0 uploads them, NCCL broadcast to the rest */
/* Query slice of rank r: rows [min(Q, r * ceil(Q/G)), min(Q, (r+1) * ceil(Q/G))).
 * queries / corpus_shard: host descriptors, or device-resident ones (reserved = PMM_MATRIX_ON_DEVICE, fixed-size rows).
 * corpus_shard may have zero rows (more ranks than rows); index_base = global row number of the shard's row 0;
 * n_total = rows of the whole corpus; k_eff = min(k, n_total). Collective and synchronous. */
PMM_API int pmm_group_topk_shard(pmm_group_t *group, const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard,
                                 int64_t index_base, int64_t n_total, int64_t k, int32_t metric, int32_t flags,
                                 uint32_t *out_index, double *out_score);

/* ------------------------------------------------------------------ diagnostics of the tensor-core filter
 * The top-k path selects candidates with a reduced-precision tensor-core FILTER and proves per query that nothing
 * relevant was dropped; that proof rests on an error bound.  These two entry points let a test measure the filter's
 * real error against the very bound the proof uses (tests/test_gpu_bound.py).
 *   level 0: the default first level (operands rounded to f16, kind::f16; exact planes when both inputs are f16),
 *         1: one TF32 MMA per K-step, 3: the 3xTF32 split.
 * pmm_dev_filter_candidates: d_kept [Q * kp] packed candidates whose key is the FILTER value (a float in "filter
 * units": dot q.c; cosine q.c/|c|; euclidean -|q-c|^2), sorted best first; kp in {32,64,128,256}; device pointers.
 * pmm_filter_error_bound: the bound E on |filter value - exact value| in the same units for a query of norm q_norm
 * against a corpus whose row norms lie in [c_norm_min, c_norm_max]; *max_norm (may be NULL) receives the operand
 * norm above which the level's proof refuses to trust the filter at all (0 = no limit). */
PMM_API int pmm_dev_filter_candidates(const pmm_matrix_t *d_queries, const pmm_matrix_t *d_corpus, int32_t metric, int32_t level,
                                      int32_t kp, int64_t index_base, uint64_t *d_kept, void *stream);
PMM_API int pmm_filter_error_bound(int32_t level, int32_t q_dtype, int32_t c_dtype, int64_t dim, int32_t metric, float q_norm,
                                   float c_norm_max, float c_norm_min, float *bound, float *max_norm);

/* ------------------------------------------------------------------ runtime */
PMM_API const char *pmm_last_error(void);     /* thread-local, never NULL */
PMM_API const char *pmm_version(void);
PMM_API int pmm_device_count(void);           /* 0 when no CUDA device is usable */
PMM_API int pmm_set_device(int32_t device);   /* device used by the calling thread's subsequent calls */

/* The CUDA stream (cudaStream_t) the host entry points and the group calls of the CALLING THREAD enqueue their work on,
 * for the thread's current device - e.g. to record timing events around synchronous calls. NULL without a device. */
PMM_API void *pmm_thread_stream(void);

/* Page-locked host memory for result buffers. The reference returns freshly allocated Vecs (from_vec,
 * src/matmul.rs:100-125, :497-518); a binding that hands this memory to pmm_topk / pmm_matmul as out_index /
 * out_score / out gets the device->host copy at PCIe rate instead of through the driver's pageable staging
 * (C3: 23 ms of a 165 ms call). cudaHostAlloc / cudaFreeHost underneath; allocation is slow (tens of ms per
 * 100 MB), so bindings should pool the blocks (polars_matmul_b200/_native.py does). */
PMM_API int pmm_host_alloc(int64_t bytes, void **out);
PMM_API int pmm_host_free(void *p);

/* Number of kernels this library has launched since load / since the last reset (process-wide). */
PMM_API int64_t pmm_kernel_launch_count(void);
PMM_API void pmm_reset_kernel_launch_count(void);

/* Tuning / diagnostics.  pmm_set_option changes the process-wide DEFAULTS; pmm_set_thread_option overrides a key for
 * the calling thread only (key == NULL drops the thread's overrides).  Every compute call takes one consistent snapshot
 * (defaults + the calling thread's overrides) when it starts, so options never change under a running call and
 * concurrent callers may use different settings.  Known keys:
 *   "force_generic" (0/1)      route top-k through the exact SIMT scores+select path (a Q x N slab in chunks)
 *   "profile" (0/1)            bracket kernels with CUDA events on the launching stream and accumulate per-kernel
 *                              milliseconds, read back with pmm_get_stat("<kernel>_ms")
 *   "verify" (0/1, default 1)  per-query losslessness proof of the tensor-core filter and its fallbacks
 *   "tc_levels" (1|2|3)        3 (default): first level on operands rounded to f16 (kind::f16), 3xTF32 on demand;
 *                              2: TF32 x1 first level; 1: 3xTF32 only
 *   "f16r_wide" (0/1, default 1)  queries the f16-rounded first level cannot prove are first re-run against the same
 *                              planes with 256-entry lists, then with 3xTF32
 *   "seed_retry" (0/1, default 1) re-query levels start from thresholds seeded by the exact k-th scores at hand
 *   "pipeline" (0/1/2, default 0) large query batches: one filter launch per round of query tiles, the merge and exact
 *                              re-scoring of a round overlap the filter of the next one (second stream).  Off by default:
 *                              the part runs against its power cap during the filter, so the overlapped re-scoring slows
 *                              the filter by what it saves (measured: 127.2 -> 126.7 ms at C3, profiles/sweep_r2.md);
 *                              2 = per-round launches without the overlap (measurement);
 *                              "pipeline_min_gflop" (default 2000): rounds below this much work are not split off;
 *                              "rescore_stream_loads" (default 1): the overlapped re-scoring gathers with evict-first loads
 *   "multipass" (0/1, default 1)  f32 top-k with 248 < k <= 2000: several passes of the fused filter (256 candidates each,
 *                              pass p+1 below the worst candidate pass p kept) instead of the score-slab path
 *   "f64_tc" (0/1, default 1)  f64 top-k: tensor-core filter + exact f64 re-scoring; 0: DMMA score slab + select
 *   "tc_cg" (1|2)              tcgen05 cta_group of the fused kernels (default 2)
 *   "tc_group"                 CTA groups sharing a query tile (0 = auto)
 *   "tc_sync_tiles"            corpus tiles between the producers' pacing barriers (default 32, 0 = off)
 *   "tc_sync_slack"            pacing: wait for the sync point this many points back (default 0)
 *   "tc_max_flush"             list merges per epilogue warp and tile once thresholds settled (0 = auto)
 *   "tc_soft_at"               staged candidates of a row that trigger its end-of-tile merge (0 = 48)
 *   "tc_clm", "tc_cluster4", "tc_max_units"   experimental cluster shapes, see profiles/sweep_r1.md (default off)
 *   "tc_debug_skip"            8: correct results + wait-cycle counters readable as pmm_get_stat("tc_dbg_wait<i>"), i = 0..51.
 *                              1..3 (kernel timing experiments that return WRONG results) exist only in builds made
 *                              with -DPMM_DIAG; the default build answers PMM_ERR_UNSUPPORTED
 *   "host_chunked" (0/1)       overlapped chunked corpus upload of the host entry points (default 1)
 *   "host_chunk_first_div", "host_chunk_ratio_pct"   first chunk = N / div (default 32); growth ratio in % (0 = auto)
 *   "host_chunk_min_rows", "host_chunk_min_mb"       smallest chunk (default 16384 rows) and smallest corpus (default
 *                              64 MB) of the chunked upload; tests lower both to drive tiny corpora through it
 *   "f64_simt" (0/1)           f64 contraction (raw matmul, slab path) on FP64 FMA instead of DMMA
 *   "generic_workspace_mb"     score slab of the SIMT path
 *   "matmul_tc_max_dim"        longest vector the raw matmul sends through the tensor cores (0 = auto: 256 for f32,
 *                              1024 for f16 storage; longer vectors use the exact sequential-FMA kernel)
 *   "matmul_split16" (0/1, default 1)  f32 operands with 32 < D <= 256: row-scaled hi/lo f16 planes, small terms accumulated
 *                              first (1), or the 3xTF32 planes (0); D <= 32 always uses the TF32 planes
 *   "matmul_exact_max_dim"     f32 vectors this short (default 8) take the exact kernel: a 22-bit split does not average out
 *   "matmul_flat" (-1/0/1/2)   tile schedule of the tensor-core matmul: classic (0), flat (1), hybrid (2), automatic (-1, default)
 *   "warm_seed" (0/1, default 1)  large top-k calls start the first filter level from thresholds of a sample pre-pass
 *                              (DESIGN 4.2); "warm_rows" (default 4096: cap of the sample size), "warm_rank" (0 = auto)
 *   "d2h_direct" (0/1, default 1)  host top-k with page-locked result buffers: the re-scoring kernel stores straight into them
 *   "multi_gpu" (0/1, default 1), "multi_gpu_min_gflop" (default 4000)   pmm_topk / pmm_matmul spread one call over all
 *                              visible GPUs when the call has at least that many GFLOP of contraction work
 * Acting immediately, process-wide (not part of the snapshot):
 *   "stage" (0/1, default 1)   pageable host buffers go through the library's page-locked staging ring (pmm_stage.h);
 *                              0: plain cudaMemcpyAsync, staged by the driver
 *   "stage_threads"            host threads per staged copy (0 = auto: half the cores, at most 8)
 *   "stage_nt" (0/1, default 1)  non-temporal stores for the copies into the ring (a third less host-memory traffic)
 *   "f64_dmma_async"           variant of the f64 DMMA matmul kernel (default 3: 16 x 32 warp tiles, four blocks per SM,
 *                              cp.async operand ring; 0: the register-staged round-1 kernel; see pmm_generic.cu)
 *   "stage_slot_mb", "stage_slots"   ring geometry (default 4 slots of 32 MB per calling thread)
 *   "prep_fast" (0/1, default 1)  128-bit loads / stores in the plane-building pass where the layout allows
 *   "rescore_fixed" (0/1, default 0)  cp.async gather in the exact re-scoring of plain f32 corpora (faster kernel, slower
 *                              step on a power-capped part: profiles/sweep_r2.md)
 *   "workspace_cache_mb"       device blocks (>= 32 MB) a thread keeps parked between calls for reuse (default 24576)
 *   "release_workspace"        return the calling thread's parked device blocks and its staging ring
 * Unknown keys return PMM_ERR_INVALID. */
PMM_API int pmm_set_option(const char *key, int64_t value);
PMM_API int pmm_set_thread_option(const char *key, int64_t value);
/* Accumulated statistics since the last pmm_reset_stats(): "<kernel>_ms", "<kernel>_launches",
 * "h2d_bytes", "d2h_bytes", "staged_h2d_bytes", "staged_d2h_bytes" (the part that went through the staging ring),
 * "requeried_f16_wide", "requeried_tf32x3", "fallback_queries". Unknown name -> 0. */
PMM_API double pmm_get_stat(const char *name);
PMM_API void pmm_reset_stats(void);

#ifdef __cplusplus
}
#endif
#endif /* PMM_H_ */
