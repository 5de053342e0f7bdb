// pmm_kernels.h — internal launcher interface between the kernel translation units and pmm_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pmm {

// ---- prep (pmm_prep.cu) -------------------------------------------------------------------------
struct PrepArgs {
    const void *values;           // device, storage dtype
    const int64_t *offsets;       // device or NULL
    const uint8_t *validity;      // device or NULL
    const uint8_t *row_validity;  // device or NULL
    int64_t n_rows;               // real rows
    int64_t dim;                  // real vector length
    int64_t rows_out;             // rows written (>= n_rows; padding rows are zeros)
    int64_t ld_out;               // leading dimension of the planes (>= dim; padding columns zeros)
    void *out0;                   // dense / hi plane / f16 plane
    void *out1;                   // lo plane (MODE_TF32) or NULL
    void *norm_out;               // [rows_out] working type or NULL
    void *sqnorm_out;             // [rows_out] working type or NULL
    float *norm32_out;            // [rows_out] f32 copy of the norms (f64 working precision + planes) or NULL
    float *sqnorm32_out;          // [rows_out] f32 copy of the squared norms or NULL
    float zero_guard_sq;          // rows with a squared norm at or below this do not enter the "smallest norm" statistic
    unsigned int *max_sq_out;     // f32 working type only: atomicMax of the squared norms' bit patterns, or NULL
    int *error_flag;              // set to 1 when a list row is longer than dim
    unsigned char *nonfinite_rows;  // [n_rows] set to 1 for rows holding an inf / NaN element, or NULL
    unsigned int *nonfinite_count;  // number of such rows (atomicAdd), with nonfinite_rows
    float *scale_out;               // PREP_SPLIT16: [rows_out] factor 2^-e that undoes the row's power-of-two scaling
};
// PREP_SPLIT16 (raw f32 matmul): every row scaled by a power of two so that its largest element lands in [2^14, 2^15),
// then split into two f16 planes hi = f16(x), lo = f16(x - hi): 22 significant bits like the TF32 split, at the f16 rate.
// Rows the scaling cannot serve (inf / NaN elements, a largest element outside [2^-60, 2^60]) are written as zeros and
// marked in nonfinite_rows; the caller recomputes them with IEEE arithmetic.
enum { PREP_DENSE = 0, PREP_TF32 = 1, PREP_F16 = 2, PREP_F16R = 3, PREP_SPLIT16 = 5 };  // = MODE_* of pmm_prep.cu
cudaError_t launch_prep(const PrepArgs &a, int src_dtype, int mode, int work_f64, cudaStream_t s);
cudaError_t launch_norms(const PrepArgs &a, int src_dtype, cudaStream_t s);
void prep_set_fast(bool on);   // vectorised fast path of the plane modes (default on)

// ---- generic SIMT path (pmm_generic.cu) -----------------------------------------------------------
// scores[i*ldo + j] = metric(dot(q_i, c_j)); metric < 0 => raw dot. Sequential FMA over the vector
// dimension per output (same order as the oracle).
cudaError_t launch_scores_f32(const float *q, const float *c, const float *qa, const float *ca, int64_t nq,
                              int64_t n, int64_t d, int metric, float *out, int64_t ldo, cudaStream_t s);
cudaError_t launch_scores_f64(const double *q, const double *c, const double *qa, const double *ca, int64_t nq,
                              int64_t n, int64_t d, int metric, double *out, int64_t ldo, cudaStream_t s);
// Same contract on the FP64 tensor path (mma.sync.m8n8k4.f64); agrees to ~1e-15 relative, not bit for bit.
cudaError_t launch_scores_f64_dmma(const double *q, const double *c, const double *qa, const double *ca, int64_t nq,
                                   int64_t n, int64_t d, int metric, double *out, int64_t ldo, cudaStream_t s);
void dmma_set_async(int mode);   // variant of the DMMA kernel (pmm_generic.cu; default 3: 16 x 32 warp tiles, four blocks per SM)
// Per-row exact top-k of a score slab (radix select + ordered tie collection + bitonic sort).
// scratch: nq * kpad * 12 bytes when kpad > select_smem_kpad_limit(), else unused.
int select_kpad(int64_t k);
int select_smem_kpad_limit(bool f64);
cudaError_t launch_select_f32(const float *scores, int64_t ld, int64_t nq, int64_t n, int64_t k, bool higher,
                              int64_t index_base, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                              void *scratch, cudaStream_t s);
cudaError_t launch_select_f64(const double *scores, int64_t ld, int64_t nq, int64_t n, int64_t k, bool higher,
                              int64_t index_base, uint32_t *out_idx, double *out_score, void *scratch,
                              cudaStream_t s);

// ---- candidate merge (pmm_merge.cu) ---------------------------------------------------------------
// Regular layout: list l of query q starts at lists + l*list_stride + q*row_stride, k_in entries.
cudaError_t launch_merge_regular(const uint64_t *lists, int64_t n_lists, int64_t list_stride, int64_t row_stride,
                                 int64_t nq, int k_in, int k_out, bool higher, uint32_t *out_idx,
                                 double *out_score, uint64_t *out_cand, cudaStream_t s);
// ---- exact re-scoring of kept candidates (pmm_rescore.cu) ---------------------------------------------
struct RawMatrix {               // a device-resident column in Arrow layout (pmm_matrix_t with device pointers)
    const void *values;
    const int64_t *offsets;
    const uint8_t *validity;
    const uint8_t *row_validity;
    int64_t n_rows, dim;
    int dtype;                   // 0 f16, 1 f32, 2 f64
};
// ---- error model of the tensor-core filter (shared by the re-scoring kernel, the host driver and the diagnostic
// entry points, so that tests measure the SAME bound the proof uses) -------------------------------------------------
// Relative error bound (per |q||c|) of a filter value against the exact working-precision score:
//   operand rounding: an 11-bit significand (TF32 via cvt.rna, or f32/f64 rounded to f16 inside the f16 normal range)
//                     has unit roundoff u = 2^-11 per operand, so one product is off by <= 2u + u^2 ~ 2^-10 and so is the
//                     sum (Cauchy-Schwarz): 9.8e-4.  3xTF32: each operand keeps a residue <= 2^-22 and the lo*lo term
//                     (<= 2^-22) is dropped: <= 3 * 2^-22 = 7.5e-7.  f16 planes of f16 input are exact: 0.
//   accumulation    : <= one f32 ulp (2^-23, truncation) of the running sum per tcgen05 accumulate step; a step
//                     covers 8 (TF32) or 16 (f16) elements of K and there are `terms` MMAs per step:
//                     D * 1.5e-8 * terms covers D/8 * 2^-23 per term (measured against adversarial inputs by
//                     tests/test_gpu_bound.py up to D = 8192);
//   the exact sum   : worst-case rounding of the sequential-FMA reference itself, D * 2^-24 = D * 6e-8.
inline float filter_eps(int64_t dim, int terms, bool exact_operands) {
    const float split = exact_operands ? 0.0f : terms == 1 ? 9.8e-4f : 7.5e-7f;
    return split + (float)dim * (1.5e-8f * (float)terms + 6.0e-8f) + 1e-6f;
}
// Rounding to f16 below the f16 normal range (|x| < 2^-14) is absolute, <= 2^-25 per element: <= sqrt(D) 2^-25 per row.
inline float f16r_abs_err(int64_t dim) { return sqrtf((float)dim) * 2.98023224e-8f * 1.0001f; }
// f64 sources at the TF32 level: elements below the f32 normal range (2^-126) may be flushed: <= sqrt(D) 2^-126 per row.
inline float f32_flush_abs_err(int64_t dim) { return sqrtf((float)dim) * 1.1754944e-38f * 1.0001f; }

// Bound E on |filter value - its exact counterpart| for one query, in the units of the filter value
// (dot: q.c; cosine: q.c / |c|; euclidean: squared distance).  eps = relative operand/accumulation error per |q||c|;
// s = absolute rounding error of one operand row: |q'.c' - q.c| <= eps |q||c| + s (|q| + |c|) + s^2.
// The small extra terms absorb the rounding of the metric pass.
__host__ __device__ inline float filter_error_bound(float eps, float s, int metric, float qn, float cmax, float cmin) {
    if (metric == 1 /* dot */) return eps * qn * cmax + s * (qn + cmax) + s * s;
    if (metric == 0 /* cosine */) return (eps + 1e-6f) * qn + (s > 0.0f ? s * qn / cmin + s + s * s / cmin : 0.0f);
    return 2.0f * (eps * qn * cmax + s * (qn + cmax) + s * s) + 1e-6f * (qn * qn + cmax * cmax);
}

// Inputs of the "was the filter lossless for this query" check (all device pointers; flags == NULL: no check).
struct RescoreCheck {
    const float *q_sq;             // [n_queries] squared query norms (f32; from f64 sources: rounded, inf when too large)
    const unsigned int *c_max_sq;  // [0] bits of the largest squared corpus norm (float >= 0, compared as uint),
                                   // [1] of the smallest one above 1e-12 (+inf bits if none)
    float eps;                     // relative error bound of the filter value vs the exact score, per |q||c|
    float abs_err;                 // absolute rounding error bound of one operand ROW (f16 subnormals), 0 if none
    float max_norm;                // operand rows with a larger norm may have overflowed the filter's format (0: no limit)
    const float *seed;             // [n_queries] seed thresholds the filter launch started from (filter units, NaN = none)
                                   // or NULL: everything at or below a row's seed was dropped, full list or not
    unsigned char *flags;          // [n_queries] set to 1 when not provable
    unsigned int *flag_count;
    float *kth_units;              // [n_queries] out (flagged queries): the exact k-th score in filter units, or NULL
    int stream_loads;              // gather the candidate rows with evict-first loads (re-scoring beside the fused kernel)
};
// cand [n_queries][kp_in] packed candidates (approximate keys, global indices) -> exact top-k_out.
cudaError_t launch_rescore(const uint64_t *cand, int kp_in, const RawMatrix &qm, const RawMatrix &cm,
                           const float *q_aux, const float *c_aux, int metric, int64_t index_base, int k_out,
                           uint32_t *out_idx, double *out_score, uint64_t *out_cand, const RescoreCheck &chk,
                           cudaStream_t s);
void rescore_set_fixed(bool on);   // cp.async gather of the re-scoring for plain f32 corpora (default on)
// Seeds for a re-query level: out[r] = kth_units[ids[r]] - E_next(query) - margin for r < n_ids, NaN for the padding
// rows [n_ids, n_pad).  `next` carries the NEXT level's eps / abs_err / max_norm and the norm inputs.
cudaError_t launch_seeds_from_lists(const uint64_t *lists, int kp, int r, int64_t n_queries, int64_t n_pad, float *out, cudaStream_t s);
cudaError_t launch_make_seeds(const int64_t *ids, int64_t n_ids, int64_t n_pad, const float *kth_units,
                              const RescoreCheck &next, int metric, float *out, cudaStream_t s);
// f64 working precision: exact f64 scores of the kept candidates (raw columns may be f16 / f32 / f64).
cudaError_t launch_rescore_f64(const uint64_t *cand, int kp_in, const RawMatrix &qm, const RawMatrix &cm, const double *q_aux,
                               const double *c_aux, int metric, int64_t index_base, int k_out, uint32_t *out_idx,
                               double *out_score, const RescoreCheck &chk, cudaStream_t s);
// Raw f32 matmul through the 3xTF32 split turns an INFINITE input element into NaN (0 * inf in the lo*hi term) where the
// reference propagates +-inf (src/metrics.rs:160-202).  The prep pass marks rows with non-finite elements; this kernel
// recomputes every output that involves a marked row with the reference's arithmetic (one FMA per element, sequential
// in d).  Launched unconditionally with a small grid; exits at once when nothing is marked - no host synchronisation.
cudaError_t launch_matmul_nonfinite_fixup(const RawMatrix &left, const RawMatrix &right, const unsigned char *nf_left,
                                          const unsigned char *nf_right, const unsigned int *nf_count, float *out, cudaStream_t s);
// Multi-pass top-k (k > 248): ceilings of the next pass = the worst candidate each query kept in this one
// (kept [nq][kp], sorted best first; 0 = list not full -> ceiling 0: nothing more to collect); rows [nq, n_pad) get 0.
cudaError_t launch_next_ceilings(const uint64_t *kept, int kp, int64_t nq, int64_t n_pad, uint64_t *ceil_out, cudaStream_t s);
// Final stage of the multi-pass top-k: per query the `n_lists` x `kp` EXACT packed candidates (each list sorted) are
// sorted as one (n_lists * kp <= 4096), the best k_out are emitted, and the losslessness of the whole collection is
// checked against the LAST pass's worst filter value (kept_last [nq][kp]): same proof as rescore_kernel.
cudaError_t launch_sort_lists(const uint64_t *lists, int n_lists, int64_t list_stride, int kp, int64_t nq, int k_out, int metric,
                              const uint64_t *kept_last, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                              const RescoreCheck &chk, cudaStream_t s);
cudaError_t launch_gather_rows(const RawMatrix &qm, const int64_t *ids, int64_t n_ids, void *out, int out_f64, cudaStream_t s);
cudaError_t launch_scatter_results(const int64_t *ids, int64_t n_ids, int k, const uint32_t *si, const double *ss,
                                   const uint64_t *sc, uint32_t *di, double *ds, uint64_t *dc, cudaStream_t s);

// ---- tensor-core path (pmm_tc_kernels.cu) ----------------------------------------------------------
constexpr int TC_TILE_M = 128;   // query rows per CTA tile
constexpr int TC_TILE_N = 256;   // corpus rows per tile

// Closed-form persistent schedule shared by the fused kernel and the merge that follows it.
// Work is done in rounds. In a full round `mc` query tiles are in flight, each scanned by `g` CTAs
// that take corpus tiles rank, rank+g, ... (so all CTAs sweep the corpus in lockstep and a corpus tile
// is fetched from HBM once per round and re-read from L2). The last round spreads the remaining
// `m_rem` query tiles over `g_rem` CTAs each.
struct TcSchedule {
    int m_tiles, n_tiles;
    int num_ctas;   // scheduling units launched (CTAs, or CTA pairs for cta_group::2)
    int g;          // CTAs per query tile in full rounds
    int mc;         // query tiles in flight in a full round (= num_ctas / g)
    int rounds;     // full rounds
    int m_full;     // = rounds * mc
    int m_rem;      // = m_tiles - m_full
    int g_rem;      // CTAs per query tile in the last round (0 when m_rem == 0)
    int flat;       // raw matmul (no lists to merge).  1: unit u takes tiles [T u / U, T (u+1) / U) of the row-major list of all
                    // T = m_tiles x n_tiles tiles; `rounds` = the most query tiles one unit touches, m_rem = 0.
                    // 2 (hybrid, G/2 < m_tiles < G units): unit m < m_tiles sweeps corpus tiles [0, n_main) of query tile m -
                    // all of them in the same order, like the classic schedule - and the remaining units share the tails
                    // [n_main, n_tiles) of all query tiles, so that no unit idles
    int n_main;     // flat == 2: corpus tiles per main sweep
    __host__ __device__ int pieces(int m_tile) const { return m_tile < m_full ? g : g_rem; }
    __host__ __device__ int64_t slot_base(int m_tile) const {
        return m_tile < m_full ? (int64_t)m_tile * g : (int64_t)m_full * g + (int64_t)(m_tile - m_full) * g_rem;
    }
    __host__ __device__ int64_t total_slots() const { return (int64_t)m_full * g + (int64_t)m_rem * g_rem; }
};
TcSchedule make_tc_schedule(int64_t q_rows, int64_t c_rows, int num_units, int group, int cg, int64_t layout_rows = 0);
TcSchedule make_tc_schedule_flat(int64_t q_rows, int64_t c_rows, int num_units, int cg);   // raw matmul: every unit busy, equal shares
// raw matmul, hybrid (TcSchedule::flat == 2); returns a classic schedule when the shape does not qualify
TcSchedule make_tc_schedule_hybrid(int64_t q_rows, int64_t c_rows, int num_units, int cg);
// Diagnostics: reads and clears the wait-cycle counters filled when TcArgs::debug_skip == 8.
void tc_debug_wait_cycles(unsigned long long out[52]);
int tc_epilogue_sets(int f16, int terms);  // epilogue warp sets of the top-k kernel variant: lists per (slot, row)
int64_t tc_staged_bytes(int num_ctas, int esets);  // size of TcArgs::staged
int64_t tc_sync_counters(const TcSchedule &s, int sync_tiles);  // number of pacing counters a launch needs

// Partial lists written by the fused kernel: [slot][cta of the group (cg)][row_in_tile (128)][kp].
cudaError_t launch_merge_tiles(const uint64_t *lists, TcSchedule sched, int cg, int esets, int kp, int64_t nq, int k_out, bool higher,
                               uint32_t *out_idx, double *out_score, uint64_t *out_cand, cudaStream_t s);

struct TcArgs {
    const void *q_hi, *q_lo;       // planes [q_rows_pad x dim_pad]  (f16 mode: q_hi only)
    const void *c_hi, *c_lo;       // planes [c_rows_pad x dim_pad]
    int64_t q_rows_pad, c_rows_pad, dim_pad;
    int64_t nq, n;                 // real rows
    int f16;                       // 1: kind::f16 (one plane; or the hi/lo f16 split with terms == 2), 0: TF32
    int cg;                        // tcgen05 cta_group: 1, or 2 (CTA pairs, UMMA M=256)
    int cluster4;                  // cg == 2, clm == 1: launch two independent pairs per cluster of 4
    int soft_at;                   // end-of-tile merge threshold in staged candidates (0 = default 48)
    int resume;                    // 1: keep the lists in `partial` (same schedule layout) and continue from their thresholds
    int sync_slack;                // pacing: wait for the sync point this many points back (0 = strict lockstep)
    int max_flush;                 // list merges per epilogue warp and tile (0 = auto from the tile's MMA time)
    int debug_skip;                // measurement only: 1 = epilogue loads TMEM but does not filter, 2 = one load per tile
    int clm;                       // CTA pairs per cluster: 1, or 2 (corpus tile multicast; needs cg == 2)
    int terms;                     // f32 top-k: 3 = 3xTF32 split, 1 = hi*hi only (first-level filter, needs cg == 2);
                                   // raw matmul with f16 == 1: 2 = hi/lo f16 split (PREP_SPLIT16 planes, q_aux / c_aux = the
                                   // rows' scale factors, needs cg == 2)
    TcSchedule sched;
    // top-k mode
    const float *q_aux, *c_aux;    // norms (cosine) / squared norms (euclidean) / NULL (dot)
    int64_t index_base;
    int metric;
    int k;                         // candidates kept per query and piece: the list position that sets the threshold (<= kp)
    int kp;                        // list capacity: 32, 64, 128 or 256
    uint64_t *partial;             // [sched.total_slots()][tc_epilogue_sets()][cg][128][kp]
    float norm_guard;              // cosine zero-norm guard of the filter (0 = 1e-6, the f32 reference constant)
    const float *seed_thr;         // top-k: [q_rows_pad] initial per-query thresholds in filter units (NaN = none) or NULL.
                                   // Everything with a filter value <= the seed is dropped from the start: the caller
                                   // must know that nothing at or below it can matter (re-query levels, pmm_api.cu)
    const uint64_t *ceil;          // top-k: [q_rows_pad] per-query ceilings (packed candidates) or NULL: only candidates strictly
                                   // below the ceiling are admitted - pass p+1 of a multi-pass top-k collects "the next
                                   // list-full below what pass p kept" (k > 248)
    int tile_stride;               // top-k: the schedule's corpus tile t is tile t * tile_stride of the planes (0 / 1: contiguous);
                                   // > 1 = a strided sample of the corpus (warm seeds): sched covers the SAMPLE's tiles, n the planes' rows
    uint64_t *staged;              // top-k: scratch of tc_staged_bytes(grid CTAs) bytes (unsorted candidates per CTA and row)
    // matmul mode
    float *out;                    // [nq x n] row-major
    unsigned int *round_sync;      // zeroed device counters for the producers' pacing barriers (tc_sync_counters()) or NULL
    int sync_tiles;                // corpus tiles between two pacing barriers
};
bool tc_supported();               // driver exposes cuTensorMapEncodeTiled and the device is sm_100
cudaError_t launch_tc_topk(const TcArgs &a, cudaStream_t s);
cudaError_t launch_tc_matmul(const TcArgs &a, cudaStream_t s);
const char *tc_last_error();

}  // namespace pmm
