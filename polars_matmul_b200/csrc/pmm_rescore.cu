// pmm_rescore.cu — exact f32 re-scoring of the candidates the tensor-core kernel kept.
//
// Why: tcgen05.mma accumulates in f32 with truncation; over the 3 x D/8 accumulate steps of a 3xTF32
// contraction the bias reaches ~1e-5 relative at D = 768 (measured), i.e. the stated
// tolerance.  So the fused kernel is used as a FILTER that keeps KP >= k + 8 candidates per query
// under its approximate scores, and this kernel recomputes the score of each kept candidate with the
// reference's arithmetic: one FMA per element, sequential in the vector dimension
// (src/metrics.rs:204-255 as restated by the oracle), then the metric pass of src/metrics.rs:323-362
// with the exact norms from pmm_prep.cu, then the final best-first order of src/topk.rs:42-75 under
// (score, lower index first).  Output scores are therefore bit-identical to the SIMT path's.
//
// One block per query, one thread per candidate; the query row sits in shared memory, each thread
// streams its candidate's corpus row (L1 keeps the sector remainder between iterations).
#include "pmm_common.cuh"
#include "pmm_kernels.h"

namespace pmm {

// floats per candidate row of a warp's transpose tile: 16-byte aligned rows whose float4 reads (one row per lane)
// and per-element writes (one column per lane) are both free of bank conflicts (pitch = 4 mod 32)
constexpr int RS_PITCH32 = 36;  // words per row: 32 f32 elements per step, or 32 f16 pairs (64 elements) for f16 rows

template <typename SRC> struct RsLoad;
template <> struct RsLoad<float> { static __device__ __forceinline__ float get(const void *v, int64_t p) { return __ldg((const float *)v + p); } };
template <> struct RsLoad<__half> { static __device__ __forceinline__ float get(const void *v, int64_t p) { return __half2float(__ldg((const __half *)v + p)); } };

template <typename SRC>
__device__ __forceinline__ float raw_fetch(const RawMatrix &m, int64_t base, int64_t len, int64_t i) {
    if (i >= len) return 0.0f;
    const int64_t p = base + i;
    if (m.validity && !((m.validity[p >> 3] >> (p & 7)) & 1)) return 0.0f;
    return RsLoad<SRC>::get(m.values, p);
}

__device__ __forceinline__ void raw_row(const RawMatrix &m, int64_t row, int64_t &base, int64_t &len) {
    if (m.offsets) {
        base = m.offsets[row];
        len = m.offsets[row + 1] - base;
        if (len > m.dim) len = m.dim;
    } else {
        base = row * m.dim;
        len = m.dim;
    }
    if (m.row_validity && !((m.row_validity[row >> 3] >> (row & 7)) & 1)) len = 0;
}

// One 32-element step of the gather: element `e` of candidate rows [0, n_act) of the warp, in batches of 8 rows
// (a warp-uniform guard per batch, no branch per row: every batch is in flight before the first value is consumed).
// PLAIN: no element validity bitmap (the usual case) -> one predicated load per row.
// STREAM: evict-first loads (ld.global.cs) - the gathered rows are used once; keeps them from displacing the fused
// kernel's L2-resident operand planes when the re-scoring runs beside it (pipelined first level).
template <typename CSRC, bool PLAIN, bool STREAM>
__device__ __forceinline__ void gather_step(float (&x)[32], const RawMatrix &cm, const int64_t *rowb, const int *rowl, int w0,
                                            int n_act, int e) {
#pragma unroll
    for (int g = 0; g < 32; g += 8) {
        if (g < n_act) {
#pragma unroll
            for (int i = g; i < g + 8; ++i) {
                const int64_t cbi = rowb[w0 + i];
                const int cli = rowl[w0 + i];
                if (PLAIN && STREAM && sizeof(CSRC) == 4) x[i] = e < cli ? __ldcs((const float *)cm.values + cbi + e) : 0.0f;
                else if (PLAIN) x[i] = e < cli ? RsLoad<CSRC>::get(cm.values, cbi + e) : 0.0f;
                else x[i] = raw_fetch<CSRC>(cm, cbi, cli, e);
            }
        }
    }
}

// FIXED (f32 corpus, fixed-size rows, no bitmaps, dim % 4 == 0, 16-byte aligned rows - the C3 / C4 case): the gather runs
// on cp.async.  Eight lanes copy one 128-byte line of a candidate row (16 bytes each) straight into the warp's
// shared-memory tile, so ONE instruction moves the lines of four candidates and a lane only ever needs the row pointers
// of its own 8 candidates - they live in registers.  Per 32-element step a warp issues 8 copies instead of 32 loads +
// 64 shared-memory reads of row offsets + 32 stores: about 80 instead of 280 instructions.  Two tile stages per warp keep
// the next step's lines in flight under the FMAs.  Measured at C3 (profiles/sweep_r2.md): the kernel itself gets faster
// (9.8 -> 9.2 ms in the step) but the NEXT step's filter launch slows down by more (112.5 -> 114.7 ms) - the part runs
// against its power cap and the denser re-scoring eats the headroom the filter starts with - so the step is 1.3-1.8 ms
// SLOWER.  Kept behind pmm_set_option("rescore_fixed", 1), off by default.
constexpr int RS_STAGES = 2;
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void *src, unsigned src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename CSRC, int NT, bool FIXED = false>
__global__ void __launch_bounds__(NT) rescore_kernel(const uint64_t *__restrict__ cand, int kp_in, RawMatrix qm,
                                                     RawMatrix cm, const float *__restrict__ q_aux,
                                                     const float *__restrict__ c_aux, int metric,
                                                     int64_t index_base, int k_out, uint32_t *out_idx,
                                                     double *out_score, uint64_t *out_cand, RescoreCheck chk) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *sortbuf = (uint64_t *)rs_smem;            // NT entries; during the gather: the candidates' row offsets
    int *rowl = (int *)(rs_smem + NT * 8);              // NT row lengths
    float *qs = (float *)(rs_smem + NT * 12);           // dim floats, then one transpose tile per warp
    const int64_t q = blockIdx.x;
    const int t = threadIdx.x;
    const int dim = (int)qm.dim;
    {
        int64_t qb, ql;
        raw_row(qm, q, qb, ql);
        for (int i = t; i < dim; i += NT)
            qs[i] = qm.dtype == 0 ? raw_fetch<__half>(qm, qb, ql, i) : raw_fetch<float>(qm, qb, ql, i);
    }
    __syncthreads();
    const bool higher = higher_is_better(metric);
    const int lane = t & 31, wrp = t >> 5;
    uint64_t c_in = (t < kp_in) ? cand[q * kp_in + t] : 0ull;
    // Candidates that cannot reach the top k are not re-scored (their corpus rows are never fetched): the list is
    // sorted by filter value f, |exact - f| <= E (the same bound the losslessness check below uses), so a candidate
    // with f < f_k - 2E (f_k = k-th best filter value) is strictly worse than k others.  With exact-product filters
    // (f16 planes) this leaves about k of the KP candidates.
    if (chk.q_sq && chk.c_max_sq && k_out > 0 && t >= k_out && k_out <= kp_in && c_in != 0ull) {
        const uint64_t ck = cand[q * kp_in + k_out - 1];
        if (ck != 0ull) {
            const float f_k = key_score(candidate_key(ck), true), f_j = key_score(candidate_key(c_in), true);
            const float qn = sqrtf(chk.q_sq[q]);
            const float cmax = sqrtf(__uint_as_float(chk.c_max_sq[0])), cmin = sqrtf(__uint_as_float(chk.c_max_sq[1]));
            const float e = filter_error_bound(chk.eps, chk.abs_err, metric, qn, cmax, cmin);
            // the bound only holds while no operand left the filter format's range (f16-rounded level: a row norm
            // above 65504 may have become inf there, and the filter values are then meaningless)
            const bool in_range = chk.max_norm <= 0.0f || (qn <= chk.max_norm && cmax <= chk.max_norm);
            if (in_range && fabsf(f_k) <= 3.0e38f && f_j < f_k - 2.0f * e - 1e-6f * fabsf(f_k)) c_in = 0ull;  // (NaN anywhere: keep)
        }
    }
    const uint64_t c = c_in;
    const uint32_t gidx = candidate_index(c);
    const int64_t row = (int64_t)gidx - index_base;
    int64_t cb = 0, cl = 0;
    if (c != 0ull) raw_row(cm, row, cb, cl);
    // Each warp walks the vector dimension 32 (f16: 64) elements at a time: candidate rows are read with
    // coalesced 128-byte requests (lane = element), transposed through shared memory, and every thread then
    // accumulates ITS candidate sequentially in d — the reference's order.  Candidates are sorted best first and
    // the skipped ones sit at the end, so the warp only fetches rows [0, n_act).
    const int n_act = 32 - __clz(__ballot_sync(0xffffffffu, c != 0ull));
    // row offset / length of every candidate of the warp, read back as shared-memory broadcasts in the loops
    int64_t *rowb = (int64_t *)sortbuf;
    rowb[t] = cb;
    rowl[t] = (int)cl;
    __syncwarp();
    const int w0 = wrp * 32;
    float acc = 0.0f;
    unsigned char *tile_base = rs_smem + NT * 12 + (size_t)((dim + 3) & ~3) * 4;
    // f16 rows without nulls: 64 elements per step, one half2 per lane (128-byte requests per row)
    const bool wide16 = sizeof(CSRC) == 2 && !cm.offsets && !cm.validity && (dim & 1) == 0;
    if (FIXED) {
        float (*tiles)[32][RS_PITCH32] = (float (*)[32][RS_PITCH32])(tile_base + (size_t)wrp * RS_STAGES * 32 * RS_PITCH32 * 4);
        const int seg = lane & 7, cgrp = lane >> 3;       // my 16-byte segment of a line; my candidates are 4 j + cgrp
        const char *rp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rp[j] = (const char *)cm.values + (size_t)rowb[w0 + 4 * j + cgrp] * 4 + seg * 16;
        const int n_steps = (dim + 31) >> 5;
        auto issue = [&](int step, int stage) {
            const int e0 = step * 32 + seg * 4;             // first element of my segment (dim % 4 == 0: all or nothing)
            const unsigned sz = e0 < dim ? 16u : 0u;        // beyond the row: zero fill, nothing is read
            const size_t off = sz ? (size_t)step * 128 : 0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (4 * j + cgrp < n_act) cp_async_16(smem_addr(&tiles[stage][4 * j + cgrp][seg * 4]), rp[j] + off, sz);
            cp_async_commit();
        };
        issue(0, 0);
        for (int step = 0; step < n_steps; ++step) {
            const int d0 = step * 32;
            if (step + 1 < n_steps) {
                issue(step + 1, (step + 1) & 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float (*tile)[RS_PITCH32] = tiles[step & 1];
            const int jn = dim - d0 < 32 ? dim - d0 : 32;
            if (jn == 32) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 qv = *(const float4 *)&qs[d0 + j], cv = *(const float4 *)&tile[lane][j];
                    acc = __fmaf_rn(qv.x, cv.x, acc);
                    acc = __fmaf_rn(qv.y, cv.y, acc);
                    acc = __fmaf_rn(qv.z, cv.z, acc);
                    acc = __fmaf_rn(qv.w, cv.w, acc);
                }
            } else {
                for (int j = 0; j < jn; ++j) acc = __fmaf_rn(qs[d0 + j], tile[lane][j], acc);
            }
            __syncwarp();   // the stage is overwritten by the copies of step + 2
        }
    } else if (wide16) {
        // the tile keeps the rows as f16 pairs (36-word pitch like the f32 tile: half the shared memory and registers
        // of an upcast tile, so more blocks are resident); the upcast happens at the FMA
        __half2 (*tileh)[RS_PITCH32] = (__half2 (*)[RS_PITCH32])(tile_base + (size_t)wrp * 32 * RS_PITCH32 * 4);
        const __half2 *vals = (const __half2 *)cm.values;
        for (int d0 = 0; d0 < dim; d0 += 64) {
            __half2 x[32];
            // row reads are issued in batches of 8 candidates (a warp-uniform guard per batch, no branch per row):
            // every batch is in flight before the first value is consumed
#pragma unroll
            for (int g = 0; g < 32; g += 8) {
                if (g < n_act) {
#pragma unroll
                    for (int i = g; i < g + 8; ++i) {
                        const int64_t cbi = rowb[w0 + i];
                        const int cli = rowl[w0 + i];
                        const int e = d0 + 2 * lane;
                        x[i] = e < cli ? __ldg(vals + ((cbi + e) >> 1)) : __floats2half2_rn(0.0f, 0.0f);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < 32; g += 8) {
                if (g < n_act) {
#pragma unroll
                    for (int i = g; i < g + 8; ++i) tileh[i][lane] = x[i];
                }
            }
            __syncwarp();
            const int jn = dim - d0 < 64 ? dim - d0 : 64;
            if (jn == 64) {
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    const float4 q0 = *(const float4 *)&qs[d0 + j], q1 = *(const float4 *)&qs[d0 + j + 4];
                    const uint4 raw = *(const uint4 *)&tileh[lane][j >> 1];   // 8 consecutive f16 of my candidate
                    const float2 c0 = __half22float2(*(const __half2 *)&raw.x), c1 = __half22float2(*(const __half2 *)&raw.y);
                    const float2 c2 = __half22float2(*(const __half2 *)&raw.z), c3 = __half22float2(*(const __half2 *)&raw.w);
                    acc = __fmaf_rn(q0.x, c0.x, acc);
                    acc = __fmaf_rn(q0.y, c0.y, acc);
                    acc = __fmaf_rn(q0.z, c1.x, acc);
                    acc = __fmaf_rn(q0.w, c1.y, acc);
                    acc = __fmaf_rn(q1.x, c2.x, acc);
                    acc = __fmaf_rn(q1.y, c2.y, acc);
                    acc = __fmaf_rn(q1.z, c3.x, acc);
                    acc = __fmaf_rn(q1.w, c3.y, acc);
                }
            } else {
                const __half *th = (const __half *)&tileh[lane][0];
                for (int j = 0; j < jn; ++j) acc = __fmaf_rn(qs[d0 + j], __half2float(th[j]), acc);
            }
            __syncwarp();
        }
    } else {
        float (*tile)[RS_PITCH32] = (float (*)[RS_PITCH32])(tile_base + (size_t)wrp * 32 * RS_PITCH32 * 4);
        const bool plain = !cm.validity;
        for (int d0 = 0; d0 < dim; d0 += 32) {
            float x[32];
            if (plain && chk.stream_loads) gather_step<CSRC, true, true>(x, cm, rowb, rowl, w0, n_act, d0 + lane);
            else if (plain) gather_step<CSRC, true, false>(x, cm, rowb, rowl, w0, n_act, d0 + lane);
            else gather_step<CSRC, false, false>(x, cm, rowb, rowl, w0, n_act, d0 + lane);
#pragma unroll
            for (int g = 0; g < 32; g += 8) {
                if (g < n_act) {
#pragma unroll
                    for (int i = g; i < g + 8; ++i) tile[i][lane] = x[i];
                }
            }
            __syncwarp();
            const int jn = dim - d0 < 32 ? dim - d0 : 32;
            if (jn == 32) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 qv = *(const float4 *)&qs[d0 + j], cv = *(const float4 *)&tile[lane][j];
                    acc = __fmaf_rn(qv.x, cv.x, acc);
                    acc = __fmaf_rn(qv.y, cv.y, acc);
                    acc = __fmaf_rn(qv.z, cv.z, acc);
                    acc = __fmaf_rn(qv.w, cv.w, acc);
                }
            } else {
                for (int j = 0; j < jn; ++j) acc = __fmaf_rn(qs[d0 + j], tile[lane][j], acc);
            }
            __syncwarp();
        }
    }
    __syncwarp();  // the row offsets in sortbuf are dead from here on
    uint64_t packed = 0ull;
    if (c != 0ull) {
        float sc = acc;
        if (metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN) sc = metric_finish(acc, metric, q_aux[q], c_aux[row]);
        packed = pack_candidate(score_key(sc, higher), gidx);
    }
    sortbuf[t] = packed;
    __syncthreads();
    for (int size = 2; size <= NT; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int p = t ^ stride;
            if (p > t) {
                const bool desc = (t & size) == 0;
                const uint64_t a = sortbuf[t], b = sortbuf[p];
                if ((a > b) != desc) { sortbuf[t] = b; sortbuf[p] = a; }
            }
            __syncthreads();
        }
    }
    // ---- was the filter provably lossless for this query?  Every candidate the filter dropped has a filter
    // value <= f_last (the worst kept one). Its exact score can exceed what f_last maps to by at most the
    // filter's error bound; if the exact k-th score is not strictly better than that, a dropped candidate
    // might belong to (or tie into) the top-k: flag the query, the host recomputes it on the exact path.
    if (t == 0 && chk.flags) {
        // worst filter value that was kept = upper bound of everything dropped.  A list that is not full dropped
        // nothing - unless the launch started from a seed threshold, then everything at or below the seed is gone.
        const uint64_t last = cand[q * kp_in + kp_in - 1];
        float f_last = 0.0f;
        bool dropped = false;
        if (last != 0ull) {
            f_last = key_score(candidate_key(last), true);
            dropped = true;
        } else if (chk.seed && chk.seed[q] == chk.seed[q]) {
            f_last = chk.seed[q];
            dropped = true;
        }
        bool ok = true;
        const float tk = k_out > 0 ? key_score(candidate_key(sortbuf[k_out - 1]), higher) : 0.0f;
        const float qn = sqrtf(chk.q_sq[q]);
        if (dropped && k_out > 0) {
            const float cmax = sqrtf(__uint_as_float(chk.c_max_sq[0])), cmin = sqrtf(__uint_as_float(chk.c_max_sq[1]));
            const float e = filter_error_bound(chk.eps, chk.abs_err, metric, qn, cmax, cmin);
            if (metric == METRIC_DOT) {
                ok = tk > f_last + e;
            } else if (metric == METRIC_COSINE) {
                ok = qn > 1e-6f && tk > (f_last + e) / qn;
            } else {  // filter value = -(squared distance)
                const float sq_floor = -f_last - e;
                ok = sq_floor > 0.0f && tk < sqrtf(sq_floor) * (1.0f - 1e-6f);
            }
            // operands beyond the filter format's range (f32 rounded to f16: 65504) made the filter value meaningless
            if (chk.max_norm > 0.0f && !(qn <= chk.max_norm && cmax <= chk.max_norm)) ok = false;
            if (!(ok)) ok = false;  // NaN anywhere -> not provable
        }
        if (!ok && chk.kth_units)   // what a re-query level may seed its thresholds from (filter units)
            chk.kth_units[q] = metric == METRIC_DOT ? tk : metric == METRIC_COSINE ? tk * qn : -(tk * tk);
        if (!ok) {
            chk.flags[q] = 1;
            atomicAdd(chk.flag_count, 1u);
        }
    }
    if (t < k_out) {
        const uint64_t r = sortbuf[t];
        if (out_idx) out_idx[q * k_out + t] = candidate_index(r);
        if (out_score) out_score[q * k_out + t] = (double)key_score(candidate_key(r), higher);
        if (out_cand) out_cand[q * k_out + t] = r;
    }
}

static bool g_rescore_fixed = false;  // rescore_set_fixed(): off by default - see the note at rescore_kernel (measured: no net gain)
void rescore_set_fixed(bool on) { g_rescore_fixed = on; }

template <typename CSRC>
static cudaError_t launch_rescore_t(const uint64_t *cand, int kp_in, const RawMatrix &qm, const RawMatrix &cm,
                                    const float *q_aux, const float *c_aux, int metric, int64_t index_base, int k_out,
                                    uint32_t *out_idx, double *out_score, uint64_t *out_cand, const RescoreCheck &chk,
                                    cudaStream_t s) {
    const unsigned grid = (unsigned)qm.n_rows;
    const size_t smem_q = (size_t)((qm.dim + 3) & ~(int64_t)3) * 4;
    const bool fixed = sizeof(CSRC) == 4 && g_rescore_fixed && !cm.offsets && !cm.validity && !cm.row_validity && (cm.dim % 4) == 0 &&
                       (((uintptr_t)cm.values) & 15) == 0;
#define PMM_RS_LAUNCH(NT, FX, TILES)                                                                                 \
    {                                                                                                               \
        size_t smem = NT * 12 + smem_q + (size_t)(NT / 32) * (TILES) * 32 * RS_PITCH32 * 4;                         \
        if (smem > 48 * 1024) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(rescore_kernel<CSRC, NT, FX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                         \
        }                                                                                                           \
        rescore_kernel<CSRC, NT, FX><<<grid, NT, smem, s>>>(cand, kp_in, qm, cm, q_aux, c_aux, metric, index_base, k_out, \
                                                            out_idx, out_score, out_cand, chk);                     \
    }
#define PMM_RS(NT)                                                                                                  \
    {                                                                                                               \
        if (fixed) PMM_RS_LAUNCH(NT, true, RS_STAGES)                                                               \
        else PMM_RS_LAUNCH(NT, false, 1)                                                                            \
    }
    if (kp_in <= 32) PMM_RS(32)
    else if (kp_in <= 64) PMM_RS(64)
    else if (kp_in <= 128) PMM_RS(128)
    else if (kp_in <= 256) PMM_RS(256)
    else return cudaErrorInvalidValue;
#undef PMM_RS_LAUNCH
#undef PMM_RS
    return cudaGetLastError();
}

cudaError_t launch_rescore(const uint64_t *cand, int kp_in, const RawMatrix &qm, const RawMatrix &cm,
                           const float *q_aux, const float *c_aux, int metric, int64_t index_base, int k_out,
                           uint32_t *out_idx, double *out_score, uint64_t *out_cand, const RescoreCheck &chk, cudaStream_t s) {
    if (qm.n_rows <= 0 || k_out <= 0) return cudaSuccess;
    if (qm.dim * 4 > 160 * 1024) return cudaErrorInvalidValue;  // query row must fit shared memory
    if (cm.dtype == 0)
        return launch_rescore_t<__half>(cand, kp_in, qm, cm, q_aux, c_aux, metric, index_base, k_out, out_idx, out_score, out_cand, chk, s);
    return launch_rescore_t<float>(cand, kp_in, qm, cm, q_aux, c_aux, metric, index_base, k_out, out_idx, out_score, out_cand, chk, s);
}



// ------------------------------------------------------------------------------------------------------------------
// f64 working precision (what Polars hands the reference by default: every list-of-floats column is Float64).
// Same scheme: the tensor-core filter (operands rounded to f16 / TF32) only SELECTS; this kernel recomputes each kept
// candidate with the reference's f64 arithmetic - one FMA per element, sequential in the vector dimension
// (src/metrics.rs:40-97 as restated by the oracle), the metric pass of src/metrics.rs:267-308 with exact f64 norms,
// best-first order of src/topk.rs:6-39 under (score, lower index) - so f64 top-k results are bit-identical to the
// oracle's, and no Q x N score slab is ever written.  The raw columns may be f16 / f32 / f64 (mixed dtypes are
// widened exactly, src/matmul.rs:308).
__device__ __forceinline__ double raw_fetch_f64(const RawMatrix &m, int64_t base, int64_t len, int64_t i) {
    if (i >= len) return 0.0;
    const int64_t p = base + i;
    if (m.validity && !((m.validity[p >> 3] >> (p & 7)) & 1)) return 0.0;
    if (m.dtype == 2) return __ldg((const double *)m.values + p);
    if (m.dtype == 1) return (double)__ldg((const float *)m.values + p);
    return (double)__half2float(__ldg((const __half *)m.values + p));
}

constexpr int RS_PITCH64 = 33;  // doubles per candidate row of a warp's transpose tile

template <int NT>
__global__ void __launch_bounds__(NT) rescore_f64_kernel(const uint64_t *__restrict__ cand, int kp_in, RawMatrix qm, RawMatrix cm,
                                                         const double *__restrict__ q_aux, const double *__restrict__ c_aux,
                                                         int metric, int64_t index_base, int k_out, uint32_t *out_idx,
                                                         double *out_score, RescoreCheck chk) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *skey = (uint64_t *)rs_smem;                    // NT exact score keys; during the gather: row offsets
    uint32_t *sidx = (uint32_t *)(rs_smem + NT * 8);         // NT corpus indices;   during the gather: row lengths
    double *qs = (double *)(rs_smem + NT * 12 + ((NT * 12) % 8 ? 4 : 0));
    const int64_t q = blockIdx.x;
    const int t = threadIdx.x;
    const int dim = (int)qm.dim;
    {
        int64_t qb, ql;
        raw_row(qm, q, qb, ql);
        for (int i = t; i < dim; i += NT) qs[i] = raw_fetch_f64(qm, qb, ql, i);
    }
    __syncthreads();
    const bool higher = higher_is_better(metric);
    const int lane = t & 31, wrp = t >> 5;
    uint64_t c_in = (t < kp_in) ? cand[q * kp_in + t] : 0ull;
    // skip rule: as in rescore_kernel (the filter values and their bound are f32 quantities in both cases)
    if (chk.q_sq && chk.c_max_sq && k_out > 0 && t >= k_out && k_out <= kp_in && c_in != 0ull) {
        const uint64_t ck = cand[q * kp_in + k_out - 1];
        if (ck != 0ull) {
            const float f_k = key_score(candidate_key(ck), true), f_j = key_score(candidate_key(c_in), true);
            const float qn = sqrtf(chk.q_sq[q]);
            const float cmax = sqrtf(__uint_as_float(chk.c_max_sq[0])), cmin = sqrtf(__uint_as_float(chk.c_max_sq[1]));
            const float e = filter_error_bound(chk.eps, chk.abs_err, metric, qn, cmax, cmin);
            const bool in_range = chk.max_norm <= 0.0f || (qn <= chk.max_norm && cmax <= chk.max_norm);
            if (in_range && fabsf(f_k) <= 3.0e38f && f_j < f_k - 2.0f * e - 1e-6f * fabsf(f_k)) c_in = 0ull;
        }
    }
    const uint64_t c = c_in;
    const uint32_t gidx = candidate_index(c);
    const int64_t row = (int64_t)gidx - index_base;
    int64_t cb = 0, cl = 0;
    if (c != 0ull) raw_row(cm, row, cb, cl);
    const int n_act = 32 - __clz(__ballot_sync(0xffffffffu, c != 0ull));
    int64_t *rowb = (int64_t *)skey;
    int *rowl = (int *)sidx;
    rowb[t] = cb;
    rowl[t] = (int)cl;
    __syncwarp();
    const int w0 = wrp * 32;
    double acc = 0.0;
    double (*tile)[RS_PITCH64] = (double (*)[RS_PITCH64])((unsigned char *)(qs + ((dim + 1) & ~1)) + (size_t)wrp * 32 * RS_PITCH64 * 8);
    for (int d0 = 0; d0 < dim; d0 += 32) {
        // coalesced row reads (lane = element), 8 candidate rows in flight per batch, transposed through shared
        // memory; every thread then accumulates ITS candidate sequentially in d - the reference's order
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            if (g < n_act) {
                double x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = raw_fetch_f64(cm, rowb[w0 + g + i], rowl[w0 + g + i], d0 + lane);
#pragma unroll
                for (int i = 0; i < 8; ++i) tile[g + i][lane] = x[i];
            }
        }
        __syncwarp();
        const int jn = dim - d0 < 32 ? dim - d0 : 32;
        if (lane < n_act)
            for (int j = 0; j < jn; ++j) acc = __fma_rn(qs[d0 + j], tile[lane][j], acc);
        __syncwarp();
    }
    __syncwarp();  // row offsets / lengths are dead from here on
    uint64_t key = 0ull;
    uint32_t idx = 0xffffffffu;
    if (c != 0ull) {
        double sc = acc;
        if (metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN) sc = metric_finish(acc, metric, q_aux[q], c_aux[row]);
        key = score_key(sc, higher);
        idx = gidx;
    }
    __syncthreads();
    // an empty slot must rank after every real candidate, NaN scores (key 0) included: real ones get bit 0 of a
    // side flag through the index order (empty: index 2^32-1, and candidates never carry that index)
    skey[t] = key;
    sidx[t] = idx;
    __syncthreads();
    for (int size = 2; size <= NT; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int p = t ^ stride;
            if (p > t) {
                const bool desc = (t & size) == 0;
                const uint64_t ka = skey[t], kb = skey[p];
                const uint32_t ia = sidx[t], ib = sidx[p];
                const bool a_first = ka > kb || (ka == kb && ia < ib);
                if (a_first != desc) { skey[t] = kb; skey[p] = ka; sidx[t] = ib; sidx[p] = ia; }
            }
            __syncthreads();
        }
    }
    if (t == 0 && chk.flags) {
        const uint64_t last = cand[q * kp_in + kp_in - 1];
        float f_last = 0.0f;
        bool dropped = false;
        if (last != 0ull) {
            f_last = key_score(candidate_key(last), true);
            dropped = true;
        } else if (chk.seed && chk.seed[q] == chk.seed[q]) {
            f_last = chk.seed[q];
            dropped = true;
        }
        bool ok = true;
        const double tk = k_out > 0 ? key_score(skey[k_out - 1], higher) : 0.0;
        const float qn = sqrtf(chk.q_sq[q]);
        if (dropped && k_out > 0) {
            const float cmax = sqrtf(__uint_as_float(chk.c_max_sq[0])), cmin = sqrtf(__uint_as_float(chk.c_max_sq[1]));
            const double e = (double)filter_error_bound(chk.eps, chk.abs_err, metric, qn, cmax, cmin);
            if (metric == METRIC_DOT) {
                ok = tk > (double)f_last + e;
            } else if (metric == METRIC_COSINE) {
                const double th = ((double)f_last + e) / (double)qn;
                ok = qn > 1e-6f && tk > th + 1e-6 * fabs(th);
            } else {
                const double sq_floor = -(double)f_last - e;
                ok = sq_floor > 0.0 && tk < sqrt(sq_floor) * (1.0 - 1e-6);
            }
            if (chk.max_norm > 0.0f && !(qn <= chk.max_norm && cmax <= chk.max_norm)) ok = false;
            if (!(ok)) ok = false;
        }
        if (!ok && chk.kth_units)
            chk.kth_units[q] = (float)(metric == METRIC_DOT ? tk : metric == METRIC_COSINE ? tk * (double)qn : -(tk * tk));
        if (!ok) {
            chk.flags[q] = 1;
            atomicAdd(chk.flag_count, 1u);
        }
    }
    if (t < k_out) {
        if (out_idx) out_idx[q * k_out + t] = sidx[t];
        if (out_score) out_score[q * k_out + t] = key_score(skey[t], higher);
    }
}

cudaError_t launch_rescore_f64(const uint64_t *cand, int kp_in, const RawMatrix &qm, const RawMatrix &cm, const double *q_aux,
                               const double *c_aux, int metric, int64_t index_base, int k_out, uint32_t *out_idx,
                               double *out_score, const RescoreCheck &chk, cudaStream_t s) {
    if (qm.n_rows <= 0 || k_out <= 0) return cudaSuccess;
    if (qm.dim * 8 > 128 * 1024) return cudaErrorInvalidValue;  // query row must fit shared memory
    const unsigned grid = (unsigned)qm.n_rows;
    const size_t smem_q = (size_t)((qm.dim + 1) & ~(int64_t)1) * 8;
#define PMM_RS64(NT)                                                                                                       \
    {                                                                                                                      \
        size_t smem = NT * 12 + 8 + smem_q + (size_t)(NT / 32) * 32 * RS_PITCH64 * 8;                                     \
        if (smem > 48 * 1024) {                                                                                            \
            cudaError_t e = cudaFuncSetAttribute(rescore_f64_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                                \
        }                                                                                                                  \
        rescore_f64_kernel<NT><<<grid, NT, smem, s>>>(cand, kp_in, qm, cm, q_aux, c_aux, metric, index_base, k_out, out_idx, \
                                                      out_score, chk);                                                     \
    }
    if (kp_in <= 32) PMM_RS64(32)
    else if (kp_in <= 64) PMM_RS64(64)
    else if (kp_in <= 128) PMM_RS64(128)
    else if (kp_in <= 256) PMM_RS64(256)
    else return cudaErrorInvalidValue;
#undef PMM_RS64
    return cudaGetLastError();
}

// ---- multi-pass top-k (k > 248) ----------------------------------------------------------------------------------
__global__ void next_ceilings_kernel(const uint64_t *__restrict__ kept, int kp, int64_t nq, int64_t n_pad, uint64_t *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    out[r] = r < nq ? kept[r * kp + kp - 1] : 0ull;
}
cudaError_t launch_next_ceilings(const uint64_t *kept, int kp, int64_t nq, int64_t n_pad, uint64_t *ceil_out, cudaStream_t s) {
    if (n_pad <= 0) return cudaSuccess;
    next_ceilings_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(kept, kp, nq, n_pad, ceil_out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sort_lists_kernel(const uint64_t *__restrict__ lists, int n_lists, int64_t list_stride, int kp, int npad,
                                                         int k_out, int metric, const uint64_t *__restrict__ kept_last, uint32_t *out_idx,
                                                         double *out_score, uint64_t *out_cand, RescoreCheck chk) {
    extern __shared__ __align__(16) unsigned char sl_smem[];
    uint64_t *buf = (uint64_t *)sl_smem;
    const int64_t q = blockIdx.x;
    const int t = threadIdx.x, n = n_lists * kp;
    const bool higher = higher_is_better(metric);
    for (int i = t; i < npad; i += 256) buf[i] = i < n ? lists[(int64_t)(i / kp) * list_stride + q * kp + (i % kp)] : 0ull;
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < npad; i += 256) {
                const int p = i ^ stride;
                if (p > i) {
                    const bool desc = (i & size) == 0;
                    const uint64_t a = buf[i], b = buf[p];
                    if ((a > b) != desc) { buf[i] = b; buf[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (t == 0 && chk.flags) {
        // everything dropped by ALL passes has a filter value <= the last pass's worst kept one (0: its list did not
        // fill, i.e. every candidate that exists was collected)
        const uint64_t last = kept_last[q * kp + kp - 1];
        bool ok = true;
        if (last != 0ull && k_out > 0) {
            const float f_last = key_score(candidate_key(last), true);
            const float tk = key_score(candidate_key(buf[k_out - 1]), higher);
            const float qn = sqrtf(chk.q_sq[q]);
            const float cmax = sqrtf(__uint_as_float(chk.c_max_sq[0])), cmin = sqrtf(__uint_as_float(chk.c_max_sq[1]));
            const float e = filter_error_bound(chk.eps, chk.abs_err, metric, qn, cmax, cmin);
            if (metric == METRIC_DOT) {
                ok = tk > f_last + e;
            } else if (metric == METRIC_COSINE) {
                ok = qn > 1e-6f && tk > (f_last + e) / qn;
            } else {
                const float sq_floor = -f_last - e;
                ok = sq_floor > 0.0f && tk < sqrtf(sq_floor) * (1.0f - 1e-6f);
            }
            if (chk.max_norm > 0.0f && !(qn <= chk.max_norm && cmax <= chk.max_norm)) ok = false;
            if (!(ok)) ok = false;
        }
        if (!ok) {
            chk.flags[q] = 1;
            atomicAdd(chk.flag_count, 1u);
        }
    }
    for (int i = t; i < k_out; i += 256) {
        const uint64_t r = buf[i];
        if (out_idx) out_idx[q * k_out + i] = candidate_index(r);
        if (out_score) out_score[q * k_out + i] = (double)key_score(candidate_key(r), higher);
        if (out_cand) out_cand[q * k_out + i] = r;
    }
}
cudaError_t launch_sort_lists(const uint64_t *lists, int n_lists, int64_t list_stride, int kp, int64_t nq, int k_out, int metric,
                              const uint64_t *kept_last, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                              const RescoreCheck &chk, cudaStream_t s) {
    if (nq <= 0 || k_out <= 0) return cudaSuccess;
    int npad = 64;
    while (npad < n_lists * kp) npad <<= 1;
    if (npad > 4096 || k_out > n_lists * kp) return cudaErrorInvalidValue;
    sort_lists_kernel<<<(unsigned)nq, 256, (size_t)npad * 8, s>>>(lists, n_lists, list_stride, kp, npad, k_out, metric, kept_last, out_idx,
                                                                  out_score, out_cand, chk);
    return cudaGetLastError();
}

// ---- seeds of a re-query level --------------------------------------------------------------------------------
// A flagged query's exact k-th score t (from the candidates the previous level kept) is a LOWER bound of its true
// k-th score, so a candidate that belongs to the top k (or ties into it) has an exact score >= t and therefore a
// filter value >= t - E at the next level (E = that level's error bound).  The next level starts each row's
// threshold just below that: no list warm-up, and usually fewer candidates than the list holds, so nothing at all is
// dropped above the seed.  Margins: 8e-6 relative covers the float rounding of the checks in rescore_kernel.
__global__ void make_seeds_kernel(const int64_t *__restrict__ ids, int64_t n_ids, int64_t n_pad,
                                  const float *__restrict__ kth_units, RescoreCheck next, int metric, float *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    float seed = __uint_as_float(0x7fc00000u);
    if (r < n_ids) {
        const int64_t q = ids[r];
        const float v = kth_units[q];
        const float qn = sqrtf(next.q_sq[q]);
        const float cmax = sqrtf(__uint_as_float(next.c_max_sq[0])), cmin = sqrtf(__uint_as_float(next.c_max_sq[1]));
        const float e = filter_error_bound(next.eps, next.abs_err, metric, qn, cmax, cmin);
        const bool in_range = next.max_norm <= 0.0f || (qn <= next.max_norm && cmax <= next.max_norm);
        const float sd = (v - e) - 8e-6f * fabsf(v) - 1e-30f;
        if (in_range && sd == sd && fabsf(sd) <= 3.0e38f) seed = sd;
    }
    out[r] = seed;
}
cudaError_t launch_make_seeds(const int64_t *ids, int64_t n_ids, int64_t n_pad, const float *kth_units,
                              const RescoreCheck &next, int metric, float *out, cudaStream_t s) {
    if (n_pad <= 0) return cudaSuccess;
    make_seeds_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(ids, n_ids, n_pad, kth_units, next, metric, out);
    return cudaGetLastError();
}

// ---- warm seeds of a first filter level: the r-th best filter value of a sample pre-pass, per query ---------------------
// lists [n_queries][kp]: merged candidate lists of the pre-pass (sorted, best first; 0 = empty slot).  out [n_pad]: the
// filter value of entry r - 1, NaN (no seed) for rows whose sample list is shorter, holds a NaN score, or is padding.
__global__ void seeds_from_lists_kernel(const uint64_t *__restrict__ lists, int kp, int r, int64_t n_queries, int64_t n_pad,
                                        float *__restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pad) return;
    float seed = __uint_as_float(0x7fc00000u);
    if (q < n_queries) {
        const uint64_t c = lists[q * kp + (r - 1)];
        if (c != 0ull && candidate_key(c) != 0u) {
            const float v = key_score(candidate_key(c), true);
            if (v == v && fabsf(v) <= 3.0e38f) seed = v;
        }
    }
    out[q] = seed;
}
cudaError_t launch_seeds_from_lists(const uint64_t *lists, int kp, int r, int64_t n_queries, int64_t n_pad, float *out, cudaStream_t s) {
    if (n_pad <= 0) return cudaSuccess;
    seeds_from_lists_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(lists, kp, r, n_queries, n_pad, out);
    return cudaGetLastError();
}

// ---- raw matmul: outputs that involve a row with inf / NaN elements, recomputed with IEEE arithmetic ---------------
__global__ void __launch_bounds__(256) matmul_nonfinite_fixup_kernel(RawMatrix lm, RawMatrix rm, const unsigned char *__restrict__ nf_left,
                                                                      const unsigned char *__restrict__ nf_right,
                                                                      const unsigned int *__restrict__ nf_count, float *__restrict__ out) {
    if (nf_count[0] == 0 && nf_count[1] == 0) return;   // the common case: nothing to do
    const int64_t Q = lm.n_rows, N = rm.n_rows, D = lm.dim;
    auto elem = [&](const RawMatrix &m, int64_t b, int64_t l, int64_t i) {
        return m.dtype == 0 ? raw_fetch<__half>(m, b, l, i) : raw_fetch<float>(m, b, l, i);
    };
    // marked LEFT rows: the whole output row; marked RIGHT rows: the whole output column
    for (int64_t i = blockIdx.x; i < Q + N; i += gridDim.x) {
        const bool is_left = i < Q;
        const int64_t r = is_left ? i : i - Q;
        if (!(is_left ? nf_left[r] : nf_right[r])) continue;
        int64_t ab, al;
        raw_row(is_left ? lm : rm, r, ab, al);
        const int64_t others = is_left ? N : Q;
        for (int64_t j = threadIdx.x; j < others; j += blockDim.x) {
            int64_t bb, bl;
            raw_row(is_left ? rm : lm, j, bb, bl);
            float acc = 0.0f;
            for (int64_t d = 0; d < D; ++d) {
                const float x = elem(is_left ? lm : rm, ab, al, d), y = elem(is_left ? rm : lm, bb, bl, d);
                acc = is_left ? __fmaf_rn(x, y, acc) : __fmaf_rn(y, x, acc);
            }
            if (is_left) out[r * N + j] = acc;
            else out[j * N + r] = acc;
        }
    }
}
cudaError_t launch_matmul_nonfinite_fixup(const RawMatrix &left, const RawMatrix &right, const unsigned char *nf_left,
                                          const unsigned char *nf_right, const unsigned int *nf_count, float *out, cudaStream_t s) {
    if (left.n_rows <= 0 || right.n_rows <= 0) return cudaSuccess;
    matmul_nonfinite_fixup_kernel<<<296, 256, 0, s>>>(left, right, nf_left, nf_right, nf_count, out);
    return cudaGetLastError();
}

// ---- fallback plumbing: gather flagged query rows into a dense f32 matrix, scatter their results back ----
template <typename OUT>
__global__ void gather_rows_kernel(RawMatrix qm, const int64_t *__restrict__ ids, int64_t n_ids, OUT *__restrict__ out) {
    const int64_t r = blockIdx.x;
    if (r >= n_ids) return;
    int64_t b, l;
    raw_row(qm, ids[r], b, l);
    for (int64_t i = threadIdx.x; i < qm.dim; i += blockDim.x) out[r * qm.dim + i] = (OUT)raw_fetch_f64(qm, b, l, i);  // widening, then exact back
}
// out: [n_ids x dim] dense rows in f32 (out_f64 = 0; the source is f16 or f32 then) or f64.
cudaError_t launch_gather_rows(const RawMatrix &qm, const int64_t *ids, int64_t n_ids, void *out, int out_f64, cudaStream_t s) {
    if (n_ids <= 0) return cudaSuccess;
    if (out_f64) gather_rows_kernel<double><<<(unsigned)n_ids, 128, 0, s>>>(qm, ids, n_ids, (double *)out);
    else gather_rows_kernel<float><<<(unsigned)n_ids, 128, 0, s>>>(qm, ids, n_ids, (float *)out);
    return cudaGetLastError();
}
__global__ void scatter_results_kernel(const int64_t *__restrict__ ids, int64_t n_ids, int k, const uint32_t *si, const double *ss,
                                       const uint64_t *sc, uint32_t *di, double *ds, uint64_t *dc) {
    const int64_t r = blockIdx.x;
    if (r >= n_ids) return;
    const int64_t q = ids[r];
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        if (di) di[q * k + t] = si[r * k + t];
        if (ds) ds[q * k + t] = ss[r * k + t];
        if (dc) dc[q * k + t] = sc[r * k + t];
    }
}
cudaError_t launch_scatter_results(const int64_t *ids, int64_t n_ids, int k, const uint32_t *si, const double *ss,
                                   const uint64_t *sc, uint32_t *di, double *ds, uint64_t *dc, cudaStream_t s) {
    if (n_ids <= 0 || k <= 0) return cudaSuccess;
    scatter_results_kernel<<<(unsigned)n_ids, 128, 0, s>>>(ids, n_ids, k, si, ss, sc, di, ds, dc);
    return cudaGetLastError();
}

}  // namespace pmm
