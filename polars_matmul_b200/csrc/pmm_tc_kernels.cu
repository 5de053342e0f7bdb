// pmm_tc_kernels.cu — the hot path on 5th-gen tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma -> TMEM -> fused epilogue.
//
// One persistent, warp-specialised kernel, two epilogues:
//   * top-k  : replaces matmul_f32 + the cosine/euclidean pass + select_topk_with_scores_f32
//              (src/metrics.rs:204-255, :314-365, src/topk.rs:42-75).  The Q x N score matrix the
//              reference materialises (src/metrics.rs:211, :354) never leaves the SM: each epilogue
//              thread owns one query row of the 128 x 256 accumulator tile in TMEM, applies the metric
//              in f32 exactly as the reference does, packs (score key, ~index) into a u64 and keeps
//              only candidates that beat the row's current k-th best.
//   * matmul : replaces matmul_slice_f32 (src/metrics.rs:160-202): the tile goes registers -> swizzled shared tile -> TMA
//              store (full 128-byte lines), on eight epilogue warps.  f32 operands with 32 < D <= 256 arrive as row-scaled
//              hi/lo f16 planes (pmm_prep.cu: prep_split16_kernel) and are contracted with three kind::f16 MMAs per 16
//              elements, the small terms swept first, the query planes of the item resident in shared memory (SPLIT16
//              below); shorter f32 vectors use the 3xTF32 planes, f16-stored input one exact plane.
//
// Precision: the top-k epilogue is a FILTER - final scores come from the exact re-scoring kernel (pmm_rescore.cu)
// and every query carries a proof that the filter dropped nothing relevant - so the operand format is an internal
// choice per level: f32 (or f64) inputs rounded to one f16 plane + ONE kind::f16 MMA per K-step (default first
// level, 11 significant bits at the full f16 rate), one TF32 plane (TERMS = 1), or the 3xTF32 split
// hi*hi + hi*lo + lo*hi on hi/lo planes from pmm_prep.cu (re-query level).  The raw matmul's result IS the output and
// must stay within 1e-5: it uses the f16 hi/lo split in the order that keeps tcgen05's accumulate truncations small
// (DESIGN.md 4.5).  f16-stored inputs use kind::f16 on exact planes (products of two f16 values are exact in f32, so
// only the summation order differs from the reference's upcast path).  f32 accumulation in TMEM.
//
// Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM lane group = warp % 4); the single-plane top-k kernels and the matmul kernels run a
// second set of four epilogue warps (6..9), each set owning half of the columns of every tile.  Two accumulator buffers of 256 TMEM
// columns let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Running top-k: per row a threshold (the k-th best packed candidate so far).  Scores that beat it are appended to a
// 128-slot per-row staging area in global memory (L2-resident, one per CTA and row); the row's sorted list of
// KP = 32/64/128/256 candidates lives in global memory too (KP/32 registers per lane when it is merged).  Staged
// candidates are merged in batches - at the end of a tile, after the TMEM buffer went back to the MMA warp - by a
// shuffle-based bitonic sort of the batch plus one bitonic merge with the list, which also refreshes the threshold.
// A launch may start from per-row SEED thresholds (TcArgs::seed_thr: "collect everything above this value").
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "pmm_common.cuh"
#include "pmm_kernels.h"
#include "pmm_tc.cuh"

namespace pmm {
using namespace tc;

namespace {

constexpr int BM = TC_TILE_M;
constexpr int BN = TC_TILE_N;
// Epilogue warp sets. The top-k kernels with ONE operand plane (f16, TF32 x1) spend about as long filtering a tile
// as multiplying it, so they run two sets of four epilogue warps: set e owns columns [e*128, e*128+128) of every
// accumulator tile and keeps its own candidate lists (merged with the others by pmm_merge.cu).  Two warps per
// scheduler also hide each other's TMEM-load and dependent-issue latencies.  The 3xTF32 and matmul kernels are
// MMA- resp. store-bound and keep one set (and the deeper operand pipeline).
// The matmul epilogue is a latency chain per warp and 32-column chunk (TMEM load -> registers -> swizzled shared tile ->
// TMA store); four warps could not keep the result stream at HBM rate (4.3 us per 256 x 256 tile against 3.0 at the
// write rate), eight - two per scheduler, half of the columns each, one store tile per warp - can.
__host__ __device__ constexpr int tc_esets(bool f16, int epi, int terms) { return ((epi == 0 && f16) || epi == 1) ? 2 : 1; }
__host__ __device__ constexpr int tc_threads(int esets) { return 64 + 128 * esets; }
// Candidates that beat a row's threshold are APPENDED to a per-row staging area in global memory (L2-resident,
// one per CTA and row, reused by every item of the CTA) and merged into the row's sorted list in batches: the
// cost of a merge (sort the batch, one bitonic merge with the list) is the same for 10 or 100 staged
// candidates, so batches should be large.  Thresholds only move at a merge; a stale threshold admits more
// candidates, never fewer.
constexpr int LOOK_PITCH = 36;             // floats per lane in the hit-lookup area (16-byte aligned, 4-way bank spread)
constexpr int SC = 128;                    // staging capacity per row
constexpr int HARD_AT = SC - 32;           // a 32-column chunk adds at most 32: merge inside the tile above this
constexpr int URGENT_AT = 80;              // rows above this are always merged at the end of a tile
// end-of-tile merge above this (rate limited after the first tiles), AFTER the TMEM buffer was released.  Smaller
// batches for short lists were measured and lose: a merge has a high fixed cost (f16 C5, k=10: 192 -> 197 ms).
constexpr int SOFT_AT = 48;

// ROWB = bytes of K per shared-memory row (= the swizzle span): 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B).
// CG   = tcgen05 cta_group: 1 = one CTA per 128 x 256 tile; 2 = a CTA pair (cluster of 2) computes a
//        256 x 256 tile with UMMA M=256: each CTA holds 128 query rows and HALF of the corpus tile
//        (128 rows), so corpus bytes per SM halve and a third pipeline stage fits.
// TERMS = tcgen05 MMAs per K-step for f32 data: 3 = the 3xTF32 split (hi/lo planes), 1 = hi*hi only (TF32
//         precision, |error| <= 2^-11 |q||c|; used as the first-level filter, see pmm_api.cu).
//         With F16: 2 = the hi/lo f16 split of the raw f32 matmul (row-scaled planes from prep_split16_kernel, see
//         the MMA issuer), any other value = one exact / rounded f16 plane.
// CLM  = CTA pairs per cluster (cta_group::2 only): 2 = a cluster of 4 CTAs works on two query tiles against the
//        SAME corpus tile; each CTA fetches a quarter of the corpus tile and TMA-multicasts it to the CTA of the
//        other pair that needs the same half, so corpus bytes L2 -> shared memory halve again.
template <bool F16, int ROWB, int CG, int TERMS, int CLM = 1, int ESETS = 1, int EPI = 0>
struct TcCfg {
    static constexpr int PLANES = (F16 || TERMS == 1) ? 1 : 2;   // operand planes per matrix IN A STAGE (the f16 hi/lo split
                                                                 // stages one plane of each matrix at a time)
    static constexpr int BK = ROWB / (F16 ? 2 : 4);  // elements of K per stage
    static constexpr int KSTEPS = ROWB / 32;         // 32 bytes of K per tcgen05.mma
    static constexpr int B_ROWS = BN / CG;           // corpus rows this CTA stages per tile
    static constexpr int A_BYTES = BM * ROWB;
    static constexpr int B_BYTES = B_ROWS * ROWB;
    static constexpr bool SPLIT16 = F16 && TERMS == 2;   // raw f32 matmul on hi/lo f16 planes: the query planes of the item stay
                                                         // resident in shared memory, a stage holds one K-block of one corpus plane
    static constexpr int STAGE_BYTES = SPLIT16 ? B_BYTES : PLANES * (A_BYTES + B_BYTES);
    static constexpr int STAGING_BYTES = ESETS * 4 * 32 * LOOK_PITCH * 4;  // per epilogue warp: one chunk of filter values (hit lookup)
    static constexpr int AUX_BYTES = 4 * BN * 4;     // per epilogue warp: the corpus aux values of its columns of the tile
    static constexpr int STORE_BYTES = 4 * 2 * 4096; // matmul epilogue: per warp two 32x32 f32 TMA-store tiles
    static constexpr int EPI_BYTES = EPI == 1 ? STORE_BYTES : (STAGING_BYTES + AUX_BYTES);   // matmul: the store tiles only
    static constexpr int COLF_BYTES = (F16 && TERMS == 2) ? BN * 4 : 0;   // f16 split: the tile's column scale factors (one copy per CTA)
    // f16 split: [epilogue][barriers, column factors: 2 KB][resident query planes: 2 x num_kb x A_BYTES][stage ring]; the
    // ring takes what is left of the 227 KB (12 - 2 num_kb stages, at most STAGES = 8: the barrier arrays' size)
    static constexpr int SPLIT_FIXED = EPI_BYTES + 2048;
    static constexpr int STAGES = SPLIT16 ? 8 : (232448 - EPI_BYTES - 256 - COLF_BYTES - 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = SPLIT16 ? 232448 : STAGES * STAGE_BYTES + EPI_BYTES + 256 + COLF_BYTES + 1024;
};

struct TcKParams {
    TcSchedule sched;
    int num_kb;
    const float *q_aux, *c_aux;
    int64_t nq, n;
    int64_t index_base;
    int metric;
    int k;
    uint64_t *partial;
    uint64_t *staged;  // top-k epilogue: grid x 128 rows x SC staged candidates
    float *out;
    int out_tma;   // 1: matmul epilogue stores through TMA (row pitch is a multiple of 16 bytes)
    unsigned int *round_sync;  // zeroed counters: producers of all CTAs meet every `sync_tiles` corpus tiles (NULL: off)
    int sync_tiles;
    int debug_skip;  // measurement only (TcArgs::debug_skip)
    int sync_slack;
    int soft_at;     // end-of-tile merge threshold (SOFT_AT unless overridden for experiments)
    int resume;      // lists already hold the candidates of earlier launches over other corpus rows
    int max_flush;   // row merges per warp at the end of a tile (rate limit; rows above URGENT_AT always go)
    float norm_guard;       // cosine: norms at or below this count as zero (1e-6 f32; just under 1e-10 for f64 sources)
    const float *seed_thr;  // per query row (padded like q_aux): initial threshold in filter units, NaN = none; or NULL
    const uint64_t *ceil;   // per query row (padded): only candidates strictly BELOW this packed value are admitted; or NULL
    int tile_stride;        // corpus tile nt of the schedule is tile nt * tile_stride of the planes (1; > 1: strided sample)
};

enum { EPI_TOPK = 0, EPI_MATMUL = 1 };

// Timing experiments that return WRONG results (debug_skip 1..3: epilogue without filter / without merges) exist only
// in -DPMM_DIAG builds; debug_skip == 8 (correct results + wait-cycle counters) is always available.
#ifdef PMM_DIAG
#define PMM_DSKIP(p, x) ((p).debug_skip == (x))
#else
#define PMM_DSKIP(p, x) false
#endif

// Diagnostics (option tc_debug_skip = 8): cycles the MMA warps spent waiting for [0] a free accumulator buffer,
// [1] a filled operand stage, [2] in total; [3] cycles epilogue warp 2 of the leader CTAs spent in list flushes.
__device__ unsigned long long g_tc_wait[4 + 48];  // [20 + b] / [36 + b]: epilogue warp 2 filter / flush cycles  // [4 + b]: accumulator-buffer waits at tile 2^b..2^(b+1)-1 of an item

// Work of CTA `cta` in round `it`. Returns false when the CTA idles in that round.
__device__ __forceinline__ bool tc_round_item(const TcSchedule &s, int cta, int it, int &m_tile, int &n_start,
                                              int &n_step, int64_t &slot, int &n_end) {
    n_end = s.n_tiles;
    if (s.flat == 2) {   // raw matmul, hybrid: main sweeps + helpers on the tails
        n_step = 1;
        slot = 0;
        if (cta < s.m_tiles) {
            m_tile = cta;
            n_start = 0;
            n_end = s.n_main;
            return it == 0;
        }
        const int H = s.num_ctas - s.m_tiles, h = cta - s.m_tiles, tail = s.n_tiles - s.n_main;
        const int64_t total = (int64_t)s.m_tiles * tail;
        const int64_t t0 = total * h / H, t1 = total * (h + 1) / H;
        if (cta >= s.num_ctas || t0 >= t1) return false;
        const int m0 = (int)(t0 / tail);
        m_tile = m0 + it;
        if ((int64_t)m_tile * tail >= t1) return false;
        n_start = s.n_main + (it == 0 ? (int)(t0 - (int64_t)m0 * tail) : 0);
        const int64_t left = t1 - (int64_t)m_tile * tail;
        n_end = s.n_main + (int)(left < tail ? left : tail);
        return true;
    }
    if (s.flat) {   // raw matmul: this unit's share of the row-major tile list, one query tile per "round"
        const int64_t total = (int64_t)s.m_tiles * s.n_tiles;
        const int64_t t0 = total * cta / s.num_ctas, t1 = total * (cta + 1) / s.num_ctas;
        if (cta >= s.num_ctas || t0 >= t1) return false;
        const int m0 = (int)(t0 / s.n_tiles);
        m_tile = m0 + it;
        if ((int64_t)m_tile * s.n_tiles >= t1) return false;
        n_start = it == 0 ? (int)(t0 - (int64_t)m0 * s.n_tiles) : 0;
        const int64_t left = t1 - (int64_t)m_tile * s.n_tiles;
        if (left < n_end) n_end = (int)left;
        n_step = 1;
        slot = 0;
        return true;
    }
    if (it < s.rounds) {
        if (cta >= s.mc * s.g) return false;
        int grp = cta / s.g, rank = cta - grp * s.g;
        m_tile = it * s.mc + grp;
        n_start = rank;
        n_step = s.g;
        slot = (int64_t)m_tile * s.g + rank;
        return true;
    }
    if (cta >= s.m_rem * s.g_rem) return false;
    int grp = cta / s.g_rem, rank = cta - grp * s.g_rem;
    m_tile = s.m_full + grp;
    n_start = rank;
    n_step = s.g_rem;
    slot = (int64_t)s.m_full * s.g + (int64_t)grp * s.g_rem + rank;
    return true;
}

// Merge the staged candidates of the rows in `rows` (bit i = lane i's row) into their lists.
// thr = packed k-th best of the row's list (0 while the list is not full); thr_f = its filter value
// as a float, NaN while the list is not full (so that `!(f <= thr_f)` admits everything).
template <int R>
__device__ __forceinline__ void flush_rows(unsigned rows, const uint64_t *stg /* the warp's 32 staging rows */,
                                           uint64_t *list_base /* lane group's 32 lists */, int lane, int k, uint64_t &thr,
                                           float &thr_f, int &cnt, uint64_t seed_c /* this lane's seed threshold, 0 = none */) {
    constexpr int KP = 32 * R;
    __syncwarp();
    if (k < 0) {  // measurement only (debug_skip == 3): drop the staged candidates
        cnt = 0;
        return;
    }
    while (rows) {
        const int src = __ffs(rows) - 1;
        rows &= rows - 1;
        const int c = __shfl_sync(0xffffffffu, cnt, src);
        uint64_t *list = list_base + (int64_t)src * KP;
        const uint64_t *sr = stg + (int64_t)src * SC;
        uint64_t L[R], M[R], S[SC / 32];
#pragma unroll
        for (int r = 0; r < R; ++r) L[r] = list[32 * r + lane];  // issued first: the L2 latency overlaps the sort below
#pragma unroll
        for (int r = 0; r < SC / 32; ++r) S[r] = (32 * r + lane < c) ? sr[32 * r + lane] : 0ull;
        // M = the staged candidates, sorted descending, cut or zero-padded to the list length and REVERSED:
        // M[r] at lane l holds element 32R-1-(32r+l), i.e. register R-1-r, lane 31-l of the sorted staging area.
        if (c <= 32) {
            S[0] = warp_sort_desc(S[0], lane);
#pragma unroll
            for (int r = 0; r < R - 1; ++r) M[r] = 0ull;
            M[R - 1] = __shfl_sync(0xffffffffu, S[0], 31 - lane);
        } else {
            if (c <= 64) {  // half-size network; registers 2, 3 are empty (zeros sort last)
                uint64_t S2[2] = {S[0], S[1]};
                warp_sort_regs_desc<2>(S2, lane);
                S[0] = S2[0];
                S[1] = S2[1];
            } else {
                warp_sort_regs_desc<SC / 32>(S, lane);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) M[r] = (R - 1 - r) < SC / 32 ? __shfl_sync(0xffffffffu, S[(R - 1 - r) < SC / 32 ? (R - 1 - r) : 0], 31 - lane) : 0ull;
        }
        warp_merge_topk_desc<R>(L, M, lane);
#pragma unroll
        for (int r = 0; r < R; ++r) list[32 * r + lane] = L[r];
        uint64_t kreg = L[R - 1];  // register holding list element k-1 (k <= 32*R)
#pragma unroll
        for (int r = 0; r < R - 1; ++r)
            if (r == ((k - 1) >> 5)) kreg = L[r];
        const uint64_t kth = __shfl_sync(0xffffffffu, kreg, (k - 1) & 31);
        if (lane == src) {
            thr = kth > seed_c ? kth : seed_c;   // a seeded row never opens up again below its seed
            thr_f = thr == 0ull ? __uint_as_float(0x7fc00000u) : key_score(candidate_key(thr), true);
            cnt = 0;
        }
    }
    __syncwarp();
}

// Filter value of one accumulator: a float that is LARGER for a BETTER candidate of this query row and
// monotone in the reference's score (the exact score is recomputed later by pmm_rescore.cu):
//   dot       f = acc
//   cosine    f = acc * inv_cn[col] * rowmul          (row constant 1/qn dropped; zero norms -> 0)
//   euclidean f = -max((csq[col] - 2 acc) + qsq, 0)   (sqrt dropped)
template <int METRIC>
__device__ __forceinline__ float filter_value(float acc, float aux, float rowc) {
    if (METRIC == METRIC_COSINE) return (acc * aux) * rowc;
    if (METRIC == METRIC_EUCLIDEAN) return -fmaxf(fmaf(-2.0f, acc, aux) + rowc, 0.0f);
    return acc;
}

// (x, y) <- ((x r) c0, (y r) c1) with two packed f32x2 multiplications (FMUL2): the matmul epilogue of the f16 split.
__device__ __forceinline__ void scale2(uint32_t &x, uint32_t &y, float r, float c0, float c1) {
    asm("{ .reg .b64 va, vr, vc; mov.b64 va, {%0, %1}; mov.b64 vr, {%2, %2}; mov.b64 vc, {%3, %4};\n\t"
        "mul.rn.f32x2 va, va, vr; mul.rn.f32x2 va, va, vc; mov.b64 {%0, %1}, va; }"
        : "+r"(x), "+r"(y) : "f"(r), "f"(c0), "f"(c1));
}

// One 32-column chunk of the accumulator tile for this thread's row.
// Common case: no score of the chunk beats any row's threshold. It costs the filter values, a max tree
// (one FMNMX per score; fmaxf drops NaN, and NaN can only matter while a list is not full, when thr_f is NaN
// and every test below is true) and ONE warp vote. Otherwise every lane collects the bit mask of its own hits and
// the lanes append their hits side by side (see below).
template <int METRIC, int R>
__device__ __forceinline__ void filter_chunk(const uint32_t (&v)[32], uint32_t aux_s /* shared address: 32 floats */,
                                             uint32_t look_s /* shared address: the warp's 32 x LOOK_PITCH floats */, float rowc,
                                             int64_t col0, int64_t n, int64_t index_base, uint64_t *stg /* warp's staging rows */,
                                             uint64_t *list_base, int lane, int k, uint64_t &thr, float &thr_f, int &cnt,
                                             uint64_t seed_c, uint64_t ceil_c) {
    float f[32];
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        if (METRIC != METRIC_DOT)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3) : "r"(aux_s + 4u * j4));
        f[j4 + 0] = filter_value<METRIC>(__uint_as_float(v[j4 + 0]), a0, rowc);
        f[j4 + 1] = filter_value<METRIC>(__uint_as_float(v[j4 + 1]), a1, rowc);
        f[j4 + 2] = filter_value<METRIC>(__uint_as_float(v[j4 + 2]), a2, rowc);
        f[j4 + 3] = filter_value<METRIC>(__uint_as_float(v[j4 + 3]), a3, rowc);
    }
    float gmax[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float m01 = fmaxf(f[8 * g + 0], f[8 * g + 1]), m23 = fmaxf(f[8 * g + 2], f[8 * g + 3]);
        float m45 = fmaxf(f[8 * g + 4], f[8 * g + 5]), m67 = fmaxf(f[8 * g + 6], f[8 * g + 7]);
        gmax[g] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
    }
    const float cmax = fmaxf(fmaxf(gmax[0], gmax[1]), fmaxf(gmax[2], gmax[3]));
    if (!__any_sync(0xffffffffu, !(cmax <= thr_f))) return;
    // Some row of the warp has a hit.  Each lane builds the bit mask of ITS hits without branches; lanes with
    // hits park their 32 filter values in shared memory (so that a hit can be fetched by its run-time position)
    // and walk their masks side by side: the loop runs max-hits-per-lane times, usually once, instead of once
    // per column position.  A hit: better than the k-th best, or the list is not full (thr_f NaN), or NaN.
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) m |= !(f[j] <= thr_f) ? (1u << j) : 0u;
    if (m) {
        const uint32_t mine = look_s + (uint32_t)lane * (LOOK_PITCH * 4);
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mine + 4u * j4), "f"(f[j4]), "f"(f[j4 + 1]), "f"(f[j4 + 2]),
                         "f"(f[j4 + 3])
                         : "memory");
        do {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            float fv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fv) : "r"(mine + 4u * j) : "memory");
            const int64_t col = col0 + j;
            if (col < n) {
                const uint64_t cand = pack_candidate(score_key(fv, true), (uint32_t)(index_base + col));
                if (cand > thr && cand < ceil_c) {   // (ceil_c: multi-pass top-k, "the next 256 below what is already kept")
                    stg[lane * SC + cnt] = cand;
                    ++cnt;
                }
            }
        } while (m);
    }
    const unsigned over = __ballot_sync(0xffffffffu, cnt > HARD_AT);  // room for the next chunk's (at most) 32
    if (over) flush_rows<R>(over, stg, list_base, lane, k, thr, thr_f, cnt, seed_c);
}

template <bool F16, int EPI, int R, int ROWB, int CG, int TERMS, int CLM>
__global__ void __launch_bounds__(tc_threads(tc_esets(F16, EPI, TERMS)), 1)
tc_kernel(const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo,
          const __grid_constant__ CUtensorMap tm_chi, const __grid_constant__ CUtensorMap tm_clo,
          const __grid_constant__ CUtensorMap tm_out, const TcKParams p) {
    constexpr int ESETS = tc_esets(F16, EPI, TERMS);
    constexpr int CPS = BN / 32 / ESETS;      // 32-column chunks of a tile per epilogue warp
    typedef TcCfg<F16, ROWB, CG, TERMS, CLM, ESETS, EPI> Cfg;
    constexpr bool ONE = Cfg::PLANES == 1;    // one operand plane per matrix, one MMA per K-step
    // Raw f32 matmul on f16 hi/lo planes: x = hi + lo to 22 bits, products hi*hi + hi*lo + lo*hi at the f16 rate.
    // tcgen05 truncates the f32 accumulator once per MMA, and a truncation costs up to one ulp of the RUNNING SUM, so
    // the order matters: the K range is swept TWICE - first all lo*hi and hi*lo terms (the sum is ~2^-11 of the result
    // while they are added, their truncations vanish), then the hi*hi terms.  Only D/16 truncations hit the full-size
    // sum (the 3xTF32 order: 3 D/8).  The unit of work sweeps all corpus tiles for ONE query tile, so both query planes
    // of the item (2 x num_kb x 16 KB, D <= 256) stay resident in shared memory and only corpus planes stream through
    // the stage ring, one K-block of one plane (16 KB) per stage: per K-block (c hi), (c lo) in the first sweep -
    // multiplied with q lo and q hi - and (c hi) again in the second.  Operand bytes L2 -> SM per tile: 12 x 16 KB per
    // CTA at D = 256 instead of 24 - the kernel had been bound by the SM <-> L2 fabric (operands in, results out).
    constexpr bool SPLIT16 = Cfg::SPLIT16;
    constexpr int GS = CG * CLM;              // CTAs per scheduling unit (= cluster size)
    static_assert(CLM == 1 || CG == 2, "corpus multicast is built on CTA pairs");
    static_assert(!SPLIT16 || (CG == 2 && CLM == 1 && EPI == EPI_MATMUL), "the f16 split is the CTA-pair matmul kernel");
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    // offsets of the epilogue area and of the stage ring (f16 split: see TcCfg)
    const int a_bytes = SPLIT16 ? 2 * p.num_kb * Cfg::A_BYTES : 0;
    const uint32_t epi_off = SPLIT16 ? 0u : (uint32_t)(Cfg::STAGES * Cfg::STAGE_BYTES);
    const uint32_t a_off = (uint32_t)Cfg::SPLIT_FIXED;
    const uint32_t ring_off = SPLIT16 ? a_off + (uint32_t)a_bytes : 0u;
    const int nst = SPLIT16 ? (12 - 2 * p.num_kb < Cfg::STAGES ? 12 - 2 * p.num_kb : Cfg::STAGES) : Cfg::STAGES;   // stages in the ring
    float *aux_tiles = (float *)(smem + epi_off + Cfg::STAGING_BYTES);
    uint64_t *bars = (uint64_t *)(smem + epi_off + Cfg::EPI_BYTES);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + b); };
    auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + s); };  // CLM == 2: "peer CTA's stage landed"
    volatile uint32_t *tmem_ptr_smem = (volatile uint32_t *)(bars + 3 * Cfg::STAGES + 4);
    const uint32_t colf_sa = bar_base + 256u;   // f16 split: BN floats, 256-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank4 = GS > 1 ? cluster_ctarank() : 0u;    // rank inside the cluster
    const uint32_t crank = crank4 & (uint32_t)(CG - 1);          // rank inside the CTA pair; 0 = leader (issues the MMAs)
    const uint32_t pairc = CLM == 2 ? (crank4 >> 1) : 0u;        // which pair of the cluster
    const uint32_t leader_rank = crank4 & ~1u;                   // cluster rank of this pair's leader
    const int cta = (int)(blockIdx.x / GS);                      // scheduling unit: CTA, CTA pair or cluster of two pairs
    const TcSchedule &S = p.sched;
    const int total_rounds = S.rounds + (S.m_rem > 0 ? 1 : 0);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_qhi);
        prefetch_tensormap(&tm_chi);
        if (!ONE || SPLIT16) {
            prefetch_tensormap(&tm_qlo);
            prefetch_tensormap(&tm_clo);
        }
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CLM);   // every pair that reads (or multicasts into) the slot releases it
            mbar_init(pfull_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4 * ESETS * CG);  // one arrival per epilogue warp of every CTA of the group
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc<CG>(smem_u32((const void *)tmem_ptr_smem), 512);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if (GS > 1) cluster_sync(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        // The whole warp walks the schedule (warp-uniform control flow and addresses); one elected lane issues.
        const bool issuer = elect_one();
        int stage = 0;
        uint32_t phase = 0, aphase = 0;
        const int n_sync_full = p.round_sync ? ((S.n_tiles + S.g - 1) / S.g + p.sync_tiles - 1) / p.sync_tiles : 0;
        bool pace = true;  // cleared by the first timeout: this CTA keeps arriving but no longer waits (the grid is
                           // evidently not co-resident, e.g. another kernel holds SMs; waiting again would cost 4 ms a time)
        for (int it = 0; it < total_rounds; ++it) {
            // Pacing. All CTAs are co-resident (persistent grid <= #SMs), so their producers can meet: every
            // `sync_tiles` corpus tiles of a round all producers wait for each other, which keeps the groups
            // sweeping the corpus in lockstep so that a corpus tile is read from HBM once per round instead of
            // once per group.  Sync points of a round are numbered identically in every CTA; a CTA that has
            // fewer tiles (or none) in the round still arrives at all of them.
            const int step_r = it < S.rounds ? S.g : S.g_rem;
            const int n_sync = p.round_sync ? ((S.n_tiles + step_r - 1) / step_r + p.sync_tiles - 1) / p.sync_tiles : 0;
            unsigned int *sync_base = p.round_sync + (it < S.rounds ? it * n_sync_full : S.rounds * n_sync_full);
            int m_tile, n_start, n_step, n_end;
            int64_t slot;
            if (!tc_round_item(S, cta, it, m_tile, n_start, n_step, slot, n_end)) {
                if (issuer)
                    for (int sp = 0; sp < n_sync; ++sp) atomicAdd(sync_base + sp, 1u);
                __syncwarp();
                continue;
            }
            const int32_t arow = ((m_tile * CLM + (int)pairc) * CG + (int)crank) * BM;   // this CTA's 128 query rows
            if (SPLIT16) {
                // the item's query planes, resident for all of its tiles: hi K-blocks, then lo K-blocks.  The region is
                // free once the MMAs of the previous item have retired (aempty: committed by the MMA warp).
                mbar_wait(pfull_bar(1), aphase ^ 1u);
                if (issuer) {
                    const uint32_t afb = mapa_u32(pfull_bar(0), leader_rank);
                    if (crank == 0) mbar_arrive_expect_tx(pfull_bar(0), 2u * (uint32_t)a_bytes);
                    for (int kb = 0; kb < p.num_kb; ++kb) {
                        tma_load_2d_pair(smem_base + a_off + (uint32_t)(kb * Cfg::A_BYTES), &tm_qhi, afb, kb * Cfg::BK, arow);
                        tma_load_2d_pair(smem_base + a_off + (uint32_t)((p.num_kb + kb) * Cfg::A_BYTES), &tm_qlo, afb, kb * Cfg::BK, arow);
                    }
                }
                __syncwarp();
                aphase ^= 1u;
            }
            int j = 0;  // tile counter of this CTA inside the round
            for (int nt = n_start; nt < n_end; nt += n_step, ++j) {
                const int32_t brow = nt * p.tile_stride * BN + (int)crank * Cfg::B_ROWS;   // this CTA's part of the corpus tile
                if (n_sync && j % p.sync_tiles == 0) {
                    if (issuer) {
                        unsigned int *ctr = sync_base + j / p.sync_tiles;
                        atomicAdd(ctr, 1u);
                        // wait until every producer has reached the sync point `sync_slack` points back (0: this one):
                        // CTAs may drift apart by sync_slack + 1 segments, which absorbs epilogue jitter
                        unsigned int *wctr = ctr - p.sync_slack;
                        if (pace && wctr >= p.round_sync) {
                            const long long t0 = clock64();
                            // best-effort, never a correctness dependency: give up after ~4 ms (e.g. when another
                            // kernel holds some SMs and part of this grid is not resident yet)
                            while (ld_acquire_u32(wctr) < gridDim.x) {
                                __nanosleep(100);
                                if (clock64() - t0 > 8000000ll) {
                                    pace = false;
                                    break;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                const int kb_total = SPLIT16 ? 3 * p.num_kb : p.num_kb;   // f16 split: stage uses, see above
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + ring_off + stage * Cfg::STAGE_BYTES;
                    if (issuer) {
                        if (SPLIT16) {   // one K-block of one corpus plane
                            const bool first = kb < 2 * p.num_kb;
                            const int k0 = (first ? (kb >> 1) : kb - 2 * p.num_kb) * Cfg::BK;
                            const uint32_t fb = mapa_u32(full_bar(stage), leader_rank);
                            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                            tma_load_2d_pair(sa, (first && (kb & 1)) ? &tm_clo : &tm_chi, fb, k0, brow);
                        } else if (CG == 1) {
                            const uint32_t fb = full_bar(stage);
                            mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
                            tma_load_2d(sa, &tm_qhi, fb, kb * Cfg::BK, arow);
                            if (!ONE) tma_load_2d(sa + Cfg::A_BYTES, &tm_qlo, fb, kb * Cfg::BK, arow);
                            tma_load_2d(sa + Cfg::PLANES * Cfg::A_BYTES, &tm_chi, fb, kb * Cfg::BK, brow);
                            if (!ONE) tma_load_2d(sa + Cfg::PLANES * Cfg::A_BYTES + Cfg::B_BYTES, &tm_clo, fb, kb * Cfg::BK, brow);
                        } else if (CLM == 1) {
                            // both CTAs of the pair load into their own shared memory; all bytes are counted on the
                            // LEADER's full barrier, which its MMA warp waits on
                            const uint32_t fb = mapa_u32(full_bar(stage), leader_rank);
                            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                            tma_load_2d_pair(sa, &tm_qhi, fb, kb * Cfg::BK, arow);
                            if (!ONE) tma_load_2d_pair(sa + Cfg::A_BYTES, &tm_qlo, fb, kb * Cfg::BK, arow);
                            tma_load_2d_pair(sa + Cfg::PLANES * Cfg::A_BYTES, &tm_chi, fb, kb * Cfg::BK, brow);
                            if (!ONE)
                                tma_load_2d_pair(sa + Cfg::PLANES * Cfg::A_BYTES + Cfg::B_BYTES, &tm_clo, fb, kb * Cfg::BK, brow);
                        } else {
                            // cluster of two pairs: every CTA arms ITS OWN full barrier with the bytes that land in its
                            // shared memory (own query tile + the whole corpus half, a quarter of which arrives from the
                            // other pair's CTA by multicast) and fetches its quarter of the corpus tile for both.
                            const uint32_t fb = full_bar(stage);
                            mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
                            tma_load_2d(sa, &tm_qhi, fb, kb * Cfg::BK, arow);
                            if (!ONE) tma_load_2d(sa + Cfg::A_BYTES, &tm_qlo, fb, kb * Cfg::BK, arow);
                            const uint16_t mask = (uint16_t)((1u << crank) | (1u << (2 + crank)));   // same half, both pairs
                            const uint32_t qoff = pairc * (Cfg::B_ROWS / 2) * ROWB;                  // my quarter inside the half
                            const int32_t qrow_b = brow + (int)pairc * (Cfg::B_ROWS / 2);
                            tma_load_2d_mc(sa + Cfg::PLANES * Cfg::A_BYTES + qoff, &tm_chi, fb, kb * Cfg::BK, qrow_b, mask);
                            if (!ONE)
                                tma_load_2d_mc(sa + Cfg::PLANES * Cfg::A_BYTES + Cfg::B_BYTES + qoff, &tm_clo, fb, kb * Cfg::BK,
                                               qrow_b, mask);
                        }
                    }
                    __syncwarp();
                    if (++stage == nst) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (n_sync) {  // sync points of this round that lie beyond this CTA's last tile
                if (issuer)
                    for (int sp = (j + p.sync_tiles - 1) / p.sync_tiles; sp < n_sync; ++sp) atomicAdd(sync_base + sp, 1u);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // Warp-uniform loop (all lanes wait on the barriers and compute descriptors in uniform registers);
        // one elected lane issues tcgen05.mma / tcgen05.commit.  Only the leader CTA of a pair issues.
        if (crank == 0) {
            constexpr uint32_t idesc = umma_instr_desc(F16 ? 0 : 2, BM * CG, BN);
            const bool issuer = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            int abuf = 0;
            uint32_t aphase = 0, qphase = 0;
            const bool dbg = p.debug_skip == 8;
            long long w_tempty = 0, w_full = 0;
            const long long w_start = dbg ? clock64() : 0ll;
            for (int it = 0; it < total_rounds; ++it) {
                int m_tile, n_start, n_step, n_end;
                int64_t slot;
                if (!tc_round_item(S, cta, it, m_tile, n_start, n_step, slot, n_end)) continue;
                if (SPLIT16) {   // the item's query planes have landed (both CTAs' bytes count on the leader's barrier)
                    mbar_wait(pfull_bar(0), qphase);
                    qphase ^= 1u;
                    tc_fence_after();
                }
                int jt = 0;
                for (int nt = n_start; nt < n_end; nt += n_step, ++jt) {
                    const long long w0 = dbg ? clock64() : 0ll;
                    mbar_wait(tempty_bar(abuf), aphase ^ 1u);
                    if (dbg) {
                        const long long w = clock64() - w0;
                        w_tempty += w;
                        if (issuer) atomicAdd(&g_tc_wait[4 + (31 - __clz(jt + 1))], (unsigned long long)w);
                    }
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(abuf * BN);
                    const int kb_total = SPLIT16 ? 3 * p.num_kb : p.num_kb;
                    for (int kb = 0; kb < kb_total; ++kb) {
                        const long long w1 = dbg ? clock64() : 0ll;
                        mbar_wait(full_bar(stage), phase);
                        if (dbg) w_full += clock64() - w1;
                        if (CLM == 2) mbar_wait(pfull_bar(stage), phase);   // the peer CTA's half of the stage landed too
                        tc_fence_after();
                        const uint32_t sa = smem_base + ring_off + stage * Cfg::STAGE_BYTES;
                        // descriptors of the stage's four tiles; a K-step of 32 bytes adds 2 to the address field
                        const uint64_t d_ah = umma_smem_desc<ROWB>(sa);
                        const uint64_t d_al = d_ah + (uint64_t)(Cfg::A_BYTES >> 4);
                        const uint64_t d_bh = d_ah + (uint64_t)((Cfg::PLANES * Cfg::A_BYTES) >> 4);
                        const uint64_t d_bl = d_bh + (uint64_t)(Cfg::B_BYTES >> 4);
                        if (issuer && SPLIT16) {
                            // the stage holds a corpus plane's K-block; its partner is the resident query plane:
                            // first sweep (c hi) x q lo, (c lo) x q hi; second sweep (c hi) x q hi
                            const bool first = kb < 2 * p.num_kb;
                            const int kq = first ? (kb >> 1) : kb - 2 * p.num_kb;
                            const bool q_lo = first && !(kb & 1);
                            const uint64_t d_q = umma_smem_desc<ROWB>(smem_base + a_off + (uint32_t)(((q_lo ? p.num_kb : 0) + kq) * Cfg::A_BYTES));
#pragma unroll
                            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) umma<CG, true>(tmem_d, d_q + 2 * ks, d_ah + 2 * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                            umma_commit_pair(empty_bar(stage), (uint16_t)(3u << leader_rank));
                        } else if (issuer) {
#pragma unroll
                            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                                const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                                if (F16) {
                                    umma<CG, true>(tmem_d, d_ah + 2 * ks, d_bh + 2 * ks, idesc, acc);
                                } else if (TERMS == 1) {
                                    umma<CG, false>(tmem_d, d_ah + 2 * ks, d_bh + 2 * ks, idesc, acc);
                                } else {
                                    umma<CG, false>(tmem_d, d_al + 2 * ks, d_bh + 2 * ks, idesc, acc);  // small terms first
                                    umma<CG, false>(tmem_d, d_ah + 2 * ks, d_bl + 2 * ks, idesc, 1u);
                                    umma<CG, false>(tmem_d, d_ah + 2 * ks, d_bh + 2 * ks, idesc, 1u);
                                }
                            }
                            // frees the smem slot when these MMAs retire: in both CTAs of the pair, and with multicast
                            // in all four CTAs of the cluster (the other pair's CTAs write into this pair's slots)
                            if (CG == 1) umma_commit(empty_bar(stage));
                            else umma_commit_pair(empty_bar(stage), CLM == 2 ? (uint16_t)0xF : (uint16_t)(3u << leader_rank));
                        }
                        __syncwarp();
                        if (++stage == nst) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (issuer) {  // accumulator tile complete (each CTA of a pair holds its 128 rows of it)
                        if (CG == 1) umma_commit(tfull_bar(abuf)); else umma_commit_pair(tfull_bar(abuf), (uint16_t)(3u << leader_rank));
                    }
                    __syncwarp();
                    abuf ^= 1;
                    if (abuf == 0) aphase ^= 1u;
                }
                if (SPLIT16) {   // every MMA that reads the resident query planes has been issued: free them on retirement
                    if (issuer) umma_commit_pair(pfull_bar(1), (uint16_t)(3u << leader_rank));
                    __syncwarp();
                }
            }
            if (dbg && issuer) {
                atomicAdd(&g_tc_wait[0], (unsigned long long)w_tempty);
                atomicAdd(&g_tc_wait[1], (unsigned long long)w_full);
                atomicAdd(&g_tc_wait[2], (unsigned long long)(clock64() - w_start));
            }
        } else if (CLM == 2) {
            // Non-leader CTA of a pair, multicast mode: its stage bytes are counted on ITS OWN full barrier; relay
            // "stage landed" to the pair leader, which issues the MMAs that read both CTAs' shared memory.
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total_rounds; ++it) {
                int m_tile, n_start, n_step, n_end;
                int64_t slot;
                if (!tc_round_item(S, cta, it, m_tile, n_start, n_step, slot, n_end)) continue;
                for (int nt = n_start; nt < n_end; nt += n_step) {
                    for (int kb = 0; kb < p.num_kb; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        if (lane == 0) mbar_arrive_cluster(pfull_bar(stage), leader_rank);
                        __syncwarp();
                        if (++stage == nst) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else {
        // ============================== epilogue (warps 2..5) ==============================
        const int lg = warp & 3;          // TMEM lane group this warp may read
        const int row0 = lg * 32;         // first tile row of the warp
        const int row = row0 + lane;      // tile row owned by this thread
        const int eset = (warp - 2) >> 2;  // epilogue set: columns [eset * BN / ESETS, (eset + 1) * BN / ESETS) of every tile
        const int ch0 = eset * CPS;
        float *aux_s = aux_tiles + (warp - 2) * (BN / ESETS);
        const uint32_t look_sa = smem_base + epi_off + (uint32_t)((warp - 2) * 32 * LOOK_PITCH * 4);
        // this warp's 32 staging rows
        uint64_t *stg = EPI == EPI_TOPK ? p.staged + (((int64_t)blockIdx.x * ESETS + eset) * BM + row0) * SC : nullptr;
        const uint32_t aux_sa = smem_u32(aux_s);
        int abuf = 0;
        uint32_t aphase = 0;
        for (int it = 0; it < total_rounds; ++it) {
            int m_tile, n_start, n_step, n_end;
            int64_t slot;
            if (!tc_round_item(S, cta, it, m_tile, n_start, n_step, slot, n_end)) continue;
            const int64_t qrow = (((int64_t)m_tile * CLM + pairc) * CG + crank) * BM + row;
            uint64_t thr = 0ull;
            float thr_f = PMM_DSKIP(p, 3) ? 103.0f : __uint_as_float(0x7fc00000u);
            if (PMM_DSKIP(p, 3)) thr = 1ull;
            const int kk = PMM_DSKIP(p, 3) ? -1 : p.k;
            int cnt = 0;
            unsigned rot = 0;
            const int max_flush = p.max_flush;
            float rowc = 0.0f;
            uint64_t seed_c = 0ull, ceil_c = ~0ull;
            uint64_t *list_base = nullptr;
            if (EPI == EPI_TOPK) {
                constexpr int KP = 32 * R;
                list_base = p.partial + (((slot * ESETS + eset) * GS + (crank4 & (uint32_t)(GS - 1))) * BM + row0) * KP;
                if (!p.resume) {
                    for (int i = lane; i < 32 * KP; i += 32) list_base[i] = 0ull;  // this warp's 32 empty lists
                } else if (!PMM_DSKIP(p, 3)) {  // continue: the threshold is the list's k-th entry
                    const uint64_t kth = list_base[(int64_t)lane * KP + (p.k - 1)];
                    thr = kth;
                    thr_f = kth == 0ull ? __uint_as_float(0x7fc00000u) : key_score(candidate_key(kth), true);
                }
                if (p.ceil) ceil_c = p.ceil[qrow];
                if (p.seed_thr) {  // "collect everything above this filter value" (NaN: no seed for this row)
                    const float sf = p.seed_thr[qrow];   // padded to the tile grid like q_aux
                    if (sf == sf) seed_c = pack_candidate(score_key(sf, true), 0xffffffffu);
                    if (seed_c > thr) {
                        thr = seed_c;
                        thr_f = key_score(candidate_key(seed_c), true);
                    }
                }
                // q_aux is padded to the tile grid. cosine: 1 unless the query norm is ~0; euclidean: |q|^2
                if (p.metric == METRIC_COSINE) rowc = p.q_aux[qrow] > p.norm_guard ? 1.0f : 0.0f;
                if (p.metric == METRIC_EUCLIDEAN) rowc = p.q_aux[qrow];
                __syncwarp();
            }
            if (SPLIT16) rowc = p.q_aux[qrow];   // this row's scale factor (padded to the tile grid)
            for (int nt = n_start; nt < n_end; nt += n_step) {
                const int64_t col_tile = (int64_t)nt * p.tile_stride * BN;
                if (EPI == EPI_TOPK && p.metric != METRIC_DOT) {  // stage this tile's corpus aux values (per warp)
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < CPS; ++i) {
                        const float a = __ldg(p.c_aux + col_tile + (ch0 + i) * 32 + lane);  // c_aux is padded to the tile grid
                        aux_s[i * 32 + lane] = p.metric == METRIC_COSINE ? (a > p.norm_guard ? __frcp_rn(a) : 0.0f) : a;
                    }
                    __syncwarp();
                }
                if (SPLIT16) {
                    // The tile's 256 column scale factors, ONE copy per CTA: each of the eight epilogue warps fetches an
                    // eighth (the load is in flight across the first barrier), the warps meet before the copy is
                    // overwritten (everyone is done with the previous tile's factors) and after it is complete.
                    static_assert(ESETS == 2 || !SPLIT16, "eight epilogue warps stage 32 column factors each");
                    const float mine = __ldg(p.c_aux + col_tile + (warp - 2) * 32 + lane);   // padded to the tile grid
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(colf_sa + (uint32_t)(((warp - 2) * 32 + lane) * 4)), "f"(mine) : "memory");
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                mbar_wait(tfull_bar(abuf), aphase);
                tc_fence_after();
                const bool edbg = p.debug_skip == 8 && warp == 2 && crank == 0;
                const long long e0 = edbg ? clock64() : 0ll;
#pragma unroll 1
                for (int ch = ch0; ch < ch0 + CPS; ++ch) {
                    uint32_t v[32];
                    if (EPI == EPI_TOPK && PMM_DSKIP(p, 2) && ch != ch0 + CPS - 1) continue;
                    tmem_ld_32x32(tmem_base + ((uint32_t)row0 << 16) + (uint32_t)(abuf * BN + ch * 32), v);
                    tmem_ld_wait();
                    if (ch == ch0 + CPS - 1) {  // this warp's TMEM reads of the buffer are done: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1) mbar_arrive(tempty_bar(abuf)); else mbar_arrive_cluster_relaxed(tempty_bar(abuf), leader_rank);
                        }
                    }
                    const int64_t col0 = col_tile + ch * 32;
                    if (EPI == EPI_TOPK && (PMM_DSKIP(p, 1) || PMM_DSKIP(p, 2))) continue;
                    if (EPI == EPI_MATMUL && SPLIT16) {   // undo the rows' power-of-two scaling: two exact multiplications
#pragma unroll
                        for (int i = 0; i < 8; ++i) {   // broadcast reads of the chunk's 32 factors; packed f32x2 multiplies
                            float c0, c1, c2, c3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c0), "=f"(c1), "=f"(c2), "=f"(c3) : "r"(colf_sa + (uint32_t)(ch * 128 + i * 16)));
                            scale2(v[4 * i + 0], v[4 * i + 1], rowc, c0, c1);
                            scale2(v[4 * i + 2], v[4 * i + 3], rowc, c2, c3);
                        }
                    }
                    if (EPI == EPI_MATMUL) {
                        if (p.out_tma) {
                            // registers -> swizzled 32x32 smem tile -> one TMA store per warp and chunk: full 128-byte
                            // lines, bounds clipped by the hardware
                            // one 4 KB store tile per warp (eight warps) or two per warp (four)
                            const uint32_t sbuf = smem_base + epi_off + (uint32_t)((ESETS == 2 ? (warp - 2) : (lg * 2 + (ch & 1))) * 4096);
                            if (lane == 0) tma_store_wait_read<ESETS == 2 ? 0 : 1>();  // the store that last used this buffer has read it
                            __syncwarp();
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const uint32_t dst = sbuf + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4));
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                                             "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                                             : "memory");
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tm_out, sbuf, (int32_t)col0, (int32_t)(((m_tile * CLM + (int)pairc) * CG + (int)crank) * BM + row0));
                                tma_store_commit();
                            }
                        } else if (qrow < p.nq) {
                            float *dst = p.out + qrow * p.n + col0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.n) dst[j] = __uint_as_float(v[j]);
                        }
                    } else if (p.metric == METRIC_DOT) {
                        filter_chunk<METRIC_DOT, R>(v, aux_sa + (ch - ch0) * 128, look_sa, rowc, col0, p.n, p.index_base, stg, list_base, lane, kk,
                                                    thr, thr_f, cnt, seed_c, ceil_c);
                    } else if (p.metric == METRIC_COSINE) {
                        filter_chunk<METRIC_COSINE, R>(v, aux_sa + (ch - ch0) * 128, look_sa, rowc, col0, p.n, p.index_base, stg, list_base, lane,
                                                       kk, thr, thr_f, cnt, seed_c, ceil_c);
                    } else {
                        filter_chunk<METRIC_EUCLIDEAN, R>(v, aux_sa + (ch - ch0) * 128, look_sa, rowc, col0, p.n, p.index_base, stg, list_base,
                                                          lane, kk, thr, thr_f, cnt, seed_c, ceil_c);
                    }
                }
                if (edbg && lane == 0)
                    atomicAdd(&g_tc_wait[20 + (31 - __clz((nt - n_start) / n_step + 1))], (unsigned long long)(clock64() - e0));
                if (EPI == EPI_TOPK) {  // regular flush: the TMEM buffer is already back with the MMA warp
                    // Rows of a warp fill their staging slots at the same rate, so they come due in bursts; a burst
                    // of 32 row merges takes several tiles' worth of MMA time and stalls the accumulator ring.
                    // Spread it: at most `max_flush` rows per tile (rows that are nearly full first), the rest
                    // keep staging (EMERGENCY_AT still bounds them).
                    unsigned due = __ballot_sync(0xffffffffu, cnt > p.soft_at);
                    if (__popc(due) > max_flush && nt - n_start >= 32 * n_step) {  // (early tiles: thresholds move fast, merge all)
                        unsigned pick = __ballot_sync(0xffffffffu, cnt > URGENT_AT);
                        unsigned rest = due & ~pick;
                        rest = __funnelshift_r(rest, rest, rot);  // rotate so that no row is always last
                        int room = max_flush - __popc(pick);
                        while (room > 0 && rest) {
                            const unsigned low = rest & (0u - rest);
                            pick |= __funnelshift_l(low, low, rot);
                            rest ^= low;
                            --room;
                        }
                        due = pick;
                        rot = (rot + 7) & 31;
                    }
                    if (due) {
                        const long long f0 = p.debug_skip == 8 ? clock64() : 0ll;
                        flush_rows<R>(due, stg, list_base, lane, kk, thr, thr_f, cnt, seed_c);
                        if (edbg && lane == 0) {
                            atomicAdd(&g_tc_wait[3], (unsigned long long)(clock64() - f0));
                            atomicAdd(&g_tc_wait[36 + (31 - __clz((nt - n_start) / n_step + 1))], (unsigned long long)(clock64() - f0));
                        }
                    }
                }
                abuf ^= 1;
                if (abuf == 0) aphase ^= 1u;
            }
            if (EPI == EPI_TOPK) {
                const unsigned pending = __ballot_sync(0xffffffffu, cnt > 0);
                if (pending) flush_rows<R>(pending, stg, list_base, lane, kk, thr, thr_f, cnt, seed_c);
            }
        }
    }

    if (EPI == EPI_MATMUL && warp >= 2 && lane == 0) tma_store_wait_all<0>();  // smem must outlive the bulk stores
    __syncwarp();
    tc_fence_before();
    if (GS > 1) cluster_sync(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, 512);
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

thread_local char g_tc_err[256] = "";

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// [rows x cols] row-major plane, box = box_rows x 128 bytes, SWIZZLE_128B.
bool make_plane_map(CUtensorMap *m, const void *base, int64_t rows, int64_t cols, int box_rows, bool f16, int rowb) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled not available from the driver");
        return false;
    }
    const int esz = f16 ? 2 : 4;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
    cuuint32_t box[2] = {(cuuint32_t)(rowb / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base),
                    gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return false;
    }
    return true;
}

// [rows x cols] f32 row-major output, box 32 x 32, SWIZZLE_128B (matmul epilogue TMA store).
bool make_out_map(CUtensorMap *m, const void *base, int64_t rows, int64_t cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled(out) failed with CUresult %d", (int)r);
        return false;
    }
    return true;
}

template <bool F16, int EPI, int R, int ROWB, int CG, int TERMS, int CLM = 1>
cudaError_t launch_t2(const TcArgs &a, cudaStream_t s) {
    constexpr int ESETS = tc_esets(F16, EPI, TERMS);
    typedef TcCfg<F16, ROWB, CG, TERMS, CLM, ESETS, EPI> Cfg;
    constexpr bool ONE = Cfg::PLANES == 1 && !(F16 && TERMS == 2);   // the f16 split needs the lo planes' maps as well
    CUtensorMap tq_hi, tq_lo, tc_hi, tc_lo;
    if (!make_plane_map(&tq_hi, a.q_hi, a.q_rows_pad, a.dim_pad, BM, F16, ROWB)) return cudaErrorInvalidValue;
    if (!make_plane_map(&tc_hi, a.c_hi, a.c_rows_pad, a.dim_pad, Cfg::B_ROWS / CLM, F16, ROWB)) return cudaErrorInvalidValue;
    if (ONE) {
        tq_lo = tq_hi;
        tc_lo = tc_hi;
    } else {
        if (!make_plane_map(&tq_lo, a.q_lo, a.q_rows_pad, a.dim_pad, BM, F16, ROWB)) return cudaErrorInvalidValue;
        if (!make_plane_map(&tc_lo, a.c_lo, a.c_rows_pad, a.dim_pad, Cfg::B_ROWS / CLM, F16, ROWB)) return cudaErrorInvalidValue;
    }
    CUtensorMap t_out = tq_hi;  // placeholder for the top-k kernels
    int out_tma = 0;
    if (EPI == EPI_MATMUL && (a.n % 4) == 0 && (((uintptr_t)a.out) & 15) == 0) {
        if (!make_out_map(&t_out, a.out, a.nq, a.n)) return cudaErrorInvalidValue;
        out_tma = 1;
    }
    TcKParams p;
    p.out_tma = out_tma;
    p.sched = a.sched;
    p.num_kb = (int)(a.dim_pad / Cfg::BK);
    p.q_aux = a.q_aux;
    p.c_aux = a.c_aux;
    p.nq = a.nq;
    p.n = a.n;
    p.index_base = a.index_base;
    p.metric = a.metric;
    p.k = a.k;
    p.partial = a.partial;
    p.staged = a.staged;
    p.out = a.out;
    p.round_sync = a.round_sync;
    p.sync_tiles = a.sync_tiles > 0 ? a.sync_tiles : 1;
    p.debug_skip = a.debug_skip;
    p.resume = a.resume;
    p.seed_thr = a.seed_thr;
    p.ceil = a.ceil;
    p.tile_stride = a.tile_stride > 1 ? a.tile_stride : 1;
    p.norm_guard = a.norm_guard > 0.0f ? a.norm_guard : 1e-6f;
    p.soft_at = a.soft_at > 0 ? a.soft_at : SOFT_AT;
    // merges that fit beside one tile's MMA time: a merge costs about as much as 8 k-blocks of one plane
    p.max_flush = a.max_flush > 0 ? a.max_flush : (p.num_kb * a.terms / 8 > 1 ? p.num_kb * a.terms / 8 : 1);
    p.sync_slack = a.sync_slack > 0 ? a.sync_slack : 0;
    auto kern = tc_kernel<F16, EPI, R, ROWB, CG, TERMS, CLM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(a.sched.num_ctas * CG * CLM));
    cfg.blockDim = dim3(tc_threads(ESETS));
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    // cluster_x: CG*CLM, or 4 for independent CTA pairs co-scheduled two per cluster (a.cluster4)
    attr[0].val.clusterDim.x = (CG == 2 && CLM == 1 && a.cluster4 && (a.sched.num_ctas % 2) == 0) ? 4 : CG * CLM;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, tq_hi, tq_lo, tc_hi, tc_lo, t_out, p);
}

template <bool F16, int EPI, int R>
cudaError_t launch_t(const TcArgs &a, cudaStream_t s) {
    if (EPI == EPI_TOPK && a.cg == 2 && a.clm == 2) {   // cluster of two pairs with corpus multicast
        if (!F16 && a.terms == 1) return launch_t2<false, EPI, R, 128, 2, 1, 2>(a, s);
        return launch_t2<F16, EPI, R, 128, 2, 3, 2>(a, s);
    }
    if (!F16 && EPI == EPI_TOPK && a.terms == 1) {   // first-level filter: TF32 x1 (cta_group::2 only)
        return launch_t2<false, EPI, R, 128, 2, 1>(a, s);
    }
    if (a.cg == 2) return launch_t2<F16, EPI, R, 128, 2, 3>(a, s);
    return launch_t2<F16, EPI, R, 128, 1, 3>(a, s);
}

}  // namespace

const char *tc_last_error() { return g_tc_err; }

bool tc_supported() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return false;
    return prop.major == 10 && get_encode_fn() != nullptr;
}

TcSchedule make_tc_schedule(int64_t q_rows, int64_t c_rows, int num_units, int group, int cg, int64_t layout_rows) {
    TcSchedule s;
    const int tile_m = BM * cg;   // query rows per scheduling unit; cg = CTAs per unit (1, 2, or 4 for a cluster of two pairs)
    s.m_tiles = (int)((q_rows + tile_m - 1) / tile_m);
    s.n_tiles = (int)((c_rows + BN - 1) / BN);
    // the sharing factors (and with them the list layout) are capped by the corpus tiles of `layout_rows` rows:
    // launches over different corpus chunks that carry their lists along pass the smallest chunk here
    const int cap_tiles = layout_rows > 0 ? (int)((layout_rows + BN - 1) / BN) : s.n_tiles;
    int G = num_units > 0 ? num_units : 1;
    s.g = group < 1 ? 1 : group;
    if (s.g > cap_tiles) s.g = cap_tiles;
    if (s.g > G) s.g = G;
    s.mc = G / s.g;
    s.rounds = s.m_tiles / s.mc;
    s.m_full = s.rounds * s.mc;
    s.m_rem = s.m_tiles - s.m_full;
    s.g_rem = 0;
    if (s.m_rem > 0) {
        s.g_rem = G / s.m_rem;
        if (s.g_rem > cap_tiles) s.g_rem = cap_tiles;
        if (s.g_rem < 1) s.g_rem = 1;
    }
    s.flat = 0;
    s.n_main = 0;
    int used_full = s.rounds > 0 ? s.mc * s.g : 0;
    int used_rem = s.m_rem * s.g_rem;
    s.num_ctas = used_full > used_rem ? used_full : used_rem;
    if (s.num_ctas < 1) s.num_ctas = 1;
    return s;
}

TcSchedule make_tc_schedule_flat(int64_t q_rows, int64_t c_rows, int num_units, int cg) {
    TcSchedule s = make_tc_schedule(q_rows, c_rows, num_units, 1, cg);
    const int64_t total = (int64_t)s.m_tiles * s.n_tiles;
    s.flat = 1;
    s.num_ctas = (int)(total < (int64_t)(num_units > 0 ? num_units : 1) ? (total > 0 ? total : 1) : (num_units > 0 ? num_units : 1));
    const int64_t per = (total + s.num_ctas - 1) / s.num_ctas;   // the largest share; it may start anywhere inside a query tile
    s.rounds = (int)((per + s.n_tiles - 2) / s.n_tiles) + 1;
    if (s.rounds > s.m_tiles) s.rounds = s.m_tiles;
    s.g = 1;
    s.mc = s.num_ctas;
    s.m_full = s.m_tiles;
    s.m_rem = 0;
    s.g_rem = 0;
    return s;
}

TcSchedule make_tc_schedule_hybrid(int64_t q_rows, int64_t c_rows, int num_units, int cg) {
    TcSchedule s = make_tc_schedule(q_rows, c_rows, num_units, 1, cg);
    const int G = num_units > 0 ? num_units : 1;
    if (s.rounds != 0 || s.g_rem != 1 || s.m_tiles >= G || s.m_tiles < 1) return s;
    const int n_main = (int)(((int64_t)s.n_tiles * s.m_tiles + G - 1) / G);
    if (n_main < 1 || n_main >= s.n_tiles) return s;
    const int H = G - s.m_tiles, tail = s.n_tiles - n_main;
    const int64_t per = ((int64_t)s.m_tiles * tail + H - 1) / H;   // the largest helper share; it may start inside a tail
    s.flat = 2;
    s.n_main = n_main;
    s.num_ctas = G;
    s.rounds = (int)((per + tail - 2) / tail) + 1;
    if (s.rounds > s.m_tiles) s.rounds = s.m_tiles;
    if (s.rounds < 1) s.rounds = 1;
    s.g = 1;
    s.mc = G;
    s.m_full = s.m_tiles;
    s.m_rem = 0;
    s.g_rem = 0;
    return s;
}

void tc_debug_wait_cycles(unsigned long long out[52]) {
    unsigned long long zero[52] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_tc_wait, sizeof(zero));
    cudaMemcpyToSymbol(g_tc_wait, zero, sizeof(zero));
}

int tc_epilogue_sets(int f16, int terms) { return tc_esets(f16 != 0, EPI_TOPK, terms); }

int64_t tc_staged_bytes(int num_ctas, int esets) { return (int64_t)num_ctas * esets * BM * SC * 8; }

int64_t tc_sync_counters(const TcSchedule &s, int sync_tiles) {
    if (sync_tiles < 1) sync_tiles = 1;
    auto per_round = [&](int step) { return step > 0 ? ((s.n_tiles + step - 1) / step + sync_tiles - 1) / sync_tiles : 0; };
    return (int64_t)s.rounds * per_round(s.g) + (s.m_rem > 0 ? per_round(s.g_rem) : 0) + 1;
}

cudaError_t launch_tc_topk(const TcArgs &a, cudaStream_t s) {
    if (a.k < 1 || a.k > a.kp) return cudaErrorInvalidValue;
    if (a.f16) {
        if (a.kp == 32) return launch_t<true, EPI_TOPK, 1>(a, s);
        if (a.kp == 64) return launch_t<true, EPI_TOPK, 2>(a, s);
        if (a.kp == 128) return launch_t<true, EPI_TOPK, 4>(a, s);
        if (a.kp == 256) return launch_t<true, EPI_TOPK, 8>(a, s);
    } else {
        if (a.kp == 32) return launch_t<false, EPI_TOPK, 1>(a, s);
        if (a.kp == 64) return launch_t<false, EPI_TOPK, 2>(a, s);
        if (a.kp == 128) return launch_t<false, EPI_TOPK, 4>(a, s);
        if (a.kp == 256) return launch_t<false, EPI_TOPK, 8>(a, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_tc_matmul(const TcArgs &a, cudaStream_t s) {
    if (a.f16 && a.terms == 2) {   // hi/lo f16 split of f32 operands (CTA pairs only; resident query planes: D <= 256)
        if (a.cg != 2 || !a.q_lo || !a.c_lo || !a.q_aux || !a.c_aux || a.dim_pad > 256) return cudaErrorInvalidValue;
        return launch_t2<true, EPI_MATMUL, 1, 128, 2, 2>(a, s);
    }
    if (a.f16) return launch_t<true, EPI_MATMUL, 1>(a, s);
    return launch_t<false, EPI_MATMUL, 1>(a, s);
}

}  // namespace pmm
