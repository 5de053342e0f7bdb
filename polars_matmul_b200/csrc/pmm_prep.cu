// pmm_prep.cu — the norm-precompute / marshalling kernel (HBM-bound, streaming).
//
// Replaces, on device, the reference's
//   * Series -> dense matrix marshalling: array_chunked_to_matrix_* / list_chunked_to_matrix_*
//     (src/matmul.rs:167-286): fixed-size rows or i64 list offsets, null element -> 0, null row ->
//     zeros, short row zero padded (a LONGER row sets an error flag; the reference panics);
//   * row norms: compute_norms_* / compute_squared_norms_* (src/metrics.rs:367-393).  The reduction
//     follows ndarray 0.16 `unrolled_dot` exactly: 8 partial sums p0..p7 over chunks of 8 with a
//     separate multiply and add, combined as (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail
//     sequentially — so norms are bit-identical to the oracle's.
// and emits, in the same pass, the operand planes the tensor-core contraction reads:
//   MODE_DENSE : dense [n_rows x dim] copy in the working type (f32 or f64) for the SIMT path;
//   MODE_TF32  : hi/lo TF32 planes [rows_pad x dim_pad] for the 3xTF32 split
//                (hi = rna_tf32(x), lo = rna_tf32(x - hi); x = hi + lo to ~2^-24 relative);
//   MODE_SPLIT16: row-scaled hi/lo f16 planes + per-row scale factors for the raw f32 matmul (prep_split16_kernel);
//   MODE_F16   : f16 plane [rows_pad x dim_pad] for f16-stored input (exact upcast, kind::f16 MMA);
//   MODE_F16R  : f16 plane of f32 input ROUNDED to f16 (round to nearest even): 11 significant bits, the same
//                unit roundoff 2^-11 as TF32, at twice the tensor rate and half the bytes - the first-level
//                filter of the f32 top-k.  Values beyond the f16 range become inf and tiny ones subnormal;
//                the losslessness check accounts for both (pmm_rescore.cu).
// Padding rows/columns are written as zeros so TMA tiles never see garbage.
//
// Thread mapping: 8 lanes own one row (lane j of the group owns partial sum p_j), 4 rows per warp,
// 8 warps per block.  Each step the 8 lanes touch one 32-byte sector of the row; steps are unrolled
// so several sectors per row are in flight.
#include "pmm_common.cuh"
#include "pmm_kernels.h"

namespace pmm {

template <typename SRC> struct SrcLoad;
template <> struct SrcLoad<float> { template <typename W> static __device__ __forceinline__ W get(const float *p) { return (W)__ldg(p); } };
template <> struct SrcLoad<double> { template <typename W> static __device__ __forceinline__ W get(const double *p) { return (W)__ldg(p); } };
template <> struct SrcLoad<__half> { template <typename W> static __device__ __forceinline__ W get(const __half *p) { return (W)__half2float(__ldg(p)); } };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}

// Split x into two TF32-representable floats. Non-finite x: hi = x, lo = 0.
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo) {
    uint32_t xb = __float_as_uint(x);
    if ((xb & 0x7f800000u) == 0x7f800000u) { hi = x; lo = 0.0f; return; }
    uint32_t hb = tf32_rna(x);
    if ((hb & 0x7f800000u) == 0x7f800000u) hb = xb & 0xffffe000u;  // rounding overflowed: truncate
    hi = __uint_as_float(hb);
    lo = __uint_as_float(tf32_rna(__fsub_rn(x, hi)));
}

enum { MODE_DENSE = 0, MODE_TF32 = 1, MODE_F16 = 2, MODE_F16R = 3, MODE_NONE = 4 /* prep_fast_kernel only: norms, no planes */,
       MODE_SPLIT16 = 5 /* prep_split16_kernel */ };
static bool g_prep_fast = true;   // prep_set_fast(): A/B switch for measurements and tests
void prep_set_fast(bool on) { g_prep_fast = on; }

template <typename SRC, typename W, int MODE>
__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7;
    const int64_t row = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const unsigned gmask = 0xffu << (lane & 24);
    if (row >= a.rows_out) return;  // whole 8-lane group leaves together

    const SRC *values = (const SRC *)a.values;
    const int64_t dim = a.dim;
    int64_t base = 0, len = 0;
    if (row < a.n_rows) {
        bool row_ok = !a.row_validity || ((a.row_validity[row >> 3] >> (row & 7)) & 1);
        if (a.offsets) {
            base = a.offsets[row];
            len = a.offsets[row + 1] - base;
            if (len > dim) {  // reference: ndarray index out of bounds panic
                if (sub == 0) atomicExch(a.error_flag, 1);
                len = dim;
            }
        } else {
            base = row * dim;
            len = dim;
        }
        if (!row_ok) len = 0;
    }

    W *dense = (W *)a.out0;
    float *hi = (float *)a.out0, *lo = (float *)a.out1;
    __half *hp = (__half *)a.out0;
    const int64_t ld = a.ld_out;

    auto fetch = [&](int64_t i) -> W {
        if (i >= len) return (W)0;
        int64_t p = base + i;
        if (a.validity && !((a.validity[p >> 3] >> (p & 7)) & 1)) return (W)0;
        return SrcLoad<SRC>::template get<W>(values + p);
    };
    bool bad = false;   // this lane saw an inf / NaN element (raw matmul only: a.nonfinite_rows)
    auto emit = [&](int64_t i, W x) {
        if (MODE == MODE_TF32) bad |= !(fabs((double)x) <= 3.4028234663852886e38);
        if (MODE == MODE_DENSE) {
            dense[row * ld + i] = x;
        } else if (MODE == MODE_TF32) {
            float h, l;
            tf32_split((float)x, h, l);
            hi[row * ld + i] = h;
            lo[row * ld + i] = l;
        } else if (sizeof(W) == 8) {
            hp[row * ld + i] = __double2half((double)x);   // f64 source: one rounding, straight to f16
        } else {
            hp[row * ld + i] = __float2half_rn((float)x);  // MODE_F16: exact, x came from an f16; MODE_F16R: rounds
        }
    };

    const int64_t d8 = dim & ~(int64_t)7;
    W p = (W)0;
    int64_t t = 0;
    for (; t + 32 <= d8; t += 32) {  // 4 sectors in flight
        W x0 = fetch(t + sub), x1 = fetch(t + 8 + sub), x2 = fetch(t + 16 + sub), x3 = fetch(t + 24 + sub);
        p = add_rn(p, mul_rn(x0, x0));
        p = add_rn(p, mul_rn(x1, x1));
        p = add_rn(p, mul_rn(x2, x2));
        p = add_rn(p, mul_rn(x3, x3));
        emit(t + sub, x0);
        emit(t + 8 + sub, x1);
        emit(t + 16 + sub, x2);
        emit(t + 24 + sub, x3);
    }
    for (; t < d8; t += 8) {
        W x = fetch(t + sub);
        p = add_rn(p, mul_rn(x, x));
        emit(t + sub, x);
    }
    // (p0+p4), (p1+p5), (p2+p6), (p3+p7) then a sequential sum, as ndarray's unrolled_dot
    const int gbase = lane & 24;
    W other = __shfl_sync(gmask, p, gbase + ((sub + 4) & 7));
    W pair = add_rn(p, other);  // valid on sub 0..3
    W s0 = __shfl_sync(gmask, pair, gbase + 0), s1 = __shfl_sync(gmask, pair, gbase + 1);
    W s2 = __shfl_sync(gmask, pair, gbase + 2), s3 = __shfl_sync(gmask, pair, gbase + 3);
    W sum = add_rn((W)0, s0);
    sum = add_rn(sum, s1);
    sum = add_rn(sum, s2);
    sum = add_rn(sum, s3);
    for (int64_t i = d8; i < dim; ++i) {  // tail (< 8 elements), every lane redundantly
        W x = fetch(i);
        sum = add_rn(sum, mul_rn(x, x));
        if (sub == (int)(i - d8)) emit(i, x);
    }
    if (MODE != MODE_DENSE) {
        for (int64_t i = dim + sub; i < ld; i += 8) {  // zero the padding columns
            if (MODE == MODE_TF32) { hi[row * ld + i] = 0.0f; lo[row * ld + i] = 0.0f; }
            else hp[row * ld + i] = __float2half_rn(0.0f);
        }
    }
    if (MODE == MODE_TF32 && a.nonfinite_rows) {
        const bool any_bad = __any_sync(gmask, bad);
        if (sub == 0 && row < a.n_rows) {
            a.nonfinite_rows[row] = any_bad ? 1 : 0;
            if (any_bad) atomicAdd(a.nonfinite_count, 1u);
        }
    }
    if (sub == 0) {
        if (a.sqnorm_out) ((W *)a.sqnorm_out)[row] = sum;
        if (a.norm_out) ((W *)a.norm_out)[row] = sqrt_rn(sum);
        // f64 working precision on the tensor-core path: f32 copies for the filter kernel and its proof
        if (a.sqnorm32_out) a.sqnorm32_out[row] = (float)sum;
        if (a.norm32_out) a.norm32_out[row] = (float)sqrt_rn(sum);
    }
    if (a.max_sq_out) {  // norm range of the column (f32; an f64 norm beyond the f32 range becomes inf): one atomic per warp
        const float fs = (float)sum;
        unsigned int bits = (sub == 0 && row < a.n_rows && fs == fs) ? __float_as_uint(fs) : 0u;  // >= 0: orders like uint
        const unsigned am = __activemask();
        bits = __reduce_max_sync(am, bits);
        if (lane == __ffs(am) - 1 && bits) atomicMax(a.max_sq_out, bits);
        // [1]: smallest squared norm among the rows cosine does not treat as zero (norm > 1e-6)
        unsigned int lo_bits = (sub == 0 && row < a.n_rows && fs > a.zero_guard_sq) ? __float_as_uint(fs) : 0x7f800000u;
        lo_bits = __reduce_min_sync(am, lo_bits);
        if (lane == __ffs(am) - 1 && lo_bits != 0x7f800000u) atomicMin(a.max_sq_out + 1, lo_bits);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fast path of the plane modes for the common layout: fixed-size rows, no bitmaps, f32 working precision, row length a
// multiple of 64, 16-byte aligned rows.  Same thread mapping (8 lanes per row, lane j owns partial sum p_j) and therefore
// the same norm bits as prep_kernel, but every lane moves 8 consecutive elements at a time: 128-bit loads, and plane stores
// of 16 bytes per lane (f16 plane: the 8 lanes of a row write one whole 128-byte line per step; TF32 hi/lo: 256 bytes each)
// instead of 2- / 4-byte scalars; the elements a lane needs for ITS residue class come back through a per-warp
// shared-memory tile (the piece is written as loaded and read back transposed).
template <typename SRC, int MODE>
__global__ void __launch_bounds__(256) prep_fast_kernel(PrepArgs a) {
    constexpr int EPL = 8;                              // elements per lane and step: 32 bytes of f32 or 16 bytes of f16
    constexpr int STEP = 8 * EPL;                       // 64 elements per row and step: 256 / 128 bytes, whole lines
    constexpr int PITCH = STEP + 8;                     // floats per tile row: rows start 8 banks apart
    constexpr int NL = sizeof(SRC) == 4 ? 2 : 1;        // 16-byte loads per lane and step
    __shared__ __align__(16) float tiles[8][2][4][PITCH];   // per warp: double-buffered 4-row tile (as floats)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, rw = lane >> 3;
    const int64_t row = ((int64_t)blockIdx.x * 8 + warp) * 4 + rw;
    const unsigned gmask = 0xffu << (lane & 24);
    if (row >= a.rows_out) return;  // whole 8-lane group leaves together
    const int64_t dim = a.dim, ld = a.ld_out;
    const bool live = row < a.n_rows;
    const SRC *src = (const SRC *)a.values + row * dim;
    float *hi = (float *)a.out0 + row * ld, *lo = (float *)a.out1 + row * ld;
    __half *hp = (__half *)a.out0 + row * ld;
    float p = 0.0f;
    bool bad = false;
    int buf = 0;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    struct Piece {
        uint4 v[NL];
    };
    auto load = [&](int64_t t) {   // this lane's 8 consecutive elements [t + 8 sub, t + 8 sub + 8) of the row
        Piece pc;
        const uint4 *s16 = (const uint4 *)(src + t + EPL * sub);
#pragma unroll
        for (int u = 0; u < NL; ++u) pc.v[u] = live ? __ldg(s16 + u) : zero4;
        return pc;
    };
    // one piece of the row (already loaded) at element offset t: planes + this lane's share of the norm
    auto process = [&](const Piece &pc, int64_t t) {
        float x[EPL];
        if (sizeof(SRC) == 4) {
#pragma unroll
            for (int u = 0; u < NL; ++u) {
                x[4 * u] = __uint_as_float(pc.v[u].x); x[4 * u + 1] = __uint_as_float(pc.v[u].y);
                x[4 * u + 2] = __uint_as_float(pc.v[u].z); x[4 * u + 3] = __uint_as_float(pc.v[u].w);
            }
        } else {
            const __half2 *h2 = (const __half2 *)&pc.v[0];
#pragma unroll
            for (int u = 0; u < EPL / 2; ++u) {
                const float2 f = __half22float2(h2[u]);
                x[2 * u] = f.x;
                x[2 * u + 1] = f.y;
            }
            if (MODE == MODE_F16 && a.out0) *((uint4 *)(hp + t) + sub) = pc.v[0];   // exact plane of f16 input: the line as loaded
        }
        if (MODE == MODE_F16R || (MODE == MODE_F16 && sizeof(SRC) == 4)) {
            __half2 h[EPL / 2];   // 8 halves = 16 bytes per lane: the 8 lanes of a row write one whole 128-byte line
#pragma unroll
            for (int u = 0; u < EPL / 2; ++u) h[u] = __floats2half2_rn(x[2 * u], x[2 * u + 1]);
            *((uint4 *)(hp + t) + sub) = *(const uint4 *)h;
        } else if (MODE == MODE_TF32) {
            float h4[EPL], l4[EPL];
#pragma unroll
            for (int u = 0; u < EPL; ++u) {
                tf32_split(x[u], h4[u], l4[u]);
                bad |= !(fabsf(x[u]) <= 3.4028234663852886e38f);
            }
#pragma unroll
            for (int u = 0; u < EPL; u += 4) {
                *((float4 *)(hi + t + EPL * sub + u)) = make_float4(h4[u], h4[u + 1], h4[u + 2], h4[u + 3]);
                *((float4 *)(lo + t + EPL * sub + u)) = make_float4(l4[u], l4[u + 1], l4[u + 2], l4[u + 3]);
            }
        }
        // norm: lane `sub` adds the elements of ITS residue class (mod 8) in index order, as ndarray's unrolled_dot
        float *tile = tiles[warp][buf][rw];
#pragma unroll
        for (int u = 0; u < EPL; u += 4) *(float4 *)(tile + EPL * sub + u) = make_float4(x[u], x[u + 1], x[u + 2], x[u + 3]);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < STEP / 8; ++c) {
            const float y = tile[8 * c + sub];
            p = add_rn(p, mul_rn(y, y));
        }
        buf ^= 1;   // (no second barrier: the next piece writes the OTHER buffer, the one after that is behind the next __syncwarp)
    };
    int64_t t = 0;
    for (; t + 4 * STEP <= dim; t += 4 * STEP) {   // four pieces (1 KB of f32 per row) in flight per lane group
        Piece r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = load(t + u * STEP);
#pragma unroll
        for (int u = 0; u < 4; ++u) process(r[u], t + u * STEP);
    }
    for (; t < dim; t += STEP) process(load(t), t);
    // (dim is a multiple of 64 here, and so is every plane's leading dimension: no padding columns to zero)
    // (p0+p4), (p1+p5), (p2+p6), (p3+p7) then a sequential sum, as ndarray's unrolled_dot (no tail: dim % 8 == 0)
    const int gbase = lane & 24;
    const float other = __shfl_sync(gmask, p, gbase + ((sub + 4) & 7));
    const float pair = add_rn(p, other);
    const float s0 = __shfl_sync(gmask, pair, gbase + 0), s1 = __shfl_sync(gmask, pair, gbase + 1);
    const float s2 = __shfl_sync(gmask, pair, gbase + 2), s3 = __shfl_sync(gmask, pair, gbase + 3);
    float sum = add_rn(0.0f, s0);
    sum = add_rn(sum, s1);
    sum = add_rn(sum, s2);
    sum = add_rn(sum, s3);
    if (MODE == MODE_TF32 && a.nonfinite_rows) {
        const bool any_bad = __any_sync(gmask, bad);
        if (sub == 0 && live) {
            a.nonfinite_rows[row] = any_bad ? 1 : 0;
            if (any_bad) atomicAdd(a.nonfinite_count, 1u);
        }
    }
    if (sub == 0) {
        if (a.sqnorm_out) ((float *)a.sqnorm_out)[row] = sum;
        if (a.norm_out) ((float *)a.norm_out)[row] = sqrt_rn(sum);
    }
    if (a.max_sq_out) {
        unsigned int bits = (sub == 0 && live && sum == sum) ? __float_as_uint(sum) : 0u;
        const unsigned am = __activemask();
        bits = __reduce_max_sync(am, bits);
        if (lane == __ffs(am) - 1 && bits) atomicMax(a.max_sq_out, bits);
        unsigned int lo_bits = (sub == 0 && live && sum > a.zero_guard_sq) ? __float_as_uint(sum) : 0x7f800000u;
        lo_bits = __reduce_min_sync(am, lo_bits);
        if (lane == __ffs(am) - 1 && lo_bits != 0x7f800000u) atomicMin(a.max_sq_out + 1, lo_bits);
    }
}

// ---- MODE_SPLIT16: row-scaled hi/lo f16 planes for the raw f32 matmul -----------------------------------------------
// One warp per output row.  Pass 1: largest |element| of the row (and: any inf / NaN?).  The row is scaled by 2^e with
// e = 14 - floor(log2(max)), an exact operation, so that the largest element lands in [2^14, 2^15) - inside the f16
// range with four binades of head room for the 16-element sums of the MMA - and every element within 2^-17 of the
// largest keeps a NORMAL lo part.  Pass 2 (the row is still in L1/L2): hi = f16(x 2^e), lo = f16(x 2^e - hi); smaller
// elements round with an absolute error <= 2^-25 in scaled units, i.e. <= 2^-39 of the row's largest element.
// scale_out[row] = 2^-e undoes the scaling in the matmul epilogue (two exact multiplications per output).
// Rows holding inf / NaN, or whose largest element lies outside [2^-60, 2^60] (the two factors of an output must stay
// representable), are written as zeros and marked for the IEEE fix-up pass like the non-finite rows of the TF32 split.
// LANES lanes per row (32, or 16 for short rows: two rows per warp, half as many warps to schedule).
template <typename SRC, int LANES>
__global__ void __launch_bounds__(256) prep_split16_kernel(PrepArgs a) {
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned gmask = LANES == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
    const int64_t row = (int64_t)blockIdx.x * (256 / LANES) + (threadIdx.x / LANES);
    if (row >= a.rows_out) return;   // (whole lane groups leave together)
    const SRC *values = (const SRC *)a.values;
    int64_t base = 0, len = 0;
    if (row < a.n_rows) {
        const bool row_ok = !a.row_validity || ((a.row_validity[row >> 3] >> (row & 7)) & 1);
        if (a.offsets) {
            base = a.offsets[row];
            len = a.offsets[row + 1] - base;
            if (len > a.dim) {
                if (lane == 0) *a.error_flag = 1;
                len = a.dim;
            }
        } else {
            base = row * a.dim;
            len = a.dim;
        }
        if (!row_ok) len = 0;
    }
    auto fetch = [&](int64_t i) -> float {
        if (i >= len) return 0.0f;
        const int64_t p = base + i;
        if (a.validity && !((a.validity[p >> 3] >> (p & 7)) & 1)) return 0.0f;
        return SrcLoad<SRC>::template get<float>(values + p);
    };
    float mx = 0.0f;
    bool bad = false;
    for (int64_t i = lane; i < len; i += LANES) {
        const float x = fetch(i);
        bad |= (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;
        mx = fmaxf(mx, fabsf(x));
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, o));
    bad = (__ballot_sync(gmask, bad) & gmask) != 0u;
    int e = 0;
    if (mx > 0.0f) {
        const int lg = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;   // floor(log2(mx)) for normal mx; -127 for subnormals
        if (lg < -60 || lg > 60) bad = true;
        e = 14 - lg;
    }
    if (bad) e = 0;
    const float up = __uint_as_float((uint32_t)(127 + e) << 23);      // 2^e   (|e| <= 74)
    const float down = __uint_as_float((uint32_t)(127 - e) << 23);    // 2^-e
    __half2 *hi = (__half2 *)((__half *)a.out0 + row * a.ld_out);
    __half2 *lo = (__half2 *)((__half *)a.out1 + row * a.ld_out);
    for (int64_t i = 2 * lane; i < a.ld_out; i += 2 * LANES) {   // ld_out is a multiple of 64: 128- / 64-byte stores per lane group
        float x0 = 0.0f, x1 = 0.0f;
        if (!bad) {
            x0 = fetch(i) * up;
            x1 = fetch(i + 1) * up;
        }
        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        const __half l0 = __float2half_rn(__fsub_rn(x0, __half2float(h0))), l1 = __float2half_rn(__fsub_rn(x1, __half2float(h1)));
        hi[i >> 1] = __halves2half2(h0, h1);
        lo[i >> 1] = __halves2half2(l0, l1);
    }
    if (lane == 0) {
        if (a.scale_out) a.scale_out[row] = down;
        if (a.nonfinite_rows && row < a.n_rows) {
            a.nonfinite_rows[row] = bad ? 1 : 0;
            if (bad) atomicAdd(a.nonfinite_count, 1u);
        }
    }
}

template <typename SRC>
static bool prep_fast_ok(const PrepArgs &a) {
    const int step = 64;
    return !a.offsets && !a.validity && !a.row_validity && !a.norm32_out && !a.sqnorm32_out && a.dim > 0 && (a.dim % step) == 0 &&
           a.ld_out == a.dim && (((uintptr_t)a.values) & 15) == 0 && (((uintptr_t)a.out0) & 15) == 0 && (((uintptr_t)a.out1) & 15) == 0;
}

template <typename SRC, int MODE>
static cudaError_t launch_prep_fast(const PrepArgs &a, int64_t groups, cudaStream_t s) {
    prep_fast_kernel<SRC, MODE><<<(unsigned)groups, 256, 0, s>>>(a);
    return cudaGetLastError();
}

template <typename SRC, typename W, int MODE>
static cudaError_t launch_prep_t(const PrepArgs &a, cudaStream_t s) {
    int64_t groups = (a.rows_out + 31) / 32;
    if (groups <= 0) return cudaSuccess;
    if (MODE != MODE_DENSE && sizeof(W) == 4 && sizeof(SRC) <= 4 && g_prep_fast && prep_fast_ok<SRC>(a)) {
        if (sizeof(SRC) == 4) return launch_prep_fast<float, MODE == MODE_DENSE ? MODE_F16R : MODE>(a, groups, s);
        return launch_prep_fast<__half, MODE == MODE_DENSE ? MODE_F16R : MODE>(a, groups, s);
    }
    prep_kernel<SRC, W, MODE><<<(unsigned)groups, 256, 0, s>>>(a);
    return cudaGetLastError();
}

// src_dtype: PMM_DTYPE_* of `values`; mode selects the product; `work_f64` (0: f32, 1: f64) is the type the norms
// are computed and written in (and of the DENSE copy).  Plane modes with f64 working precision: the planes are the
// tensor-core FILTER's operands (rounded to f16 / split to TF32 from the f64 value), the norms stay exact f64.
cudaError_t launch_prep(const PrepArgs &a, int src_dtype, int mode, int work_f64, cudaStream_t s) {
    if (mode == MODE_SPLIT16) {   // raw f32 matmul operands: planes + scale factors, no norms
        if (work_f64 || src_dtype > 1 || !a.out0 || !a.out1 || (a.ld_out & 63)) return cudaErrorInvalidValue;
        if (a.rows_out <= 0) return cudaSuccess;
        if (a.ld_out <= 256) {   // short rows: 16 lanes per row
            const int64_t blocks = (a.rows_out + 15) / 16;
            if (src_dtype == 1) prep_split16_kernel<float, 16><<<(unsigned)blocks, 256, 0, s>>>(a);
            else prep_split16_kernel<__half, 16><<<(unsigned)blocks, 256, 0, s>>>(a);
        } else {
            const int64_t blocks = (a.rows_out + 7) / 8;
            if (src_dtype == 1) prep_split16_kernel<float, 32><<<(unsigned)blocks, 256, 0, s>>>(a);
            else prep_split16_kernel<__half, 32><<<(unsigned)blocks, 256, 0, s>>>(a);
        }
        return cudaGetLastError();
    }
    if (work_f64 && (mode == MODE_F16R || mode == MODE_TF32)) {
        if (mode == MODE_F16R) {
            if (src_dtype == 0) return launch_prep_t<__half, double, MODE_F16R>(a, s);
            if (src_dtype == 1) return launch_prep_t<float, double, MODE_F16R>(a, s);
            return launch_prep_t<double, double, MODE_F16R>(a, s);
        }
        if (src_dtype == 0) return launch_prep_t<__half, double, MODE_TF32>(a, s);
        if (src_dtype == 1) return launch_prep_t<float, double, MODE_TF32>(a, s);
        return launch_prep_t<double, double, MODE_TF32>(a, s);
    }
    if (mode == MODE_TF32) {
        if (src_dtype == 1) return launch_prep_t<float, float, MODE_TF32>(a, s);
        if (src_dtype == 0) return launch_prep_t<__half, float, MODE_TF32>(a, s);
        return cudaErrorInvalidValue;
    }
    if (mode == MODE_F16) {
        if (src_dtype == 0) return launch_prep_t<__half, float, MODE_F16>(a, s);
        return cudaErrorInvalidValue;
    }
    if (mode == MODE_F16R) {
        if (src_dtype == 1) return launch_prep_t<float, float, MODE_F16R>(a, s);
        if (src_dtype == 0) return launch_prep_t<__half, float, MODE_F16R>(a, s);
        return cudaErrorInvalidValue;
    }
    if (work_f64) {
        if (src_dtype == 0) return launch_prep_t<__half, double, MODE_DENSE>(a, s);
        if (src_dtype == 1) return launch_prep_t<float, double, MODE_DENSE>(a, s);
        return launch_prep_t<double, double, MODE_DENSE>(a, s);
    }
    if (src_dtype == 0) return launch_prep_t<__half, float, MODE_DENSE>(a, s);
    if (src_dtype == 1) return launch_prep_t<float, float, MODE_DENSE>(a, s);
    return cudaErrorInvalidValue;
}

// f16 rows, fixed size, no bitmaps, even dim: each lane loads one f16 PAIR per 16 elements (4-byte requests, a full
// 32-byte sector per row and step like the f32 path) and the 8 lanes of the row hand each other the elements of
// their own residue class mod 8 by shuffles, so every partial sum p_j still adds its elements in index order.
// Consumes whole 16-element steps from *t on; returns the partial sum of this lane.
template <typename W>
__device__ __forceinline__ W norms_f16_pairs(const void *values, int64_t base, bool row_live, int64_t d8, int64_t *t, int sub,
                                             int gbase, unsigned gmask, W p) {
    const __half2 *v2 = (const __half2 *)values;
    int64_t i = *t;
    for (; i + 64 <= d8; i += 64) {  // 4 sectors in flight per row
        unsigned bits[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            __half2 h = row_live ? __ldg(v2 + ((base + i + 16 * u + 2 * sub) >> 1)) : __floats2half2_rn(0.0f, 0.0f);
            bits[u] = *(unsigned *)&h;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned lo = __shfl_sync(gmask, bits[u], gbase + (sub >> 1));      // holds element i+16u+sub
            unsigned hi = __shfl_sync(gmask, bits[u], gbase + 4 + (sub >> 1));  // holds element i+16u+8+sub
            const __half2 a2 = *(__half2 *)&lo, b2 = *(__half2 *)&hi;
            const W x0 = (W)((sub & 1) ? __high2float(a2) : __low2float(a2));
            const W x1 = (W)((sub & 1) ? __high2float(b2) : __low2float(b2));
            p = add_rn(p, mul_rn(x0, x0));
            p = add_rn(p, mul_rn(x1, x1));
        }
    }
    *t = i;
    return p;
}

// Norms only (pmm_dev_norms): same reduction, no planes written.
template <typename SRC, typename W>
__global__ void __launch_bounds__(256) norms_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7;
    const int64_t row = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const unsigned gmask = 0xffu << (lane & 24);
    if (row >= a.n_rows) return;
    const SRC *values = (const SRC *)a.values;
    const int64_t dim = a.dim;
    int64_t base, len;
    bool row_ok = !a.row_validity || ((a.row_validity[row >> 3] >> (row & 7)) & 1);
    if (a.offsets) { base = a.offsets[row]; len = a.offsets[row + 1] - base; if (len > dim) len = dim; }
    else { base = row * dim; len = dim; }
    if (!row_ok) len = 0;
    auto fetch = [&](int64_t i) -> W {
        if (i >= len) return (W)0;
        int64_t p = base + i;
        if (a.validity && !((a.validity[p >> 3] >> (p & 7)) & 1)) return (W)0;
        return SrcLoad<SRC>::template get<W>(values + p);
    };
    const int64_t d8 = dim & ~(int64_t)7;
    W p = (W)0;
    int64_t t = 0;
    if (sizeof(SRC) == 2 && !a.offsets && !a.validity && (dim & 1) == 0)
        p = norms_f16_pairs<W>(a.values, base, len > 0, d8, &t, sub, lane & 24, gmask, p);
    for (; t + 64 <= d8; t += 64) {  // 8 sectors in flight per row
        W x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = fetch(t + 8 * u + sub);
#pragma unroll
        for (int u = 0; u < 8; ++u) p = add_rn(p, mul_rn(x[u], x[u]));
    }
    for (; t < d8; t += 8) { W x = fetch(t + sub); p = add_rn(p, mul_rn(x, x)); }
    const int gbase = lane & 24;
    W other = __shfl_sync(gmask, p, gbase + ((sub + 4) & 7));
    W pair = add_rn(p, other);
    W s0 = __shfl_sync(gmask, pair, gbase + 0), s1 = __shfl_sync(gmask, pair, gbase + 1);
    W s2 = __shfl_sync(gmask, pair, gbase + 2), s3 = __shfl_sync(gmask, pair, gbase + 3);
    W sum = add_rn((W)0, s0);
    sum = add_rn(sum, s1);
    sum = add_rn(sum, s2);
    sum = add_rn(sum, s3);
    for (int64_t i = d8; i < dim; ++i) { W x = fetch(i); sum = add_rn(sum, mul_rn(x, x)); }
    if (sub == 0) {
        if (a.sqnorm_out) ((W *)a.sqnorm_out)[row] = sum;
        if (a.norm_out) ((W *)a.norm_out)[row] = sqrt_rn(sum);
    }
}

cudaError_t launch_norms(const PrepArgs &a, int src_dtype, cudaStream_t s) {
    int64_t groups = (a.n_rows + 31) / 32;
    if (groups <= 0) return cudaSuccess;
    // f16 rows of a plain layout: the 16-byte-load pass of the plane builder without its plane stores (4-byte pair loads
    // keep too few bytes in flight: 0.61 of HBM)
    PrepArgs b = a;
    b.ld_out = a.dim;
    b.out0 = b.out1 = nullptr;
    b.zero_guard_sq = 1e-12f;
    if (src_dtype == 0 && g_prep_fast && a.rows_out == a.n_rows && prep_fast_ok<__half>(b)) return launch_prep_fast<__half, MODE_NONE>(b, groups, s);
    if (src_dtype == 0) norms_kernel<__half, float><<<(unsigned)groups, 256, 0, s>>>(a);
    else if (src_dtype == 1) norms_kernel<float, float><<<(unsigned)groups, 256, 0, s>>>(a);
    else norms_kernel<double, double><<<(unsigned)groups, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace pmm
