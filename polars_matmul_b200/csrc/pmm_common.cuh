// pmm_common.cuh — shared device helpers: metric post-pass, ordered score keys, packed candidates,
// warp-level bitonic primitives.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pmm {

constexpr int METRIC_COSINE = 0;
constexpr int METRIC_DOT = 1;
constexpr int METRIC_EUCLIDEAN = 2;

__host__ __device__ inline bool higher_is_better(int metric) { return metric != METRIC_EUCLIDEAN; }

// ------------------------------------------------------------------------------------------------
// Metric post-pass. Mirrors the reference op for op (src/metrics.rs:323-362 f32, :267-308 f64):
// cosine = dot / (qn*cn) with the product formed first and ONE divide, zero-norm guards 1e-6 / 1e-10;
// euclidean = sqrt(max((qsq + csq) - 2*dot, 0)).  Explicit _rn intrinsics: never contracted to FMA.
// `qa`/`ca` are norms for cosine, squared norms for euclidean, ignored for dot.
__device__ __forceinline__ float metric_finish(float dot, int metric, float qa, float ca) {
    if (metric == METRIC_COSINE) {
        if (qa > 1e-6f) {
            if (ca > 1e-6f) return __fdiv_rn(dot, __fmul_rn(qa, ca));
            return 0.0f;
        }
        return 0.0f;
    }
    if (metric == METRIC_EUCLIDEAN) {
        float sq = __fsub_rn(__fadd_rn(qa, ca), __fmul_rn(2.0f, dot));
        return __fsqrt_rn(fmaxf(sq, 0.0f));  // fmaxf drops NaN like Rust's f32::max
    }
    return dot;
}
__device__ __forceinline__ double metric_finish(double dot, int metric, double qa, double ca) {
    if (metric == METRIC_COSINE) {
        if (qa > 1e-10) {
            if (ca > 1e-10) return __ddiv_rn(dot, __dmul_rn(qa, ca));
            return 0.0;
        }
        return 0.0;
    }
    if (metric == METRIC_EUCLIDEAN) {
        double sq = __dsub_rn(__dadd_rn(qa, ca), __dmul_rn(2.0, dot));
        return __dsqrt_rn(fmax(sq, 0.0));
    }
    return dot;
}

// ------------------------------------------------------------------------------------------------
// Ordered keys: an unsigned integer that is LARGER for a BETTER score.
//   higher-is-better: monotone increasing in the score; lower-is-better: monotone decreasing.
//   NaN -> 0 (ranks last); -0.0 is folded into +0.0 first so numerically equal scores tie.
__device__ __forceinline__ uint32_t score_key(float s, bool higher) {
    if (s != s) return 0u;
    s = __fadd_rn(s, 0.0f);
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return higher ? u : ~u;
}
__host__ __device__ __forceinline__ float key_score(uint32_t key, bool higher) {
    if (key == 0u) {
#ifdef __CUDA_ARCH__
        return __uint_as_float(0x7fc00000u);
#else
        union { uint32_t u; float f; } c; c.u = 0x7fc00000u; return c.f;
#endif
    }
    uint32_t u = higher ? key : ~key;
    u = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ uint64_t score_key(double s, bool higher) {
    if (s != s) return 0ull;
    s = __dadd_rn(s, 0.0);
    uint64_t u = (uint64_t)__double_as_longlong(s);
    u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    return higher ? u : ~u;
}
__device__ __forceinline__ double key_score(uint64_t key, bool higher) {
    if (key == 0ull) return __longlong_as_double(0x7ff8000000000000ll);
    uint64_t u = higher ? key : ~key;
    u = (u & 0x8000000000000000ull) ? (u ^ 0x8000000000000000ull) : ~u;
    return __longlong_as_double((long long)u);
}

// Packed candidate (f32 working precision): (key << 32) | ~index.  A plain u64 compare implements the
// total order "better score first, then lower index first".  0 is the empty slot (no real candidate
// packs to 0 because indices stay below 2^32 - 1).
__host__ __device__ __forceinline__ uint64_t pack_candidate(uint32_t key, uint32_t index) {
    return ((uint64_t)key << 32) | (uint64_t)(~index);
}
__host__ __device__ __forceinline__ uint32_t candidate_index(uint64_t c) { return ~(uint32_t)c; }
__host__ __device__ __forceinline__ uint32_t candidate_key(uint64_t c) { return (uint32_t)(c >> 32); }

// ------------------------------------------------------------------------------------------------
// Warp-level bitonic primitives on u64 (descending = best first).
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
    return __shfl_xor_sync(0xffffffffu, v, mask);
}

// Full bitonic sort of one value per lane, descending across lanes 0..31.
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            uint64_t o = shfl_xor_u64(v, stride);
            bool lower = (lane & stride) == 0;            // I hold the lower-indexed slot of the pair
            bool desc = (lane & size) == 0;               // this sub-sequence sorts descending
            bool keep_max = (lower == desc);
            v = ((v > o) != keep_max) ? o : v;            // one compare: take the partner's value unless mine is the one to keep
        }
    }
    return v;
}

// Full bitonic sort of 32*RS values, element e at (reg e/32, lane e%32), descending in e.
template <int RS>
__device__ __forceinline__ void warp_sort_regs_desc(uint64_t (&S)[RS], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * RS; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {                            // register-to-register; direction depends on r only
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    if ((r & rs) == 0) {
                        const bool desc = ((32 * r) & size) == 0;
                        const uint64_t a = S[r], b = S[r + rs];
                        const uint64_t hi = a > b ? a : b, lo = a > b ? b : a;
                        S[r] = desc ? hi : lo;
                        S[r + rs] = desc ? lo : hi;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const uint64_t o = shfl_xor_u64(S[r], stride);
                    const bool lower = (lane & stride) == 0;
                    const bool desc = (((32 * r) | lane) & size) == 0;
                    const bool keep_max = (lower == desc);
                    S[r] = ((S[r] > o) != keep_max) ? o : S[r];
                }
            }
        }
    }
}

// L holds a BITONIC sequence of 32*R values, element e at (reg e/32, lane e%32).
// Sorts it descending in place.
template <int R>
__device__ __forceinline__ void warp_bitonic_merge_desc(uint64_t (&L)[R], int lane) {
#pragma unroll
    for (int rs = R >> 1; rs > 0; rs >>= 1) {              // strides >= 32: register-to-register
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if ((r & rs) == 0) {
                uint64_t a = L[r], b = L[r + rs];
                L[r] = a > b ? a : b;
                L[r + rs] = a > b ? b : a;
            }
        }
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {      // strides < 32: lane shuffles
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint64_t o = shfl_xor_u64(L[r], stride);
            bool lower = (lane & stride) == 0;
            L[r] = ((L[r] > o) != lower) ? o : L[r];
        }
    }
}

// L: sorted descending (32*R values). Mrev: another descending list M of the same length, supplied
// REVERSED: Mrev[r] at lane l holds M[32*R - 1 - (32*r + l)].  On return L = the 32*R best of the
// union, sorted descending.
template <int R>
__device__ __forceinline__ void warp_merge_topk_desc(uint64_t (&L)[R], const uint64_t (&Mrev)[R], int lane) {
#pragma unroll
    for (int r = 0; r < R; ++r) L[r] = L[r] > Mrev[r] ? L[r] : Mrev[r];   // bitonic, holds the best 32R
    warp_bitonic_merge_desc<R>(L, lane);
}

}  // namespace pmm
