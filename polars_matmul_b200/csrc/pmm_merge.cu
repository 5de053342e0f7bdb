// pmm_merge.cu — K-way merge of packed candidate lists (one warp per query).
//
// Closes the top-k of src/topk.rs:42-75 when the corpus was scanned in pieces: corpus ranges of one
// query tile handled by different CTAs of the fused kernel, or corpus shards on different GPUs
// (SURVEY §8e).  Every list is sorted best-first under the packed total order (pmm_common.cuh), so
// merging is "keep the KP largest u64 of the union": a bitonic merge in registers, KP/32 per lane.
// Emits the reference's output layout: u32 indices (`idx as u32`, src/matmul.rs:506) and f64 scores
// (exact widening of the f32 score, src/matmul.rs:447).
#include "pmm_common.cuh"
#include "pmm_kernels.h"

namespace pmm {

template <int R, typename ListPtr>
__device__ __forceinline__ void merge_query(ListPtr list_ptr, int n_lists, int k_in, int k_out, bool higher,
                                            int64_t q, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                                            int lane) {
    constexpr int KP = 32 * R;
    uint64_t L[R];
#pragma unroll
    for (int r = 0; r < R; ++r) L[r] = 0ull;
    for (int l = 0; l < n_lists; ++l) {
        const uint64_t *p = list_ptr(l);
        uint64_t M[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int e = KP - 1 - (32 * r + lane);  // reversed read
            M[r] = (e < k_in) ? __ldg(p + e) : 0ull;
        }
        warp_merge_topk_desc<R>(L, M, lane);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int t = 32 * r + lane;
        if (t < k_out) {
            uint64_t cnd = L[r];
            if (out_idx) out_idx[q * k_out + t] = candidate_index(cnd);
            if (out_score) out_score[q * k_out + t] = (double)key_score(candidate_key(cnd), higher);
            if (out_cand) out_cand[q * k_out + t] = cnd;
        }
    }
}

template <int R>
__global__ void __launch_bounds__(256) merge_regular_kernel(const uint64_t *__restrict__ lists, int n_lists,
                                                            int64_t list_stride, int64_t row_stride, int64_t nq,
                                                            int k_in, int k_out, bool higher, uint32_t *out_idx,
                                                            double *out_score, uint64_t *out_cand) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    auto ptr = [&](int l) { return lists + (int64_t)l * list_stride + q * row_stride; };
    merge_query<R>(ptr, n_lists, k_in, k_out, higher, q, out_idx, out_score, out_cand, lane);
}

template <int R>
__global__ void __launch_bounds__(256) merge_tiles_kernel(const uint64_t *__restrict__ lists, TcSchedule sched, int cg, int esets,
                                                          int64_t nq, int k_out, bool higher, uint32_t *out_idx,
                                                          double *out_score, uint64_t *out_cand) {
    constexpr int KP = 32 * R;
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int tile_rows = TC_TILE_M * cg;
    const int mt = (int)(q / tile_rows), r = (int)(q % tile_rows);  // r = cta_in_group * 128 + row
    // lists of one query: [piece][epilogue set], a regular stride apart
    const uint64_t *base = lists + (sched.slot_base(mt) * esets * tile_rows + r) * KP;
    const int64_t piece_stride = (int64_t)tile_rows * KP;
    auto ptr = [&](int l) { return base + l * piece_stride; };
    merge_query<R>(ptr, sched.pieces(mt) * esets, KP, k_out, higher, q, out_idx, out_score, out_cand, lane);
}

// Many pieces per query (a handful of re-queried rows shared by all scheduling units: 74 lists of 256 entries each): one
// BLOCK per query, warp w merges the pieces l = w, w + 8, ... into a partial list in shared memory, warp 0 merges the
// eight partial lists.  The chain of dependent list loads shrinks from `pieces` to pieces / 8 + 8.
template <int R>
__global__ void __launch_bounds__(256) merge_tiles_block_kernel(const uint64_t *__restrict__ lists, TcSchedule sched, int cg, int esets,
                                                                int64_t nq, int k_out, bool higher, uint32_t *out_idx,
                                                                double *out_score, uint64_t *out_cand) {
    constexpr int KP = 32 * R;
    __shared__ uint64_t part[8][KP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = blockIdx.x;
    if (q >= nq) return;
    const int tile_rows = TC_TILE_M * cg;
    const int mt = (int)(q / tile_rows), r = (int)(q % tile_rows);
    const uint64_t *base = lists + (sched.slot_base(mt) * esets * tile_rows + r) * KP;
    const int64_t piece_stride = (int64_t)tile_rows * KP;
    const int n_lists = sched.pieces(mt) * esets;
    uint64_t L[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) L[rr] = 0ull;
    for (int l = w; l < n_lists; l += 8) {
        const uint64_t *p = base + l * piece_stride;
        uint64_t M[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) M[rr] = __ldg(p + (KP - 1 - (32 * rr + lane)));   // reversed read
        warp_merge_topk_desc<R>(L, M, lane);
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) part[w][32 * rr + lane] = L[rr];
    __syncthreads();
    if (w != 0) return;
    auto ptr = [&](int l) { return (const uint64_t *)part[l]; };
    // (merge_query reads with __ldg semantics on generic pointers: shared memory is fine for plain loads)
    uint64_t F[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) F[rr] = 0ull;
    for (int l = 0; l < 8; ++l) {
        const uint64_t *p = ptr(l);
        uint64_t M[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) M[rr] = p[KP - 1 - (32 * rr + lane)];
        warp_merge_topk_desc<R>(F, M, lane);
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        const int t = 32 * rr + lane;
        if (t < k_out) {
            const uint64_t cnd = F[rr];
            if (out_idx) out_idx[q * k_out + t] = candidate_index(cnd);
            if (out_score) out_score[q * k_out + t] = (double)key_score(candidate_key(cnd), higher);
            if (out_cand) out_cand[q * k_out + t] = cnd;
        }
    }
}

cudaError_t launch_merge_regular(const uint64_t *lists, int64_t n_lists, int64_t list_stride, int64_t row_stride,
                                 int64_t nq, int k_in, int k_out, bool higher, uint32_t *out_idx,
                                 double *out_score, uint64_t *out_cand, cudaStream_t s) {
    if (nq <= 0 || k_out <= 0) return cudaSuccess;
    if (k_in > 256 || k_out > k_in) return cudaErrorInvalidValue;
    unsigned grid = (unsigned)((nq + 7) / 8);
    if (k_in > 128)
        merge_regular_kernel<8><<<grid, 256, 0, s>>>(lists, (int)n_lists, list_stride, row_stride, nq, k_in, k_out, higher, out_idx, out_score, out_cand);
    else if (k_in <= 32)
        merge_regular_kernel<1><<<grid, 256, 0, s>>>(lists, (int)n_lists, list_stride, row_stride, nq, k_in, k_out, higher, out_idx, out_score, out_cand);
    else if (k_in <= 64)
        merge_regular_kernel<2><<<grid, 256, 0, s>>>(lists, (int)n_lists, list_stride, row_stride, nq, k_in, k_out, higher, out_idx, out_score, out_cand);
    else
        merge_regular_kernel<4><<<grid, 256, 0, s>>>(lists, (int)n_lists, list_stride, row_stride, nq, k_in, k_out, higher, out_idx, out_score, out_cand);
    return cudaGetLastError();
}

cudaError_t launch_merge_tiles(const uint64_t *lists, TcSchedule sched, int cg, int esets, int kp, int64_t nq, int k_out, bool higher,
                               uint32_t *out_idx, double *out_score, uint64_t *out_cand, cudaStream_t s) {
    if (nq <= 0 || k_out <= 0) return cudaSuccess;
    // few queries, many pieces each (re-query launches): one block per query
    const int max_lists = (sched.g > sched.g_rem ? sched.g : sched.g_rem) * esets;
    if (max_lists >= 16 && nq <= 8192) {
        const unsigned bg = (unsigned)nq;
        if (kp == 32) merge_tiles_block_kernel<1><<<bg, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
        else if (kp == 64) merge_tiles_block_kernel<2><<<bg, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
        else if (kp == 128) merge_tiles_block_kernel<4><<<bg, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
        else if (kp == 256) merge_tiles_block_kernel<8><<<bg, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
        else return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    unsigned grid = (unsigned)((nq + 7) / 8);
    if (kp == 32)
        merge_tiles_kernel<1><<<grid, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
    else if (kp == 64)
        merge_tiles_kernel<2><<<grid, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
    else if (kp == 128)
        merge_tiles_kernel<4><<<grid, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
    else if (kp == 256)
        merge_tiles_kernel<8><<<grid, 256, 0, s>>>(lists, sched, cg, esets, nq, k_out, higher, out_idx, out_score, out_cand);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace pmm
