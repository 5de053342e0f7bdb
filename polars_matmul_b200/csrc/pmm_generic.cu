// pmm_generic.cu — precision-exact SIMT path: score slab + per-row exact top-k.
//
// Used for f64 working precision (tcgen05 has no f64 kind), for k > 128, and as the on-device
// cross-check of the tensor-core path.  It restates, on the GPU,
//   * C = A*B^T (src/metrics.rs:40-97, :204-255) with ONE FMA per element, sequential in the vector
//     dimension for every output — the same order as the CPU oracle, so scores are bit-identical to it;
//   * the cosine / euclidean pass (src/metrics.rs:267-308, :323-362), fused into the epilogue;
//   * select_topk_with_scores[_f32] (src/topk.rs:6-75) as radix select + ordered tie collection +
//     bitonic sort under the total order (better score, then lower index; NaN last).
#include "pmm_common.cuh"
#include "pmm_kernels.h"

namespace pmm {

__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

// 64x64 output tile, 16-deep k slices, 256 threads, 4x4 outputs per thread.
template <typename T>
__global__ void __launch_bounds__(256) scores_kernel(const T *__restrict__ q, const T *__restrict__ c,
                                                     const T *__restrict__ qa, const T *__restrict__ ca,
                                                     int64_t nq, int64_t n, int64_t d, int metric,
                                                     T *__restrict__ out, int64_t ldo) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row lr, k offset lk..lk+3

    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (T)0;

    for (int64_t k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t kk = k0 + lk + u;
            T av = (T)0, bv = (T)0;
            if (kk < d) {
                if (m0 + lr < nq) av = __ldg(q + (m0 + lr) * d + kk);
                if (n0 + lr < n) bv = __ldg(c + (n0 + lr) * d + kk);
            }
            As[lk + u][lr] = av;
            Bs[lk + u][lr] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma_rn(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t r = m0 + ty * 4 + i;
        if (r >= nq) continue;
        T qav = (metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN) ? qa[r] : (T)0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t cc = n0 + tx * 4 + j;
            if (cc >= n) continue;
            T v = acc[i][j];
            if (metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN) v = metric_finish(v, metric, qav, ca[cc]);
            out[r * ldo + cc] = v;
        }
    }
}

template <typename T>
static cudaError_t launch_scores_t(const T *q, const T *c, const T *qa, const T *ca, int64_t nq, int64_t n,
                                   int64_t d, int metric, T *out, int64_t ldo, cudaStream_t s) {
    if (nq <= 0 || n <= 0) return cudaSuccess;
    // gridDim.y <= 65535: callers chunk queries so nq/64 stays below that
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((nq + 63) / 64));
    scores_kernel<T><<<grid, 256, 0, s>>>(q, c, qa, ca, nq, n, d, metric, out, ldo);
    return cudaGetLastError();
}
// f32, larger shapes: 128 x 128 x 16 block tile, 8 x 8 outputs per thread (two 4-wide quadrants per side so that the
// shared-memory reads are conflict-free float4s), global -> register -> shared double buffering.  Still ONE FMA per
// element and output, k ascending: bit-identical to scores_kernel and to the oracle.  This is the exact path of the raw
// f32 matmul for vector lengths where the tensor-core split no longer holds 1e-5 (pmm_api.cu: matmul_tc_max_dim) and
// of the exact top-k fallback; about 3x the rate of the 64 x 64 kernel.
template <bool VEC>
__global__ void __launch_bounds__(256) scores_f32_tiled_kernel(const float *__restrict__ q, const float *__restrict__ c,
                                                               const float *__restrict__ qa, const float *__restrict__ ca,
                                                               int64_t nq, int64_t n, int64_t d, int metric,
                                                               float *__restrict__ out, int64_t ldo) {
    constexpr int BM = 128, BN = 128, BK = 16, P = BM + 4;
    __shared__ __align__(16) float As[2][BK][P];
    __shared__ __align__(16) float Bs[2][BK][P];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: rows lr and lr + 64, k offset lk..lk+3

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float ra[2][4], rb[2][4];
    auto gload = [&](int64_t k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t ar = m0 + lr + 64 * h, br = n0 + lr + 64 * h, kk = k0 + lk;
            if (VEC) {   // d % 4 == 0 and 16-byte aligned bases: whole float4s are in or out
                float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                if (kk < d) {
                    if (ar < nq) va = __ldg((const float4 *)(q + ar * d + kk));
                    if (br < n) vb = __ldg((const float4 *)(c + br * d + kk));
                }
                ra[h][0] = va.x; ra[h][1] = va.y; ra[h][2] = va.z; ra[h][3] = va.w;
                rb[h][0] = vb.x; rb[h][1] = vb.y; rb[h][2] = vb.z; rb[h][3] = vb.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ra[h][u] = (ar < nq && kk + u < d) ? __ldg(q + ar * d + kk + u) : 0.0f;
                    rb[h][u] = (br < n && kk + u < d) ? __ldg(c + br * d + kk + u) : 0.0f;
                }
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                As[buf][lk + u][lr + 64 * h] = ra[h][u];
                Bs[buf][lk + u][lr + 64 * h] = rb[h][u];
            }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = 0; k0 < d; k0 += BK) {
        const bool more = k0 + BK < d;
        if (more) gload(k0 + BK);   // next slice's global loads fly while this one is multiplied
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *(const float4 *)&As[buf][kk][ty * 4], a1 = *(const float4 *)&As[buf][kk][64 + ty * 4];
            const float4 b0 = *(const float4 *)&Bs[buf][kk][tx * 4], b1 = *(const float4 *)&Bs[buf][kk][64 + tx * 4];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    const bool aux = metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= nq) continue;
        const float qav = aux ? qa[r] : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t cc = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (cc >= n) continue;
            float v = acc[i][j];
            if (aux) v = metric_finish(v, metric, qav, ca[cc]);
            out[r * ldo + cc] = v;
        }
    }
}

cudaError_t launch_scores_f32(const float *q, const float *c, const float *qa, const float *ca, int64_t nq,
                              int64_t n, int64_t d, int metric, float *out, int64_t ldo, cudaStream_t s) {
    if (nq <= 0 || n <= 0) return cudaSuccess;
    if (nq < 96 || n < 96) return launch_scores_t<float>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);   // small: finer tiles
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)((nq + 127) / 128));
    const bool vec = (d % 4) == 0 && (((uintptr_t)q | (uintptr_t)c) & 15) == 0;
    if (vec) scores_f32_tiled_kernel<true><<<grid, 256, 0, s>>>(q, c, qa, ca, nq, n, d, metric, out, ldo);
    else scores_f32_tiled_kernel<false><<<grid, 256, 0, s>>>(q, c, qa, ca, nq, n, d, metric, out, ldo);
    return cudaGetLastError();
}
cudaError_t launch_scores_f64(const double *q, const double *c, const double *qa, const double *ca, int64_t nq,
                              int64_t n, int64_t d, int metric, double *out, int64_t ldo, cudaStream_t s) {
    return launch_scores_t<double>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);
}

// ------------------------------------------------------------------------------------------------
// f64 contraction on the FP64 tensor path: mma.sync.m8n8k4.f64 (DMMA; tcgen05 has no f64 kind).
// Replaces matmul_f64 / matmul_slice_f64 (src/metrics.rs:40-157) with the same fused metric epilogue as
// scores_kernel. 128 x 64 x 16 block tile, 8 warps of 32 x 32, register-staged double buffering.
// Scores agree with the sequential-FMA order to ~1e-15 relative (tolerance 1e-12), not bit for bit.
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) scores_f64_dmma_kernel(const double *__restrict__ q, const double *__restrict__ c,
                                                              const double *__restrict__ qa, const double *__restrict__ ca,
                                                              int64_t nq, int64_t n, int64_t d, int metric,
                                                              double *__restrict__ out, int64_t ldo) {
    constexpr int BM = 128, BN = 64, BK = 16, LD = BK + 4;  // row stride 20 doubles: conflict-free fragment loads
    __shared__ double As[BM * LD];
    __shared__ double Bs[BN * LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;              // 4 x 2 warps
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int ar = tid >> 1, ak = (tid & 1) * 8;           // A loader: row ar, k offset ak..ak+7
    const int br = tid >> 2, bk = (tid & 3) * 4;           // B loader: row br, k offset bk..bk+3

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double ra[8], rb[4];
    auto gload = [&](int64_t k0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t kk = k0 + ak + u;
            ra[u] = (m0 + ar < nq && kk < d) ? __ldg(q + (m0 + ar) * d + kk) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t kk = k0 + bk + u;
            rb[u] = (n0 + br < n && kk < d) ? __ldg(c + (n0 + br) * d + kk) : 0.0;
        }
    };
    gload(0);
    for (int64_t k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
        for (int u = 0; u < 8; ++u) As[ar * LD + ak + u] = ra[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) Bs[br * LD + bk + u] = rb[u];
        __syncthreads();
        if (k0 + BK < d) gload(k0 + BK);  // next tile's global loads fly while this tile is multiplied
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(wm * 32 + i * 8 + (lane >> 2)) * LD + ks + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wn * 32 + j * 8 + (lane >> 2)) * LD + ks + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
    const bool aux = metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (r >= nq) continue;
        const double qav = aux ? qa[r] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t cc = n0 + wn * 32 + j * 8 + (lane & 3) * 2 + e;
                if (cc >= n) continue;
                double v = acc[i][j][e];
                if (aux) v = metric_finish(v, metric, qav, ca[cc]);
                out[r * ldo + cc] = v;
            }
        }
    }
}

// Same tile and fragment mapping, operands through a three-stage cp.async ring (8-byte copies: any row pitch; rows and
// K positions beyond the matrix are zero-filled by the copy itself): one block barrier per K-tile instead of two, two
// K-tiles of loads in flight, no staging registers (two blocks per SM stay resident).
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// MI = 8-row fragments per warp in M: 4 (warp tile 32 x 32, block tile 128 x 64, two blocks per SM) or 8 (64 x 32, block
// tile 256 x 64, one block per SM: 12 instead of 8 fragment loads per 32 resp. 16 DMMAs, 32 independent accumulators).
// MI = 2 (warp tile 16 x 32, block tile 64 x 64, <= 64 registers): three or four blocks per SM - DMMA has a long latency
// and a warp keeps few of them in flight, so resident WARPS are what fills the FP64 pipe (measured: 8 / 16 warps per SM
// -> 7.7 / 5.0 ms at 8192 x 8192 x 1024).  ST = ring stages (2: the copy of tile kt+1 overlaps tile kt only).
template <int MI, int ST, int MINB>
__global__ void __launch_bounds__(256, MINB) scores_f64_dmma_async_kernel(const double *__restrict__ q, const double *__restrict__ c,
                                                                                 const double *__restrict__ qa, const double *__restrict__ ca,
                                                                                 int64_t nq, int64_t n, int64_t d, int metric,
                                                                                 double *__restrict__ out, int64_t ldo) {
    constexpr int BM = 32 * MI, BN = 64, BK = 16, LD = BK + 4, APT = BM * BK / 256;   // APT: A elements per thread and K-tile
    extern __shared__ __align__(16) double dmma_smem[];
    double *As = dmma_smem;                       // [ST][BM * LD]
    double *Bs = dmma_smem + ST * BM * LD;        // [ST][BN * LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int ar = tid / (BK / APT), ak = (tid % (BK / APT)) * APT;
    const int br = tid >> 2, bk = (tid & 3) * 4;
    const bool a_row = m0 + ar < nq, b_row = n0 + br < n;
    const double *a_src = q + (a_row ? (m0 + ar) : 0) * d;
    const double *b_src = c + (b_row ? (n0 + br) : 0) * d;
    const int nkt = (int)((d + BK - 1) / BK);
    auto issue = [&](int kt) {
        if (kt < nkt) {
            const int slot = kt % ST;
            const int64_t k0 = (int64_t)kt * BK;
            double *ad = As + slot * BM * LD + ar * LD + ak;
            double *bd = Bs + slot * BN * LD + br * LD + bk;
#pragma unroll
            for (int u = 0; u < APT; ++u) {
                const bool ok = a_row && k0 + ak + u < d;
                cp_async8(ad + u, a_src + (ok ? k0 + ak + u : 0), ok);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = b_row && k0 + bk + u < d;
                cp_async8(bd + u, b_src + (ok ? k0 + bk + u : 0), ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // (an empty group keeps the wait counts uniform)
    };
    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    issue(0);
    if (ST > 2) issue(1);
    for (int kt = 0; kt < nkt; ++kt) {
        if (ST > 2) asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's copies of tile kt have landed ...
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                        // ... everyone's have, and tile kt-1's slot is free
        issue(kt + ST - 1);
        const double *at = As + (kt % ST) * BM * LD, *bt = Bs + (kt % ST) * BN * LD;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double a[MI], b[4];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = at[(wm * 8 * MI + i * 8 + (lane >> 2)) * LD + ks + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bt[(wn * 32 + j * 8 + (lane >> 2)) * LD + ks + (lane & 3)];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const bool aux = metric == METRIC_COSINE || metric == METRIC_EUCLIDEAN;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t r = m0 + wm * 8 * MI + i * 8 + (lane >> 2);
        if (r >= nq) continue;
        const double qav = aux ? qa[r] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t cc = n0 + wn * 32 + j * 8 + (lane & 3) * 2 + e;
                if (cc >= n) continue;
                double v = acc[i][j][e];
                if (aux) v = metric_finish(v, metric, qav, ca[cc]);
                out[r * ldo + cc] = v;
            }
        }
    }
}

static int g_dmma_async = 3;   // dmma_set_async(): 0 register-staged kernel; cp.async ring with 1: 32 x 32 warp tiles, 2 blocks per SM,
                                // 2: 64 x 32, 1 block, 3 (default): 16 x 32, 4 blocks, two stages, 4: 16 x 32, 3 blocks, three stages
void dmma_set_async(int mode) { g_dmma_async = mode; }

template <int MI, int ST, int MINB>
static cudaError_t launch_dmma_async(const double *q, const double *c, const double *qa, const double *ca, int64_t nq, int64_t n,
                                     int64_t d, int metric, double *out, int64_t ldo, cudaStream_t s) {
    constexpr int smem = ST * (32 * MI + 64) * 20 * 8;
    static bool attr_set[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(scores_f64_dmma_async_kernel<MI, ST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set[dev & 15] = true;
    }
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((nq + 32 * MI - 1) / (32 * MI)));
    scores_f64_dmma_async_kernel<MI, ST, MINB><<<grid, 256, smem, s>>>(q, c, qa, ca, nq, n, d, metric, out, ldo);
    return cudaGetLastError();
}

cudaError_t launch_scores_f64_dmma(const double *q, const double *c, const double *qa, const double *ca, int64_t nq,
                                   int64_t n, int64_t d, int metric, double *out, int64_t ldo, cudaStream_t s) {
    if (nq <= 0 || n <= 0) return cudaSuccess;
    if (g_dmma_async == 4) return launch_dmma_async<2, 3, 3>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);
    if (g_dmma_async == 3) return launch_dmma_async<2, 2, 4>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);
    if (g_dmma_async == 2) return launch_dmma_async<8, 3, 1>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);
    if (g_dmma_async == 1) return launch_dmma_async<4, 3, 2>(q, c, qa, ca, nq, n, d, metric, out, ldo, s);
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((nq + 127) / 128));
    scores_f64_dmma_kernel<<<grid, 256, 0, s>>>(q, c, qa, ca, nq, n, d, metric, out, ldo);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Exact per-row top-k. One 256-thread block per query row.
template <typename T> struct KeyOf;
template <> struct KeyOf<float> { typedef uint32_t type; static constexpr int bits = 32; };
template <> struct KeyOf<double> { typedef uint64_t type; static constexpr int bits = 64; };

int select_kpad(int64_t k) {
    int p = 32;
    while (p < k) p <<= 1;
    return p;
}
int select_smem_kpad_limit(bool f64) { return f64 ? 2048 : 4096; }  // 12 B resp. 8+4 B per entry

template <typename K>
__device__ __forceinline__ bool entry_before(K ka, uint32_t ia, K kb, uint32_t ib) {
    return ka > kb || (ka == kb && ia < ib);
}

template <typename T>
__global__ void __launch_bounds__(256) select_topk_kernel(const T *__restrict__ scores, int64_t ld, int64_t n,
                                                          int k, int kpad, bool higher, int64_t index_base,
                                                          uint32_t *__restrict__ out_idx,
                                                          double *__restrict__ out_score,
                                                          uint64_t *__restrict__ out_cand, void *scratch,
                                                          int use_global) {
    typedef typename KeyOf<T>::type K;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int warp_tot[8];
    __shared__ unsigned int sh_gt, sh_eq_base;
    __shared__ K sh_prefix;
    __shared__ unsigned int sh_remaining;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t row = blockIdx.x;
    const T *s = scores + row * ld;

    K *skeys;
    uint32_t *sidx;
    if (use_global) {
        skeys = (K *)((unsigned char *)scratch + (size_t)row * kpad * (sizeof(K) + 4));
        sidx = (uint32_t *)(skeys + kpad);
    } else {
        skeys = (K *)dyn_smem;
        sidx = (uint32_t *)(skeys + kpad);
    }

    // ---- 1. radix select: K-th largest key, 8 bits per pass from the top
    K prefix = 0, mask = 0;
    unsigned int remaining = (unsigned int)k;
    for (int shift = KeyOf<T>::bits - 8; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int64_t j = tid; j < n; j += 256) {
            K key = score_key(s[j], higher);
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)((key >> shift) & 0xff)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int cum = 0;
            int digit = 0;
            for (int dgt = 255; dgt >= 0; --dgt) {
                unsigned int h = hist[dgt];
                if (cum + h >= remaining) { digit = dgt; break; }
                cum += h;
            }
            sh_remaining = remaining - cum;
            sh_prefix = prefix | ((K)digit << shift);
        }
        __syncthreads();
        remaining = sh_remaining;
        prefix = sh_prefix;
        mask |= (K)0xff << shift;
        __syncthreads();
    }
    const K kth = prefix;                 // key of the k-th best entry
    const unsigned int need_eq = remaining;  // how many entries equal to kth belong to the result (>= 1)
    const unsigned int n_gt = (unsigned int)k - need_eq;

    // ---- 2. collect: every key > kth, plus the first need_eq keys == kth in index order
    if (tid == 0) { sh_gt = 0; sh_eq_base = 0; }
    for (int t = tid; t < kpad; t += 256) { skeys[t] = 0; sidx[t] = 0xffffffffu; }
    __syncthreads();
    for (int64_t j0 = 0; j0 < n; j0 += 256) {
        int64_t j = j0 + tid;
        K key = 0;
        bool valid = j < n;
        if (valid) key = score_key(s[j], higher);
        bool is_gt = valid && key > kth;
        bool is_eq = valid && key == kth;
        if (is_gt) {
            unsigned int pos = atomicAdd(&sh_gt, 1u);
            skeys[pos] = key;
            sidx[pos] = (uint32_t)j;
        }
        unsigned int bal = __ballot_sync(0xffffffffu, is_eq);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        unsigned int before = sh_eq_base, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            unsigned int c = warp_tot[w];
            if (w < wid) before += c;
            total += c;
        }
        if (is_eq) {
            unsigned int pos = before + __popc(bal & ((1u << lane) - 1u));
            if (pos < need_eq) { skeys[n_gt + pos] = key; sidx[n_gt + pos] = (uint32_t)j; }
        }
        __syncthreads();
        if (tid == 0) sh_eq_base += total;
        __syncthreads();
    }

    // ---- 3. bitonic sort of kpad entries, best first
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < kpad; t += 256) {
                int p = t ^ stride;
                if (p > t) {
                    bool desc = (t & size) == 0;
                    K ka = skeys[t], kb = skeys[p];
                    uint32_t ia = sidx[t], ib = sidx[p];
                    bool a_first = entry_before<K>(ka, ia, kb, ib);
                    if (a_first != desc) { skeys[t] = kb; skeys[p] = ka; sidx[t] = ib; sidx[p] = ia; }
                }
            }
            __syncthreads();
        }
    }

    // ---- 4. emit
    for (int t = tid; t < k; t += 256) {
        K key = skeys[t];
        uint32_t idx = sidx[t] + (uint32_t)index_base;
        if (out_idx) out_idx[row * k + t] = idx;
        if (out_score) out_score[row * k + t] = (double)key_score(key, higher);
        if (sizeof(K) == 4 && out_cand) out_cand[row * k + t] = pack_candidate((uint32_t)key, idx);
    }
}

template <typename T>
static cudaError_t launch_select_t(const T *scores, int64_t ld, int64_t nq, int64_t n, int64_t k, bool higher,
                                   int64_t index_base, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                                   void *scratch, cudaStream_t s) {
    typedef typename KeyOf<T>::type K;
    if (nq <= 0 || k <= 0) return cudaSuccess;
    int kpad = select_kpad(k);
    int use_global = kpad > select_smem_kpad_limit(sizeof(T) == 8);
    size_t smem = use_global ? 0 : (size_t)kpad * (sizeof(K) + 4);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(select_topk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    select_topk_kernel<T><<<(unsigned)nq, 256, smem, s>>>(scores, ld, n, (int)k, kpad, higher, index_base, out_idx,
                                                          out_score, out_cand, scratch, use_global);
    return cudaGetLastError();
}
cudaError_t launch_select_f32(const float *scores, int64_t ld, int64_t nq, int64_t n, int64_t k, bool higher,
                              int64_t index_base, uint32_t *out_idx, double *out_score, uint64_t *out_cand,
                              void *scratch, cudaStream_t s) {
    return launch_select_t<float>(scores, ld, nq, n, k, higher, index_base, out_idx, out_score, out_cand, scratch, s);
}
cudaError_t launch_select_f64(const double *scores, int64_t ld, int64_t nq, int64_t n, int64_t k, bool higher,
                              int64_t index_base, uint32_t *out_idx, double *out_score, void *scratch,
                              cudaStream_t s) {
    return launch_select_t<double>(scores, ld, nq, n, k, higher, index_base, out_idx, out_score, nullptr, scratch, s);
}

}  // namespace pmm
