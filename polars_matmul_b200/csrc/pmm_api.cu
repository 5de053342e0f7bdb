// pmm_api.cu — the C ABI (include/pmm.h): validation in the reference's order, host<->device staging,
// path selection and kernel orchestration.  Host-side counterpart of src/matmul.rs:288-519
// (matmul_impl / compute_topk_indices_scores / topk_impl) and src/lib.rs:15-55.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pmm.h"
#include "pmm_kernels.h"
#include "pmm_nccl.h"
#include "pmm_stage.h"

using namespace pmm;

namespace {

// ------------------------------------------------------------------------------------------------ errors
thread_local std::string g_err = "";

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) return fail(PMM_ERR_CUDA, "CUDA error: %s (%s)", cudaGetErrorString(e__), #expr); \
    } while (0)

// ------------------------------------------------------------------------------------------------ options / stats
std::atomic<int64_t> g_launches{0};

// Tuning / diagnostic options (include/pmm.h).  pmm_set_option changes the process-wide defaults; every compute
// entry point takes ONE consistent snapshot of them when it starts (plus the calling thread's own overrides,
// pmm_set_thread_option), so a call never sees an option change half way through and concurrent callers
// (src/lib.rs:25,45 releases the GIL: the reference is re-entrant) can run with different settings.
struct Options {
    int force_generic = 0, profile = 0, tc_group = 0, tc_cg = 2, tc_sync_tiles = 32, host_chunked = 1, f64_simt = 0, verify = 1,
        tc_levels = 3, tc_clm = 1, tc_cluster4 = 0, tc_max_units = 0, tc_debug_skip = 0, tc_sync_slack = 0, tc_max_flush = 0,
        host_chunk_ratio_pct = 0, host_chunk_first_div = 0, f16r_wide = 1, host_chunk_min_rows = 16384, host_chunk_min_mb = 64,
        tc_soft_at = 0, f64_tc = 1, multi_gpu = 1, seed_retry = 1, matmul_tc_max_dim = 0, matmul_split16 = 1, matmul_exact_max_dim = 8, matmul_flat = -1, d2h_direct = 1, warm_seed = 1, warm_rows = 4096, warm_rank = 0, pipeline = 0, rescore_stream_loads = 1, multipass = 1;
    int64_t generic_ws_mb = 1024, multi_gpu_min_gflop = 4000, pipeline_min_gflop = 2000;
};
Options g_opt;                 // process-wide defaults, guarded by g_opt_mu
std::mutex g_opt_mu;
thread_local Options t_opt;    // the snapshot the current call runs with
thread_local std::vector<std::pair<std::string, int64_t>> t_overrides;

// Returns false for an unknown key.
bool apply_option(Options &o, const std::string &k, int64_t value) {
    if (k == "force_generic") o.force_generic = (int)value;
    else if (k == "profile") o.profile = (int)value;
    else if (k == "tc_group") o.tc_group = value < 0 ? 0 : (int)value;  // 0 = automatic
    else if (k == "tc_cg") o.tc_cg = value == 2 ? 2 : 1;
    else if (k == "tc_max_units") o.tc_max_units = (int)value;
    else if (k == "tc_cluster4") o.tc_cluster4 = value ? 1 : 0;
    else if (k == "tc_sync_slack") o.tc_sync_slack = value < 0 ? 0 : (int)value;
    else if (k == "host_chunk_min_rows") o.host_chunk_min_rows = (int)value;   // smallest chunk (rows, multiple of 256; default 16384)
    else if (k == "host_chunk_min_mb") o.host_chunk_min_mb = (int)value;       // corpora below this many MB are uploaded in one piece (default 64)
    else if (k == "host_chunk_ratio_pct") o.host_chunk_ratio_pct = (int)value; // 0 = auto
    else if (k == "host_chunk_first_div") o.host_chunk_first_div = (int)value; // first chunk = N / this (0 = 32)
    else if (k == "f16r_wide") o.f16r_wide = value ? 1 : 0;                     // retry of the f16-rounded level before 3xTF32
    else if (k == "seed_retry") o.seed_retry = value ? 1 : 0;                   // that retry starts from thresholds seeded by the exact k-th scores
    else if (k == "tc_soft_at") o.tc_soft_at = value < 0 ? 0 : value > 88 ? 88 : (int)value;  // staged candidates that trigger an end-of-tile merge (0 = 48)
    else if (k == "tc_max_flush") o.tc_max_flush = value < 0 ? 0 : (int)value;
    else if (k == "tc_debug_skip") o.tc_debug_skip = (int)value;               // 1..3 need a -DPMM_DIAG build (checked by pmm_set_option)
    else if (k == "tc_clm") o.tc_clm = value == 2 ? 2 : 1;                      // 2: clusters of two CTA pairs, corpus tile multicast
    else if (k == "tc_levels") o.tc_levels = value >= 3 ? 3 : value == 2 ? 2 : 1;  // 3: f16-rounded first level, 2: TF32 x1, 1: 3xTF32 only
    else if (k == "verify") o.verify = value ? 1 : 0;                           // 0: skip the filter-losslessness check (and its fallback)
    else if (k == "f64_simt") o.f64_simt = value ? 1 : 0;                       // 1: bit-exact sequential-FMA f64 contraction instead of DMMA
    else if (k == "f64_tc") o.f64_tc = value ? 1 : 0;                           // f64 top-k: tensor-core filter + exact f64 re-scoring (default) or the slab path
    else if (k == "host_chunked") o.host_chunked = value ? 1 : 0;
    else if (k == "tc_sync_tiles") o.tc_sync_tiles = value < 0 ? 0 : (int)value;  // 0 = no pacing barriers
    else if (k == "generic_workspace_mb") o.generic_ws_mb = value < 1 ? 1 : value;
    else if (k == "multipass") o.multipass = value ? 1 : 0;                     // 248 < k <= 2000 on the fused path (several filter passes)
    else if (k == "rescore_stream_loads") o.rescore_stream_loads = value ? 1 : 0;
    else if (k == "pipeline") o.pipeline = value < 0 ? 0 : value > 2 ? 2 : (int)value;   // 2: per-round launches WITHOUT the overlap (measurement)
    else if (k == "__unused_pipeline") o.pipeline = value ? 1 : 0;                       // per-round filter launches with overlapped merge + re-scoring
    else if (k == "pipeline_min_gflop") o.pipeline_min_gflop = value < 0 ? 0 : value;   // smallest round worth a launch of its own
    else if (k == "matmul_tc_max_dim") o.matmul_tc_max_dim = value < 0 ? 0 : (int)value;   // 0 = automatic (see dev_matmul_impl)
    else if (k == "matmul_exact_max_dim") o.matmul_exact_max_dim = value < 0 ? 0 : (int)value;   // f32 vectors this short: exact SIMT kernel
    else if (k == "warm_seed") o.warm_seed = value ? 1 : 0;     // first filter level starts from thresholds of a sample pre-pass
    else if (k == "warm_rows") o.warm_rows = value < 256 ? 256 : value > 65536 ? 65536 : (int)(value / 256 * 256);   // sample size (rows)
    else if (k == "warm_rank") o.warm_rank = value < 0 ? 0 : value > 32 ? 32 : (int)value;   // sample rank that becomes the seed (0 = auto)
    else if (k == "d2h_direct") o.d2h_direct = value ? 1 : 0;   // top-k results written straight into page-locked result buffers
    else if (k == "matmul_flat") o.matmul_flat = value < 0 ? -1 : value > 2 ? 2 : (int)value;   // tile schedule of the tensor-core matmul: -1 automatic, 0 classic, 1 flat, 2 hybrid
    else if (k == "matmul_split16") o.matmul_split16 = value ? 1 : 0;           // raw f32 matmul: hi/lo f16 planes (1) or the 3xTF32 split (0)
    else if (k == "multi_gpu") o.multi_gpu = value ? 1 : 0;                     // host entry points may spread one call over all GPUs
    else if (k == "multi_gpu_min_gflop") o.multi_gpu_min_gflop = value < 0 ? 0 : value;
    else return false;
    return true;
}

// Every compute entry point starts with this: one consistent option snapshot for the whole call.
void begin_call() {
    {
        std::lock_guard<std::mutex> lk(g_opt_mu);
        t_opt = g_opt;
    }
    for (const auto &kv : t_overrides) apply_option(t_opt, kv.first, kv.second);
}

std::mutex g_stat_mu;
std::map<std::string, double> g_stats;
struct PendingEvent {
    std::string name;
    cudaEvent_t a, b;
};
std::vector<PendingEvent> g_pending;

void stat_add(const std::string &name, double v) {
    std::lock_guard<std::mutex> lk(g_stat_mu);
    g_stats[name] += v;
}

void resolve_pending_locked() {
    for (auto &p : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_stats[p.name + "_ms"] += ms;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_pending.clear();
}

// Counts the launch and, when profiling, brackets it with CUDA events on the launching stream.
template <typename F>
cudaError_t launch_counted(const char *name, cudaStream_t s, F &&f) {
    g_launches.fetch_add(1);
    if (!t_opt.profile) return f();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, s);
    cudaError_t e = f();
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lk(g_stat_mu);
    g_pending.push_back({name, a, b});
    g_stats[std::string(name) + "_launches"] += 1;
    return e;
}

// ------------------------------------------------------------------------------------------------ device info
struct DevInfo {
    int num_sms = 0;
    bool tc = false;
    bool init = false;
};
DevInfo &dev_info() {
    static thread_local DevInfo info[16];
    int dev = 0;
    cudaGetDevice(&dev);
    DevInfo &d = info[dev & 15];
    if (!d.init) {
        cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
        d.tc = tc_supported();
        d.init = true;
    }
    return d;
}

// The library's own stream-ordered memory pool per device (freed blocks stay in it instead of going back to the OS).
// A private pool: the device's default pool, which other libraries in the process (torch, cuDF) may use, is left
// untouched.  NULL when the pool cannot be created: allocations then come from the default pool.
cudaMemPool_t device_pool() {
    static std::mutex mu;
    static cudaMemPool_t pools[16] = {nullptr};
    static bool tried[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    const int i = dev & 15;
    if (!tried[i]) {
        tried[i] = true;
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[i], &props) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pools[i], cudaMemPoolAttrReleaseThreshold, &thr);
        } else {
            cudaGetLastError();
            pools[i] = nullptr;
        }
    }
    return pools[i];
}

cudaError_t pool_alloc(void **p, size_t bytes, cudaStream_t s) {
    cudaMemPool_t pool = device_pool();
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, s) : cudaMallocAsync(p, bytes, s);
}

// One host call at a time per device: every call saturates the GPU anyway, and the fused kernel is a persistent grid
// whose pacing barriers assume its CTAs are co-resident.  Concurrent callers (the reference releases the GIL and is
// re-entrant, tests/test_polars_matmul.py:551-572) are therefore serialised here, per device, not rejected.
std::mutex &device_mutex() {
    static std::mutex mus[16];
    int dev = 0;
    cudaGetDevice(&dev);
    return mus[dev & 15];
}

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(PMM_ERR_CUDA, "no CUDA device available (libpmm_b200 has no CPU fallback): %s",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return PMM_OK;
}

// Large blocks (>= 32 MB) that a thread frees are parked here and handed to the next request of a similar size on
// the same stream and device, instead of going back to the CUDA memory pool.  Why: the host path allocates
// multi-GB buffers (raw corpus, operand planes) in a different interleaving every call; the pool then has the
// bytes but not as one range and re-maps physical memory to serve the request, which showed up as 0.3-1.5 s stalls
// in 1 of 5 end-to-end calls.  Reuse on the SAME stream keeps stream order (like cudaFreeAsync + cudaMallocAsync).
// pmm_set_option("release_workspace", 1) returns the parked blocks to the pool.
std::atomic<int64_t> g_block_cache_cap_mb{24576};  // option "workspace_cache_mb": parked bytes per thread
std::atomic<int64_t> g_stage_slot_mb{32};
std::atomic<int> g_stage_slots{4};
struct BlockCache {
    struct Entry {
        void *p;
        size_t cap;
        cudaStream_t s;
        int dev;
    };
    std::vector<Entry> free_blocks;
    size_t bytes = 0;
    static constexpr size_t kMinBlock = (size_t)32 << 20;
    void *take(size_t want, cudaStream_t s, int dev, size_t *cap) {
        int best = -1;
        for (int i = 0; i < (int)free_blocks.size(); ++i) {
            const Entry &e = free_blocks[i];
            if (e.s == s && e.dev == dev && e.cap >= want && e.cap <= want + want / 2 &&
                (best < 0 || e.cap < free_blocks[best].cap))
                best = i;
        }
        if (best < 0) return nullptr;
        void *p = free_blocks[best].p;
        *cap = free_blocks[best].cap;
        bytes -= free_blocks[best].cap;
        free_blocks.erase(free_blocks.begin() + best);
        return p;
    }
    bool park(void *p, size_t cap, cudaStream_t s, int dev) {
        if (cap < kMinBlock || bytes + cap > ((size_t)g_block_cache_cap_mb.load() << 20)) return false;
        free_blocks.push_back(Entry{p, cap, s, dev});
        bytes += cap;
        return true;
    }
    void clear() {
        for (const Entry &e : free_blocks) cudaFreeAsync(e.p, e.s);
        free_blocks.clear();
        bytes = 0;
    }
    ~BlockCache() { clear(); }
};
thread_local BlockCache g_block_cache;

// Stream-ordered device buffer.
// alloc() on a buffer that is already large enough (same stream) keeps it: loops over corpus chunks reuse their
// scratch instead of cycling differently sized blocks.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaStream_t s = nullptr;
    int dev = 0;
    bool owned = true;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        if (bytes == 0) bytes = 16;
        if (p && cap >= bytes && s == stream) return cudaSuccess;
        release();
        s = stream;
        cudaGetDevice(&dev);
        if (bytes >= BlockCache::kMinBlock) {
            bytes = (bytes + BlockCache::kMinBlock - 1) / BlockCache::kMinBlock * BlockCache::kMinBlock;
            if ((p = g_block_cache.take(bytes, stream, dev, &cap))) return cudaSuccess;
        }
        cudaError_t e = pool_alloc(&p, bytes, stream);
        if (e != cudaSuccess && g_block_cache.bytes) {  // out of memory with blocks parked: give them back and retry
            cudaGetLastError();
            g_block_cache.clear();
            e = pool_alloc(&p, bytes, stream);
        }
        cap = e == cudaSuccess ? bytes : 0;
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
    // A view of `bytes` bytes inside another buffer (never freed here); alloc() of up to that size keeps the view.
    void borrow(void *ptr, size_t bytes, cudaStream_t stream) {
        release();
        p = ptr;
        cap = bytes;
        s = stream;
        owned = false;
    }
    void release() {
        if (p && owned && !g_block_cache.park(p, cap, s, dev)) cudaFreeAsync(p, s);
        p = nullptr;
        cap = 0;
        owned = true;
    }
    ~DevBuf() { release(); }
    template <typename T> T *as() const { return (T *)p; }
};

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
int esize(int dtype) { return dtype == PMM_DTYPE_F16 ? 2 : dtype == PMM_DTYPE_F32 ? 4 : 8; }

// ------------------------------------------------------------------------------------------------ prepared operands
struct Prepared {
    int mode = PREP_DENSE;   // PREP_*
    bool f64 = false;        // working precision of norm / sqnorm (and of the DENSE copy)
    int64_t n_rows = 0, dim = 0, rows_pad = 0, ld = 0;
    DevBuf p0, p1, norm, sqnorm, norm32, sqnorm32, max_sq;
    DevBuf scale;            // PREP_SPLIT16: [rows_pad] factors that undo the rows' power-of-two scaling
    unsigned int *max_sq_ptr = nullptr;  // own (max_sq) or shared across corpus chunks
    // f32 views of the norms for the tensor-core filter and its proof: the working-type buffers themselves for f32
    // working precision, rounded copies written by the same prep pass for f64
    const float *norm_f32() const { return f64 ? norm32.as<float>() : norm.as<float>(); }
    const float *sq_f32() const { return f64 ? sqnorm32.as<float>() : sqnorm.as<float>(); }
};

// Norm range of a column, filled by the prep kernel: [0] largest squared norm (atomicMax on the float bits),
// [1] smallest squared norm above 1e-12 (atomicMin; +inf while empty).
cudaError_t init_norm_range(unsigned int *d, cudaStream_t s) {
    static const unsigned int init[2] = {0u, 0x7f800000u};
    return cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, s);
}

// Bytes of one operand plane prepare() allocates for `rows` x `dim` (tensor-core modes).
size_t plane_bytes(int mode, int64_t rows, int64_t dim, int64_t row_tile) {
    const bool half_plane = mode == PREP_F16 || mode == PREP_F16R || mode == PREP_SPLIT16;
    const int64_t kq = half_plane ? 64 : 32, es = half_plane ? 2 : 4;
    return (size_t)(round_up(rows, row_tile) * round_up(dim, kq) * es);
}

// Runs the prep kernel on a device-resident matrix. row_tile: pad rows to this multiple (planes only).
int prepare(const pmm_matrix_t &m, int mode, bool f64, int64_t row_tile, bool want_norm, bool want_sq, int *d_err,
            cudaStream_t s, Prepared *out, bool want_max = false, unsigned int *shared_max = nullptr,
            unsigned char *nonfinite_rows = nullptr, unsigned int *nonfinite_count = nullptr) {
    out->mode = mode;
    out->f64 = f64;
    out->n_rows = m.n_rows;
    out->dim = m.dim;
    const int64_t wsz = f64 ? 8 : 4;
    if (mode == PREP_DENSE) {
        out->rows_pad = m.n_rows;
        out->ld = m.dim;
        CUDA_TRY(out->p0.alloc((size_t)(m.n_rows * m.dim * wsz), s));
    } else {
        const bool half_plane = mode == PREP_F16 || mode == PREP_F16R || mode == PREP_SPLIT16;
        const int64_t kq = half_plane ? 64 : 32;
        const int64_t es = half_plane ? 2 : 4;
        out->rows_pad = round_up(m.n_rows, row_tile);
        out->ld = round_up(m.dim, kq);
        CUDA_TRY(out->p0.alloc((size_t)(out->rows_pad * out->ld * es), s));
        if (mode == PREP_TF32 || mode == PREP_SPLIT16) CUDA_TRY(out->p1.alloc((size_t)(out->rows_pad * out->ld * es), s));
        if (mode == PREP_SPLIT16) CUDA_TRY(out->scale.alloc((size_t)(out->rows_pad * 4), s));
    }
    if (want_norm) CUDA_TRY(out->norm.alloc((size_t)(out->rows_pad * wsz), s));
    if (want_sq) CUDA_TRY(out->sqnorm.alloc((size_t)(out->rows_pad * wsz), s));
    const bool want32 = f64 && mode != PREP_DENSE;
    if (want32 && want_norm) CUDA_TRY(out->norm32.alloc((size_t)(out->rows_pad * 4), s));
    if (want32 && want_sq) CUDA_TRY(out->sqnorm32.alloc((size_t)(out->rows_pad * 4), s));
    out->max_sq_ptr = shared_max;
    if (want_max && !shared_max) {
        CUDA_TRY(out->max_sq.alloc(2 * sizeof(unsigned int), s));
        CUDA_TRY(init_norm_range(out->max_sq.as<unsigned int>(), s));
        out->max_sq_ptr = out->max_sq.as<unsigned int>();
    }
    PrepArgs a;
    a.values = m.values;
    a.offsets = m.offsets;
    a.validity = m.validity;
    a.row_validity = m.row_validity;
    a.n_rows = m.n_rows;
    a.dim = m.dim;
    a.rows_out = out->rows_pad;
    a.ld_out = out->ld;
    a.out0 = out->p0.p;
    a.out1 = out->p1.p;
    a.norm_out = out->norm.p;
    a.sqnorm_out = out->sqnorm.p;
    a.norm32_out = want32 ? out->norm32.as<float>() : nullptr;
    a.sqnorm32_out = want32 ? out->sqnorm32.as<float>() : nullptr;
    a.zero_guard_sq = f64 ? 1e-20f : 1e-12f;   // square of the cosine zero-norm guard (1e-10 f64 / 1e-6 f32)
    a.max_sq_out = out->max_sq_ptr;
    a.error_flag = d_err;
    a.nonfinite_rows = nonfinite_rows;
    a.nonfinite_count = nonfinite_count;
    a.scale_out = out->scale.as<float>();
    CUDA_TRY(launch_counted("prep", s, [&] { return launch_prep(a, m.dtype, mode, f64 ? 1 : 0, s); }));
    return PMM_OK;
}

// ------------------------------------------------------------------------------------------------ validation
int check_matrix(const pmm_matrix_t *m, const char *what) {
    if (!m) return fail(PMM_ERR_INVALID, "%s: null matrix descriptor", what);
    if (m->dtype < PMM_DTYPE_F16 || m->dtype > PMM_DTYPE_F64) return fail(PMM_ERR_INVALID, "%s: unknown dtype %d", what, m->dtype);
    if (m->n_rows < 0 || m->dim < 0) return fail(PMM_ERR_INVALID, "%s: negative shape", what);
    return PMM_OK;
}

// Shared shape checks in the reference's order (src/matmul.rs:131-175, :433-441).
int check_pair(const pmm_matrix_t *l, const pmm_matrix_t *r) {
    if (l->n_rows == 0 || r->n_rows == 0) return fail(PMM_ERR_INVALID, "Empty series");
    if (l->dim == 0 || r->dim == 0) return fail(PMM_ERR_INVALID, "Zero-dimensional vectors");
    if (l->dim != r->dim)
        return fail(PMM_ERR_INVALID, "Dimension mismatch: left has %lld dimensional vectors, right has %lld dimensional vectors",
                    (long long)l->dim, (long long)r->dim);
    if (r->n_rows >= 0xffffffffll) return fail(PMM_ERR_UNSUPPORTED, "corpus of 2^32-1 or more rows per call is not supported");
    return PMM_OK;
}

int finish_error_flag(int *d_err, cudaStream_t s) {
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (h)
        return fail(PMM_ERR_INVALID,
                    "ragged list column: a row is longer than row 0 (the reference panics here: ndarray index out of bounds)");
    return PMM_OK;
}

// ------------------------------------------------------------------------------------------------ top-k on device
struct TopkOut {
    uint32_t *index;
    double *score;
    uint64_t *cand;
};

// Generic SIMT path on prepared DENSE operands.
int topk_generic(const Prepared &q, const Prepared &c, int64_t keff, int metric, int64_t index_base, TopkOut o,
                 cudaStream_t s) {
    const bool f64 = q.f64;
    if (f64 && o.cand) return fail(PMM_ERR_UNSUPPORTED, "packed candidates exist for f32 working precision only");
    const int64_t Q = q.n_rows, N = c.n_rows, D = q.dim;
    const int64_t wsz = f64 ? 8 : 4;
    int64_t chunk = (t_opt.generic_ws_mb << 20) / (N * wsz);
    chunk = chunk / 64 * 64;
    if (chunk < 64) chunk = 64;
    if (chunk > 1 << 20) chunk = 1 << 20;
    if (chunk > Q) chunk = Q;
    DevBuf slab, scratch;
    CUDA_TRY(slab.alloc((size_t)(chunk * N * wsz), s));
    const int kpad = select_kpad(keff);
    if (kpad > select_smem_kpad_limit(f64)) CUDA_TRY(scratch.alloc((size_t)chunk * kpad * 12, s));
    const bool higher = metric != PMM_METRIC_EUCLIDEAN;
    const void *qa = metric == PMM_METRIC_COSINE ? q.norm.p : metric == PMM_METRIC_EUCLIDEAN ? q.sqnorm.p : nullptr;
    const void *ca = metric == PMM_METRIC_COSINE ? c.norm.p : metric == PMM_METRIC_EUCLIDEAN ? c.sqnorm.p : nullptr;
    for (int64_t q0 = 0; q0 < Q; q0 += chunk) {
        const int64_t nq = (Q - q0 < chunk) ? Q - q0 : chunk;
        uint32_t *oi = o.index ? o.index + q0 * keff : nullptr;
        double *os = o.score ? o.score + q0 * keff : nullptr;
        uint64_t *oc = o.cand ? o.cand + q0 * keff : nullptr;
        if (f64) {
            const bool simt = t_opt.f64_simt != 0;
            CUDA_TRY(launch_counted(simt ? "scores_f64" : "scores_f64_dmma", s, [&] {
                return (simt ? launch_scores_f64 : launch_scores_f64_dmma)(
                    q.p0.as<double>() + q0 * D, c.p0.as<double>(), qa ? (const double *)qa + q0 : nullptr, (const double *)ca, nq,
                    N, D, metric, slab.as<double>(), N, s);
            }));
            CUDA_TRY(launch_counted("select_f64", s, [&] {
                return launch_select_f64(slab.as<double>(), N, nq, N, keff, higher, index_base, oi, os, scratch.p, s);
            }));
        } else {
            CUDA_TRY(launch_counted("scores_f32", s, [&] {
                return launch_scores_f32(q.p0.as<float>() + q0 * D, c.p0.as<float>(), qa ? (const float *)qa + q0 : nullptr,
                                         (const float *)ca, nq, N, D, metric, slab.as<float>(), N, s);
            }));
            CUDA_TRY(launch_counted("select_f32", s, [&] {
                return launch_select_f32(slab.as<float>(), N, nq, N, keff, higher, index_base, oi, os, oc, scratch.p, s);
            }));
        }
    }
    return PMM_OK;
}

RawMatrix raw_of(const pmm_matrix_t &m) {
    RawMatrix r;
    r.values = m.values;
    r.offsets = m.offsets;
    r.validity = m.validity;
    r.row_validity = m.row_validity;
    r.n_rows = m.n_rows;
    r.dim = m.dim;
    r.dtype = m.dtype;
    return r;
}

// CTA pairs that share one query tile (they take corpus tiles rank, rank+g, ...). More sharers shrink the
// set of query tiles in flight — whose operand planes must stay L2-resident, they are re-read for every
// corpus tile — but every sharer pays the warm-up of its own candidate lists. Rule (measured with the
// pacing barriers on, profiles/sweep_r1.md): the smallest g in {1,2,4} that keeps the in-flight query
// planes under 64 MB, as long as every sharer still sweeps >= 64 corpus tiles.
int tc_group_for(int64_t corpus_rows, int64_t dim_pad, bool f16, int units, int cg) {
    int g = t_opt.tc_group;
    if (g > 0) return g;
    const int64_t n_tiles = (corpus_rows + TC_TILE_N - 1) / TC_TILE_N;
    const int64_t tile_bytes = (int64_t)TC_TILE_M * cg * dim_pad * (f16 ? 2 : 8);  // hi+lo planes for f32
    g = 1;
    while (g < 4 && (units / g) * tile_bytes > (64ll << 20) && n_tiles / (2 * g) >= 64) g *= 2;
    return g;
}

// List capacity of the tensor-core filter: at least 8 more candidates than requested are kept, so the
// exact re-scoring can reorder near-ties across the k-th position.  k <= 248.
constexpr int64_t TC_MAX_K = 248;
int tc_list_capacity(int64_t keff) { return keff <= 24 ? 32 : keff <= 56 ? 64 : keff <= 120 ? 128 : 256; }

// Tensor-core filter on prepared PLANES: fused kernel -> merge of the corpus pieces of every query tile.
// kept [Q x kp]: per query the kp best candidates under the kernel's filter value (approximate keys,
// comparable across corpus chunks and shards of the same query), indices = index_base + corpus row.
// Corpus in chunks (host path): `carry` keeps the per-row candidate lists between the launches, so a later chunk
// starts with full lists and tight thresholds instead of paying the list warm-up again, and the lists of all
// chunks never have to be merged.  All launches of one carry use the schedule group of the WHOLE corpus
// (`corpus_rows_total`), which fixes the list layout.  phase bit 0: first launch (fresh lists), bit 1: last launch
// (merge the pieces of every query tile into `kept`).
// seed: per query row (padded like the planes) an initial threshold in filter units, or NULL (see make_seeds_kernel).
struct TcCarry {
    DevBuf partial;
    int64_t corpus_rows_total = 0;
    int64_t layout_rows = 0;  // rows of the smallest chunk (caps the sharing factors of every launch alike)
};

const char *tc_kernel_stat_name(const Prepared &q, int terms, int kp, bool seeded) {
    if (q.mode == PREP_F16R) return seeded ? "tc_topk_f16r_seeded" : kp == 256 ? "tc_topk_f16r_kp256" : "tc_topk_f16r";
    if (q.mode == PREP_F16) return "tc_topk_f16";
    return terms == 1 ? "tc_topk_tf32x1" : "tc_topk_tf32x3";
}

// pipe: the merge of the launch's corpus pieces runs on pipe->s2 (after an event on s) out of pipe->partial, a list buffer
// the CALLER owns and recycles, instead of on s out of a buffer of this launch - the next filter launch can then start
// on s while s2 still merges (filter_rescore_pipelined).
struct TcPipe {
    cudaStream_t s2 = nullptr;
    DevBuf *partial = nullptr;
    cudaEvent_t filter_done = nullptr;
};

cudaStream_t aux_stream() {
    static thread_local cudaStream_t s = nullptr;
    static thread_local int dev_of_stream = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!s || dev_of_stream != dev) {
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        dev_of_stream = dev;
    }
    return s;
}

int tc_filter(const Prepared &q, const Prepared &c, int kp, int metric, int64_t index_base, uint64_t *kept,
              cudaStream_t s, int terms, TcCarry *carry = nullptr, int phase = 3, const float *seed = nullptr,
              TcPipe *pipe = nullptr, const uint64_t *ceil = nullptr, int k_thr = 0, const char *stat_name = nullptr,
              int sample_tiles = 0) {
    DevInfo &di = dev_info();
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.q_hi = q.p0.p;
    a.q_lo = q.p1.p;
    a.c_hi = c.p0.p;
    a.c_lo = c.p1.p;
    a.q_rows_pad = q.rows_pad;
    a.c_rows_pad = c.rows_pad;
    a.dim_pad = q.ld;
    a.nq = q.n_rows;
    a.n = c.n_rows;
    a.f16 = (q.mode == PREP_F16 || q.mode == PREP_F16R) ? 1 : 0;
    a.terms = a.f16 ? 1 : terms;
    a.cg = (a.terms == 1 && !a.f16) ? 2 : t_opt.tc_cg;
    a.clm = (a.cg == 2 && t_opt.tc_clm == 2) ? 2 : 1;
    a.cluster4 = t_opt.tc_cluster4;
    a.debug_skip = t_opt.tc_debug_skip;
    a.sync_slack = t_opt.tc_sync_slack;
    a.max_flush = t_opt.tc_max_flush;
    a.soft_at = t_opt.tc_soft_at;
    const int gs = a.cg * a.clm;   // CTAs per scheduling unit
    int units = di.num_sms / gs;
    if (t_opt.tc_max_units > 0 && units > t_opt.tc_max_units) units = t_opt.tc_max_units;
    // sample_tiles > 0: the launch covers a strided SAMPLE of the corpus tiles (warm seeds): every stride-th whole tile
    int64_t sched_rows = c.n_rows;
    if (sample_tiles > 0) {
        const int64_t whole = c.n_rows / TC_TILE_N;
        if (whole < sample_tiles) sample_tiles = (int)whole;
        if (sample_tiles < 1) return fail(PMM_ERR_INVALID, "tc_filter: corpus smaller than one tile cannot be sampled");
        a.tile_stride = (int)(whole / sample_tiles);
        sched_rows = (int64_t)sample_tiles * TC_TILE_N;
    }
    const int64_t group_rows = carry ? carry->corpus_rows_total : sched_rows;
    a.sched = make_tc_schedule(q.n_rows, sched_rows, units, tc_group_for(group_rows, q.ld, a.f16 != 0 || a.terms == 1, units, gs), gs,
                               carry ? carry->layout_rows : 0);
    // the filter runs on f32 copies of the norms (f64 working precision keeps the exact ones for the re-scoring)
    a.q_aux = metric == PMM_METRIC_COSINE ? q.norm_f32() : metric == PMM_METRIC_EUCLIDEAN ? q.sq_f32() : nullptr;
    a.c_aux = metric == PMM_METRIC_COSINE ? c.norm_f32() : metric == PMM_METRIC_EUCLIDEAN ? c.sq_f32() : nullptr;
    // f64: the reference's guard is 1e-10 (src/metrics.rs:275); a shade lower on the rounded norm, so that a row the
    // exact pass treats as non-zero is never zeroed by the filter (the other way round only costs a list slot)
    a.norm_guard = q.f64 ? 0.999999e-10f : 0.0f;
    a.seed_thr = seed;
    a.ceil = ceil;
    a.index_base = index_base;
    a.metric = metric;
    a.kp = kp;
    a.k = (k_thr > 0 && k_thr <= kp) ? k_thr : kp;   // list position that sets a row's threshold
    DevBuf own_partial, rsync, staged;
    DevBuf &partial = pipe ? *pipe->partial : carry ? carry->partial : own_partial;
    const int esets = tc_epilogue_sets(a.f16, a.terms);
    if (!carry || (phase & 1)) CUDA_TRY(partial.alloc((size_t)a.sched.total_slots() * esets * gs * TC_TILE_M * a.kp * 8, s));
    a.resume = (carry && !(phase & 1)) ? 1 : 0;
    CUDA_TRY(staged.alloc((size_t)tc_staged_bytes(a.sched.num_ctas * gs, esets), s));
    a.staged = staged.as<uint64_t>();
    if (t_opt.tc_sync_tiles > 0) {
        a.sync_tiles = t_opt.tc_sync_tiles;
        const size_t nb = (size_t)tc_sync_counters(a.sched, a.sync_tiles) * sizeof(unsigned int);
        CUDA_TRY(rsync.alloc(nb, s));
        CUDA_TRY(cudaMemsetAsync(rsync.p, 0, nb, s));
        a.round_sync = rsync.as<unsigned int>();
    }
    a.partial = partial.as<uint64_t>();
    // (256-entry lists and seeded launches are the retry levels: statistics of their own)
    cudaError_t e = launch_counted(stat_name ? stat_name : tc_kernel_stat_name(q, a.terms, kp, seed != nullptr), s, [&] { return launch_tc_topk(a, s); });
    if (e != cudaSuccess)
        return fail(PMM_ERR_CUDA, "tensor-core top-k launch failed: %s %s", cudaGetErrorString(e), tc_last_error());
    cudaStream_t ms = s;
    if (pipe) {
        CUDA_TRY(cudaEventRecord(pipe->filter_done, s));
        CUDA_TRY(cudaStreamWaitEvent(pipe->s2, pipe->filter_done, 0));
        ms = pipe->s2;
    }
    if (phase & 2)
        CUDA_TRY(launch_counted("merge", ms, [&] {
            return launch_merge_tiles(a.partial, a.sched, gs, esets, a.kp, q.n_rows, a.kp, true, nullptr, nullptr, kept, ms);
        }));
    return PMM_OK;
}

// Error model of one filter level (formulas and their derivation: pmm_kernels.h).
struct LevelErr {
    float eps, abs_err, max_norm;
};
LevelErr level_err(int mode, int terms, bool f64_source, int64_t dim) {
    if (mode == PREP_F16) return LevelErr{filter_eps(dim, 1, true), 0.0f, 0.0f};   // exact planes of f16 input
    if (mode == PREP_F16R)                                                        // rounded to f16: 11 bits + range terms
        return LevelErr{filter_eps(dim, 1, false) + (f64_source ? 2.0e-7f : 0.0f), f16r_abs_err(dim), 65504.0f};
    // TF32 planes. From f64 sources the value is first rounded to f32 (2^-24 per operand) and may leave the f32 range.
    return LevelErr{filter_eps(dim, terms, false) + (f64_source ? 1.3e-7f : 0.0f), f64_source ? f32_flush_abs_err(dim) : 0.0f,
                    f64_source ? 1.0e38f : 0.0f};
}

// What all levels of one top-k call share: the raw corpus on the device (the exact re-scoring gathers candidate rows
// from it), the corpus norms in the working type (exact metric pass) and as f32 (proof), and the call's parameters.
struct VerifyCtx {
    pmm_matrix_t raw_c;
    bool f64 = false;
    const void *c_norm = nullptr, *c_sq = nullptr;   // working type, whole corpus
    const unsigned int *c_range = nullptr;           // norm range from the prep pass (f32 bits)
    int metric = 0;
    int64_t index_base = 0, keff = 0;
    cudaStream_t s = nullptr;
};

// seed: thresholds of a RE-QUERY level (guaranteed lower bounds).  warm: thresholds a first level started from (warm
// seeds, below) when the caller ran the filter itself (kept_in): the losslessness check has to know them.
int tc_topk_verified(const VerifyCtx &vc, const Prepared &q, const pmm_matrix_t &raw_q, const Prepared *c, int terms,
                     const uint64_t *kept_in, TopkOut o, int kp_override = 0, const float *seed = nullptr, const float *warm = nullptr);

// Rows [r0, r0 + rows) of a device-resident matrix as a matrix of its own. r0 must be a multiple of 256 (element
// validity bitmaps are re-based by whole bytes).
pmm_matrix_t slice_rows(const pmm_matrix_t &m, int64_t r0, int64_t rows) {
    pmm_matrix_t dm = m;
    dm.n_rows = rows;
    if (dm.offsets) {
        dm.offsets += r0;  // offsets hold absolute child positions: values / validity stay as they are
    } else {
        dm.values = (const char *)dm.values + (size_t)r0 * m.dim * esize(m.dtype);
        if (dm.validity) dm.validity += (r0 * m.dim) / 8;
    }
    if (dm.row_validity) dm.row_validity += r0 / 8;
    return dm;
}

// Exact re-scoring of the kept candidates + the losslessness check.  Queries the check cannot clear are gathered and
// recomputed one level up: next_terms = 1: the same f16-rounded filter again, 256 candidates per query; 3: the 3xTF32
// filter; 0: the exact SIMT path.  Re-query levels on the tensor cores start from SEEDED thresholds (option
// "seed_retry"): the flagged query's exact k-th score among the candidates at hand bounds what can still matter.
// May synchronise the stream.
struct RescoreJob {   // buffers and parameters of one level's exact re-scoring + losslessness check
    DevBuf flags, count, kth;
    RescoreCheck chk;          // pointers for query 0 of the level
    bool verify = false;
    const void *q_aux = nullptr, *c_aux = nullptr;   // working type
    int kp = 0;
};

int rescore_setup(const VerifyCtx &vc, const Prepared &q, int64_t Q, int kp, LevelErr le, const float *seed, RescoreJob *job) {
    cudaStream_t s = vc.s;
    job->kp = kp;
    job->q_aux = vc.metric == PMM_METRIC_COSINE ? q.norm.p : vc.metric == PMM_METRIC_EUCLIDEAN ? q.sqnorm.p : nullptr;
    job->c_aux = vc.metric == PMM_METRIC_COSINE ? vc.c_norm : vc.metric == PMM_METRIC_EUCLIDEAN ? vc.c_sq : nullptr;
    memset(&job->chk, 0, sizeof(job->chk));
    const float *q_sq32 = q.sq_f32();
    job->verify = t_opt.verify && q_sq32 && vc.c_range;
    if (job->verify) {
        CUDA_TRY(job->flags.alloc((size_t)Q, s));
        CUDA_TRY(job->count.alloc(sizeof(unsigned int), s));
        CUDA_TRY(job->kth.alloc((size_t)Q * 4, s));
        CUDA_TRY(cudaMemsetAsync(job->flags.p, 0, (size_t)Q, s));
        CUDA_TRY(cudaMemsetAsync(job->count.p, 0, sizeof(unsigned int), s));
        job->chk.q_sq = q_sq32;
        job->chk.c_max_sq = vc.c_range;
        job->chk.eps = le.eps;
        job->chk.abs_err = le.abs_err;
        job->chk.max_norm = le.max_norm;
        job->chk.seed = seed;
        job->chk.flags = job->flags.as<unsigned char>();
        job->chk.flag_count = job->count.as<unsigned int>();
        job->chk.kth_units = job->kth.as<float>();
    }
    return PMM_OK;
}

// Re-scores the queries [q0, q0 + rows) of the level (q0 a multiple of 256) on `stream`; `kept` and `o` are the level's
// whole buffers.
int rescore_launch(const VerifyCtx &vc, const RescoreJob &job, const pmm_matrix_t &raw_q, const uint64_t *kept, int64_t q0, int64_t rows,
                   TopkOut o, cudaStream_t stream) {
    const int64_t keff = vc.keff;
    const pmm_matrix_t rq = (q0 == 0 && rows == raw_q.n_rows) ? raw_q : slice_rows(raw_q, q0, rows);
    RescoreCheck chk = job.chk;
    chk.stream_loads = (stream != vc.s && t_opt.rescore_stream_loads) ? 1 : 0;
    if (job.verify) {
        chk.q_sq += q0;
        chk.flags += q0;
        chk.kth_units += q0;
        if (chk.seed) chk.seed += q0;
    }
    const int64_t wsz = vc.f64 ? 8 : 4;
    const void *qa = job.q_aux ? (const char *)job.q_aux + q0 * wsz : nullptr;
    uint32_t *oi = o.index ? o.index + q0 * keff : nullptr;
    double *os = o.score ? o.score + q0 * keff : nullptr;
    uint64_t *oc = o.cand ? o.cand + q0 * keff : nullptr;
    if (vc.f64) {
        if (o.cand) return fail(PMM_ERR_UNSUPPORTED, "packed candidates exist for f32 working precision only");
        CUDA_TRY(launch_counted("rescore_f64", stream, [&] {
            return launch_rescore_f64(kept + q0 * job.kp, job.kp, raw_of(rq), raw_of(vc.raw_c), (const double *)qa, (const double *)job.c_aux,
                                      vc.metric, vc.index_base, (int)keff, oi, os, chk, stream);
        }));
    } else {
        CUDA_TRY(launch_counted("rescore", stream, [&] {
            return launch_rescore(kept + q0 * job.kp, job.kp, raw_of(rq), raw_of(vc.raw_c), (const float *)qa, (const float *)job.c_aux,
                                  vc.metric, vc.index_base, (int)keff, oi, os, oc, chk, stream);
        }));
    }
    return PMM_OK;
}

int rescore_finish(const VerifyCtx &vc, RescoreJob &job, const pmm_matrix_t &raw_q, const Prepared *c_planes, int next_terms, TopkOut o);

int rescore_and_verify(const VerifyCtx &vc, const Prepared &q, const pmm_matrix_t &raw_q, const uint64_t *kept, int kp,
                       const Prepared *c_planes, LevelErr le, int next_terms, const float *seed, TopkOut o) {
    RescoreJob job;
    int rc;
    if ((rc = rescore_setup(vc, q, raw_q.n_rows, kp, le, seed, &job))) return rc;
    if ((rc = rescore_launch(vc, job, raw_q, kept, 0, raw_q.n_rows, o, vc.s))) return rc;
    return rescore_finish(vc, job, raw_q, c_planes, next_terms, o);
}

// The flagged queries of a level -> the next level (see rescore_and_verify's comment above).
int rescore_finish(const VerifyCtx &vc, RescoreJob &job, const pmm_matrix_t &raw_q, const Prepared *c_planes, int next_terms, TopkOut o) {
    cudaStream_t s = vc.s;
    const int metric = vc.metric;
    const int64_t Q = raw_q.n_rows, keff = vc.keff;
    const bool verify = job.verify;
    DevBuf &flags = job.flags, &count = job.count, &kth = job.kth;
    RescoreCheck &chk = job.chk;
    if (!verify) return PMM_OK;
    unsigned int n_flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_flag, count.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (n_flag == 0) return PMM_OK;
    stat_add(next_terms == 1 ? "requeried_f16_wide" : next_terms == 3 ? "requeried_tf32x3" : "fallback_queries", (double)n_flag);
    // ---- gather the flagged queries into a dense matrix of the working type
    std::vector<unsigned char> hflags((size_t)Q);
    CUDA_TRY(cudaMemcpy(hflags.data(), flags.p, (size_t)Q, cudaMemcpyDeviceToHost));
    std::vector<int64_t> ids;
    ids.reserve(n_flag);
    for (int64_t i = 0; i < Q; ++i)
        if (hflags[(size_t)i]) ids.push_back(i);
    const int64_t F = (int64_t)ids.size();
    const int64_t wsz = vc.f64 ? 8 : 4;
    DevBuf d_ids, dense_q, t_idx, t_sc, t_cand, err, seeds;
    CUDA_TRY(d_ids.alloc((size_t)F * 8, s));
    CUDA_TRY(cudaMemcpyAsync(d_ids.p, ids.data(), (size_t)F * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(dense_q.alloc((size_t)F * raw_q.dim * wsz, s));
    CUDA_TRY(launch_counted("gather", s, [&] { return launch_gather_rows(raw_of(raw_q), d_ids.as<int64_t>(), F, dense_q.p, vc.f64 ? 1 : 0, s); }));
    pmm_matrix_t qd;
    memset(&qd, 0, sizeof(qd));
    qd.values = dense_q.p;
    qd.n_rows = F;
    qd.dim = raw_q.dim;
    qd.dtype = vc.f64 ? PMM_DTYPE_F64 : PMM_DTYPE_F32;
    CUDA_TRY(err.alloc(sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    CUDA_TRY(t_idx.alloc((size_t)F * keff * 4, s));
    CUDA_TRY(t_sc.alloc((size_t)F * keff * 8, s));
    if (o.cand) CUDA_TRY(t_cand.alloc((size_t)F * keff * 8, s));
    TopkOut t{t_idx.as<uint32_t>(), t_sc.as<double>(), o.cand ? t_cand.as<uint64_t>() : nullptr};
    const bool want_norm = metric == PMM_METRIC_COSINE, want_sq = metric == PMM_METRIC_EUCLIDEAN;
    // the few re-queried rows are padded to ONE scheduling unit's query tile (256 rows for CTA pairs), not to the 512 the
    // bulk path uses: a second, all-padding query tile would make half of the units stream the corpus planes for nothing
    const int64_t q_tile = (int64_t)TC_TILE_M * std::max(2, t_opt.tc_cg * (t_opt.tc_cg == 2 && t_opt.tc_clm == 2 ? 2 : 1));
    const float *seed_next = nullptr;
    if (next_terms != 0 && t_opt.seed_retry) {
        const int64_t n_pad = round_up(F, q_tile);
        CUDA_TRY(seeds.alloc((size_t)n_pad * 4, s));
        const LevelErr ne = level_err(next_terms == 1 ? PREP_F16R : PREP_TF32, next_terms, vc.f64, raw_q.dim);
        RescoreCheck nx = chk;
        nx.eps = ne.eps;
        nx.abs_err = ne.abs_err;
        nx.max_norm = ne.max_norm;
        CUDA_TRY(launch_counted("seeds", s, [&] {
            return launch_make_seeds(d_ids.as<int64_t>(), F, n_pad, kth.as<float>(), nx, metric, seeds.as<float>(), s);
        }));
        seed_next = seeds.as<float>();
    }
    int rc;
    if (next_terms == 1) {
        // same f16-rounded filter against the planes at hand, 256 candidates per query
        Prepared qf;
        if ((rc = prepare(qd, PREP_F16R, vc.f64, q_tile, want_norm, true, err.as<int>(), s, &qf))) return rc;
        if ((rc = tc_topk_verified(vc, qf, qd, c_planes, 1, nullptr, t, 256, seed_next))) return rc;
    } else if (next_terms == 3) {
        // one level up on the tensor cores: 3xTF32 planes of the flagged queries against the corpus planes
        Prepared qf;
        if ((rc = prepare(qd, PREP_TF32, vc.f64, q_tile, want_norm, true, err.as<int>(), s, &qf))) return rc;
        if (c_planes) {
            if ((rc = tc_topk_verified(vc, qf, qd, c_planes, 3, nullptr, t, 0, seed_next))) return rc;
        } else {
            // The corpus planes at hand are f16 planes: rebuild TF32 planes from the resident raw corpus piece by piece
            // (at most 128k rows at a time: bounded, equally sized scratch that the memory pool hands back without
            // mapping new memory - a 2 x corpus-size request after the chunk-sized frees stalled for 0.3-1.5 s),
            // carrying the candidate lists from piece to piece.
            const int64_t N = vc.raw_c.n_rows, step = 131072;
            const int kp1 = tc_list_capacity(keff);
            DevBuf kept1;
            CUDA_TRY(kept1.alloc((size_t)F * kp1 * 8, s));
            TcCarry carry;
            carry.corpus_rows_total = N;
            carry.layout_rows = N < step ? N : (N % step ? N % step : step);
            Prepared cpiece;  // plane buffers of the first (largest) piece are reused by the others
            for (int64_t r0 = 0; r0 < N; r0 += step) {
                const int64_t rows = N - r0 < step ? N - r0 : step;
                if ((rc = prepare(slice_rows(vc.raw_c, r0, rows), PREP_TF32, vc.f64, TC_TILE_N, want_norm, want_sq, err.as<int>(), s, &cpiece)))
                    return rc;
                if ((rc = tc_filter(qf, cpiece, kp1, metric, vc.index_base + r0, kept1.as<uint64_t>(), s, 3, &carry,
                                    (r0 == 0 ? 1 : 0) | (r0 + rows >= N ? 2 : 0), seed_next)))
                    return rc;
            }
            if ((rc = tc_topk_verified(vc, qf, qd, nullptr, 3, kept1.as<uint64_t>(), t, 0, seed_next))) return rc;
        }
    } else {
        Prepared qf, cf;
        if ((rc = prepare(qd, PREP_DENSE, vc.f64, 1, want_norm, want_sq, err.as<int>(), s, &qf))) return rc;
        if ((rc = prepare(vc.raw_c, PREP_DENSE, vc.f64, 1, want_norm, want_sq, err.as<int>(), s, &cf))) return rc;
        if ((rc = topk_generic(qf, cf, keff, metric, vc.index_base, t, s))) return rc;
    }
    CUDA_TRY(launch_counted("scatter", s, [&] {
        return launch_scatter_results(d_ids.as<int64_t>(), F, (int)keff, t.index, t.score, t.cand, o.index, o.score, o.cand, s);
    }));
    CUDA_TRY(cudaStreamSynchronize(s));  // ids (host vector) and temporaries stay alive until the work is done
    return PMM_OK;
}

// ---- pipelined first level --------------------------------------------------------------------------------------
// Query rows per filter launch of the pipelined first level = one ROUND of the persistent schedule (every scheduling
// unit takes one query tile and sweeps the whole corpus), or 0 when pipelining is off / pointless.
int64_t pipeline_part_rows(const Prepared &q, const Prepared &c, int terms) {
    if (!t_opt.pipeline) return 0;
    DevInfo &di = dev_info();
    const bool f16 = q.mode == PREP_F16 || q.mode == PREP_F16R;
    const int cg = (terms == 1 && !f16) ? 2 : t_opt.tc_cg;
    const int gs = cg * ((cg == 2 && t_opt.tc_clm == 2) ? 2 : 1);
    int units = di.num_sms / gs;
    if (t_opt.tc_max_units > 0 && units > t_opt.tc_max_units) units = t_opt.tc_max_units;
    const int g = tc_group_for(c.n_rows, q.ld, f16 || terms == 1, units, gs);
    const int mc = units / (g < 1 ? 1 : g);
    if (mc < 1) return 0;
    // a round must be long enough to hide a launch boundary (~20 us) and a merge + re-scoring of the previous one
    const double round_flop = 2.0 * (double)mc * TC_TILE_M * gs * (double)c.n_rows * (double)q.dim;
    if (round_flop < 1.0e9 * (double)t_opt.pipeline_min_gflop) return 0;
    return (int64_t)mc * TC_TILE_M * gs;
}

// Rows [q0, q0 + rows) of prepared query operands as an operand set of its own (views, nothing is copied).
void view_query_rows(const Prepared &q, int64_t q0, int64_t rows, cudaStream_t s, Prepared *v) {
    const bool half_plane = q.mode == PREP_F16 || q.mode == PREP_F16R;
    const int64_t es = half_plane ? 2 : 4, wsz = q.f64 ? 8 : 4;
    v->mode = q.mode;
    v->f64 = q.f64;
    v->n_rows = rows;
    v->dim = q.dim;
    v->ld = q.ld;
    v->rows_pad = q.rows_pad - q0;
    const size_t pb = (size_t)v->rows_pad * q.ld * es;
    v->p0.borrow((char *)q.p0.p + (size_t)q0 * q.ld * es, pb, s);
    if (q.p1.p) v->p1.borrow((char *)q.p1.p + (size_t)q0 * q.ld * es, pb, s);
    if (q.norm.p) v->norm.borrow((char *)q.norm.p + q0 * wsz, (size_t)v->rows_pad * wsz, s);
    if (q.sqnorm.p) v->sqnorm.borrow((char *)q.sqnorm.p + q0 * wsz, (size_t)v->rows_pad * wsz, s);
    if (q.norm32.p) v->norm32.borrow((char *)q.norm32.p + q0 * 4, (size_t)v->rows_pad * 4, s);
    if (q.sqnorm32.p) v->sqnorm32.borrow((char *)q.sqnorm32.p + q0 * 4, (size_t)v->rows_pad * 4, s);
    v->max_sq_ptr = q.max_sq_ptr;
}

int filter_rescore_pipelined(const VerifyCtx &vc, const Prepared &q, const pmm_matrix_t &raw_q, const Prepared &c, int terms, int kp,
                             int64_t part_rows, uint64_t *kept, const RescoreJob &job, TopkOut o) {
    cudaStream_t s = vc.s, s2 = t_opt.pipeline == 2 ? vc.s : aux_stream();
    const int64_t Q = q.n_rows;
    const int n_parts = (int)((Q + part_rows - 1) / part_rows);
    DevBuf partial[2];
    struct Events {
        std::vector<cudaEvent_t> ev;
        ~Events() {
            for (cudaEvent_t e : ev)
                if (e) cudaEventDestroy(e);
        }
        cudaError_t make(cudaEvent_t *e) {
            cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
            if (r == cudaSuccess) ev.push_back(*e);
            return r;
        }
    } events;
    // every way out joins the second stream into s first: buffers freed on s afterwards must not be in use on s2
    struct Join {
        cudaStream_t s, s2;
        ~Join() {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) {
                cudaEventRecord(e, s2);
                cudaStreamWaitEvent(s, e, 0);
                cudaEventDestroy(e);
            } else {
                cudaStreamSynchronize(s2);
            }
        }
    } join{s, s2};
    cudaEvent_t start;
    CUDA_TRY(events.make(&start));
    CUDA_TRY(cudaEventRecord(start, s));                 // everything prepared on s so far (planes, norms, flag buffers)
    CUDA_TRY(cudaStreamWaitEvent(s2, start, 0));
    std::vector<cudaEvent_t> part_done(n_parts, nullptr);
    for (int p = 0; p < n_parts; ++p) {
        const int64_t q0 = (int64_t)p * part_rows, rows = std::min<int64_t>(part_rows, Q - q0);
        Prepared qv;
        view_query_rows(q, q0, rows, s, &qv);
        if (p >= 2) CUDA_TRY(cudaStreamWaitEvent(s, part_done[p - 2], 0));   // its list buffer is being recycled
        TcPipe pipe;
        pipe.s2 = s2;
        pipe.partial = &partial[p & 1];
        CUDA_TRY(events.make(&pipe.filter_done));
        int rc = tc_filter(qv, c, kp, vc.metric, vc.index_base, kept + q0 * kp, s, terms, nullptr, 3, nullptr, &pipe);
        if (rc) return rc;
        if ((rc = rescore_launch(vc, job, raw_q, kept, q0, rows, o, s2))) return rc;
        CUDA_TRY(events.make(&part_done[p]));
        CUDA_TRY(cudaEventRecord(part_done[p], s2));
    }
    return PMM_OK;   // ~Join: s waits for s2
}

// Filter at `terms` (unless the kept lists are supplied) -> exact re-scoring -> verification -> next level.
// c may be NULL only when kept_in is given.
// ---- warm seeds: thresholds for the first filter level from a sample pre-pass -----------------------------------------
// The first tiles of every query tile's sweep are dominated by list maintenance: with open thresholds every score is a
// candidate, and the tensor cores wait for the epilogue (measured at C3, scripts/filter_wait_histogram.py: 85 % of the
// MMA warps' accumulator waits fall into the first 512 of 3906 tiles, the first 64 tiles alone cost ~6 % of the launch).
// A pre-pass filters the first 1024-4096 corpus rows with 32-entry lists whose r-th entry sets the threshold (cheap to
// maintain), and the r-th best sample value of a query seeds its row in the real launch: "collect what beats the best
// 0.2-0.8 % of a sample" (C3: 114.7 -> 108.7 ms for a 2.0 ms pre-pass; sweep in profiles/sweep_r2.md).  The seed is a lower bound of the final k-th best only with overwhelming probability (it needs k
// corpus rows above the r-th best of the sample) - which is all it has to be: a seeded row whose list does not fill is dropped
// by the losslessness check like any other unprovable row and re-queried one level up.  Applied when the corpus is at
// least 16 k / r samples long, so that the expected number of rows above the seed is >= 16 k.
// Sample size: the pre-pass costs sample / corpus of the launch's MMA work plus its own list warm-up, so short corpora
// (C5: 125k rows per GPU, a million queries) get a smaller sample: N / 128 rows, within [1024, "warm_rows" = 4096].
int64_t warm_rows_for(int64_t c_rows_total) {
    int64_t s = c_rows_total / 128 / 256 * 256;
    if (s < 1024) s = 1024;
    if (s > t_opt.warm_rows) s = t_opt.warm_rows;
    return s;
}
// The sample rank that becomes the seed: the smallest r >= 8 (a stable order statistic) for which the corpus is expected
// to hold >= 16 k rows above the r-th best of the samples; 0 = no seeds (r would exceed the 32-entry sample lists).
int warm_seed_rank(int64_t c_rows_total, int64_t keff) {
    if (t_opt.warm_rank > 0) return t_opt.warm_rank;
    const int64_t need = (16 * keff * warm_rows_for(c_rows_total) + c_rows_total - 1) / (c_rows_total > 0 ? c_rows_total : 1);
    return need > 32 ? 0 : need < 8 ? 8 : (int)need;
}
bool warm_seed_applies(int q_mode, int64_t q_rows, int64_t c_rows_total, int64_t c_rows_at_hand, int64_t keff, int kp) {
    if (!t_opt.warm_seed || !(q_mode == PREP_F16R || q_mode == PREP_F16) || kp > 256) return false;
    const int64_t rows = warm_rows_for(c_rows_total);
    if (q_rows < 2048 || c_rows_at_hand < rows) return false;
    return c_rows_total >= 16 * rows && warm_seed_rank(c_rows_total, keff) > 0;
}
bool warm_seed_applies(const Prepared &q, int64_t c_rows_total, int64_t c_rows_at_hand, int64_t keff, int kp) {
    return warm_seed_applies(q.mode, q.n_rows, c_rows_total, c_rows_at_hand, keff, kp);
}
// c: prepared planes holding at least warm_rows_for(c_rows_total) rows (the corpus, its first chunk, or a host-side sample).
// out: [q.rows_pad] floats.
int warm_seeds(const Prepared &q, const Prepared &c, int64_t c_rows_total, int64_t keff, int metric, int terms, cudaStream_t s, DevBuf *out) {
    // The sample is STRIDED over the rows at hand (every stride-th whole corpus tile): a corpus sorted by norm, topic or
    // popularity would otherwise hand every query a sample far better (or worse) than the rest, and seeds that are too
    // high leave the lists unfilled - correct, but every such row is re-queried.
    const int sample_tiles = (int)(warm_rows_for(c_rows_total) / TC_TILE_N);
    DevBuf kept_s;
    CUDA_TRY(kept_s.alloc((size_t)q.n_rows * 32 * 8, s));
    const int r = warm_seed_rank(c_rows_total, keff);
    int rc = tc_filter(q, c, 32, metric, 0, kept_s.as<uint64_t>(), s, terms, nullptr, 3, nullptr, nullptr, nullptr, r, "tc_topk_warm", sample_tiles);
    if (rc) return rc;
    CUDA_TRY(out->alloc((size_t)q.rows_pad * 4, s));
    CUDA_TRY(launch_counted("seeds", s, [&] { return launch_seeds_from_lists(kept_s.as<uint64_t>(), 32, r, q.n_rows, q.rows_pad, out->as<float>(), s); }));
    return PMM_OK;
}

int tc_topk_verified(const VerifyCtx &vc, const Prepared &q, const pmm_matrix_t &raw_q, const Prepared *c, int terms,
                     const uint64_t *kept_in, TopkOut o, int kp_override, const float *seed, const float *warm) {
    const int kp = kp_override ? kp_override : tc_list_capacity(vc.keff);
    const bool f16 = q.mode == PREP_F16;     // exact f16 planes
    const bool f16r = q.mode == PREP_F16R;   // input rounded to f16: TF32-x1-like error, then 3xTF32 on demand
    DevBuf kept, warm_buf;
    const uint64_t *kept_ptr = kept_in;
    const LevelErr le = level_err(q.mode, f16r ? 1 : terms, vc.f64, raw_q.dim);
    // next level for the queries the proof rejects (see below)
    const bool wide_ = f16r && c && kp < 256 && !seed && t_opt.f16r_wide;
    const int next_terms_ = wide_ ? 1 : (f16r || (!f16 && terms == 1)) ? 3 : 0;
    if (!kept_ptr) {
        CUDA_TRY(kept.alloc((size_t)q.n_rows * kp * 8, vc.s));
        // Large query batches: one filter launch per ROUND of query tiles, and the merge + exact re-scoring of a round
        // run on a second stream while the tensor cores filter the next one (the filter leaves DRAM ~98 % idle, the
        // re-scoring is a DRAM gather; its blocks fit beside the persistent filter CTAs).
        int rc = PMM_OK;
        const int64_t part_rows = pipeline_part_rows(q, *c, terms);
        if (part_rows > 0 && !seed && !kp_override && q.n_rows >= 2 * part_rows) {
            RescoreJob job;
            if ((rc = rescore_setup(vc, q, raw_q.n_rows, kp, le, seed, &job))) return rc;
            if ((rc = filter_rescore_pipelined(vc, q, raw_q, *c, terms, kp, part_rows, kept.as<uint64_t>(), job, o))) return rc;
            return rescore_finish(vc, job, raw_q, (f16r && !wide_) ? nullptr : c, next_terms_, o);
        }
        if (!seed && !kp_override && warm_seed_applies(q, c->n_rows, c->n_rows, vc.keff, kp)) {
            if ((rc = warm_seeds(q, *c, c->n_rows, vc.keff, vc.metric, terms, vc.s, &warm_buf))) return rc;
            warm = warm_buf.as<float>();
        }
        // (a warm-seeded first level keeps the first level's statistics name; "..._seeded" are the re-query levels)
        rc = tc_filter(q, *c, kp, vc.metric, vc.index_base, kept.as<uint64_t>(), vc.s, terms, nullptr, 3, seed ? seed : warm, nullptr, nullptr, 0,
                       seed ? nullptr : tc_kernel_stat_name(q, f16r ? 1 : terms, kp, false));
        if (rc) return rc;
        kept_ptr = kept.as<uint64_t>();
    }
    // Next level for the queries the proof rejects.  f16-rounded level with its planes at hand: first the SAME filter
    // with 256-entry lists (next_terms = 1) - the proof needs the exact k-th score to clear the worst kept filter
    // value by the error bound; seeded thresholds (or, without them, 128 more ranks of margin) almost always do it,
    // for one small launch instead of rebuilding TF32 planes of the whole corpus.  Then 3xTF32 (3), then the exact
    // SIMT path (0).
    const bool wide = f16r && c && kp < 256 && !seed && t_opt.f16r_wide;
    const int next_terms = wide ? 1 : (f16r || (!f16 && terms == 1)) ? 3 : 0;
    // f16 planes cannot serve the 3xTF32 level: it rebuilds TF32 planes piece by piece (c_planes = NULL)
    return rescore_and_verify(vc, q, raw_q, kept_ptr, kp, (f16r && !wide) ? nullptr : c, level_err(q.mode, f16r ? 1 : terms, vc.f64, raw_q.dim),
                              next_terms, seed ? seed : warm, o);
}

// First filter level for f32 planes: TF32 x1 unless switched off ("tc_levels" = 1) or cta_group::1 was forced.
int first_level_terms(const Prepared &q) {
    return (q.mode == PREP_F16R || (q.mode == PREP_TF32 && t_opt.tc_levels >= 2 && t_opt.tc_cg == 2)) ? 1 : 3;
}

VerifyCtx verify_ctx(const Prepared &c, const pmm_matrix_t &raw_c, int64_t keff, int metric, int64_t index_base, cudaStream_t s) {
    VerifyCtx vc;
    vc.raw_c = raw_c;
    vc.f64 = c.f64;
    vc.c_norm = c.norm.p;
    vc.c_sq = c.sqnorm.p;
    vc.c_range = c.max_sq_ptr;
    vc.metric = metric;
    vc.index_base = index_base;
    vc.keff = keff;
    vc.s = s;
    return vc;
}

// Tensor-core path: filter -> exact re-scoring of the kept candidates -> verification (-> next level).
int topk_tc(const Prepared &q, const Prepared &c, const pmm_matrix_t &raw_q, const pmm_matrix_t &raw_c, int64_t keff,
            int metric, int64_t index_base, TopkOut o, cudaStream_t s) {
    return tc_topk_verified(verify_ctx(c, raw_c, keff, metric, index_base, s), q, raw_q, &c, first_level_terms(q), nullptr, o);
}

// k > 248 on the fused path (f32 working precision): P passes of the first-level filter with 256-entry lists; pass p+1
// admits only candidates strictly BELOW the worst one pass p kept (per-query ceilings), so the passes collect
// consecutive, disjoint ranks of the filter order - 256 P candidates per query in all.  Every pass's candidates are
// re-scored exactly; the P sorted lists are sorted as one, the best k are emitted, and the usual proof is made against
// the LAST pass's worst filter value; queries it cannot clear go to the exact SIMT path.  No Q x N slab.
constexpr int64_t TC_MAX_K_MULTIPASS = 2000;
int topk_tc_multipass(const Prepared &q, const Prepared &c, const pmm_matrix_t &raw_q, const pmm_matrix_t &raw_c, int64_t keff,
                      int metric, int64_t index_base, TopkOut o, cudaStream_t s) {
    const int kp = 256;
    const int P = (int)((keff + 8 + kp - 1) / kp);
    const int64_t Q = q.n_rows;
    const int terms = first_level_terms(q);
    VerifyCtx vc = verify_ctx(c, raw_c, keff, metric, index_base, s);
    DevBuf kept_all, exact, ceil;
    CUDA_TRY(kept_all.alloc((size_t)P * Q * kp * 8, s));
    CUDA_TRY(exact.alloc((size_t)P * Q * kp * 8, s));
    CUDA_TRY(ceil.alloc((size_t)q.rows_pad * 8, s));
    const void *q_aux = metric == PMM_METRIC_COSINE ? q.norm.p : metric == PMM_METRIC_EUCLIDEAN ? q.sqnorm.p : nullptr;
    const void *c_aux = metric == PMM_METRIC_COSINE ? vc.c_norm : metric == PMM_METRIC_EUCLIDEAN ? vc.c_sq : nullptr;
    RescoreCheck none;
    memset(&none, 0, sizeof(none));
    int rc;
    for (int p = 0; p < P; ++p) {
        uint64_t *kept_p = kept_all.as<uint64_t>() + (size_t)p * Q * kp;
        if ((rc = tc_filter(q, c, kp, metric, index_base, kept_p, s, terms, nullptr, 3, nullptr, nullptr, p > 0 ? ceil.as<uint64_t>() : nullptr)))
            return rc;
        CUDA_TRY(launch_counted("rescore", s, [&] {   // exact scores of ALL candidates of the pass (k_out = kp), as packed candidates
            return launch_rescore(kept_p, kp, raw_of(raw_q), raw_of(raw_c), (const float *)q_aux, (const float *)c_aux, metric, index_base, kp,
                                  nullptr, nullptr, exact.as<uint64_t>() + (size_t)p * Q * kp, none, s);
        }));
        if (p + 1 < P)
            CUDA_TRY(launch_counted("ceilings", s, [&] { return launch_next_ceilings(kept_p, kp, Q, q.rows_pad, ceil.as<uint64_t>(), s); }));
    }
    RescoreJob job;
    if ((rc = rescore_setup(vc, q, Q, kp, level_err(q.mode, q.mode == PREP_F16R ? 1 : terms, false, raw_q.dim), nullptr, &job))) return rc;
    CUDA_TRY(launch_counted("sort_lists", s, [&] {
        return launch_sort_lists(exact.as<uint64_t>(), P, Q * kp, kp, Q, (int)keff, metric, kept_all.as<uint64_t>() + (size_t)(P - 1) * Q * kp,
                                 o.index, o.score, o.cand, job.chk, s);
    }));
    if (!job.verify) return PMM_OK;
    return rescore_finish(vc, job, raw_q, nullptr, 0, o);
}

struct PathChoice {
    bool tc;
    bool f64;
    int mode;  // prep mode for both operands
    bool multipass = false;  // tc && k > 248: topk_tc_multipass
};
// for_topk: the top-k starts with f16-rounded planes ("tc_levels" >= 3, the default) whatever the working precision -
// the filter only selects, the exact scores come from the re-scoring in f32 or f64; raw matmul needs 3xTF32 (f32) or
// DMMA (f64) because its result IS the output.
PathChoice choose_path(int q_dtype, int c_dtype, int64_t keff, bool for_topk = true) {
    PathChoice pc;
    pc.f64 = pmm_working_dtype(q_dtype, c_dtype) == PMM_DTYPE_F64;
    pc.tc = keff <= TC_MAX_K && !t_opt.force_generic && dev_info().tc && (!pc.f64 || (for_topk && t_opt.f64_tc));
    if (!pc.tc && for_topk && !pc.f64 && keff > TC_MAX_K && keff <= TC_MAX_K_MULTIPASS && t_opt.multipass && !t_opt.force_generic && dev_info().tc) {
        pc.tc = true;
        pc.multipass = true;
    }
    pc.mode = !pc.tc ? PREP_DENSE : (q_dtype == PMM_DTYPE_F16 && c_dtype == PMM_DTYPE_F16) ? PREP_F16 : PREP_TF32;
    if (pc.mode == PREP_TF32 && for_topk && t_opt.tc_levels >= 3 && t_opt.tc_cg == 2) pc.mode = PREP_F16R;
    return pc;
}

// All pointers device-resident. `pc_corpus`: an already prepared corpus or NULL.
int dev_topk_impl(const pmm_matrix_t *dq, const pmm_matrix_t *dc, const Prepared *pc_corpus, int corpus_dtype, int64_t k,
                  int metric, int64_t index_base, TopkOut o, cudaStream_t s) {
    // dc: the raw corpus column on the device (always needed: the exact re-scoring reads it);
    // pc_corpus: optional operands already prepared from it (resident corpus handle).
    const int64_t N = pc_corpus ? pc_corpus->n_rows : dc->n_rows;
    const int64_t keff = k < N ? k : N;
    if (keff == 0) return PMM_OK;
    PathChoice pc = choose_path(dq->dtype, corpus_dtype, keff);
    // a resident corpus prepared for another path (query dtype, k > 248 or an option changed): prepare afresh from
    // the raw column, which the handle keeps on the device as well
    if (pc_corpus && (pc_corpus->mode != pc.mode || pc_corpus->f64 != pc.f64)) pc_corpus = nullptr;
    const bool want_norm = metric == PMM_METRIC_COSINE, want_sq = metric == PMM_METRIC_EUCLIDEAN;
    DevBuf err;
    CUDA_TRY(err.alloc(sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    Prepared q, c_local;
    // the tensor-core path always needs squared norms (error bound of the losslessness check)
    int rc = prepare(*dq, pc.mode, pc.f64, 4 * TC_TILE_M, want_norm, want_sq || pc.tc, err.as<int>(), s, &q);
    if (rc) return rc;
    const Prepared *c = pc_corpus;
    if (!c) {
        rc = prepare(*dc, pc.mode, pc.f64, TC_TILE_N, want_norm, want_sq, err.as<int>(), s, &c_local, pc.tc);
        if (rc) return rc;
        c = &c_local;
    }
    rc = pc.multipass ? topk_tc_multipass(q, *c, *dq, *dc, keff, metric, index_base, o, s)
         : pc.tc      ? topk_tc(q, *c, *dq, *dc, keff, metric, index_base, o, s)
                      : topk_generic(q, *c, keff, metric, index_base, o, s);
    if (rc) return rc;
    if (dq->offsets || dc->offsets) return finish_error_flag(err.as<int>(), s);
    return PMM_OK;
}

// Longest vector the raw matmul still sends through the tensor cores.  The matmul's result IS the output (no exact
// re-scoring behind it), and tcgen05 accumulates with truncation: measured against the oracle (tests/test_gpu_bound.py,
// profiles/matmul_precision_r2.md) the 3xTF32 result is off by ~1e-8 * D relative on same-sign data and by up to
// ~1e-6 |q||c| on Gaussian data at D >= 384 - beyond the stated 1e-5 * max(|x|, 0.05 |q||c|) - while exact f16 planes
// (one MMA per 16 elements, exact products) hold it up to D = 1024.  Longer vectors take the exact SIMT kernel
// (sequential FMA: bit-identical to the oracle), which raw matmul can afford: end to end the path is bound by the
// device->host copy of the Q x N result, not by the contraction.
int matmul_tc_dim_limit(int mode) {
    if (t_opt.matmul_tc_max_dim > 0) return t_opt.matmul_tc_max_dim;
    return mode == PREP_F16 ? 1024 : 256;
}
// f32 operands of the tensor-core matmul: row-scaled hi/lo f16 planes (PREP_SPLIT16; CTA-pair kernel), three kind::f16
// MMAs per 16 elements with the small terms swept first - twice the rate of the 3xTF32 split and a sixth of its
// accumulate truncations on the full-size sum (profiles/matmul_soak_r2*.json: worst |error| / tolerance 1.13 -> see
// DESIGN 4.5).  `matmul_split16 = 0` or cta_group::1 keep the 3xTF32 planes.
int matmul_f32_mode() { return (t_opt.matmul_split16 && t_opt.tc_cg == 2) ? PREP_SPLIT16 : PREP_TF32; }

int dev_matmul_impl(const pmm_matrix_t *dl, const pmm_matrix_t *dr, void *d_out, cudaStream_t s) {
    PathChoice pc = choose_path(dl->dtype, dr->dtype, 1, false);
    if (pc.tc && dl->dim > matmul_tc_dim_limit(pc.mode)) {
        pc.tc = false;
        pc.mode = PREP_DENSE;
    }
    // Very short f32 vectors: a 22-bit split is off by up to 2^-21 |q||c| per product and nothing averages out over two or
    // three elements (measured: 0.6 of the tolerance at D = 2, 3), while one FMA per element costs nothing against the
    // Q x N writes - the exact kernel serves them.  Up to 32 elements the TF32 planes (32-element K-blocks) are the
    // cheaper operands; beyond, the f16 split (64-element K-blocks) wins on rate and on accumulate truncations.
    if (pc.tc && pc.mode == PREP_TF32 && dl->dim <= t_opt.matmul_exact_max_dim) {
        pc.tc = false;
        pc.mode = PREP_DENSE;
    }
    if (pc.tc && pc.mode == PREP_TF32 && dl->dim > 32 && dl->dim <= 256) pc.mode = matmul_f32_mode();   // (resident query planes: D <= 256)
    DevBuf err;
    CUDA_TRY(err.alloc(sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    Prepared l, r;
    // f32 on the tensor cores: rows with inf / NaN elements (f16 split: also rows beyond the scalable range) are marked
    // by the prep pass and fixed up afterwards
    DevBuf nf;
    const bool track = pc.tc && (pc.mode == PREP_TF32 || pc.mode == PREP_SPLIT16);
    if (track) {
        CUDA_TRY(nf.alloc((size_t)(dl->n_rows + dr->n_rows) + 16, s));
        CUDA_TRY(cudaMemsetAsync(nf.p, 0, 16, s));   // [0], [1]: counts for left / right (the row flags are written by every row)
    }
    unsigned int *nf_count = nf.as<unsigned int>();
    unsigned char *nf_left = track ? nf.as<unsigned char>() + 16 : nullptr, *nf_right = track ? nf_left + dl->n_rows : nullptr;
    int rc = prepare(*dl, pc.mode, pc.f64, 2 * TC_TILE_M, false, false, err.as<int>(), s, &l, false, nullptr, nf_left, track ? nf_count : nullptr);
    if (rc) return rc;
    rc = prepare(*dr, pc.mode, pc.f64, TC_TILE_N, false, false, err.as<int>(), s, &r, false, nullptr, nf_right, track ? nf_count + 1 : nullptr);
    if (rc) return rc;
    const int64_t Q = dl->n_rows, N = dr->n_rows, D = dl->dim;
    if (pc.tc) {
        DevInfo &di = dev_info();
        TcArgs a;
        memset(&a, 0, sizeof(a));
        a.q_hi = l.p0.p;
        a.q_lo = l.p1.p;
        a.c_hi = r.p0.p;
        a.c_lo = r.p1.p;
        a.q_rows_pad = l.rows_pad;
        a.c_rows_pad = r.rows_pad;
        a.dim_pad = l.ld;
        a.nq = Q;
        a.n = N;
        a.f16 = (pc.mode == PREP_F16 || pc.mode == PREP_SPLIT16) ? 1 : 0;
        a.terms = pc.mode == PREP_SPLIT16 ? 2 : 3;
        if (pc.mode == PREP_SPLIT16) {
            a.q_aux = l.scale.as<float>();
            a.c_aux = r.scale.as<float>();
        }
        a.cg = t_opt.tc_cg;
        a.debug_skip = t_opt.tc_debug_skip == 8 ? 8 : 0;   // wait-cycle counters of the MMA warps (diagnostics)
        // Classic schedule: whole query tiles per unit, all units sweep the corpus tiles in the same order (a corpus tile
        // is fetched from HBM once and hit in L2 by everyone else).  When the number of query tiles leaves many units
        // idle, the flat schedule (equal shares of the tile list, unaligned sweeps: measured 15-30 % slower per tile)
        // wins anyway.
        a.sched = make_tc_schedule(Q, N, di.num_sms / a.cg, 1, a.cg);
        {
            const TcSchedule &c = a.sched;
            const int64_t classic = (int64_t)c.rounds * ((c.n_tiles + c.g - 1) / c.g) + (c.m_rem > 0 ? (c.n_tiles + c.g_rem - 1) / c.g_rem : 0);
            const TcSchedule f = make_tc_schedule_flat(Q, N, di.num_sms / a.cg, a.cg);
            const int64_t flat = ((int64_t)f.m_tiles * f.n_tiles + f.num_ctas - 1) / f.num_ctas;
            if (t_opt.matmul_flat == 1 || (t_opt.matmul_flat < 0 && flat * 13 < classic * 10)) a.sched = f;
            // between half and all of the units holding a query tile: the idle ones take over the tails of the sweeps -
            // unless the shape is write-bound and few units idle (measured at 64 of 74: D = 32 classic 0.76 / hybrid 0.86 ms,
            // D = 64 equal, D = 128 0.98 / 0.92, D = 256 1.63 / 1.58; at 47 of 74, D = 128: 1.11 / 0.74 ms)
            else if (t_opt.matmul_flat == 2 ||
                     (t_opt.matmul_flat < 0 && (D > 64 || (di.num_sms / a.cg - c.m_tiles) * 5 > di.num_sms / a.cg)))
                a.sched = make_tc_schedule_hybrid(Q, N, di.num_sms / a.cg, a.cg);
        }
        a.metric = PMM_METRIC_DOT;
        a.k = 1;
        a.kp = 32;
        a.out = (float *)d_out;
        cudaError_t e = launch_counted(pc.mode == PREP_SPLIT16 ? "tc_matmul_f16x3" : a.f16 ? "tc_matmul_f16" : "tc_matmul_tf32x3", s,
                                       [&] { return launch_tc_matmul(a, s); });
        if (e != cudaSuccess)
            return fail(PMM_ERR_CUDA, "tensor-core matmul launch failed: %s %s", cudaGetErrorString(e), tc_last_error());
        if (track)   // +-inf inputs: 0 * inf inside the split gives NaN where the reference propagates the infinity
            CUDA_TRY(launch_counted("matmul_fixup", s, [&] {
                return launch_matmul_nonfinite_fixup(raw_of(*dl), raw_of(*dr), nf_left, nf_right, nf_count, (float *)d_out, s);
            }));
    } else if (pc.f64) {
        for (int64_t q0 = 0; q0 < Q; q0 += (1 << 20)) {
            int64_t nq = Q - q0 < (1 << 20) ? Q - q0 : (1 << 20);
            const bool simt = t_opt.f64_simt != 0;
            CUDA_TRY(launch_counted(simt ? "scores_f64" : "scores_f64_dmma", s, [&] {
                return (simt ? launch_scores_f64 : launch_scores_f64_dmma)(l.p0.as<double>() + q0 * D, r.p0.as<double>(), nullptr,
                                                                           nullptr, nq, N, D, PMM_METRIC_DOT,
                                                                           (double *)d_out + q0 * N, N, s);
            }));
        }
    } else {
        for (int64_t q0 = 0; q0 < Q; q0 += (1 << 20)) {
            int64_t nq = Q - q0 < (1 << 20) ? Q - q0 : (1 << 20);
            CUDA_TRY(launch_counted("scores_f32", s, [&] {
                return launch_scores_f32(l.p0.as<float>() + q0 * D, r.p0.as<float>(), nullptr, nullptr, nq, N, D, PMM_METRIC_DOT,
                                         (float *)d_out + q0 * N, N, s);
            }));
        }
    }
    if (dl->offsets || dr->offsets) return finish_error_flag(err.as<int>(), s);
    return PMM_OK;
}

// ------------------------------------------------------------------------------------------------ host staging
cudaStream_t host_stream() {
    static thread_local cudaStream_t s = nullptr;
    static thread_local int dev_of_stream = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!s || dev_of_stream != dev) {
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        dev_of_stream = dev;
    }
    return s;
}

// A host matrix uploaded to the device. `dm` describes it with device pointers.
struct Uploaded {
    DevBuf values, offsets, validity, row_validity;
    pmm_matrix_t dm;
};

// ---- multi-chunk host columns (PMM_MATRIX_CHUNKED): a Polars Series with several chunks keeps its rows in several
// buffers; the reference's zero-copy path gives up there (`cont_slice`, src/matmul.rs:53) and copies.  Here the chunks
// are uploaded one after the other into ONE device buffer - no host-side concatenation.
bool is_chunked(const pmm_matrix_t *m) { return (m->reserved & PMM_MATRIX_CHUNKED) != 0; }
const pmm_chunks_t *chunks_of(const pmm_matrix_t *m) { return (const pmm_chunks_t *)m->values; }

int check_chunked(const pmm_matrix_t *m, const char *what) {
    if (!is_chunked(m)) return PMM_OK;
    if (m->offsets || m->validity || m->row_validity || (m->reserved & PMM_MATRIX_ON_DEVICE))
        return fail(PMM_ERR_UNSUPPORTED, "%s: a chunked column must be fixed-size rows in host memory without bitmaps", what);
    const pmm_chunks_t *ch = chunks_of(m);
    if (!ch || ch->n_chunks < 0 || (ch->n_chunks > 0 && !ch->chunks)) return fail(PMM_ERR_INVALID, "%s: bad chunk list", what);
    int64_t rows = 0;
    for (int64_t i = 0; i < ch->n_chunks; ++i) {
        if (ch->chunks[i].n_rows < 0 || (ch->chunks[i].n_rows > 0 && !ch->chunks[i].values)) return fail(PMM_ERR_INVALID, "%s: bad chunk %lld", what, (long long)i);
        rows += ch->chunks[i].n_rows;
    }
    if (rows != m->n_rows) return fail(PMM_ERR_INVALID, "%s: the chunks hold %lld rows, the column says %lld", what, (long long)rows, (long long)m->n_rows);
    return PMM_OK;
}

// Rows [r0, r1) of a fixed-size-row host column -> dst (device), on stream s; walks the chunk list when there is one.
cudaError_t copy_host_rows(const pmm_matrix_t *m, int64_t r0, int64_t r1, void *dst, cudaStream_t s) {
    const size_t row_bytes = (size_t)m->dim * esize(m->dtype);
    if (r1 <= r0) return cudaSuccess;
    if (!is_chunked(m)) return stage_h2d(dst, (const char *)m->values + (size_t)r0 * row_bytes, (size_t)(r1 - r0) * row_bytes, s);
    const pmm_chunks_t *ch = chunks_of(m);
    int64_t base = 0;
    for (int64_t i = 0; i < ch->n_chunks && base < r1; ++i) {
        const int64_t n = ch->chunks[i].n_rows, lo = std::max(r0, base), hi = std::min(r1, base + n);
        if (hi > lo) {
            cudaError_t e = stage_h2d((char *)dst + (size_t)(lo - r0) * row_bytes, (const char *)ch->chunks[i].values + (size_t)(lo - base) * row_bytes,
                                      (size_t)(hi - lo) * row_bytes, s);
            if (e != cudaSuccess) return e;
        }
        base += n;
    }
    return cudaSuccess;
}

const void *first_host_values(const pmm_matrix_t *m) {
    if (!is_chunked(m)) return m->values;
    const pmm_chunks_t *ch = chunks_of(m);
    for (int64_t i = 0; i < ch->n_chunks; ++i)
        if (ch->chunks[i].n_rows > 0) return ch->chunks[i].values;
    return nullptr;
}

int upload(const pmm_matrix_t *m, cudaStream_t s, Uploaded *u) {
    u->dm = *m;
    const int es = esize(m->dtype);
    if (is_chunked(m)) {
        const size_t vbytes = (size_t)m->n_rows * m->dim * es;
        CUDA_TRY(u->values.alloc(vbytes, s));
        CUDA_TRY(copy_host_rows(m, 0, m->n_rows, u->values.p, s));
        stat_add("h2d_bytes", (double)vbytes);
        u->dm.values = u->values.p;
        u->dm.reserved = 0;
        return PMM_OK;
    }
    int64_t first = 0, last = m->n_rows * m->dim;
    if (m->offsets) {
        first = m->offsets[0];
        last = m->offsets[m->n_rows];
        if (last < first) return fail(PMM_ERR_INVALID, "list offsets are not monotonic");
        CUDA_TRY(u->offsets.alloc((size_t)(m->n_rows + 1) * 8, s));
        CUDA_TRY(stage_h2d(u->offsets.p, m->offsets, (size_t)(m->n_rows + 1) * 8, s));
        u->dm.offsets = u->offsets.as<int64_t>();
        stat_add("h2d_bytes", (double)(m->n_rows + 1) * 8);
    }
    const size_t vbytes = (size_t)(last - first) * es;
    CUDA_TRY(u->values.alloc(vbytes, s));
    if (vbytes)
        CUDA_TRY(stage_h2d(u->values.p, (const char *)m->values + (size_t)first * es, vbytes, s));
    stat_add("h2d_bytes", (double)vbytes);
    // kernels index child values by absolute position: rebase so that position `first` is byte 0 of the buffer
    u->dm.values = (const char *)u->values.p - (size_t)first * es;
    if (m->validity) {
        const size_t nb = (size_t)((last + 7) / 8);
        CUDA_TRY(u->validity.alloc(nb, s));
        CUDA_TRY(stage_h2d(u->validity.p, m->validity, nb, s));
        u->dm.validity = u->validity.as<uint8_t>();
        stat_add("h2d_bytes", (double)nb);
    }
    if (m->row_validity) {
        const size_t nb = (size_t)((m->n_rows + 7) / 8);
        CUDA_TRY(u->row_validity.alloc(nb, s));
        CUDA_TRY(stage_h2d(u->row_validity.p, m->row_validity, nb, s));
        u->dm.row_validity = u->row_validity.as<uint8_t>();
        stat_add("h2d_bytes", (double)nb);
    }
    return PMM_OK;
}

int list_dim_check(const pmm_matrix_t *m, const char *) {
    // src/matmul.rs:237-239: a List column takes its dimension from row 0; a null first row is an error
    if (m->offsets && m->n_rows > 0 && m->row_validity && !(m->row_validity[0] & 1))
        return fail(PMM_ERR_INVALID, "First element is null");
    return PMM_OK;
}

cudaStream_t copy_stream() {
    static thread_local cudaStream_t s = nullptr;
    static thread_local int dev_of_stream = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!s || dev_of_stream != dev) {
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        dev_of_stream = dev;
    }
    return s;
}

// Host top-k on the tensor-core path with the corpus upload overlapped with compute (north_star item 5: the Arrow
// buffer feeds the H2D copies directly - through the page-locked staging ring when it is pageable).  The corpus is cut
// into row chunks; chunk i+1 is copied on a second stream while chunk i runs prep + the fused filter.  The candidate
// lists are CARRIED from launch to launch (TcCarry), so nothing is merged across chunks and only the first chunk pays
// the list warm-up; every chunk is prepared into its slice of whole-corpus plane buffers, so the re-query levels find the
// full planes afterwards; the kept candidates are re-scored once against the whole (now resident) corpus.
// Outputs: host index/score buffers and/or exact packed candidates left on the device (d_cand, for the multi-GPU
// exchange); corpus row j is reported as index_base + j.
// Sustained algorithmic FLOP/s of the first filter level on f32 planes (TF32 x1 with two levels, else 3xTF32).
double terms0_rate(int mode) { return (mode == PREP_TF32 && t_opt.tc_levels >= 2 && t_opt.tc_cg == 2) ? 7.5e14 : 2.6e14; }

int host_topk_chunked(const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t keff, int metric, PathChoice pc,
                      int64_t index_base, uint32_t *out_index, double *out_score, uint64_t *d_cand) {
    cudaStream_t s = host_stream(), cs = copy_stream();
    // Everything that owns device memory is declared first and the guard last, so that on EVERY way out (errors
    // included) both streams are drained before a buffer is released: the copy stream may still be writing into the
    // corpus buffer, and the caller's host buffers must not be in use by a DMA after we return.
    Uploaded uq, uc;
    DevBuf err, kept, c_aux_all, c_aux32_all, d_idx, d_sc, c_max, warm_buf, sample_vals;
    Prepared q, call, sample_prep;
    TcCarry carry;
    struct Drain {
        cudaStream_t s, cs;
        cudaEvent_t ready = nullptr;
        std::vector<cudaEvent_t> ev;
        ~Drain() {
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(s);
            if (ready) cudaEventDestroy(ready);
            for (cudaEvent_t e : ev)
                if (e) cudaEventDestroy(e);
        }
    } drain{s, cs};
    const int64_t Q = queries->n_rows, N = corpus->n_rows, D = corpus->dim;
    const int es = esize(corpus->dtype);
    // Chunk boundaries in rows (multiples of 256 so bitmaps can be re-based by whole bytes).  The filter of chunk i
    // runs while chunk i+1 is copied, so only the first copy is exposed: start small (N/32) and let the chunks grow
    // by the ratio of filter time to copy time per row (2 Q / tensor rate vs element size / host-link rate: about 1.6
    // at Q = 100k f32 from page-locked memory, below 1 through the staging ring), the first term of a geometric series
    // of at most eight chunks that sums to N.  Copy-bound shapes (few queries, pageable sources) get equal chunks.
    const bool copy_up_front = !stage_enabled() || host_ptr_is_pinned(first_host_values(corpus));
    std::vector<int64_t> cut{0};
    {
        // sustained filter FLOP/s, measured: f16 planes (f16 input, or f32 rounded to f16) / TF32 x1 / 3xTF32
        const double rate = (pc.mode == PREP_F16 || pc.mode == PREP_F16R) ? 1.25e15 : terms0_rate(pc.mode);
        // host -> device: ~40 GB/s from page-locked memory, ~22 GB/s through the staging ring (pageable source); 20 % margin
        const double bw = copy_up_front ? 4.0e10 : 2.2e10;
        double ratio = 0.8 * (2.0 * (double)Q / rate) / ((double)es / bw);
        if (t_opt.host_chunk_ratio_pct > 0) ratio = t_opt.host_chunk_ratio_pct / 100.0;
        if (ratio < 1.0) ratio = 1.0;
        if (ratio > 4.0) ratio = 4.0;
        const int max_chunks = 8;
        // first chunk of a geometric series of max_chunks terms that sums to N (equal chunks when ratio = 1) ...
        double first = ratio > 1.001 ? (double)N * (ratio - 1.0) / (pow(ratio, max_chunks) - 1.0) : (double)N / max_chunks;
        // ... but not below N / first_div: tiny first chunks buy nothing, the query upload comes first anyway
        const int first_div = t_opt.host_chunk_first_div > 0 ? t_opt.host_chunk_first_div : 32;
        if (first < (double)N / first_div) first = (double)N / first_div;
        int64_t size = (int64_t)first / 256 * 256;
        const int64_t min_rows = (int64_t)t_opt.host_chunk_min_rows / 256 * 256 >= 256 ? (int64_t)t_opt.host_chunk_min_rows / 256 * 256 : 256;
        if (size < min_rows) size = min_rows;
        int64_t at = 0;
        while (at + size < N && (int)cut.size() < max_chunks) {
            at += size;
            cut.push_back(at);
            size = (int64_t)((double)size * ratio) / 256 * 256;
        }
        // a small remainder joins the previous chunk
        if (cut.size() > 2 && N - cut.back() < (cut.back() - cut[cut.size() - 2]) / 4) cut.pop_back();
        cut.push_back(N);
    }
    const int n_chunks = (int)cut.size() - 1;

    int rc = PMM_OK;
    if (queries->reserved & PMM_MATRIX_ON_DEVICE) {  // pmm_topk_shard: the driver broadcast the queries over NVLink
        uq.dm = *queries;
        uq.dm.reserved = 0;
    } else if ((rc = upload(queries, s, &uq))) {
        return rc;
    }
    // corpus metadata now, values chunk by chunk
    uc.dm = *corpus;
    uc.dm.reserved = 0;
    int64_t pos0 = 0, pos1 = N * D;
    if (corpus->offsets) {
        pos0 = corpus->offsets[0];
        pos1 = corpus->offsets[N];
        if (pos1 < pos0) return fail(PMM_ERR_INVALID, "list offsets are not monotonic");
        CUDA_TRY(uc.offsets.alloc((size_t)(N + 1) * 8, s));
        CUDA_TRY(stage_h2d(uc.offsets.p, corpus->offsets, (size_t)(N + 1) * 8, s));
        uc.dm.offsets = uc.offsets.as<int64_t>();
        stat_add("h2d_bytes", (double)(N + 1) * 8);
    }
    CUDA_TRY(uc.values.alloc((size_t)(pos1 - pos0) * es, s));
    uc.dm.values = (const char *)uc.values.p - (size_t)pos0 * es;
    if (corpus->validity) {
        const size_t nb = (size_t)((pos1 + 7) / 8);
        CUDA_TRY(uc.validity.alloc(nb, s));
        CUDA_TRY(stage_h2d(uc.validity.p, corpus->validity, nb, s));
        uc.dm.validity = uc.validity.as<uint8_t>();
    }
    if (corpus->row_validity) {
        const size_t nb = (size_t)((N + 7) / 8);
        CUDA_TRY(uc.row_validity.alloc(nb, s));
        CUDA_TRY(stage_h2d(uc.row_validity.p, corpus->row_validity, nb, s));
        uc.dm.row_validity = uc.row_validity.as<uint8_t>();
    }
    // Warm seeds (see warm_seeds) need a sample of the WHOLE corpus, and only its first chunk will be on the device when
    // the first launch starts: for plain fixed-size rows a strided sample of the host buffer (every (N / sample)-th row,
    // ~12 MB at C3) goes up FIRST - ahead of the chunk copies in the DMA queue - and its pre-pass runs while the first
    // chunk is still crossing PCIe.  Other layouts (lists, bitmaps, multi-chunk columns) take a strided sample of the
    // first chunk instead (in the chunk loop).
    const bool warm_host_sample = warm_seed_applies(pc.mode, Q, N, N, keff, tc_list_capacity(keff)) && !corpus->offsets && !corpus->validity &&
                                  !corpus->row_validity && !(corpus->reserved & PMM_MATRIX_CHUNKED);
    if (warm_host_sample) {
        const int64_t rows_s = warm_rows_for(N), step = N / rows_s;
        CUDA_TRY(sample_vals.alloc((size_t)rows_s * D * es, s));
        CUDA_TRY(stage_h2d_rows(sample_vals.p, corpus->values, (size_t)D * es, (size_t)step * D * es, (size_t)rows_s, s));
        stat_add("h2d_bytes", (double)rows_s * D * es);
    }
    CUDA_TRY(cudaEventCreateWithFlags(&drain.ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(drain.ready, s));
    CUDA_TRY(cudaStreamWaitEvent(cs, drain.ready, 0));  // the copy stream may touch the buffers once they exist
    std::vector<cudaEvent_t> &ev = drain.ev;
    ev.assign(n_chunks, nullptr);
    // Chunk i of the corpus values -> its place in the device buffer, on the copy stream; ev[i] fires when it has landed.
    // Page-locked source: plain DMA, all chunks are queued up front.  Pageable source (what Polars hands over): the
    // bytes go through the page-locked staging ring (pmm_stage.h), which keeps this thread busy for the duration of
    // the host-side copy - so chunk i+1 is staged AFTER the kernels of chunk i have been enqueued and the host copy
    // overlaps the GPU's work on the previous chunk.
    auto copy_chunk = [&](int i) -> int {
        const int64_t a = corpus->offsets ? corpus->offsets[cut[i]] : cut[i] * D;
        const int64_t b = corpus->offsets ? corpus->offsets[cut[i + 1]] : cut[i + 1] * D;
        if (b > a) {
            if (corpus->offsets)
                CUDA_TRY(stage_h2d((char *)uc.values.p + (size_t)(a - pos0) * es, (const char *)corpus->values + (size_t)a * es,
                                   (size_t)(b - a) * es, cs));
            else   // fixed-size rows: one buffer, or the chunk list of a multi-chunk column
                CUDA_TRY(copy_host_rows(corpus, cut[i], cut[i + 1], (char *)uc.values.p + (size_t)a * es, cs));
        }
        stat_add("h2d_bytes", (double)(b - a) * es);
        CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(ev[i], cs));
        return PMM_OK;
    };
    if (copy_up_front)
        for (int i = 0; i < n_chunks; ++i)
            if ((rc = copy_chunk(i))) return rc;

    const bool want_norm = metric == PMM_METRIC_COSINE, want_sq = metric == PMM_METRIC_EUCLIDEAN;
    CUDA_TRY(err.alloc(sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    if (pc.f64 && d_cand) return fail(PMM_ERR_UNSUPPORTED, "packed candidates exist for f32 working precision only");
    const int64_t wsz = pc.f64 ? 8 : 4;   // working type of the norms; f64 keeps f32 copies for the filter beside them
    if ((rc = prepare(uq.dm, pc.mode, pc.f64, 4 * TC_TILE_M, want_norm, true, err.as<int>(), s, &q))) return rc;
    CUDA_TRY(c_max.alloc(2 * sizeof(unsigned int), s));
    CUDA_TRY(init_norm_range(c_max.as<unsigned int>(), s));
    const int kp = tc_list_capacity(keff);
    const int terms0 = first_level_terms(q);
    const float *warm = nullptr;
    carry.corpus_rows_total = N;
    carry.layout_rows = N;
    for (int i = 0; i < n_chunks; ++i) carry.layout_rows = std::min<int64_t>(carry.layout_rows, cut[i + 1] - cut[i]);
    CUDA_TRY(kept.alloc((size_t)Q * kp * 8, s));
    if (want_norm || want_sq) {   // chunks write their padded tails too
        CUDA_TRY(c_aux_all.alloc((size_t)round_up(N, TC_TILE_N) * wsz, s));
        if (pc.f64) CUDA_TRY(c_aux32_all.alloc((size_t)round_up(N, TC_TILE_N) * 4, s));
    }
    if (warm_host_sample) {   // the strided host sample went up before the chunks (above): its pre-pass runs while chunk 0 is in flight
        pmm_matrix_t sm;
        memset(&sm, 0, sizeof(sm));
        sm.values = sample_vals.p;
        sm.n_rows = warm_rows_for(N);
        sm.dim = D;
        sm.dtype = corpus->dtype;
        if ((rc = prepare(sm, pc.mode, pc.f64, TC_TILE_N, want_norm, want_sq, err.as<int>(), s, &sample_prep))) return rc;
        if ((rc = warm_seeds(q, sample_prep, N, keff, metric, terms0, s, &warm_buf))) return rc;
        warm = warm_buf.as<float>();
    }
    // Plane buffers for the WHOLE corpus; every chunk is prepared into its slice (chunk starts are multiples of the
    // corpus tile), so that after the last chunk the planes of the full corpus are at hand for the re-query levels.
    call.mode = pc.mode;
    call.f64 = pc.f64;
    call.n_rows = N;
    call.dim = D;
    call.rows_pad = round_up(N, TC_TILE_N);
    const int64_t plane_es = (pc.mode == PREP_F16 || pc.mode == PREP_F16R) ? 2 : 4;
    call.ld = round_up(D, plane_es == 2 ? 64 : 32);
    CUDA_TRY(call.p0.alloc(plane_bytes(pc.mode, N, D, TC_TILE_N), s));
    if (pc.mode == PREP_TF32) CUDA_TRY(call.p1.alloc(plane_bytes(pc.mode, N, D, TC_TILE_N), s));  // (f16 modes: one plane)
    for (int i = 0; i < n_chunks; ++i) {
        if (!copy_up_front && (rc = copy_chunk(i))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(s, ev[i], 0));
        const int64_t r0 = cut[i], rows = cut[i + 1] - cut[i];
        const pmm_matrix_t dm = slice_rows(uc.dm, r0, rows);  // r0 is a multiple of 256
        Prepared c;  // views into the whole-corpus buffers
        const size_t off = (size_t)r0 * call.ld * plane_es, pb = plane_bytes(pc.mode, rows, D, TC_TILE_N);
        c.p0.borrow((char *)call.p0.p + off, pb, s);
        if (pc.mode == PREP_TF32) c.p1.borrow((char *)call.p1.p + off, pb, s);
        const size_t aux_rows = (size_t)round_up(rows, TC_TILE_N);
        if (want_norm) c.norm.borrow((char *)c_aux_all.p + (size_t)r0 * wsz, aux_rows * wsz, s);
        if (want_sq) c.sqnorm.borrow((char *)c_aux_all.p + (size_t)r0 * wsz, aux_rows * wsz, s);
        if (pc.f64 && want_norm) c.norm32.borrow(c_aux32_all.as<float>() + r0, aux_rows * 4, s);
        if (pc.f64 && want_sq) c.sqnorm32.borrow(c_aux32_all.as<float>() + r0, aux_rows * 4, s);
        if ((rc = prepare(dm, pc.mode, pc.f64, TC_TILE_N, want_norm, want_sq, err.as<int>(), s, &c, true, c_max.as<unsigned int>()))) return rc;
        // warm seeds (see warm_seeds) from the first rows of the first chunk; every chunk's launch starts from them
        if (i == 0 && !warm && warm_seed_applies(q, N, rows, keff, kp)) {
            if ((rc = warm_seeds(q, c, N, keff, metric, terms0, s, &warm_buf))) return rc;
            warm = warm_buf.as<float>();
        }
        // the candidate lists are carried from chunk to chunk; the last launch merges them into `kept`
        if ((rc = tc_filter(q, c, kp, metric, index_base + r0, kept.as<uint64_t>(), s, terms0, &carry,
                            (i == 0 ? 1 : 0) | (i == n_chunks - 1 ? 2 : 0), warm, nullptr, nullptr, 0, tc_kernel_stat_name(q, terms0, kp, false))))
            return rc;
    }
    if (want_norm) call.norm.borrow(c_aux_all.p, (size_t)call.rows_pad * wsz, s);
    if (want_sq) call.sqnorm.borrow(c_aux_all.p, (size_t)call.rows_pad * wsz, s);
    if (pc.f64 && want_norm) call.norm32.borrow(c_aux32_all.p, (size_t)call.rows_pad * 4, s);
    if (pc.f64 && want_sq) call.sqnorm32.borrow(c_aux32_all.p, (size_t)call.rows_pad * 4, s);
    call.max_sq_ptr = c_max.as<unsigned int>();
    const uint64_t *kept_ptr = kept.as<uint64_t>();
    const size_t cnt = (size_t)Q * keff;
    // Page-locked result buffers (the Python shim's pooled ones) are addressable from the device: the re-scoring kernel
    // then stores its coalesced rows straight into them - the 12 bytes per result cross PCIe while the kernel runs and
    // there is no trailing copy.  Pageable buffers get device buffers and the staged copy.
    uint32_t *direct_idx = nullptr;
    double *direct_sc = nullptr;
    const bool direct = t_opt.d2h_direct && out_index && out_score && host_device_view(out_index, (void **)&direct_idx) &&
                        host_device_view(out_score, (void **)&direct_sc);
    if (!direct) {
        CUDA_TRY(d_idx.alloc(cnt * 4, s));
        CUDA_TRY(d_sc.alloc(cnt * 8, s));
    }
    TopkOut o{direct ? direct_idx : d_idx.as<uint32_t>(), direct ? direct_sc : d_sc.as<double>(), d_cand};
    // c_aux_all holds the corpus norms (cosine) or squared norms (euclidean) of the whole corpus
    VerifyCtx vc = verify_ctx(call, uc.dm, keff, metric, index_base, s);
    if ((rc = tc_topk_verified(vc, q, uq.dm, &call, terms0, kept_ptr, o, 0, nullptr, warm))) return rc;
    if (!direct && out_index) CUDA_TRY(stage_d2h(out_index, d_idx.p, cnt * 4, s));
    if (!direct && out_score) CUDA_TRY(stage_d2h(out_score, d_sc.p, cnt * 8, s));
    if (out_index || out_score) stat_add("d2h_bytes", (double)cnt * 12);
    if (queries->offsets || corpus->offsets) rc = finish_error_flag(err.as<int>(), s);
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaStreamSynchronize(cs));
    return rc;
}


// ------------------------------------------------------------------------------------------------ shard -> candidates
int check_shard_args(const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard, int64_t k, int32_t metric) {
    int rc;
    if ((rc = check_matrix(queries, "queries")) || (rc = check_matrix(corpus_shard, "corpus"))) return rc;
    if (k < 0) return fail(PMM_ERR_INVALID, "k must be non-negative");
    if (queries->n_rows == 0) return PMM_OK;
    if (metric < 0 || metric > 2) return fail(PMM_ERR_INVALID, "Unknown metric: '%d'. Supported: cosine, dot, euclidean", metric);
    if ((rc = check_pair(queries, corpus_shard))) return rc;
    const bool q_on_device = (queries->reserved & PMM_MATRIX_ON_DEVICE) != 0;
    if (q_on_device && (queries->offsets || queries->validity || queries->row_validity))
        return fail(PMM_ERR_UNSUPPORTED, "device-resident queries must be fixed-size rows without bitmaps");
    if ((!q_on_device && (rc = list_dim_check(queries, "queries"))) || (rc = list_dim_check(corpus_shard, "corpus"))) return rc;
    if ((rc = check_chunked(queries, "queries")) || (rc = check_chunked(corpus_shard, "corpus"))) return rc;
    return PMM_OK;
}

// Host shard (+ host or device-resident queries) -> this GPU's exact local top-k as packed candidates
// [Q * min(k, shard rows)] in device memory.  Synchronous.  Arguments already validated.
int shard_candidates(const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard, int64_t k, int32_t metric,
                     int64_t index_base, uint64_t *d_candidates) {
    int rc;
    const int64_t keff = k < corpus_shard->n_rows ? k : corpus_shard->n_rows;
    if (keff == 0) return PMM_OK;
    PathChoice pc = choose_path(queries->dtype, corpus_shard->dtype, keff);
    if (pc.f64) return fail(PMM_ERR_UNSUPPORTED, "packed candidates exist for f32 working precision only");
    if (pc.tc && !pc.multipass && t_opt.host_chunked &&
        (double)corpus_shard->n_rows * corpus_shard->dim * esize(corpus_shard->dtype) >= 1e6 * t_opt.host_chunk_min_mb)
        return host_topk_chunked(queries, corpus_shard, keff, metric, pc, index_base, nullptr, nullptr, d_candidates);
    cudaStream_t s = host_stream();
    Uploaded uq, uc;
    if (queries->reserved & PMM_MATRIX_ON_DEVICE) {
        uq.dm = *queries;
        uq.dm.reserved = 0;
    } else if ((rc = upload(queries, s, &uq))) {
        return rc;
    }
    if ((rc = upload(corpus_shard, s, &uc))) return rc;
    TopkOut o{nullptr, nullptr, d_candidates};
    if ((rc = dev_topk_impl(&uq.dm, &uc.dm, nullptr, corpus_shard->dtype, k, metric, index_base, o, s))) return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU groups
// SURVEY §8e / north_star item 6: the corpus partitions by rows, every GPU scans its shard for ALL queries with the
// fused kernel and emits its exact local top-k as packed candidates with GLOBAL row numbers; the GPUs then exchange
// candidates with NCCL over NVLink and merge with the same u64-max merge the single-GPU path uses between corpus
// pieces.  The exchange is an ALL-TO-ALL, not an all-gather: GPU g receives, from every GPU, only the rows of the
// queries it merges ([g Q/G, (g+1) Q/G)), so each GPU moves and merges 1/G of the candidates; the merged slices are
// either left where they are (each rank returns its query slice), written to one shared host buffer (single
// process) or broadcast back so every rank holds the full result.
//
// Two process models, one code path:
//   * single process (the plugin call, pmm_topk): ncclCommInitAll, one persistent host thread per GPU;
//   * one process per GPU (torchrun, MPI, ...): ncclCommInitRank from a unique id the host application distributes.
#define NCCL_TRY(expr)                                                                                        \
    do {                                                                                                      \
        ncclResult_t r__ = (expr);                                                                            \
        if (r__ != ncclSuccess) return fail(PMM_ERR_CUDA, "NCCL error: %s (%s)", nccl->GetErrorString(r__), #expr); \
    } while (0)

// A persistent host thread bound to one device.
class DeviceWorker {
  public:
    explicit DeviceWorker(int device) : device_(device), th_([this] { loop(); }) {}
    ~DeviceWorker() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = std::move(job);
            has_job_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return done_; });
    }

  private:
    void loop() {
        cudaSetDevice(device_);
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return has_job_ || quit_; });
                if (quit_) return;
                job = std::move(job_);
                has_job_ = false;
            }
            job();
            {
                std::lock_guard<std::mutex> lk(mu_);
                done_ = true;
            }
            cv_.notify_all();
        }
    }
    int device_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<void()> job_;
    bool has_job_ = false, done_ = true, quit_ = false;
    std::thread th_;   // last member: starts after everything above is initialised
};

}  // namespace

struct pmm_group {
    int world = 1;                        // ranks in the group
    struct Member {
        int device = 0, rank = 0;
        ncclComm_t comm = nullptr;
        std::unique_ptr<DeviceWorker> worker;   // single-process groups only
    };
    std::vector<Member> members;          // the ranks THIS process drives: all of them, or exactly one
    bool local = false;                   // single-process group (ncclCommInitAll)
    std::mutex mu;                        // one collective at a time
    // Rendezvous of the member threads of a single-process group, carrying an error status.  Besides agreeing on
    // errors it keeps CUDA calls that may synchronise the whole process (first cudaHostAlloc of a staging ring, pool
    // growth) out of the window in which another GPU already spins inside an NCCL kernel waiting for this one.
    struct HostBarrier {
        std::mutex mu;
        std::condition_variable cv;
        int count = 0, gen = 0, acc = 0, last = 0;
        int arrive(int n, int status) {
            std::unique_lock<std::mutex> lk(mu);
            acc |= status;
            const int my_gen = gen;
            if (++count == n) {
                count = 0;
                last = acc;
                acc = 0;
                ++gen;
                cv.notify_all();
                return last;
            }
            cv.wait(lk, [&] { return gen != my_gen; });
            return last;
        }
    } barrier;
};

namespace {

// Contiguous split of n rows over `parts`, boundaries rounded up to `align` rows (bitmaps slice by whole bytes).
void split_rows(int64_t n, int parts, int64_t align, std::vector<int64_t> *cut) {
    int64_t per = parts > 0 ? (n + parts - 1) / parts : n;
    per = (per + align - 1) / align * align;
    if (per < align) per = align;
    cut->assign(parts + 1, 0);
    for (int g = 0; g <= parts; ++g) (*cut)[g] = std::min<int64_t>(n, (int64_t)g * per);
}

constexpr int GROUP_OUT_HOST_SLICE = 1;    // out_* are host buffers of the FULL result; this rank writes only its query slice
constexpr int GROUP_OUT_DEVICE_FULL = 2;   // out_* are device buffers [Q x k]; every rank receives the full result
constexpr int GROUP_OUT_DEVICE_SLICE = 3;  // out_* are device buffers [Q x k]; this rank writes only its query slice
constexpr int GROUP_OUT_HOST_FULL = 4;     // out_* are host buffers; every rank reads the full result back

struct GroupTopkArgs {
    const pmm_matrix_t *queries = nullptr;   // host descriptor (replicated), or device-resident (PMM_MATRIX_ON_DEVICE)
    const pmm_matrix_t *shard = nullptr;     // this rank's corpus rows: host descriptor, or device-resident (reserved flag)
    bool queries_from_root = false;          // host queries: only rank 0 uploads, the others receive them over NVLink
    int64_t index_base = 0, n_total = 0, k = 0;
    int metric = 0;
    int out_mode = GROUP_OUT_HOST_SLICE;
    uint32_t *out_index = nullptr;
    double *out_score = nullptr;
};

// One rank's part of a sharded top-k.  Collective: every rank of the group runs it with the same Q, k, metric.
// Runs on the thread that owns the rank's device (current device = member.device), options already in t_opt.
// Structure: phases that may fail on their own (CUDA_TRY inside a lambda) separated by rendezvous points that every
// rank reaches whatever happened before, so that no rank is ever left alone inside a collective.
int group_rank_topk(pmm_group *g, pmm_group::Member &m, const GroupTopkArgs &a) {
    const NcclApi *nccl = nccl_api();
    const int G = g->world, me = m.rank;
    if (!nccl && G > 1) return fail(PMM_ERR_UNSUPPORTED, "%s", nccl_load_error());
    std::lock_guard<std::mutex> device_lock(device_mutex());
    cudaStream_t s = host_stream();
    const int64_t Q = a.queries->n_rows, D = a.queries->dim;
    const int64_t keff = a.k < a.n_total ? a.k : a.n_total;
    if (Q == 0 || keff == 0) return PMM_OK;
    const int64_t n_local = a.shard->n_rows;
    const int64_t k_local = keff < n_local ? keff : n_local;
    const bool profile = t_opt.profile != 0;
    // CUDA-event bracket around a collective (statistics "<name>_ms"), like launch_counted for kernels
    struct Bracket {
        const char *name;
        cudaStream_t s;
        bool on;
        cudaEvent_t a = nullptr, b = nullptr;
        Bracket(const char *n, cudaStream_t st, bool enabled) : name(n), s(st), on(enabled) {
            if (on) {
                cudaEventCreate(&a);
                cudaEventCreate(&b);
                cudaEventRecord(a, s);
            }
        }
        ~Bracket() {
            if (!on) return;
            cudaEventRecord(b, s);
            std::lock_guard<std::mutex> lk(g_stat_mu);
            g_pending.push_back({name, a, b});
            g_stats[std::string(name) + "_launches"] += 1;
        }
    };
    // rendezvous with error agreement: host barrier inside one process, a 4-byte all-reduce between processes
    DevBuf st;
    auto agree = [&](int rc_mine, const char *what) -> int {
        int any = rc_mine ? 1 : 0;
        if (g->local) {
            any = g->barrier.arrive(G, any);
        } else if (st.p) {
            int h = any;
            CUDA_TRY(cudaMemcpyAsync(st.p, &h, sizeof(int), cudaMemcpyHostToDevice, s));
            NCCL_TRY(nccl->AllReduce(st.p, st.p, 1, ncclInt32, ncclMax, m.comm, s));
            CUDA_TRY(cudaMemcpyAsync(&h, st.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            any = h;
        }
        if (rc_mine) return rc_mine;
        if (any) return fail(PMM_ERR_CUDA, "another GPU of the group failed (%s)", what);
        return PMM_OK;
    };
    if (G > 1 && !g->local) CUDA_TRY(st.alloc(sizeof(int), s));   // (nothing collective has started yet: a plain return is safe)

    // ---- phase 1: the replicated queries.  Dense host queries cross ONE host link (rank 0) and reach the other GPUs
    //      over NVLink; anything else (lists, bitmaps, already on the device) is taken as it is by every rank.
    DevBuf d_q;
    pmm_matrix_t q_desc = *a.queries;
    const bool bcast = G > 1 && a.queries_from_root && !(a.queries->reserved & PMM_MATRIX_ON_DEVICE) && !a.queries->offsets &&
                       !a.queries->validity && !a.queries->row_validity;
    int rc = PMM_OK;
    if (bcast) {
        const size_t qbytes = (size_t)Q * D * esize(a.queries->dtype);
        rc = [&]() -> int {
            CUDA_TRY(d_q.alloc(qbytes, s));
            if (me == 0) {
                CUDA_TRY(stage_h2d(d_q.p, a.queries->values, qbytes, s));
                stat_add("h2d_bytes", (double)qbytes);
            }
            CUDA_TRY(cudaStreamSynchronize(s));
            return PMM_OK;
        }();
        if ((rc = agree(rc, "query upload"))) return rc;
        {
            Bracket br("group_broadcast", s, profile);
            NCCL_TRY(nccl->Broadcast(d_q.p, d_q.p, qbytes, ncclUint8, 0, m.comm, s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        q_desc.values = d_q.p;
        q_desc.reserved = PMM_MATRIX_ON_DEVICE;
    }

    // ---- phase 2: local exact top-k of the shard -> packed candidates [Q x keff] (zero = empty slot), and the
    //      buffers of the exchange (allocated before the rendezvous: see HostBarrier)
    std::vector<int64_t> qcut;
    split_rows(Q, G, 1, &qcut);
    const int64_t q0 = qcut[me], qn = qcut[me + 1] - qcut[me];
    const bool full_dev = a.out_mode == GROUP_OUT_DEVICE_FULL, full_host = a.out_mode == GROUP_OUT_HOST_FULL;
    const bool to_host = a.out_mode == GROUP_OUT_HOST_SLICE || full_host;
    DevBuf cand, cand_local, recv, oi, os, full_i, full_s;
    uint32_t *mi = nullptr;   // where this rank's merged slice goes (device)
    double *ms = nullptr;
    uint32_t *fi = nullptr;   // the full [Q x keff] result on the device, when one is assembled
    double *fs = nullptr;
    rc = [&]() -> int {
        CUDA_TRY(cand.alloc((size_t)Q * keff * 8, s));
        uint64_t *local_ptr = cand.as<uint64_t>();
        if (k_local < keff) {
            CUDA_TRY(cudaMemsetAsync(cand.p, 0, (size_t)Q * keff * 8, s));
            if (k_local > 0) {
                CUDA_TRY(cand_local.alloc((size_t)Q * k_local * 8, s));
                local_ptr = cand_local.as<uint64_t>();
            }
        }
        if (k_local > 0) {
            int r2;
            if (a.shard->reserved & PMM_MATRIX_ON_DEVICE) {
                if (!(q_desc.reserved & PMM_MATRIX_ON_DEVICE)) return fail(PMM_ERR_INVALID, "a device-resident shard needs device-resident queries");
                pmm_matrix_t dq = q_desc, dc = *a.shard;
                dq.reserved = 0;
                dc.reserved = 0;
                TopkOut o{nullptr, nullptr, local_ptr};
                if ((r2 = dev_topk_impl(&dq, &dc, nullptr, dc.dtype, k_local, a.metric, a.index_base, o, s))) return r2;
            } else {
                CUDA_TRY(cudaStreamSynchronize(s));   // cand is zeroed before the shard path (its own stream syncs) fills it
                if ((r2 = shard_candidates(&q_desc, a.shard, k_local, a.metric, a.index_base, local_ptr))) return r2;
            }
            if (local_ptr != cand.as<uint64_t>())
                CUDA_TRY(cudaMemcpy2DAsync(cand.p, (size_t)keff * 8, local_ptr, (size_t)k_local * 8, (size_t)k_local * 8, (size_t)Q,
                                           cudaMemcpyDeviceToDevice, s));
        }
        const int64_t qn1 = qn > 0 ? qn : 1;
        if (G > 1) CUDA_TRY(recv.alloc((size_t)G * qn1 * keff * 8, s));
        if (full_dev) {
            fi = a.out_index;
            fs = a.out_score;
        } else if (full_host && G > 1) {
            CUDA_TRY(full_i.alloc((size_t)Q * keff * 4, s));
            CUDA_TRY(full_s.alloc((size_t)Q * keff * 8, s));
            fi = full_i.as<uint32_t>();
            fs = full_s.as<double>();
        }
        if (fi) {
            mi = fi + q0 * keff;
            ms = fs + q0 * keff;
        } else if (a.out_mode == GROUP_OUT_DEVICE_SLICE) {
            mi = a.out_index + q0 * keff;
            ms = a.out_score + q0 * keff;
        } else {
            CUDA_TRY(oi.alloc((size_t)qn1 * keff * 4, s));
            CUDA_TRY(os.alloc((size_t)qn1 * keff * 8, s));
            mi = oi.as<uint32_t>();
            ms = os.as<double>();
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        return PMM_OK;
    }();
    if (G > 1 && (rc = agree(rc, "local top-k"))) return rc;
    if (rc) return rc;

    // ---- phase 3: all-to-all of candidates (rank r merges the queries [qcut[r], qcut[r+1])), merge of my slice
    const uint64_t *lists = cand.as<uint64_t>() + q0 * keff;
    int64_t list_stride = 0;
    if (G > 1) {
        Bracket br("group_exchange", s, profile);
        NCCL_TRY(nccl->GroupStart());
        for (int r = 0; r < G; ++r) {
            const int64_t rn = qcut[r + 1] - qcut[r];
            if (rn > 0) NCCL_TRY(nccl->Send(cand.as<uint64_t>() + qcut[r] * keff, (size_t)rn * keff, ncclUint64, r, m.comm, s));
            if (qn > 0) NCCL_TRY(nccl->Recv(recv.as<uint64_t>() + (int64_t)r * qn * keff, (size_t)qn * keff, ncclUint64, r, m.comm, s));
        }
        NCCL_TRY(nccl->GroupEnd());
        lists = recv.as<uint64_t>();
        list_stride = qn * keff;
    }
    if (qn > 0)
        CUDA_TRY(launch_counted("group_merge", s, [&] {
            return launch_merge_regular(lists, G, list_stride, keff, qn, (int)keff, (int)keff, a.metric != PMM_METRIC_EUCLIDEAN, mi, ms,
                                        nullptr, s);
        }));

    // ---- phase 4: results
    if (fi && G > 1) {  // every rank ends up with the whole [Q x keff] result: one broadcast per slice
        Bracket br("group_gather", s, profile);
        NCCL_TRY(nccl->GroupStart());
        for (int r = 0; r < G; ++r) {
            const int64_t rn = qcut[r + 1] - qcut[r];
            if (rn <= 0) continue;
            NCCL_TRY(nccl->Broadcast(fi + qcut[r] * keff, fi + qcut[r] * keff, (size_t)rn * keff, ncclUint32, r, m.comm, s));
            NCCL_TRY(nccl->Broadcast(fs + qcut[r] * keff, fs + qcut[r] * keff, (size_t)rn * keff, ncclFloat64, r, m.comm, s));
        }
        NCCL_TRY(nccl->GroupEnd());
    }
    if (to_host) {
        const bool whole = full_host;   // G == 1: q0 = 0 and qn = Q, so the "slice" is everything
        const int64_t r0 = whole ? 0 : q0, rn = whole ? Q : qn;
        const uint32_t *src_i = whole && fi ? fi : mi;
        const double *src_s = whole && fs ? fs : ms;
        if (rn > 0) {
            CUDA_TRY(stage_d2h(a.out_index + r0 * keff, src_i, (size_t)rn * keff * 4, s));
            CUDA_TRY(stage_d2h(a.out_score + r0 * keff, src_s, (size_t)rn * keff * 8, s));
            stat_add("d2h_bytes", (double)rn * keff * 12);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

// ---- the process-wide single-process group behind pmm_topk / pmm_matmul -----------------------------------------------
std::mutex g_local_group_mu;
pmm_group *g_local_group = nullptr;
bool g_local_group_failed = false;

int create_local_group(int n_devices, pmm_group **out) {
    const NcclApi *nccl = nccl_api();
    if (!nccl) return fail(PMM_ERR_UNSUPPORTED, "%s", nccl_load_error());
    int have = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    if (n_devices <= 0 || n_devices > have) n_devices = have;
    if (n_devices > 16) n_devices = 16;
    std::unique_ptr<pmm_group> g(new pmm_group());
    g->world = n_devices;
    g->local = true;
    std::vector<int> devs(n_devices);
    std::vector<ncclComm_t> comms(n_devices, nullptr);
    for (int i = 0; i < n_devices; ++i) devs[i] = i;
    int cur = 0;
    cudaGetDevice(&cur);
    NCCL_TRY(nccl->CommInitAll(comms.data(), n_devices, devs.data()));
    cudaSetDevice(cur);
    g->members.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        g->members[i].device = i;
        g->members[i].rank = i;
        g->members[i].comm = comms[i];
        g->members[i].worker.reset(new DeviceWorker(i));
    }
    *out = g.release();
    return PMM_OK;
}

// NULL when the box has one GPU, NCCL is missing or the communicator could not be created (the caller then stays on
// one GPU; the reason is in pmm_get_stat-independent g_err of the creating thread only, which is fine: it is a fallback
// to the path every call used to take, not to a CPU).
pmm_group *local_group() {
    std::lock_guard<std::mutex> lk(g_local_group_mu);
    if (g_local_group || g_local_group_failed) return g_local_group;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 2) {
        cudaGetLastError();
        g_local_group_failed = true;
        return nullptr;
    }
    if (create_local_group(n, &g_local_group) != PMM_OK) {
        g_local_group = nullptr;
        g_local_group_failed = true;
    }
    return g_local_group;
}

// Runs fn(member index) on every member's worker thread with the caller's option snapshot; returns the first error.
int run_on_members(pmm_group *g, const std::function<int(int)> &fn) {
    const Options opt = t_opt;
    const int n = (int)g->members.size();
    std::vector<int> rcs(n, PMM_OK);
    std::vector<std::string> errs(n);
    for (int i = 0; i < n; ++i)
        g->members[i].worker->post([&, i] {
            t_opt = opt;
            rcs[i] = fn(i);
            if (rcs[i]) errs[i] = g_err;
        });
    for (int i = 0; i < n; ++i) g->members[i].worker->wait();
    for (int i = 0; i < n; ++i)
        if (rcs[i]) {
            g_err = errs[i];
            return rcs[i];
        }
    return PMM_OK;
}

// Whole-corpus host call spread over the GPUs of a single-process group: corpus rows sharded (boundaries multiples of
// 256 rows), queries uploaded once and broadcast, every GPU writes its slice of the result into the caller's buffers.
int group_host_topk(pmm_group *g, const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t k, int metric, uint32_t *out_index,
                    double *out_score) {
    std::lock_guard<std::mutex> lk(g->mu);
    std::vector<int64_t> cut;
    split_rows(corpus->n_rows, g->world, 256, &cut);
    std::vector<pmm_matrix_t> shards(g->world);
    for (int r = 0; r < g->world; ++r) shards[r] = slice_rows(*corpus, cut[r], cut[r + 1] - cut[r]);
    return run_on_members(g, [&](int i) {
        GroupTopkArgs a;
        a.queries = queries;
        a.shard = &shards[i];
        a.queries_from_root = true;
        a.index_base = cut[i];
        a.n_total = corpus->n_rows;
        a.k = k;
        a.metric = metric;
        a.out_mode = GROUP_OUT_HOST_SLICE;
        a.out_index = out_index;
        a.out_score = out_score;
        return group_rank_topk(g, g->members[i], a);
    });
}

// Raw matmul over a single-process group: output ROWS (left rows) are sharded, no collective is needed:
// every GPU uploads its slice of `left` and the whole of `right`, and writes its slab of the result through its own
// host link - the device->host copy of the result, which dominates this path, runs on all links at once.
int dev_matmul_impl(const pmm_matrix_t *dl, const pmm_matrix_t *dr, void *d_out, cudaStream_t s);
int group_host_matmul(pmm_group *g, const pmm_matrix_t *left, const pmm_matrix_t *right, void *out) {
    std::lock_guard<std::mutex> lk(g->mu);
    std::vector<int64_t> cut;
    split_rows(left->n_rows, g->world, 256, &cut);
    const int wd = pmm_working_dtype(left->dtype, right->dtype);
    const size_t osz = wd == PMM_DTYPE_F64 ? 8 : 4;
    return run_on_members(g, [&](int i) -> int {
        const int64_t rows = cut[i + 1] - cut[i];
        if (rows <= 0) return PMM_OK;
        std::lock_guard<std::mutex> device_lock(device_mutex());
        cudaStream_t s = host_stream();
        const pmm_matrix_t part = slice_rows(*left, cut[i], rows);
        Uploaded ul, ur;
        int rc;
        if ((rc = upload(&part, s, &ul)) || (rc = upload(right, s, &ur))) return rc;
        DevBuf d_out;
        const size_t bytes = (size_t)rows * right->n_rows * osz;
        CUDA_TRY(d_out.alloc(bytes, s));
        if ((rc = dev_matmul_impl(&ul.dm, &ur.dm, d_out.p, s))) return rc;
        CUDA_TRY(stage_d2h((char *)out + (size_t)cut[i] * right->n_rows * osz, d_out.p, bytes, s));
        stat_add("d2h_bytes", (double)bytes);
        CUDA_TRY(cudaStreamSynchronize(s));
        return PMM_OK;
    });
}

// Is this call worth spreading over the box?  (GFLOP of contraction work; the multi-GPU path costs ~1 ms of
// dispatch, broadcast and exchange.)
bool wants_multi_gpu(double q_rows, double c_rows, double dim, double out_bytes) {
    if (!t_opt.multi_gpu) return false;
    const double gflop = 2.0 * q_rows * c_rows * dim / 1e9;
    return gflop >= (double)t_opt.multi_gpu_min_gflop || out_bytes >= 1.0e9;
}

}  // namespace

struct pmm_corpus {
    Uploaded raw;   // the raw column stays on the device: the exact re-scoring reads candidate rows from it
    Prepared prep;
    int device = 0;
    int storage_dtype = PMM_DTYPE_F32;
    int query_dtype = PMM_DTYPE_F32;
    cudaStream_t stream = nullptr;
};

// ================================================================================================ C ABI
extern "C" {

const char *pmm_last_error(void) { return g_err.c_str(); }
const char *pmm_version(void) { return "0.1.4+b200.r2"; }

int pmm_metric_from_str(const char *name, int32_t *metric) {
    if (!name || !metric) return fail(PMM_ERR_INVALID, "Unknown metric: ''. Supported: cosine, dot, euclidean");
    std::string low(name);
    for (auto &ch : low) ch = (char)tolower((unsigned char)ch);
    if (low == "cosine") { *metric = PMM_METRIC_COSINE; return PMM_OK; }
    if (low == "dot") { *metric = PMM_METRIC_DOT; return PMM_OK; }
    if (low == "euclidean" || low == "l2") { *metric = PMM_METRIC_EUCLIDEAN; return PMM_OK; }
    return fail(PMM_ERR_INVALID, "Unknown metric: '%s'. Supported: cosine, dot, euclidean", name);
}

int pmm_higher_is_better(int32_t metric) { return metric != PMM_METRIC_EUCLIDEAN; }

int pmm_working_dtype(int32_t l, int32_t r) {
    const bool lf = (l == PMM_DTYPE_F32 || l == PMM_DTYPE_F16), rf = (r == PMM_DTYPE_F32 || r == PMM_DTYPE_F16);
    return (lf && rf) ? PMM_DTYPE_F32 : PMM_DTYPE_F64;
}

int pmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pmm_set_device(int32_t device) {
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    return PMM_OK;
}

int pmm_host_alloc(int64_t bytes, void **out) {
    if (!out || bytes < 0) return fail(PMM_ERR_INVALID, "pmm_host_alloc: bad arguments");
    *out = nullptr;
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaHostAlloc(out, bytes > 0 ? (size_t)bytes : 16, cudaHostAllocPortable));
    return PMM_OK;
}

int pmm_host_free(void *p) {
    if (!p) return PMM_OK;
    CUDA_TRY(cudaFreeHost(p));
    return PMM_OK;
}

void *pmm_thread_stream(void) {
    if (ensure_device()) return nullptr;
    return (void *)host_stream();
}

int64_t pmm_kernel_launch_count(void) { return g_launches.load(); }
void pmm_reset_kernel_launch_count(void) { g_launches.store(0); }

// Options that act immediately instead of being part of the per-call snapshot. Returns 1 when handled.
static int immediate_option(const std::string &k, int64_t value) {
    if (k == "release_workspace") {  // this thread's parked device blocks and staging ring go back to the system
        g_block_cache.clear();
        stage_release_thread_ring();
        return 1;
    }
    if (k == "workspace_cache_mb") { g_block_cache_cap_mb.store(value < 0 ? 0 : value); return 1; }
    if (k == "prep_fast") { prep_set_fast(value != 0); return 1; }
    if (k == "rescore_fixed") { rescore_set_fixed(value != 0); return 1; }
    if (k == "stage") { stage_set_enabled(value != 0); return 1; }
    if (k == "stage_threads") { stage_set_threads((int)value); return 1; }
    if (k == "stage_nt") { stage_set_nt_stores((int)value); return 1; }
    if (k == "f64_dmma_async") { dmma_set_async((int)value); return 1; }
    if (k == "stage_slot_mb") { stage_set_ring((size_t)(value < 1 ? 1 : value) << 20, g_stage_slots.load()); g_stage_slot_mb.store(value < 1 ? 1 : value); return 1; }
    if (k == "stage_slots") { g_stage_slots.store((int)value); stage_set_ring((size_t)g_stage_slot_mb.load() << 20, (int)value); return 1; }
    return 0;
}

static int check_diag_option(const std::string &k, int64_t value) {
#ifndef PMM_DIAG
    if (k == "tc_debug_skip" && value >= 1 && value <= 3)
        return fail(PMM_ERR_UNSUPPORTED,
                    "tc_debug_skip=%lld returns wrong results by design (kernel timing experiments) and is compiled out of this "
                    "build; rebuild with -DPMM_DIAG", (long long)value);
#endif
    (void)k;
    (void)value;
    return PMM_OK;
}

int pmm_set_option(const char *key, int64_t value) {
    if (!key) return fail(PMM_ERR_INVALID, "null option key");
    const std::string k(key);
    if (immediate_option(k, value)) return PMM_OK;
    int rc = check_diag_option(k, value);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_opt_mu);
    if (!apply_option(g_opt, k, value)) return fail(PMM_ERR_INVALID, "unknown option '%s'", key);
    return PMM_OK;
}

int pmm_set_thread_option(const char *key, int64_t value) {
    if (!key) {  // NULL key: drop all overrides of the calling thread
        t_overrides.clear();
        return PMM_OK;
    }
    const std::string k(key);
    int rc = check_diag_option(k, value);
    if (rc) return rc;
    Options probe;
    if (!apply_option(probe, k, value)) return fail(PMM_ERR_INVALID, "unknown per-thread option '%s'", key);
    for (auto &kv : t_overrides)
        if (kv.first == k) {
            kv.second = value;
            return PMM_OK;
        }
    t_overrides.emplace_back(k, value);
    return PMM_OK;
}

double pmm_get_stat(const char *name) {
    if (!name) return 0.0;
    if (!strncmp(name, "tc_dbg_wait", 11)) {  // tc_dbg_wait0..3: diagnostics of the last launches with tc_debug_skip = 8
        static unsigned long long last[52];
        const int i = atoi(name + 11);
        if (i == 0) tc_debug_wait_cycles(last);
        return i >= 0 && i < 52 ? (double)last[i] : 0.0;
    }
    std::lock_guard<std::mutex> lk(g_stat_mu);
    resolve_pending_locked();
    {
        double h2d = 0, d2h = 0;
        stage_take_counters(&h2d, &d2h);
        g_stats["staged_h2d_bytes"] += h2d;
        g_stats["staged_d2h_bytes"] += d2h;
    }
    auto it = g_stats.find(name);
    return it == g_stats.end() ? 0.0 : it->second;
}

void pmm_reset_stats(void) {
    std::lock_guard<std::mutex> lk(g_stat_mu);
    resolve_pending_locked();
    stage_take_counters(nullptr, nullptr);
    g_stats.clear();
}

// ---------------------------------------------------------------------------------------------- device entry points
int pmm_dev_topk(const pmm_matrix_t *dq, const pmm_matrix_t *dc, int64_t k, int32_t metric, int64_t index_base,
                 uint32_t *d_index, double *d_score, uint64_t *d_candidates, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if ((rc = check_matrix(dq, "queries")) || (rc = check_matrix(dc, "corpus"))) return rc;
    if (dq->n_rows == 0) return PMM_OK;
    if (metric < 0 || metric > 2) return fail(PMM_ERR_INVALID, "Unknown metric: '%d'. Supported: cosine, dot, euclidean", metric);
    if (k < 0) return fail(PMM_ERR_INVALID, "k must be non-negative");
    if ((rc = check_pair(dq, dc))) return rc;
    TopkOut o{d_index, d_score, d_candidates};
    return dev_topk_impl(dq, dc, nullptr, dc->dtype, k, metric, index_base, o, (cudaStream_t)stream);
}

int pmm_dev_merge_candidates(const uint64_t *d_lists, int64_t n_lists, int64_t n_queries, int64_t k_in, int64_t k_out,
                             int32_t metric, uint32_t *d_index, double *d_score, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if (n_lists < 1 || k_in < 1 || k_in > 256 || k_out < 0 || k_out > k_in)
        return fail(PMM_ERR_INVALID, "merge: need 1 <= k_out <= k_in <= 256 and at least one list");
    if (n_queries == 0 || k_out == 0) return PMM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(launch_counted("merge", s, [&] {
        return launch_merge_regular(d_lists, n_lists, n_queries * k_in, k_in, n_queries, (int)k_in, (int)k_out,
                                    metric != PMM_METRIC_EUCLIDEAN, d_index, d_score, nullptr, s);
    }));
    return PMM_OK;
}

int pmm_dev_matmul(const pmm_matrix_t *dl, const pmm_matrix_t *dr, void *d_out, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if ((rc = check_matrix(dl, "left")) || (rc = check_matrix(dr, "right"))) return rc;
    if (dl->n_rows == 0) return PMM_OK;
    if ((rc = check_pair(dl, dr))) return rc;
    return dev_matmul_impl(dl, dr, d_out, (cudaStream_t)stream);
}

int pmm_dev_norms(const pmm_matrix_t *dx, int32_t squared, void *d_out, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if ((rc = check_matrix(dx, "matrix"))) return rc;
    if (dx->n_rows == 0) return PMM_OK;
    PrepArgs a;
    memset(&a, 0, sizeof(a));
    a.values = dx->values;
    a.offsets = dx->offsets;
    a.validity = dx->validity;
    a.row_validity = dx->row_validity;
    a.n_rows = dx->n_rows;
    a.dim = dx->dim;
    a.rows_out = dx->n_rows;
    if (squared) a.sqnorm_out = d_out;
    else a.norm_out = d_out;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(launch_counted("norms", s, [&] { return launch_norms(a, dx->dtype, s); }));
    return PMM_OK;
}

// ---------------------------------------------------------------------------------------------- diagnostics
// Mode and MMA terms of a filter level for the given storage dtypes: level 0 = the default first level, 1 = TF32 x1,
// 3 = 3xTF32.
static void filter_level_mode(int level, int q_dtype, int c_dtype, int *mode, int *terms) {
    const bool both_f16 = q_dtype == PMM_DTYPE_F16 && c_dtype == PMM_DTYPE_F16;
    if (level == 0) {
        *mode = both_f16 ? PREP_F16 : PREP_F16R;
        *terms = 1;
    } else {
        *mode = PREP_TF32;
        *terms = level == 1 ? 1 : 3;
    }
}

int pmm_dev_filter_candidates(const pmm_matrix_t *dq, const pmm_matrix_t *dc, int32_t metric, int32_t level, int32_t kp,
                              int64_t index_base, uint64_t *d_kept, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if ((rc = check_matrix(dq, "queries")) || (rc = check_matrix(dc, "corpus"))) return rc;
    if (metric < 0 || metric > 2) return fail(PMM_ERR_INVALID, "Unknown metric: '%d'. Supported: cosine, dot, euclidean", metric);
    if ((rc = check_pair(dq, dc))) return rc;
    if (!(kp == 32 || kp == 64 || kp == 128 || kp == 256)) return fail(PMM_ERR_INVALID, "kp must be 32, 64, 128 or 256");
    if (!(level == 0 || level == 1 || level == 3)) return fail(PMM_ERR_INVALID, "level must be 0, 1 or 3");
    if (!dev_info().tc) return fail(PMM_ERR_UNSUPPORTED, "the tensor-core filter needs an sm_100 device");
    cudaStream_t s = (cudaStream_t)stream;
    const bool f64 = pmm_working_dtype(dq->dtype, dc->dtype) == PMM_DTYPE_F64;
    int mode, terms;
    filter_level_mode(level, dq->dtype, dc->dtype, &mode, &terms);
    DevBuf err;
    CUDA_TRY(err.alloc(sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    Prepared q, c;
    if ((rc = prepare(*dq, mode, f64, 4 * TC_TILE_M, true, true, err.as<int>(), s, &q))) return rc;
    if ((rc = prepare(*dc, mode, f64, TC_TILE_N, true, true, err.as<int>(), s, &c, true))) return rc;
    if ((rc = tc_filter(q, c, kp, metric, index_base, d_kept, s, terms))) return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

int pmm_filter_error_bound(int32_t level, int32_t q_dtype, int32_t c_dtype, int64_t dim, int32_t metric, float q_norm, float c_norm_max,
                           float c_norm_min, float *bound, float *max_norm) {
    if (!bound) return fail(PMM_ERR_INVALID, "null output");
    int mode, terms;
    filter_level_mode(level, q_dtype, c_dtype, &mode, &terms);
    const LevelErr le = level_err(mode, terms, pmm_working_dtype(q_dtype, c_dtype) == PMM_DTYPE_F64, dim);
    *bound = filter_error_bound(le.eps, le.abs_err, metric, q_norm, c_norm_max, c_norm_min);
    if (max_norm) *max_norm = le.max_norm;
    return PMM_OK;
}

// ---------------------------------------------------------------------------------------------- host entry points
int pmm_topk(const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t k, const char *metric, uint32_t *out_index,
             double *out_score, int64_t *k_actual) {
    int rc;
    if ((rc = check_matrix(queries, "queries")) || (rc = check_matrix(corpus, "corpus"))) return rc;
    if (k < 0) return fail(PMM_ERR_INVALID, "k must be non-negative (can't convert negative int to unsigned)");
    const int64_t keff = k < corpus->n_rows ? k : corpus->n_rows;
    if (k_actual) *k_actual = keff;
    if (queries->n_rows == 0) return PMM_OK;  // before the metric is parsed, src/matmul.rs:480-490
    int32_t m;
    if ((rc = pmm_metric_from_str(metric, &m))) return rc;
    if ((rc = check_pair(queries, corpus))) return rc;
    if ((rc = list_dim_check(queries, "queries")) || (rc = list_dim_check(corpus, "corpus"))) return rc;
    if ((rc = check_chunked(queries, "queries")) || (rc = check_chunked(corpus, "corpus"))) return rc;
    if ((rc = ensure_device())) return rc;
    begin_call();
    if (keff == 0) return PMM_OK;
    // One call, the whole box (north_star item 6): above a size threshold the corpus rows are sharded over all visible
    // GPUs by a single-process group (one host thread per GPU, NCCL candidate exchange).  f32 working precision,
    // fused path only; everything else stays on the calling thread's device.
    if (wants_multi_gpu((double)queries->n_rows, (double)corpus->n_rows, (double)corpus->dim, 0.0) &&
        pmm_working_dtype(queries->dtype, corpus->dtype) == PMM_DTYPE_F32 && keff <= TC_MAX_K && !t_opt.force_generic && dev_info().tc) {
        pmm_group *g = (is_chunked(queries) || is_chunked(corpus)) ? nullptr : local_group();
        if (g && corpus->n_rows >= (int64_t)4096 * g->world) return group_host_topk(g, queries, corpus, k, m, out_index, out_score);
    }
    std::lock_guard<std::mutex> device_lock(device_mutex());
    cudaStream_t s = host_stream();
    {
        PathChoice pc = choose_path(queries->dtype, corpus->dtype, keff);
        if (pc.tc && !pc.multipass && t_opt.host_chunked && (double)corpus->n_rows * corpus->dim * esize(corpus->dtype) >= 1e6 * t_opt.host_chunk_min_mb)
            return host_topk_chunked(queries, corpus, keff, m, pc, 0, out_index, out_score, nullptr);
    }
    Uploaded uq, uc;
    if ((rc = upload(queries, s, &uq)) || (rc = upload(corpus, s, &uc))) return rc;
    DevBuf d_idx, d_sc;
    const size_t cnt = (size_t)queries->n_rows * keff;
    CUDA_TRY(d_idx.alloc(cnt * 4, s));
    CUDA_TRY(d_sc.alloc(cnt * 8, s));
    TopkOut o{d_idx.as<uint32_t>(), d_sc.as<double>(), nullptr};
    if ((rc = dev_topk_impl(&uq.dm, &uc.dm, nullptr, corpus->dtype, k, m, 0, o, s))) return rc;
    CUDA_TRY(stage_d2h(out_index, d_idx.p, cnt * 4, s));
    CUDA_TRY(stage_d2h(out_score, d_sc.p, cnt * 8, s));
    stat_add("d2h_bytes", (double)cnt * 12);
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

int pmm_topk_shard(const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard, int64_t k, int32_t metric,
                   int64_t index_base, uint64_t *d_candidates) {
    int rc;
    if ((rc = check_shard_args(queries, corpus_shard, k, metric))) return rc;
    if (queries->n_rows == 0) return PMM_OK;
    if ((rc = ensure_device())) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    return shard_candidates(queries, corpus_shard, k, metric, index_base, d_candidates);
}

int pmm_matmul(const pmm_matrix_t *left, const pmm_matrix_t *right, void *out) {
    int rc;
    if ((rc = check_matrix(left, "left")) || (rc = check_matrix(right, "right"))) return rc;
    if (left->n_rows == 0) return PMM_OK;  // src/matmul.rs:297-305
    if ((rc = check_pair(left, right))) return rc;
    if ((rc = list_dim_check(left, "left")) || (rc = list_dim_check(right, "right"))) return rc;
    if ((rc = check_chunked(left, "left")) || (rc = check_chunked(right, "right"))) return rc;
    if ((rc = ensure_device())) return rc;
    begin_call();
    const int wd = pmm_working_dtype(left->dtype, right->dtype);
    const size_t bytes = (size_t)left->n_rows * right->n_rows * (wd == PMM_DTYPE_F64 ? 8 : 4);
    if (wants_multi_gpu((double)left->n_rows, (double)right->n_rows, (double)left->dim, (double)bytes)) {
        pmm_group *g = (is_chunked(left) || is_chunked(right)) ? nullptr : local_group();   // output rows sharded over the GPUs, no collective
        if (g && left->n_rows >= (int64_t)512 * g->world) return group_host_matmul(g, left, right, out);
    }
    std::lock_guard<std::mutex> device_lock(device_mutex());
    cudaStream_t s = host_stream();
    Uploaded ul, ur;
    if ((rc = upload(left, s, &ul)) || (rc = upload(right, s, &ur))) return rc;
    DevBuf d_out;
    CUDA_TRY(d_out.alloc(bytes, s));
    if ((rc = dev_matmul_impl(&ul.dm, &ur.dm, d_out.p, s))) return rc;
    CUDA_TRY(stage_d2h(out, d_out.p, bytes, s));
    stat_add("d2h_bytes", (double)bytes);
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

// ---------------------------------------------------------------------------------------------- groups of GPUs
int pmm_group_unique_id(void *id) {
    if (!id) return fail(PMM_ERR_INVALID, "null id buffer");
    const NcclApi *nccl = nccl_api();
    if (!nccl) return fail(PMM_ERR_UNSUPPORTED, "%s", nccl_load_error());
    ncclUniqueId u;
    NCCL_TRY(nccl->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return PMM_OK;
}

int pmm_group_init_rank(const void *id, int32_t rank, int32_t world, pmm_group_t **out) {
    if (!out || !id) return fail(PMM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(PMM_ERR_INVALID, "need 0 <= rank < world");
    int rc = ensure_device();
    if (rc) return rc;
    std::unique_ptr<pmm_group> g(new pmm_group());
    g->world = world;
    g->local = false;
    g->members.resize(1);
    cudaGetDevice(&g->members[0].device);
    g->members[0].rank = rank;
    if (world > 1) {
        const NcclApi *nccl = nccl_api();
        if (!nccl) return fail(PMM_ERR_UNSUPPORTED, "%s", nccl_load_error());
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        NCCL_TRY(nccl->CommInitRank(&g->members[0].comm, world, u, rank));
    }
    *out = g.release();
    return PMM_OK;
}

int pmm_group_init_local(int32_t n_devices, pmm_group_t **out) {
    if (!out) return fail(PMM_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = ensure_device();
    if (rc) return rc;
    return create_local_group(n_devices, out);
}

int pmm_group_destroy(pmm_group_t *g) {
    if (!g) return PMM_OK;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        const NcclApi *nccl = nccl_api();
        for (auto &m : g->members) {
            m.worker.reset();
            if (m.comm && nccl) nccl->CommDestroy(m.comm);
            m.comm = nullptr;
        }
    }
    delete g;
    return PMM_OK;
}

int pmm_group_size(const pmm_group_t *g) { return g ? g->world : 0; }

int pmm_group_topk(pmm_group_t *g, const pmm_matrix_t *queries, const pmm_matrix_t *corpus, int64_t k, const char *metric,
                   uint32_t *out_index, double *out_score, int64_t *k_actual) {
    int rc;
    if (!g || !g->local) return fail(PMM_ERR_INVALID, "pmm_group_topk needs a single-process group (pmm_group_init_local)");
    if ((rc = check_matrix(queries, "queries")) || (rc = check_matrix(corpus, "corpus"))) return rc;
    if (is_chunked(queries) || is_chunked(corpus))
        return fail(PMM_ERR_UNSUPPORTED, "the group entry points take single-buffer columns (multi-chunk columns: pmm_topk on one GPU)");
    if (k < 0) return fail(PMM_ERR_INVALID, "k must be non-negative (can't convert negative int to unsigned)");
    const int64_t keff = k < corpus->n_rows ? k : corpus->n_rows;
    if (k_actual) *k_actual = keff;
    if (queries->n_rows == 0) return PMM_OK;
    int32_t m;
    if ((rc = pmm_metric_from_str(metric, &m))) return rc;
    if ((rc = check_pair(queries, corpus))) return rc;
    if ((rc = list_dim_check(queries, "queries")) || (rc = list_dim_check(corpus, "corpus"))) return rc;
    if (keff == 0) return PMM_OK;
    if (pmm_working_dtype(queries->dtype, corpus->dtype) != PMM_DTYPE_F32 || keff > TC_MAX_K)
        return fail(PMM_ERR_UNSUPPORTED, "sharded top-k covers f32 working precision and k <= 248 (packed 8-byte candidates)");
    begin_call();
    return group_host_topk(g, queries, corpus, k, m, out_index, out_score);
}

int pmm_group_topk_shard(pmm_group_t *g, const pmm_matrix_t *queries, const pmm_matrix_t *corpus_shard, int64_t index_base,
                         int64_t n_total, int64_t k, int32_t metric, int32_t flags, uint32_t *out_index, double *out_score) {
    int rc;
    if (!g || g->local || g->members.size() != 1)
        return fail(PMM_ERR_INVALID, "pmm_group_topk_shard needs a one-rank-per-process group (pmm_group_init_rank)");
    if ((rc = check_matrix(queries, "queries")) || (rc = check_matrix(corpus_shard, "corpus"))) return rc;
    if (is_chunked(queries) || is_chunked(corpus_shard))
        return fail(PMM_ERR_UNSUPPORTED, "the group entry points take single-buffer columns (multi-chunk columns: pmm_topk on one GPU)");
    if (k < 0 || n_total < corpus_shard->n_rows || index_base < 0) return fail(PMM_ERR_INVALID, "bad k / n_total / index_base");
    if (metric < 0 || metric > 2) return fail(PMM_ERR_INVALID, "Unknown metric: '%d'. Supported: cosine, dot, euclidean", metric);
    if (queries->n_rows == 0) return PMM_OK;
    if (n_total == 0) return fail(PMM_ERR_INVALID, "Empty series");
    if (corpus_shard->n_rows > 0) {
        if ((rc = check_pair(queries, corpus_shard))) return rc;
        if ((rc = list_dim_check(corpus_shard, "corpus"))) return rc;
    }
    if (pmm_working_dtype(queries->dtype, corpus_shard->dtype) != PMM_DTYPE_F32 || (k < n_total ? k : n_total) > TC_MAX_K)
        return fail(PMM_ERR_UNSUPPORTED, "sharded top-k covers f32 working precision and k <= 248 (packed 8-byte candidates)");
    const int out_mode = flags & 7;
    if (out_mode < GROUP_OUT_HOST_SLICE || out_mode > GROUP_OUT_HOST_FULL) return fail(PMM_ERR_INVALID, "bad output mode in flags");
    if ((rc = ensure_device())) return rc;
    begin_call();
    std::lock_guard<std::mutex> lk(g->mu);
    GroupTopkArgs a;
    a.queries = queries;
    a.shard = corpus_shard;
    a.queries_from_root = (flags & PMM_GROUP_QUERIES_FROM_ROOT) != 0;
    a.index_base = index_base;
    a.n_total = n_total;
    a.k = k;
    a.metric = metric;
    a.out_mode = out_mode;
    a.out_index = out_index;
    a.out_score = out_score;
    return group_rank_topk(g, g->members[0], a);
}

// ---------------------------------------------------------------------------------------------- resident corpus
int pmm_corpus_create(const pmm_matrix_t *corpus, int32_t query_dtype, pmm_corpus_t **out) {
    int rc;
    if (!out) return fail(PMM_ERR_INVALID, "null output handle");
    *out = nullptr;
    if ((rc = check_matrix(corpus, "corpus"))) return rc;
    if (corpus->n_rows == 0) return fail(PMM_ERR_INVALID, "Empty series");
    if (corpus->dim == 0) return fail(PMM_ERR_INVALID, "Zero-dimensional vectors");
    if ((rc = list_dim_check(corpus, "corpus")) || (rc = check_chunked(corpus, "corpus"))) return rc;
    if ((rc = ensure_device())) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    cudaStream_t s = host_stream();
    pmm_corpus *h = new pmm_corpus();
    Uploaded &uc = h->raw;
    if ((rc = upload(corpus, s, &uc))) {
        delete h;
        return rc;
    }
    cudaGetDevice(&h->device);
    h->storage_dtype = corpus->dtype;
    h->query_dtype = query_dtype;
    h->stream = s;
    PathChoice pc = choose_path(query_dtype, corpus->dtype, 1);
    DevBuf err;
    cudaError_t e = err.alloc(sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(err.p, 0, sizeof(int), s);
    if (e != cudaSuccess) {
        delete h;
        return fail(PMM_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
    }
    rc = prepare(uc.dm, pc.mode, pc.f64, TC_TILE_N, true, true, err.as<int>(), s, &h->prep, pc.tc);
    if (!rc) rc = finish_error_flag(err.as<int>(), s);
    if (rc) {
        delete h;
        return rc;
    }
    *out = h;
    return PMM_OK;
}

int pmm_corpus_destroy(pmm_corpus_t *corpus) {
    if (!corpus) return PMM_OK;
    cudaStreamSynchronize(corpus->stream);
    delete corpus;
    return PMM_OK;
}

int64_t pmm_corpus_rows(const pmm_corpus_t *corpus) { return corpus ? corpus->prep.n_rows : 0; }

int pmm_topk_corpus(const pmm_matrix_t *queries, const pmm_corpus_t *corpus, int64_t k, const char *metric,
                    uint32_t *out_index, double *out_score, int64_t *k_actual) {
    int rc;
    if (!corpus) return fail(PMM_ERR_INVALID, "null corpus handle");
    if ((rc = check_matrix(queries, "queries"))) return rc;
    if (k < 0) return fail(PMM_ERR_INVALID, "k must be non-negative (can't convert negative int to unsigned)");
    const int64_t N = corpus->prep.n_rows;
    const int64_t keff = k < N ? k : N;
    if (k_actual) *k_actual = keff;
    if (queries->n_rows == 0) return PMM_OK;
    int32_t m;
    if ((rc = pmm_metric_from_str(metric, &m))) return rc;
    if (queries->dim == 0) return fail(PMM_ERR_INVALID, "Zero-dimensional vectors");
    if (queries->dim != corpus->prep.dim)
        return fail(PMM_ERR_INVALID, "Dimension mismatch: left has %lld dimensional vectors, right has %lld dimensional vectors",
                    (long long)queries->dim, (long long)corpus->prep.dim);
    if ((rc = list_dim_check(queries, "queries")) || (rc = check_chunked(queries, "queries"))) return rc;
    if ((rc = ensure_device())) return rc;
    begin_call();
    std::lock_guard<std::mutex> device_lock(device_mutex());
    if (keff == 0) return PMM_OK;
    cudaStream_t s = host_stream();
    Uploaded uq;
    if ((rc = upload(queries, s, &uq))) return rc;
    DevBuf d_idx, d_sc;
    const size_t cnt = (size_t)queries->n_rows * keff;
    CUDA_TRY(d_idx.alloc(cnt * 4, s));
    CUDA_TRY(d_sc.alloc(cnt * 8, s));
    TopkOut o{d_idx.as<uint32_t>(), d_sc.as<double>(), nullptr};
    if ((rc = dev_topk_impl(&uq.dm, &corpus->raw.dm, &corpus->prep, corpus->storage_dtype, k, m, 0, o, s))) return rc;
    CUDA_TRY(stage_d2h(out_index, d_idx.p, cnt * 4, s));
    CUDA_TRY(stage_d2h(out_score, d_sc.p, cnt * 8, s));
    stat_add("d2h_bytes", (double)cnt * 12);
    CUDA_TRY(cudaStreamSynchronize(s));
    return PMM_OK;
}

}  // extern "C"
