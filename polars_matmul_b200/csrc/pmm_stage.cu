// pmm_stage.cu — page-locked staging ring + host copy threads (see pmm_stage.h).
#include "pmm_stage.h"

#include <emmintrin.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace pmm {
namespace {

std::atomic<int> g_threads{0};                       // 0 = automatic
std::atomic<size_t> g_slot_bytes{(size_t)32 << 20};
std::atomic<int> g_slots{4};
std::atomic<uint64_t> g_staged_h2d{0}, g_staged_d2h{0};
std::atomic<int> g_enabled{1};

std::atomic<int> g_nt_stores{1};

// One range of a staged copy.  Into the ring (pageable -> page-locked) the bytes are written with non-temporal stores: the
// destination is read next by the DMA engine, not by a core, so pulling its lines into the cache first (read for
// ownership) only costs memory bandwidth - a third of the copy's traffic.  glibc's memcpy does the same, but only above a
// size threshold that the per-thread ranges (a few MB) stay below.
void copy_range(void *dst, const void *src, size_t bytes, bool stream_dst) {
    if (!stream_dst || !g_nt_stores.load(std::memory_order_relaxed) || (((uintptr_t)dst) & 15) || bytes < 4096) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t n64 = bytes / 64;
    const __m128i *s = (const __m128i *)src;
    __m128i *d = (__m128i *)dst;
    for (size_t i = 0; i < n64; ++i) {
        const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1);
        const __m128i c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
        _mm_stream_si128(d + 4 * i, a);
        _mm_stream_si128(d + 4 * i + 1, b);
        _mm_stream_si128(d + 4 * i + 2, c);
        _mm_stream_si128(d + 4 * i + 3, e);
    }
    _mm_sfence();
    if (bytes & 63) memcpy((char *)dst + n64 * 64, (const char *)src + n64 * 64, bytes & 63);
}

int auto_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    int n = (int)(hc / 2);   // leave cores to the caller's other threads (Polars runs its own pool)
    if (n > 8) n = 8;
    if (n < 1) n = 1;
    return n;
}

// A process-wide pool of helper threads; a job is one memcpy range.  Several callers (one host thread per GPU in the
// multi-GPU driver) may submit concurrently; each waits for its own ranges only.
class CopyPool {
  public:
    struct Job {
        void *dst;
        const void *src;
        size_t bytes;
        std::atomic<int> *pending;
        bool stream_dst;
    };
    static CopyPool &get() {
        static CopyPool *p = new CopyPool();  // leaked on purpose: helper threads may outlive static destructors
        return *p;
    }
    void ensure(int helpers) {
        std::lock_guard<std::mutex> lk(mu_);
        while ((int)workers_.size() < helpers) workers_.emplace_back([this] { run(); }), workers_.back().detach();
    }
    void submit(const Job &j) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(j);
        }
        cv_.notify_one();
    }
    // The caller helps while it waits: takes queued ranges (its own or another caller's) instead of sleeping.
    void wait(std::atomic<int> *pending) {
        while (pending->load(std::memory_order_acquire) > 0) {
            Job j;
            bool have = false;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!q_.empty()) {
                    j = q_.front();
                    q_.pop_front();
                    have = true;
                }
            }
            if (have) {
                copy_range(j.dst, j.src, j.bytes, j.stream_dst);
                j.pending->fetch_sub(1, std::memory_order_release);
            } else {
                std::this_thread::yield();
            }
        }
    }

  private:
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                j = q_.front();
                q_.pop_front();
            }
            copy_range(j.dst, j.src, j.bytes, j.stream_dst);
            j.pending->fetch_sub(1, std::memory_order_release);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    std::vector<std::thread> workers_;
};

// stream_dst: the destination is a ring slot (see copy_range)
void parallel_memcpy(void *dst, const void *src, size_t bytes, bool stream_dst = false) {
    const int nt = stage_threads();
    if (nt <= 1 || bytes < ((size_t)1 << 20)) {
        copy_range(dst, src, bytes, stream_dst);
        return;
    }
    CopyPool &pool = CopyPool::get();
    pool.ensure(nt - 1);
    size_t part = (bytes / nt + 4095) & ~(size_t)4095;
    std::atomic<int> pending{0};
    size_t off = part;  // the caller copies [0, part) itself
    int n_jobs = 0;
    for (; off < bytes; off += part) ++n_jobs;
    pending.store(n_jobs, std::memory_order_relaxed);
    for (off = part; off < bytes; off += part) {
        const size_t n = bytes - off < part ? bytes - off : part;
        pool.submit(CopyPool::Job{(char *)dst + off, (const char *)src + off, n, &pending, stream_dst});
    }
    copy_range(dst, src, part < bytes ? part : bytes, stream_dst);
    pool.wait(&pending);
}

struct Ring {
    struct Slot {
        void *p = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false;
    };
    std::vector<Slot> slots;
    size_t slot_bytes = 0;
    int dev = -1;
    int next = 0;
    cudaError_t ensure() {
        int cur = 0;
        cudaError_t e = cudaGetDevice(&cur);
        if (e != cudaSuccess) return e;
        const size_t want_bytes = g_slot_bytes.load();
        const int want_slots = g_slots.load();
        if (!slots.empty() && (slot_bytes != want_bytes || (int)slots.size() != want_slots)) release();
        if (slots.empty()) {
            slots.resize(want_slots);
            slot_bytes = want_bytes;
            for (auto &s : slots) {
                e = cudaHostAlloc(&s.p, slot_bytes, cudaHostAllocPortable);
                if (e != cudaSuccess) {
                    release();
                    return e;
                }
            }
            dev = -1;
        }
        if (dev != cur) {  // events belong to a device
            for (auto &s : slots) {
                if (s.ev) {
                    if (s.busy) cudaEventSynchronize(s.ev);
                    cudaEventDestroy(s.ev);
                }
                s.busy = false;
                e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming);
                if (e != cudaSuccess) return e;
            }
            dev = cur;
        }
        return cudaSuccess;
    }
    cudaError_t acquire(Slot **out) {
        Slot &s = slots[next];
        next = (next + 1) % (int)slots.size();
        if (s.busy) {
            cudaError_t e = cudaEventSynchronize(s.ev);
            if (e != cudaSuccess) return e;
            s.busy = false;
        }
        *out = &s;
        return cudaSuccess;
    }
    void release() {
        for (auto &s : slots) {
            if (s.ev) {
                if (s.busy) cudaEventSynchronize(s.ev);
                cudaEventDestroy(s.ev);
            }
            if (s.p) cudaFreeHost(s.p);
        }
        slots.clear();
        slot_bytes = 0;
        dev = -1;
        next = 0;
    }
    ~Ring() { release(); }
};
thread_local Ring t_ring;

constexpr size_t kDirectBelow = (size_t)256 << 10;  // small pageable copies: the driver's own staging is fine

}  // namespace

void stage_set_enabled(int on) { g_enabled.store(on ? 1 : 0); }
bool stage_enabled() { return g_enabled.load() != 0; }
void stage_set_nt_stores(int on) { g_nt_stores.store(on ? 1 : 0); }
void stage_set_threads(int n) { g_threads.store(n < 0 ? 0 : n > 64 ? 64 : n); }
int stage_threads() {
    const int n = g_threads.load();
    return n > 0 ? n : auto_threads();
}
void stage_set_ring(size_t slot_bytes, int slots) {
    if (slot_bytes < ((size_t)1 << 20)) slot_bytes = (size_t)1 << 20;
    if (slots < 2) slots = 2;
    if (slots > 16) slots = 16;
    g_slot_bytes.store(slot_bytes);
    g_slots.store(slots);
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// Page-locked host memory the current device can address: *dev = the device-side pointer (UVA: normally the same address).
bool host_device_view(const void *p, void **dev) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return false;
    *dev = attr.devicePointer;
    return true;
}

cudaError_t stage_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (bytes < kDirectBelow || !g_enabled.load() || host_ptr_is_pinned(src_host))
        return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream);
    cudaError_t e = t_ring.ensure();
    if (e != cudaSuccess) {  // no page-locked memory to be had: fall back to the driver's staging
        cudaGetLastError();
        return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream);
    }
    const size_t S = t_ring.slot_bytes;
    for (size_t off = 0; off < bytes; off += S) {
        const size_t n = bytes - off < S ? bytes - off : S;
        Ring::Slot *s;
        if ((e = t_ring.acquire(&s)) != cudaSuccess) return e;
        parallel_memcpy(s->p, (const char *)src_host + off, n, true);
        if ((e = cudaMemcpyAsync((char *)dst_dev + off, s->p, n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(s->ev, stream)) != cudaSuccess) return e;
        s->busy = true;
    }
    g_staged_h2d.fetch_add(bytes);
    return cudaSuccess;
}

// `rows` rows of `row_bytes` each, `src_pitch` bytes apart in host memory -> a dense block in device memory.  Through the ring
// whatever the source (the gather itself needs a page-locked landing area); small: used for a few thousand sampled rows.
cudaError_t stage_h2d_rows(void *dst_dev, const void *src_host, size_t row_bytes, size_t src_pitch, size_t rows, cudaStream_t stream) {
    if (rows == 0 || row_bytes == 0) return cudaSuccess;
    cudaError_t e = t_ring.ensure();
    if (e != cudaSuccess) {   // no page-locked memory to be had: let the driver do the strided copy
        cudaGetLastError();
        return cudaMemcpy2DAsync(dst_dev, row_bytes, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, stream);
    }
    const size_t per_slot = t_ring.slot_bytes / row_bytes;
    if (per_slot == 0) return cudaMemcpy2DAsync(dst_dev, row_bytes, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, stream);
    for (size_t r0 = 0; r0 < rows; r0 += per_slot) {
        const size_t n = rows - r0 < per_slot ? rows - r0 : per_slot;
        Ring::Slot *sl;
        if ((e = t_ring.acquire(&sl)) != cudaSuccess) return e;
        for (size_t i = 0; i < n; ++i) memcpy((char *)sl->p + i * row_bytes, (const char *)src_host + (r0 + i) * src_pitch, row_bytes);
        if ((e = cudaMemcpyAsync((char *)dst_dev + r0 * row_bytes, sl->p, n * row_bytes, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(sl->ev, stream)) != cudaSuccess) return e;
        sl->busy = true;
    }
    g_staged_h2d.fetch_add(rows * row_bytes);
    return cudaSuccess;
}

cudaError_t stage_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (bytes < kDirectBelow || !g_enabled.load() || host_ptr_is_pinned(dst_host))
        return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = t_ring.ensure();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
    }
    const size_t S = t_ring.slot_bytes;
    const int K = (int)t_ring.slots.size();
    struct InFlight {
        Ring::Slot *s;
        size_t off, n;
    };
    std::deque<InFlight> fl;
    auto drain_one = [&]() -> cudaError_t {
        InFlight f = fl.front();
        fl.pop_front();
        cudaError_t e2 = cudaEventSynchronize(f.s->ev);
        f.s->busy = false;
        if (e2 != cudaSuccess) return e2;
        parallel_memcpy((char *)dst_host + f.off, f.s->p, f.n);
        return cudaSuccess;
    };
    for (size_t off = 0; off < bytes; off += S) {
        const size_t n = bytes - off < S ? bytes - off : S;
        if ((int)fl.size() >= K - 1 && (e = drain_one()) != cudaSuccess) return e;
        Ring::Slot *s;
        if ((e = t_ring.acquire(&s)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(s->p, (const char *)src_dev + off, n, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(s->ev, stream)) != cudaSuccess) return e;
        s->busy = true;
        fl.push_back(InFlight{s, off, n});
    }
    while (!fl.empty())
        if ((e = drain_one()) != cudaSuccess) return e;
    g_staged_d2h.fetch_add(bytes);
    return cudaSuccess;
}

void stage_take_counters(double *h2d, double *d2h) {
    const uint64_t a = g_staged_h2d.exchange(0), b = g_staged_d2h.exchange(0);
    if (h2d) *h2d = (double)a;
    if (d2h) *d2h = (double)b;
}

void stage_release_thread_ring() { t_ring.release(); }

}  // namespace pmm
