// pmm_stage.h — page-locked staging for pageable host buffers (north_star item 5).
//
// A Polars / Arrow buffer is ordinary pageable memory.  cudaMemcpyAsync from pageable memory is staged by
// the driver through its own bounce buffer and blocks the calling thread until the staging is done, so it
// neither runs at PCIe rate nor overlaps with the host code that would launch the next kernel.  The
// reference pays one full copy of every input as well (`cont_slice().to_vec()`, src/matmul.rs:182-186,
// :214-218); here that copy goes into a ring of page-locked slots, piece by piece, filled by a small pool
// of host threads while the DMA engine drains the previous slot — the upload then overlaps the fused
// kernel exactly as it does for callers that pinned their buffers themselves (those are detected with
// cudaPointerGetAttributes and copied directly).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pmm {

// 0: never stage (pageable copies go through the driver as plain cudaMemcpyAsync). Default 1.
void stage_set_enabled(int on);
bool stage_enabled();
// Number of host threads one staged copy is spread over (including the caller). 0 = automatic.
void stage_set_threads(int n);
// Non-temporal stores for copies INTO the ring (default on).
void stage_set_nt_stores(int on);
int stage_threads();
// Slot size in bytes (>= 1 MB) and slot count (>= 2) of rings created afterwards.
void stage_set_ring(size_t slot_bytes, int slots);

// True when `p` is page-locked (cudaHostAlloc / cudaHostRegister) or managed memory: cudaMemcpyAsync can DMA it.
bool host_ptr_is_pinned(const void *p);
// True when `p` is page-locked host memory the current device can address; *dev = the pointer to use in kernels.
bool host_device_view(const void *p, void **dev);

// Enqueue a host -> device copy of `bytes` on `stream`.  Page-locked sources: one cudaMemcpyAsync.  Pageable sources:
// staged through the calling thread's ring; the call returns when the last piece has been ENQUEUED (the source
// buffer is no longer read after return, the device copy completes in stream order).
cudaError_t stage_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);

// Gather: `rows` rows of `row_bytes`, `src_pitch` bytes apart in host memory, into a dense device block (through the ring).
cudaError_t stage_h2d_rows(void *dst_dev, const void *src_host, size_t row_bytes, size_t src_pitch, size_t rows, cudaStream_t stream);

// Device -> host copy of `bytes`, complete on return for pageable destinations (DMA into the ring, host threads copy
// out); page-locked destinations get one cudaMemcpyAsync on `stream` (complete in stream order, like before).
cudaError_t stage_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);

// Bytes that went through the ring since the last call (statistics: "staged_h2d_bytes" / "staged_d2h_bytes").
void stage_take_counters(double *h2d, double *d2h);

// Frees the calling thread's ring (page-locked memory); the helper threads stay.
void stage_release_thread_ring();

}  // namespace pmm
