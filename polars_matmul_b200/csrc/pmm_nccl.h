// pmm_nccl.h — NCCL, loaded at run time.
//
// The multi-GPU driver (pmm_api.cu, "groups") exchanges Q x k packed candidates between the GPUs of one box with
// NCCL send/recv over NVLink (north_star item 6).  libnccl is opened with dlopen on first use instead of being a
// link-time dependency: a single-GPU user never needs it, and a process that already carries a copy (PyTorch ships
// one) keeps using exactly that copy — two NCCL builds in one address space do not mix.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>   // types and enums only; no symbol of libnccl is referenced at link time

namespace pmm {

struct NcclApi {
    ncclResult_t (*GetVersion)(int *);
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};

// The loaded API, or NULL when libnccl cannot be opened (nccl_load_error() says why).
const NcclApi *nccl_api();
const char *nccl_load_error();

}  // namespace pmm
