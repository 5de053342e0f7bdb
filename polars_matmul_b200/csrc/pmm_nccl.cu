// pmm_nccl.cu — dlopen of libnccl (see pmm_nccl.h).
#include "pmm_nccl.h"

#include <dlfcn.h>
#include <stdio.h>

#include <mutex>

namespace pmm {
namespace {
NcclApi g_api;
bool g_ok = false;
char g_err[256] = "";
std::once_flag g_once;

void load() {
    // RTLD_NOLOAD first: a copy that is already mapped (e.g. the one PyTorch loaded) is the one to use
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
    if (!h)
        for (const char *n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) {
        snprintf(g_err, sizeof(g_err), "libnccl.so.2 could not be loaded: %s", dlerror());
        return;
    }
#define PMM_SYM(field, name)                                                          \
    g_api.field = (decltype(g_api.field))dlsym(h, name);                              \
    if (!g_api.field) {                                                               \
        snprintf(g_err, sizeof(g_err), "libnccl lacks the symbol %s", name);          \
        return;                                                                       \
    }
    PMM_SYM(GetVersion, "ncclGetVersion")
    PMM_SYM(GetUniqueId, "ncclGetUniqueId")
    PMM_SYM(CommInitRank, "ncclCommInitRank")
    PMM_SYM(CommInitAll, "ncclCommInitAll")
    PMM_SYM(CommDestroy, "ncclCommDestroy")
    PMM_SYM(GroupStart, "ncclGroupStart")
    PMM_SYM(GroupEnd, "ncclGroupEnd")
    PMM_SYM(Send, "ncclSend")
    PMM_SYM(Recv, "ncclRecv")
    PMM_SYM(Broadcast, "ncclBroadcast")
    PMM_SYM(AllReduce, "ncclAllReduce")
    PMM_SYM(GetErrorString, "ncclGetErrorString")
#undef PMM_SYM
    g_ok = true;
}
}  // namespace

const NcclApi *nccl_api() {
    std::call_once(g_once, load);
    return g_ok ? &g_api : nullptr;
}
const char *nccl_load_error() { return g_err; }

}  // namespace pmm
