// nccl_dyn.h — NCCL bound at run time with dlopen (libnccl.so.2), so the library loads on hosts
// without NCCL and shares the copy PyTorch already mapped when the host process uses torch.
// Only the point-to-point subset the slab exchange needs.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t; // ncclSuccess == 0
enum { ncclChar_ = 0 };   // ncclInt8 / ncclChar

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
    const char* why = "";
};

static inline NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.why = "libnccl.so.2 not found (dlopen)";
        return api;
    }
#define NCCL_SYM(field, name)                                                   \
    *(void**)(&api.field) = dlsym(api.handle, name);                            \
    if (!api.field) {                                                           \
        api.why = "missing symbol " name;                                       \
        return api;                                                             \
    }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
    NCCL_SYM(GetVersion, "ncclGetVersion")
#undef NCCL_SYM
    api.ok = true;
    return api;
}
