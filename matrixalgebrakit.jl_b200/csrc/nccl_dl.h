// NCCL entry points resolved at run time (dlopen): libmakb200.so has no link-time dependency on NCCL, so the
// single-GPU paths load on a box without it, and inside a torch process the library shares the libnccl.so.2
// torch already mapped (RTLD_NOLOAD first).  Types come from <nccl.h>; only the symbols are late-bound.
#pragma once
#include <nccl.h>

namespace mak {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    const char* loaded_from;
};

// nullptr when no libnccl.so.2 can be found (path override: env MAKB200_NCCL_LIB); `why` receives dlerror()
const NcclApi* nccl_api(const char** why = nullptr);

}  // namespace mak
