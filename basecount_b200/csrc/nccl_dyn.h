// nccl_dyn.h -- NCCL bound at run time (dlopen), so the library neither needs NCCL to load on a one-GPU box nor
// brings a second copy into a process that already holds one (a host that imported torch has torch's NCCL
// mapped under the same soname; RTLD_NOLOAD finds that copy first).  BASECOUNT_B200_NCCL names a library file.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace bcnccl {

struct Api {
    void *lib = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok() const { return lib != nullptr && err.empty(); }
};

inline Api &api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {std::getenv("BASECOUNT_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        for (const char *n : names) {
            if (a.lib) break;
            if (n && *n) a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!a.lib) {
            a.err = std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found");
            return;
        }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(a.lib, name);
            if (!p && a.err.empty()) a.err = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.Send = (decltype(a.Send))sym("ncclSend");
        a.Recv = (decltype(a.Recv))sym("ncclRecv");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    });
    return a;
}

}  // namespace bcnccl
