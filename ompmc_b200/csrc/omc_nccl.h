// omc_nccl.h -- NCCL bound at run time (internal to libompmc_b200.so).
//
// The library's only collective is the sum of a completed batch grid over the GPUs of a run.  NCCL is
// resolved with dlopen() the first time a communicator is asked for, not at link time: single-GPU users (and the CPU-side
// ABI tests) need no NCCL at all, and a host process that already carries an NCCL (PyTorch bundles its own libnccl.so.2)
// must end up with ONE copy of it -- dlopen("libnccl.so.2") returns the copy that is already mapped, else the system one.
// Only the stable 2.x C API below is used; the two enum values are NCCL's (ncclSum = 0, ncclFloat64 = 8).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>

namespace omc {

typedef struct ncclComm *nccl_comm_t;
struct nccl_uid {
    char internal[128];     // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES)
};
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    void *dl = nullptr;
    std::string err;
    bool ok = false;
};

inline NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.dl) break;
        }
        if (!a.dl) { a.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return a; }
        bool all = true;
        auto sym = [&](const char *name) { void *p = dlsym(a.dl, name); if (!p) { all = false; a.err = std::string("NCCL symbol missing: ") + name; } return p; };
        a.GetUniqueId = (int (*)(nccl_uid *))sym("ncclGetUniqueId");
        a.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid, int))sym("ncclCommInitRank");
        a.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
        a.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclAllReduce");
        a.GroupStart = (int (*)())sym("ncclGroupStart");
        a.GroupEnd = (int (*)())sym("ncclGroupEnd");
        a.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
        a.GetVersion = (int (*)(int *))sym("ncclGetVersion");
        a.ok = all;
        return a;
    }();
    return api;
}

}  // namespace omc
