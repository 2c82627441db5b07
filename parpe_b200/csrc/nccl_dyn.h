// nccl_dyn.h -- NCCL entry points resolved at run time (dlopen), so that libparpe_b200.so
// neither links against a particular libnccl nor fails to load on a box without one.  The
// collective it serves: the ONE all-reduce of (nllh, #errors, gradient[n_opt]) that replaces
// the reference master's summation of the workers' results
// (src/parpeamici/multiConditionProblem.cpp:738-748,775-824; loadBalancerMaster.cpp:200-244).
// If a libnccl.so.2 is already loaded in the process (e.g. PyTorch's bundled copy), that one is
// used; otherwise the system library.
#ifndef PB200_NCCL_DYN_H
#define PB200_NCCL_DYN_H

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <string>

namespace pb200 {

struct NcclApi {
    // minimal mirror of nccl.h (stable ABI since NCCL 2.0)
    struct UniqueId { char internal[128]; };
    typedef struct ncclComm* Comm;
    enum { Success = 0, Sum = 0, Max = 2, Float64 = 8 };
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* dl = nullptr;
    std::string error;

    bool ok() const { return AllReduce != nullptr; }

    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }

  private:
    static NcclApi load() {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.dl) break;
        }
        if (!a.dl) {
            a.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return a;
        }
#define PB_NCCL_SYM(field, sym)                                      \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.dl, sym)); \
    if (!a.field) {                                                  \
        a.error = std::string("NCCL symbol missing: ") + sym;        \
        a.AllReduce = nullptr;                                       \
        return a;                                                    \
    }
        PB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        PB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        PB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        PB_NCCL_SYM(GroupStart, "ncclGroupStart")
        PB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        PB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        PB_NCCL_SYM(GetVersion, "ncclGetVersion")
        PB_NCCL_SYM(AllReduce, "ncclAllReduce")
#undef PB_NCCL_SYM
        return a;
    }
};

}  // namespace pb200
#endif
