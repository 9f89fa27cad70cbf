// NCCL binding of the library: the all-reduce of the MergeChains counts and of the
// ChainConvergence within/between sums at the monitor interval (cmd/root.go:498-539) runs INSIDE the C ABI, so a Go
// or C++ host needs no collective plumbing of its own.
//
// libnccl is bound at run time (dlopen + dlsym), not at link time: the shared library must load on a box without
// NCCL or without a GPU (the CPU test suite checks its exports), and inside a Python process that already carries
// torch's bundled libnccl.so.2 the same copy is reused (RTLD_NOLOAD first) instead of a second one being mapped.
// Only the handful of entry points below are used; their signatures have been stable since NCCL 2.0.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <mutex>
#include <string>

#include "host_model.hpp"

namespace gbn {

// values of nccl.h (ncclDataType_t / ncclRedOp_t), restated so that the build does not need the header
constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;
constexpr int kUniqueIdBytes = 128;
struct UniqueId { char internal[kUniqueIdBytes]; };
using Comm = void*;

struct Api {
    void* so = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommInitAll)(Comm*, int, const int*) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

inline const Api& api() {
    static Api a;
    static std::once_flag once;
    static std::string why;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.so = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // a copy the process already mapped (torch's)
            if (a.so) break;
        }
        if (!a.so)
            for (const char* n : names) {
                a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (a.so) break;
            }
        if (!a.so) {
            const char* e = dlerror();
            why = std::string("libnccl.so.2 could not be loaded: ") + (e ? e : "unknown");
            return;
        }
        auto sym = [&](const char* s) {
            void* p = dlsym(a.so, s);
            if (!p) why = std::string("libnccl lacks ") + s;
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!a.so || !why.empty()) throw gb::Err("grample_b200: NCCL is not available (" + why + ")");
    return a;
}

inline void check(int rc, const char* what) {
    if (rc != 0) throw gb::Err(std::string("NCCL error: ") + api().GetErrorString(rc) + " at " + what);
}
#define NCCL_CHECK(expr) ::gbn::check((expr), #expr)

}  // namespace gbn

// One rank of a communicator, bound to one device.  world == 1 needs no NCCL at all (nccl == nullptr).
struct gb_comm {
    gbn::Comm nccl = nullptr;
    int32_t world = 1, rank = 0;
    int device = 0;
    bool in_fleet = false;  // single-process communicator: collectives of its ranks are issued inside one NCCL group
};
