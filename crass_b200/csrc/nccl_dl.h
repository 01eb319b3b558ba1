// nccl_dl.h -- NCCL through dlopen: the library has no link-time dependency on it (engine.cu: one process, several devices,
// ncclCommInitAll; capi.cu: one process per GPU, ncclCommInitRank with an id the caller passes round).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>

namespace {
struct Nccl {
    struct UniqueId { char internal[128]; };               // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
    typedef int (*CommInitAll_t)(void** comms, int ndev, const int* devlist);
    typedef int (*CommInitRank_t)(void** comm, int nranks, UniqueId id, int rank);
    typedef int (*GetUniqueId_t)(UniqueId* id);
    typedef int (*CommDestroy_t)(void* comm);
    typedef int (*AllGather_t)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t s);
    typedef int (*Group_t)();
    typedef const char* (*ErrStr_t)(int);
    void* lib = nullptr;
    CommInitAll_t CommInitAll = nullptr; CommInitRank_t CommInitRank = nullptr; GetUniqueId_t GetUniqueId = nullptr;
    CommDestroy_t CommDestroy = nullptr; AllGather_t AllGather = nullptr;
    Group_t GroupStart = nullptr, GroupEnd = nullptr; ErrStr_t GetErrorString = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {getenv("CRASS_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (CommInitAll_t)dlsym(lib, "ncclCommInitAll");
        CommInitRank = (CommInitRank_t)dlsym(lib, "ncclCommInitRank");
        GetUniqueId = (GetUniqueId_t)dlsym(lib, "ncclGetUniqueId");
        CommDestroy = (CommDestroy_t)dlsym(lib, "ncclCommDestroy");
        AllGather = (AllGather_t)dlsym(lib, "ncclAllGather");
        GroupStart = (Group_t)dlsym(lib, "ncclGroupStart");
        GroupEnd = (Group_t)dlsym(lib, "ncclGroupEnd");
        GetErrorString = (ErrStr_t)dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommInitRank || !GetUniqueId || !CommDestroy || !AllGather || !GroupStart || !GroupEnd) { dlclose(lib); lib = nullptr; return false; }
        return true;
    }
    const char* why(int rc) const { return GetErrorString ? GetErrorString(rc) : "?"; }
};
}  // namespace
