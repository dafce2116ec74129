// NCCL resolved at run time (dlopen), so libb2s.so has no link-time NCCL dependency and, inside a process that already
// carries an NCCL (PyTorch bundles one under the same soname), uses THAT copy instead of loading a second one.
// Only the few entry points the CFR table exchange needs (ncclAllReduce over NVLink for CFR's tables).
#pragma once
#include <dlfcn.h>
#include <nccl.h>      // types and enums only; every function is looked up with dlsym

#include <string>

namespace b2s {

struct NcclApi {
  void* handle = nullptr;
  std::string error;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok() const { return handle != nullptr && error.empty(); }
};

inline const NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process, if any
      if (a.handle) break;
    }
    if (!a.handle)
      for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
      }
    if (!a.handle) { a.error = std::string("NCCL not found: ") + dlerror(); return a; }
    auto sym = [&](const char* n) { void* p = dlsym(a.handle, n); if (!p && a.error.empty()) a.error = std::string("NCCL symbol missing: ") + n; return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    return a;
  }();
  return api;
}

}  // namespace b2s
