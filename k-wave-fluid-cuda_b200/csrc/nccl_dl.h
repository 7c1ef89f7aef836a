// NCCL entry points used by the slab-decomposed transforms, resolved at run time (dlopen) so that the library has no
// link-time dependency on NCCL: single-GPU hosts (and the C++ host) never load it, and under torchrun the very
// libnccl.so.2 that torch.distributed already mapped is reused.  Only the stable C ABI of NCCL 2.x is declared here.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

namespace kw {

struct KwNcclUniqueId {
  char internal[128];
};
typedef struct ncclComm* KwNcclComm;
enum { kNcclSuccess = 0, kNcclInt32 = 2, kNcclFloat = 7, kNcclMin = 3 };

struct NcclApi {
  int (*GetUniqueId)(KwNcclUniqueId*) = nullptr;
  int (*CommInitRank)(KwNcclComm*, int, KwNcclUniqueId, int) = nullptr;
  int (*CommDestroy)(KwNcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, KwNcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, KwNcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, KwNcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  void* handle = nullptr;
  std::string error;
  bool ok = false;
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* env = getenv("KW_NCCL_LIB");
  const char* candidates[] = {env, "libnccl.so.2", "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2",
                              "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
  for (const char* name : candidates) {
    if (!name || !*name) continue;
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.error = "libnccl.so.2 not found (set KW_NCCL_LIB)";
    return api;
  }
#define KW_NCCL_SYM(field, sym)                                    \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) {                                                \
    api.error = std::string("NCCL symbol missing: ") + sym;       \
    return api;                                                    \
  }
  KW_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  KW_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  KW_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  KW_NCCL_SYM(Send, "ncclSend")
  KW_NCCL_SYM(Recv, "ncclRecv")
  KW_NCCL_SYM(AllReduce, "ncclAllReduce")
  KW_NCCL_SYM(GroupStart, "ncclGroupStart")
  KW_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  KW_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  KW_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef KW_NCCL_SYM
  api.ok = true;
  return api;
}

}  // namespace kw
