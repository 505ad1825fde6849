// NCCL, loaded at run time.  libtod_b200.so has no link-time NCCL dependency: a host process that already carries an
// NCCL (a PyTorch process brings its own libnccl.so.2) must not get a second, possibly different copy mapped over it,
// and single-GPU users need no NCCL at all.  dlopen("libnccl.so.2") returns the copy already loaded in the process if
// there is one, else the system library.  Only the handful of entry points the sharded matcher needs are bound.
#ifndef TOD_NCCL_DL_H_
#define TOD_NCCL_DL_H_

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <cstdint>

namespace tod {

// ABI-stable subset of nccl.h (values unchanged since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { kNcclSuccess = 0 };
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint32 = 3 };
enum { kNcclSum = 0, kNcclProd = 1, kNcclMax = 2, kNcclMin = 3 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  bool ok = false;
};

inline const NcclApi &nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.CommAbort = reinterpret_cast<decltype(a.CommAbort)>(dlsym(h, "ncclCommAbort"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GetErrorString;
    return a;
  }();
  return api;
}

}  // namespace tod
#endif
