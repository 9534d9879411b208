// NCCL is bound lazily with dlopen so that (a) the library loads on a box with no GPU / no NCCL and
// (b) inside a Python process that already imported torch we reuse torch's libnccl.so.2 instead of
// pulling in a second copy.  Only the handful of entry points the exchange step needs.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace speigen {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_DOUBLE = 8 };                 // ncclFloat64
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };      // ncclRedOp_t (ncclSum, ncclProd, ncclMax, ncclMin)

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommInitAll)(ncclComm_t*, int, const int*);
  int (*CommDestroy)(ncclComm_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
};

// nullptr (and set_error) if NCCL cannot be loaded.
const NcclApi* nccl_api();

}  // namespace speigen
