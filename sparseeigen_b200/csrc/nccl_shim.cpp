#include "nccl_shim.h"
#include <dlfcn.h>
#include <mutex>

namespace speigen {
void set_error(const char* fmt, ...);

namespace {
NcclApi g_api;
bool g_ok = false;
std::once_flag g_once;

void* open_nccl() {
  // prefer a copy that is already mapped into the process (torch's bundled NCCL)
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL)) return h;
  }
  for (const char* n : names) {
    if (void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) return h;
  }
  return nullptr;
}

void load() {
  void* h = open_nccl();
  if (!h) return;
#define SYM(field, name)                                                    \
  *reinterpret_cast<void**>(&g_api.field) = dlsym(h, name);                 \
  if (!g_api.field) return;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GetVersion, "ncclGetVersion")
#undef SYM
  g_ok = true;
}
}  // namespace

const NcclApi* nccl_api() {
  std::call_once(g_once, load);
  if (!g_ok) {
    const char* why = dlerror();   // dlerror() clears its state: read it once
    set_error("NCCL (libnccl.so.2) could not be loaded: %s", why ? why : "symbol lookup failed");
    return nullptr;
  }
  return &g_api;
}

}  // namespace speigen
