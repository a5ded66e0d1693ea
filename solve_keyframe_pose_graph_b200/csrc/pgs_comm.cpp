#include "pgs_comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/pgs.h"

namespace pgs {

// Minimal mirror of the NCCL 2.x C API (stable since 2.0): only what the border exchange needs.
namespace {
struct NcclUniqueId { char internal[128]; };
enum { kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string load_error;
};
NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
    if (!a.lib) { a.load_error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    a.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))dlsym(a.lib, "ncclCommInitRank");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(a.lib, "ncclAllReduce");
    a.CommDestroy = (int (*)(void*))dlsym(a.lib, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) a.load_error = "libnccl.so.2 lacks the expected symbols";
  });
  return a;
}
int nccl_fail(int rc, const char* what, std::string* err) {
  if (err) *err = std::string("NCCL error in ") + what + ": " + (api().GetErrorString ? api().GetErrorString(rc) : "?");
  return PGS_ERR_CUDA;
}
}  // namespace

Comm::~Comm() { if (comm_ && api().CommDestroy) api().CommDestroy(comm_); }

int Comm::unique_id(void* id128, std::string* err) {
  NcclApi& a = api();
  if (!a.load_error.empty()) { if (err) *err = a.load_error; return PGS_ERR_STATE; }
  NcclUniqueId id;
  const int rc = a.GetUniqueId(&id);
  if (rc) return nccl_fail(rc, "ncclGetUniqueId", err);
  std::memcpy(id128, &id, sizeof(id));
  return PGS_OK;
}

int Comm::init(int rank_, int world_, const void* id128, std::string* err) {
  NcclApi& a = api();
  if (!a.load_error.empty()) { if (err) *err = a.load_error; return PGS_ERR_STATE; }
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  const int rc = a.CommInitRank(&comm_, world_, id, rank_);
  if (rc) return nccl_fail(rc, "ncclCommInitRank", err);
  rank = rank_; world = world_;
  return PGS_OK;
}

int Comm::allreduce_sum(double* dev, size_t n, cudaStream_t st, std::string* err) {
  const int rc = api().AllReduce(dev, dev, n, kNcclFloat64, kNcclSum, comm_, st);
  ++n_collectives; bytes_reduced += (long long)(n * sizeof(double));
  return rc ? nccl_fail(rc, "ncclAllReduce(sum)", err) : PGS_OK;
}
int Comm::allreduce_max(double* dev, size_t n, cudaStream_t st, std::string* err) {
  const int rc = api().AllReduce(dev, dev, n, kNcclFloat64, kNcclMax, comm_, st);
  ++n_collectives; bytes_reduced += (long long)(n * sizeof(double));
  return rc ? nccl_fail(rc, "ncclAllReduce(max)", err) : PGS_OK;
}

}  // namespace pgs
