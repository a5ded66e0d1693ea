#include "pgs_comm.h"

#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <algorithm>

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

// ---- in-process group: `world` Comm objects of one process, one host thread each
struct LocalGroup {
  std::mutex m; std::condition_variable cv;
  int world = 0, arrived = 0; long long gen = 0;
  std::vector<double*> ptr; std::vector<int> dev;
  double* scratch = nullptr; size_t scratch_n = 0; int scratch_dev = -1;
  int status = 0;
  ~LocalGroup() { if (scratch) { cudaSetDevice(scratch_dev); cudaFree(scratch); } }
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const long long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
namespace {
std::mutex g_groups_mutex;
std::map<std::string, std::weak_ptr<LocalGroup>> g_groups;
__global__ void local_reduce_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n, int op) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = op == kNcclSum ? dst[i] + src[i] : fmax(dst[i], src[i]);
}
}  // namespace

Comm::~Comm() {
  if (comm_ && api().CommDestroy) api().CommDestroy(comm_);
  if (d_flag_) cudaFree(d_flag_);
}

int Comm::init_local(int rank_, int world_, const char* group, std::string* err) {
  if (!group || world_ < 1 || rank_ < 0 || rank_ >= world_) { if (err) *err = "init_local: bad rank/world/group"; return PGS_ERR_INVALID_ARGUMENT; }
  std::lock_guard<std::mutex> lk(g_groups_mutex);
  std::shared_ptr<LocalGroup> g = g_groups[group].lock();
  if (!g) { g = std::make_shared<LocalGroup>(); g->world = world_; g->ptr.assign(world_, nullptr); g->dev.assign(world_, 0); g_groups[group] = g; }
  if (g->world != world_) { if (err) *err = "init_local: group exists with a different world size"; return PGS_ERR_INVALID_ARGUMENT; }
  local_ = g; rank = rank_; world = world_;
  return PGS_OK;
}

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

int Comm::allreduce(double* dev, size_t n, int op, cudaStream_t st, std::string* err) {
  ++n_collectives; bytes_reduced += (long long)(n * sizeof(double));
  if (!local_) {
    const int rc = api().AllReduce(dev, dev, n, kNcclFloat64, op, comm_, st);
    return rc ? nccl_fail(rc, "ncclAllReduce", err) : PGS_OK;
  }
  // local transport: every rank's stream is drained, rank 0 folds the buffers together and hands the result back
  LocalGroup& g = *local_;
  int my_dev = 0; cudaGetDevice(&my_dev);
  cudaError_t e = cudaStreamSynchronize(st);
  g.ptr[rank] = dev; g.dev[rank] = my_dev;
  g.barrier();
  if (rank == 0 && e == cudaSuccess) {
    for (int k = 1; k < world && e == cudaSuccess; ++k) {
      const double* src = g.ptr[k];
      if (g.dev[k] != my_dev) {   // another device: stage the peer's buffer here first
        if (g.scratch_n < n) { if (g.scratch) cudaFree(g.scratch); g.scratch = nullptr; e = cudaMalloc((void**)&g.scratch, sizeof(double) * n); g.scratch_n = e == cudaSuccess ? n : 0; g.scratch_dev = my_dev; }
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(g.scratch, my_dev, g.ptr[k], g.dev[k], sizeof(double) * n, st);
        src = g.scratch;
      }
      if (e == cudaSuccess) local_reduce_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, st>>>(dev, src, n, op);
    }
    for (int k = 1; k < world && e == cudaSuccess; ++k) e = cudaMemcpyPeerAsync(g.ptr[k], g.dev[k], dev, my_dev, sizeof(double) * n, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    g.status = e == cudaSuccess ? 0 : 1;
  }
  g.barrier();
  if (e != cudaSuccess || g.status) { if (err) *err = std::string("local all-reduce failed: ") + cudaGetErrorString(e); return PGS_ERR_CUDA; }
  return PGS_OK;
}
int Comm::allreduce_sum(double* dev, size_t n, cudaStream_t st, std::string* err) { return allreduce(dev, n, kNcclSum, st, err); }
int Comm::allreduce_max(double* dev, size_t n, cudaStream_t st, std::string* err) { return allreduce(dev, n, kNcclMax, st, err); }

int Comm::agree(int status, cudaStream_t st, std::string* err) {
  if (!d_flag_) { if (cudaMalloc((void**)&d_flag_, sizeof(double)) != cudaSuccess) { d_flag_ = nullptr; if (err) *err = "agree: cudaMalloc failed"; return PGS_ERR_CUDA; } }
  double v = status ? (double)(-status) : 0.0;   // pgs_status codes are negative: the largest magnitude wins
  if (cudaMemcpyAsync(d_flag_, &v, sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) { if (err) *err = "agree: copy failed"; return PGS_ERR_CUDA; }
  cudaStreamSynchronize(st);
  if (int rc = allreduce(d_flag_, 1, kNcclMax, st, err)) return rc;
  if (cudaMemcpyAsync(&v, d_flag_, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { if (err) *err = "agree: copy failed"; return PGS_ERR_CUDA; }
  if (v != 0.0 && status == 0 && err) *err = "another rank of the multi-GPU solve failed (status " + std::to_string(-(int)v) + ")";
  return v != 0.0 ? -(int)v : PGS_OK;
}

}  // namespace pgs
