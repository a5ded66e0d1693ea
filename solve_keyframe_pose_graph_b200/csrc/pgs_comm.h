// Collectives of the multi-GPU solve.  Two transports behind one interface:
//   NCCL   one process per GPU (the production path).  libnccl is reached through dlopen("libnccl.so.2") so that
//          libpgs.so loads on machines without NCCL and shares the copy a host process (e.g. PyTorch) has loaded.
//   local  `world` solver handles of ONE process, one host thread each, on the same or on different devices: the
//          ranks meet at a host barrier and rank 0 reduces through device copies.  This is what lets the sharded
//          solve run — and be tested — with any number of ranks on a single GPU.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <memory>
#include <string>

namespace pgs {

struct LocalGroup;

class Comm {
 public:
  ~Comm();
  static int unique_id(void* id128, std::string* err);                       // ncclGetUniqueId
  int init(int rank, int world, const void* id128, std::string* err);        // ncclCommInitRank on the current device
  int init_local(int rank, int world, const char* group, std::string* err);  // join the in-process group `group`
  int allreduce_sum(double* dev, size_t n, cudaStream_t st, std::string* err);   // in place
  int allreduce_max(double* dev, size_t n, cudaStream_t st, std::string* err);   // in place
  // Error agreement: every rank passes its own status, all get the first non-zero one (0 if none).  A rank that
  // fails between two collectives calls this instead of leaving the others blocked in the next one.
  int agree(int status, cudaStream_t st, std::string* err);
  int rank = 0, world = 1;
  long long n_collectives = 0, bytes_reduced = 0;
 private:
  int allreduce(double* dev, size_t n, int op, cudaStream_t st, std::string* err);
  void* comm_ = nullptr;
  std::shared_ptr<LocalGroup> local_;
  double* d_flag_ = nullptr;
};

}  // namespace pgs
