// NCCL plumbing for the multi-GPU solve.  libnccl is reached through dlopen("libnccl.so.2") so that libpgs.so
// loads on machines without NCCL and shares the copy a host process (e.g. PyTorch) has already loaded.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <string>

namespace pgs {

class Comm {
 public:
  ~Comm();
  static int unique_id(void* id128, std::string* err);                       // ncclGetUniqueId
  int init(int rank, int world, const void* id128, std::string* err);        // ncclCommInitRank on the current device
  int allreduce_sum(double* dev, size_t n, cudaStream_t st, std::string* err);   // in place
  int allreduce_max(double* dev, size_t n, cudaStream_t st, std::string* err);   // in place
  int rank = 0, world = 1;
  long long n_collectives = 0, bytes_reduced = 0;
 private:
  void* comm_ = nullptr;
};

}  // namespace pgs
