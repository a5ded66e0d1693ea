// h2d_lab — how fast do 6 MB of pinned host memory reach the device on this box, as one copy, as the three copies
// pgs_evaluate_from_host issues (q, t, s), and split over 2 / 4 streams?  (Decides whether the end-to-end step can
// be shortened by using more copy engines.)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)
int main() {
  const size_t total = 6000000;
  char *h, *d; CK(cudaMallocHost((void**)&h, total)); CK(cudaMalloc((void**)&d, total));
  cudaStream_t st[4]; for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  auto run = [&](const char* name, int nstreams, int pieces) {
    double best = 1e9, sum = 0; const int reps = 50;
    for (int r = 0; r < reps + 5; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int p = 0; p < pieces; ++p) { const size_t a = total * p / pieces, b = total * (p + 1) / pieces; CK(cudaMemcpyAsync(d + a, h + a, b - a, cudaMemcpyHostToDevice, st[p % nstreams])); }
      for (int s = 0; s < nstreams; ++s) CK(cudaStreamSynchronize(st[s]));
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (r >= 5) { sum += us; if (us < best) best = us; }
    }
    printf("%-28s mean %7.1f us  best %7.1f us  -> %5.1f GB/s\n", name, sum / reps, best, total / (sum / reps) / 1e3);
  };
  run("1 copy, 1 stream", 1, 1);
  run("3 copies, 1 stream", 1, 3);
  run("2 copies, 2 streams", 2, 2);
  run("4 copies, 4 streams", 4, 4);
  run("8 copies, 4 streams", 4, 8);
  const size_t big = 256 << 20; char *hb, *db; CK(cudaMallocHost((void**)&hb, big)); CK(cudaMalloc((void**)&db, big));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, st[0])); CK(cudaStreamSynchronize(st[0]));
  CK(cudaEventRecord(e0, st[0])); CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, st[0])); CK(cudaEventRecord(e1, st[0])); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("256 MiB single copy: %.1f GB/s\n", big / (ms * 1e-3) / 1e9);
  return 0;
}
