// diag_lab — times the phases of K4's one-CTA diagonal-block kernel (sky_diag_kernel) with clock64 stamps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSKY_DIAG_CLOCKS -o diag_lab tools/diag_lab.cu
#include "../solve_keyframe_pose_graph_b200/csrc/pgs_skyline.cu"

#include <random>
using namespace pgs;
#define CKL(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;   // 0: sky_diag_kernel, 1: sky_diag2_kernel
  const int n = PW;
  // one panel: rows 0..95 each store columns [0, 96); SPD matrix A = M M^T + 96 I
  std::vector<double> M(n * n), A(n * n, 0.0);
  std::mt19937_64 rng(7); std::uniform_real_distribution<double> U(-1, 1);
  for (auto& x : M) x = U(rng);
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = i == j ? 96.0 : 0.0; for (int k = 0; k < n; ++k) s += M[i * n + k] * M[j * n + k]; A[i * n + j] = s; }
  std::vector<long long> ptr(n + 2); std::vector<int> start(n + 1, 0);
  for (int i = 0; i <= n + 1; ++i) ptr[i] = (long long)i * n;
  double *val, *dinv; long long* dptr; int *dstart, *fail;
  CKL(cudaMalloc((void**)&val, sizeof(double) * n * (n + 1))); CKL(cudaMalloc((void**)&dinv, sizeof(double) * n * n));
  CKL(cudaMalloc((void**)&dptr, sizeof(long long) * (n + 2))); CKL(cudaMalloc((void**)&dstart, sizeof(int) * (n + 1))); CKL(cudaMalloc((void**)&fail, 4));
  CKL(cudaMemcpy(dptr, ptr.data(), sizeof(long long) * (n + 2), cudaMemcpyHostToDevice)); CKL(cudaMemcpy(dstart, start.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
  CKL(cudaMemset(fail, 0, 4));
  CKL(cudaFuncSetAttribute(sky_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_DIAG));
  CKL(cudaFuncSetAttribute(sky_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_DIAG));
  cudaEvent_t e0, e1; CKL(cudaEventCreate(&e0)); CKL(cudaEventCreate(&e1));
  float best = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    CKL(cudaMemcpy(val, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CKL(cudaEventRecord(e0));
    if (mode == 1) sky_diag2_kernel<<<1, DG2_THREADS, SM_DIAG>>>(0, n, 0, dptr, dstart, val, dinv, fail);
    else sky_diag_kernel<<<1, 256, SM_DIAG>>>(0, n, 0, dptr, dstart, val, dinv, fail);
    CKL(cudaEventRecord(e1)); CKL(cudaEventSynchronize(e1));
    float ms; CKL(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  CKL(cudaGetLastError());
  long long clk[64]; CKL(cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof(clk)));
  printf("diag kernel (mode %d: %s): %.2f us (events, best of 5)\n", mode, mode ? "panel / update warps pipelined; 'phase A' = panel warps' step, 'phase B' = update warps' step, both since the previous stamp of thread 0 of their group" : "two barrier phases per step", best * 1e3);
  printf("load: %lld cycles\n", clk[1] - clk[0]);
  for (int I = 0; I < PW / 8; ++I) printf("step %2d: phase A %6lld  phase B %6lld cycles\n", I, clk[2 + 2 * I] - clk[1 + 2 * I], clk[3 + 2 * I] - clk[2 + 2 * I]);
  printf("finish: %lld  store: %lld cycles; total %lld cycles\n", clk[2 + 2 * (PW / 8)] - clk[1 + 2 * (PW / 8)], clk[3 + 2 * (PW / 8)] - clk[2 + 2 * (PW / 8)], clk[3 + 2 * (PW / 8)] - clk[0]);
  // check: X A X^T == I  (X = L^-1; the kernel stores only X)
  std::vector<double> Xh(n * n); int hf = 0;
  CKL(cudaMemcpy(Xh.data(), dinv, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
  CKL(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
  std::vector<double> XA(n * n, 0.0);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k <= i; ++k) s += Xh[i * n + k] * (k >= j ? A[k * n + j] : A[j * n + k]); XA[i * n + j] = s; }
  double e1m = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += XA[i * n + k] * Xh[j * n + k]; e1m = std::max(e1m, std::fabs(s - (i == j ? 1.0 : 0.0))); }
  printf("fail=%d  max|X A X^T - I| = %.3e\n", hf, e1m);
  return 0;
}
