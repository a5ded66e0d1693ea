// fp64_lab — measures what the B200 FP64 pipes deliver, to decide how K4's rank-96 trailing update should be written:
//   (1) register-resident DFMA throughput (independent accumulators, no memory),
//   (2) register-resident DMMA throughput (mma.sync.aligned.m8n8k4 f64),
//   (3) a shared-memory fed 128x64 tile SYRK-like loop, DFMA (8x4 micro-tiles) vs DMMA (warp tile 32x32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_lab tools/fp64_lab.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC> __global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double x, double y) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC> __global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double x, double y) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  double a = x + threadIdx.x * 1e-9, b = y;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-issue latency: one warp, NCH independent accumulator chains, clock64 around a long dependent sequence
template <int NCH> __global__ void dmma_latency_kernel(double* out, long long* cyc, int iters, double x, double y) {
  double c0[NCH], c1[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  const double a = x + threadIdx.x * 1e-9, b = y;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma(c0[i], c1[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += c0[i] + c1[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dfma_latency_kernel(double* out, long long* cyc, int iters, double x, double y) {
  double acc = threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) acc = fma(acc, x, y);
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rsqrt_latency_kernel(double* out, long long* cyc, int iters, double x) {
  double acc = x + threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) acc = rsqrt(acc) + 1.5;
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <typename F> static double time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  CK(cudaSetDevice(0));
  int nsm = 148; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc((void**)&out, sizeof(double) * nsm * 8 * 256));
  const int iters = 4096;
  {
    long long* cyc; CK(cudaMalloc((void**)&cyc, 8)); long long h = 0;
    dmma_latency_kernel<1><<<1, 32>>>(out, cyc, 2048, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DMMA dependent chain, 1 warp: %.1f cycles per mma\n", (double)h / 2048);
    dmma_latency_kernel<2><<<1, 32>>>(out, cyc, 2048, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DMMA 2 chains, 1 warp: %.1f cycles per mma\n", (double)h / 4096);
    dmma_latency_kernel<4><<<1, 32>>>(out, cyc, 2048, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DMMA 4 chains, 1 warp: %.1f cycles per mma\n", (double)h / 8192);
    dmma_latency_kernel<8><<<1, 32>>>(out, cyc, 2048, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DMMA 8 chains, 1 warp: %.1f cycles per mma\n", (double)h / 16384);
    dmma_latency_kernel<4><<<1, 256>>>(out, cyc, 2048, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DMMA 4 chains, 8 warps (one CTA): %.1f cycles per mma per warp\n", (double)h / 8192);
    dfma_latency_kernel<<<1, 32>>>(out, cyc, 4096, 0.999, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DFMA dependent chain: %.1f cycles per fma\n", (double)h / 4096);
    rsqrt_latency_kernel<<<1, 32>>>(out, cyc, 4096, 2.0); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("rsqrt(double)+add dependent chain: %.1f cycles per iteration\n", (double)h / 4096);
  }
  for (int ctas : {1, 2, 4, 8}) {
    {
      const double ms = time_ms([&] { dfma_kernel<16><<<nsm * ctas, 256>>>(out, iters, 0.999, 1e-3); }, 5);
      const double fl = 2.0 * 16 * iters * 256.0 * nsm * ctas;
      printf("DFMA  16 acc/thread, %d CTAs/SM x 256 thr: %8.3f ms  %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
    }
    {
      const double ms = time_ms([&] { dmma_kernel<8><<<nsm * ctas, 256>>>(out, iters, 0.999, 1e-3); }, 5);
      const double fl = 2.0 * 8 * 256 * iters * 8.0 * nsm * ctas;   // 8 mma x (8x8x4 FMA) per warp-iteration, 8 warps
      printf("DMMA   8 acc/thread, %d CTAs/SM x 256 thr: %8.3f ms  %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
    }
    {
      const double ms = time_ms([&] { dmma_kernel<16><<<nsm * ctas, 256>>>(out, iters, 0.999, 1e-3); }, 5);
      const double fl = 2.0 * 16 * 256 * iters * 8.0 * nsm * ctas;
      printf("DMMA  16 acc/thread, %d CTAs/SM x 256 thr: %8.3f ms  %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
    }
  }
  return 0;
}
