// mma_lab — what a shared-memory fed DMMA tile loop reaches on B200, by warp count, fragment shape and fragment-load
// width.  The trailing update of K4 spends its time in exactly this loop (one K chunk of 32 columns from a padded
// shared-memory tile, mma.sync.m8n8k4.f64), so this is the ceiling for sky_update_*:
//   W x (FM x FN)   warps per CTA and 8x8 accumulator fragments per warp (tile = rows 8 FM WM x cols 8 FN WN)
//   lds64 / lds128  one 64-bit fragment load per mma operand, or one 128-bit load feeding two consecutive k-steps
//                   (lane t takes columns 2t, 2t+1 of an 8-column group for A and B alike: a permutation of k)
//   sync            a __syncthreads() after every chunk, as a multi-stage pipeline has
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_lab tools/mma_lab.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int KC = 32;

template <int WM, int WN, int FM, int FN, int WIDE, int SYNC>
__global__ void __launch_bounds__(WM * WN * 32) tile_kernel(double* out, int iters) {
  constexpr int UM = WM * FM * 8, UN = WN * FN * 8;
  constexpr int LDK = WIDE ? KC + 8 : KC + 4;      // 128-bit loads want a row stride of 16 words mod 32, 64-bit ones 8 words mod 32
  extern __shared__ __align__(16) double sm[];
  double* A = sm; double* B = sm + UM * LDK;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = wid % WM, wn = wid / WM;
  for (int i = tid; i < (UM + UN) * LDK; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13);
  __syncthreads();
  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int it = 0; it < iters; ++it) {
    if (WIDE) {
      const double* a_s = A + (wm * FM * 8 + g) * LDK + 2 * t;
      const double* b_s = B + (wn * FN * 8 + g) * LDK + 2 * t;
#pragma unroll
      for (int k = 0; k < KC; k += 8) {
        double2 a[FM], b[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) a[i] = *reinterpret_cast<const double2*>(a_s + (8 * i) * LDK + k);
#pragma unroll
        for (int j = 0; j < FN; ++j) b[j] = *reinterpret_cast<const double2*>(b_s + (8 * j) * LDK + k);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
    } else {
      const double* a_s = A + (wm * FM * 8 + g) * LDK + t;
      const double* b_s = B + (wn * FN * 8 + g) * LDK + t;
#pragma unroll
      for (int k = 0; k < KC; k += 4) {
        double a[FM], b[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) a[i] = a_s[(8 * i) * LDK + k];
#pragma unroll
        for (int j = 0; j < FN; ++j) b[j] = b_s[(8 * j) * LDK + k];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    if (SYNC) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + tid] = s;
}

template <int WM, int WN, int FM, int FN, int WIDE, int SYNC>
static void run(const char* name, int nsm, int ctas_per_sm, double* out) {
  constexpr int UM = WM * FM * 8, UN = WN * FN * 8;
  constexpr int LDK = WIDE ? KC + 8 : KC + 4;
  const size_t smem = sizeof(double) * (UM + UN) * LDK;
  auto k = tile_kernel<WM, WN, FM, FN, WIDE, SYNC>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WM * WN * 32, smem));
  if (occ < ctas_per_sm) { printf("%-44s %d CTAs/SM do not fit (max %d)\n", name, ctas_per_sm, occ); return; }
  const int iters = 3000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<<<nsm * ctas_per_sm, WM * WN * 32, smem>>>(out, iters); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 3; ++r) k<<<nsm * ctas_per_sm, WM * WN * 32, smem>>>(out, iters);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  const double fl = 2.0 * UM * UN * KC * (double)iters * nsm * ctas_per_sm;
  printf("%-44s %d CTAs/SM  tile %3dx%3d  %7.3f ms  %6.2f TFLOP/s\n", name, ctas_per_sm, UM, UN, ms, fl / ms / 1e9);
}

int main() {
  CK(cudaSetDevice(0));
  int nsm = 148; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc((void**)&out, sizeof(double) * nsm * 4 * 1024));
  run<4, 2, 4, 4, 0, 0>("8 warps 4x4 frags lds64", nsm, 1, out);
  run<4, 2, 4, 4, 0, 0>("8 warps 4x4 frags lds64", nsm, 2, out);
  run<4, 2, 4, 4, 0, 1>("8 warps 4x4 frags lds64 sync", nsm, 1, out);
  run<4, 2, 4, 4, 0, 1>("8 warps 4x4 frags lds64 sync", nsm, 2, out);
  run<4, 2, 4, 4, 1, 0>("8 warps 4x4 frags lds128", nsm, 1, out);
  run<4, 2, 4, 4, 1, 0>("8 warps 4x4 frags lds128", nsm, 2, out);
  run<4, 2, 4, 4, 1, 1>("8 warps 4x4 frags lds128 sync", nsm, 1, out);
  run<4, 4, 4, 2, 0, 0>("16 warps 4x2 frags lds64", nsm, 1, out);
  run<4, 4, 4, 2, 0, 1>("16 warps 4x2 frags lds64 sync", nsm, 1, out);
  run<4, 4, 4, 2, 1, 0>("16 warps 4x2 frags lds128", nsm, 1, out);
  run<4, 4, 4, 2, 1, 1>("16 warps 4x2 frags lds128 sync", nsm, 1, out);
  run<4, 4, 4, 4, 0, 0>("16 warps 4x4 frags lds64 (128x128)", nsm, 1, out);
  run<4, 4, 4, 4, 1, 0>("16 warps 4x4 frags lds128 (128x128)", nsm, 1, out);
  run<4, 4, 4, 4, 1, 1>("16 warps 4x4 frags lds128 (128x128) sync", nsm, 1, out);
  run<2, 2, 4, 4, 0, 0>("4 warps 4x4 frags lds64 (64x64)", nsm, 1, out);
  run<2, 2, 4, 4, 0, 0>("4 warps 4x4 frags lds64 (64x64)", nsm, 4, out);
  run<2, 2, 4, 4, 1, 0>("4 warps 4x4 frags lds128 (64x64)", nsm, 4, out);
  run<4, 2, 2, 4, 1, 0>("8 warps 2x4 frags lds128 (64x64)", nsm, 2, out);
  return 0;
}
