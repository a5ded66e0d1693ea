// How much of the kernel-to-kernel latency on one stream survives a programmatic dependent launch, and what stream
// operations between the two launches do to it.  Kernel A works ~20 us; kernel B stamps %globaltimer when it starts and
// again after griddepcontrol.wait.  Printed per variant: B's start and B's "A is complete" relative to A's last instruction.
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long g_t[4];
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void kA(int early) {
  if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t0 = gt();
  while (gt() - t0 < 20000) { }
  if (threadIdx.x == 0) g_t[0] = gt();
}
__global__ void kC(int us) { const unsigned long long t0 = gt(); while (gt() - t0 < (unsigned long long)us * 1000) { } }
__global__ void kB() {
  if (threadIdx.x == 0) g_t[1] = gt();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) g_t[2] = gt();
}
int main() {
  cudaStream_t s, s2; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t done_before, rec; cudaEventCreateWithFlags(&done_before, cudaEventDisableTiming); cudaEventCreateWithFlags(&rec, cudaEventDisableTiming);
  cudaEventRecord(done_before, s2); cudaStreamSynchronize(s2);
  const char* names[] = {"plain launches", "PDL, nothing between", "PDL, satisfied cudaStreamWaitEvent between", "PDL, cudaEventRecord between",
                         "PDL, record + wait between", "plain, record + wait between", "PDL, A triggers at its start, record + wait between",
                         "PDL, wait on an event of another stream that completes 10 us into A", "same, A triggers at its start",
                         "plain, wait on an event of another stream that completes 10 us into A"};
  cudaEvent_t other; cudaEventCreateWithFlags(&other, cudaEventDisableTiming);
  for (int v = 0; v < 10; ++v) {
    double sum_start = 0, sum_go = 0; int reps = 200;
    for (int r = 0; r < reps + 20; ++r) {
      const bool pdl = v >= 1 && v != 5 && v != 9;
      if (v >= 7) { kC<<<1, 32, 0, s2>>>(10); cudaEventRecord(other, s2); }
      kA<<<1, 32, 0, s>>>((v == 6 || v == 8) ? 1 : 0);
      if (v >= 7) cudaStreamWaitEvent(s, other, 0);
      if (v == 3 || v == 4 || v == 5 || v == 6) cudaEventRecord(rec, s);
      if (v == 2 || v == 4 || v == 5 || v == 6) cudaStreamWaitEvent(s, done_before, 0);
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, kB);
      cudaStreamSynchronize(s);
      unsigned long long t[4]; cudaMemcpyFromSymbol(t, g_t, sizeof(t));
      if (r >= 20) { sum_start += (double)((long long)t[1] - (long long)t[0]); sum_go += (double)((long long)t[2] - (long long)t[0]); }
    }
    printf("%-55s B starts %+7.2f us, B past its wait %+7.2f us after A's last instruction\n", names[v], sum_start / reps / 1e3, sum_go / reps / 1e3);
  }
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
