// sweep_lab — stand-alone tuning harness for K1 (the residual + Jacobian sweep).  Not part of libpgs.so.
//
// Builds a synthetic graph with BASELINE config 3's shape (100k nodes, fan-out-3 odometry, 50k loop edges with a
// gap of 50..2000 keyframes), then times variants of the sweep kernel that differ only in how the work is
// scheduled and how wide the global loads/stores are:
//   V      planes interleaved per lane: 1 = [plane][32] (64-bit accesses), 2 = [plane/2][32][2] (128-bit),
//          4 = [plane/4][32][4] (256-bit, sm_100 LDG/STG.256)
//   SCHED  0 persistent grid, tiles strided over warps; 1 one tile per warp; 2 persistent grid, tiles handed out
//          by an atomic counter
//   BLK / MINB  block size and __launch_bounds__ min blocks
// plus pure write / copy kernels of the same byte volume as practical ceilings for a store-dominated kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o sweep_lab tools/sweep_lab.cu
//   ./sweep_lab [reps]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../solve_keyframe_pose_graph_b200/csrc/pgs_kernels.cuh"

using namespace pgs;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

template <int V> __device__ __forceinline__ void stv(double* p, const double* v);
template <> __device__ __forceinline__ void stv<1>(double* p, const double* v) { __stcs(p, v[0]); }
template <> __device__ __forceinline__ void stv<2>(double* p, const double* v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1])); }
template <> __device__ __forceinline__ void stv<4>(double* p, const double* v) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
template <int V> __device__ __forceinline__ void ldv(const double* p, double* v);
template <> __device__ __forceinline__ void ldv<1>(const double* p, double* v) { v[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<2>(const double* p, double* v) { const double2 t = __ldg(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y; }
template <> __device__ __forceinline__ void ldv<4>(const double* p, double* v) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <int V> __device__ __forceinline__ P7 load_pose_v(const double* __restrict__ pose, int i) {
  if (V == 4) {
    double a[4], b[4];
    ldv<4>(pose + 8 * (size_t)i, a); ldv<4>(pose + 8 * (size_t)i + 4, b);
    P7 r; r.q.x = a[0]; r.q.y = a[1]; r.q.z = a[2]; r.q.w = a[3]; r.tx = b[0]; r.ty = b[1]; r.tz = b[2];
    return r;
  }
  return load_pose(pose, i);
}

__host__ __device__ constexpr int pad_to(int x, int v) { return (x + v - 1) / v * v; }

struct LabArgs {
  const double* pose; const double* sw;
  const int2* o_idx; const double* o_obs; int n_odom;
  const int2* l_idx; const double* l_obs; int n_loop;
  double* o_r; double* o_J; double* l_r; double* l_J;
  double* cost_tile;   // one partial per tile (fixed summation order whatever the schedule)
  int* counter;
};

template <int V, int BLK, int MINB, int SCHED>
__global__ void __launch_bounds__(BLK, MINB) sweep_lab_kernel(LabArgs A) {
  constexpr int ORP = pad_to(OD_R, V), OJP = pad_to(OD_J, V), LRP = pad_to(LP_R, V), LJP = pad_to(LP_J, V);
  const int lane = threadIdx.x & 31;
  const int wpb = BLK / 32;
  const int gwarp = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpb;
  const int To = (A.n_odom + TILE - 1) / TILE, Tl = (A.n_loop + TILE - 1) / TILE;
  const int T = To + Tl;
  int tile = gwarp;
  if (SCHED == 2) { if (lane == 0) tile = atomicAdd(A.counter, 1); tile = __shfl_sync(0xffffffffu, tile, 0); }
  while (tile < T) {
    double cost = 0.0;
    if (tile < To) {
      const int e = tile * TILE + lane;
      if (e < A.n_odom) {
        const int2 ij = __ldg(A.o_idx + e);
        double ob[8];
        const double* obp = A.o_obs + (size_t)tile * (OBS * TILE);
#pragma unroll
        for (int g = 0; g < 8 / V; ++g) ldv<V>(obp + (g * TILE + lane) * V, ob + g * V);
        const Q4 qo{ob[0], ob[1], ob[2], ob[3]};
        const double w = ob[7];
        const P7 p1 = load_pose_v<V>(A.pose, ij.x), p2 = load_pose_v<V>(A.pose, ij.y);
        double ev[6], Rt[9], Ba[9], Bv[9], M[9];
        sixdof_core<true>(p1, p2, qo, ob[4], ob[5], ob[6], ev, Rt, Ba, Bv, M);
        double r[ORP];
#pragma unroll
        for (int i = 0; i < ORP; ++i) r[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { r[i] = ev[i] * w; cost += r[i] * r[i]; }
        double* rp = A.o_r + (size_t)tile * (ORP * TILE);
#pragma unroll
        for (int g = 0; g < ORP / V; ++g) stv<V>(rp + (g * TILE + lane) * V, r + g * V);
        double J[OJP];
        const double w2 = 2.0 * w;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            J[i * 6 + j] = -w2 * Ba[3 * i + j];
            J[i * 6 + 3 + j] = w * Rt[3 * i + j];
            J[(3 + i) * 6 + j] = w2 * M[3 * i + j];
            J[(3 + i) * 6 + 3 + j] = 0.0;
            J[36 + i * 6 + j] = w2 * Bv[3 * i + j];
            J[36 + i * 6 + 3 + j] = -w * Rt[3 * i + j];
            J[36 + (3 + i) * 6 + j] = -w2 * M[3 * i + j];
            J[36 + (3 + i) * 6 + 3 + j] = 0.0;
          }
        double* Jp = A.o_J + (size_t)tile * (OJP * TILE);
#pragma unroll
        for (int g = 0; g < OJP / V; ++g) stv<V>(Jp + (g * TILE + lane) * V, J + g * V);
      }
    } else {
      const int lt = tile - To;
      const int e = lt * TILE + lane;
      if (e < A.n_loop) {
        const int2 ij = __ldg(A.l_idx + e);
        double ob[8];
        const double* obp = A.l_obs + (size_t)lt * (OBS * TILE);
#pragma unroll
        for (int g = 0; g < 8 / V; ++g) ldv<V>(obp + (g * TILE + lane) * V, ob + g * V);
        const Q4 qo{ob[0], ob[1], ob[2], ob[3]};
        const double s = __ldg(A.sw + e);
        const P7 p1 = load_pose_v<V>(A.pose, ij.x), p2 = load_pose_v<V>(A.pose, ij.y);
        double ev[6], Rt[9], Ba[9], Bv[9], M[9];
        sixdof_core<true>(p1, p2, qo, ob[4], ob[5], ob[6], ev, Rt, Ba, Bv, M);
        double r[LRP];
#pragma unroll
        for (int i = 0; i < LRP; ++i) r[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { r[i] = s * ev[i]; cost += r[i] * r[i]; }
        r[6] = s * (1.0 - s); cost += r[6] * r[6];
        double* rp = A.l_r + (size_t)lt * (LRP * TILE);
#pragma unroll
        for (int g = 0; g < LRP / V; ++g) stv<V>(rp + (g * TILE + lane) * V, r + g * V);
        double J[LJP];
#pragma unroll
        for (int i = 0; i < LJP; ++i) J[i] = 0.0;
        const double s2 = 2.0 * s;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            J[i * 6 + j] = -s2 * Ba[3 * i + j];
            J[i * 6 + 3 + j] = s * Rt[3 * i + j];
            J[(3 + i) * 6 + j] = s2 * M[3 * i + j];
            J[42 + i * 6 + j] = s2 * Bv[3 * i + j];
            J[42 + i * 6 + 3 + j] = -s * Rt[3 * i + j];
            J[42 + (3 + i) * 6 + j] = -s2 * M[3 * i + j];
          }
#pragma unroll
        for (int i = 0; i < 6; ++i) J[84 + i] = ev[i];
        J[90] = 1.0 - 2.0 * s;
        double* Jp = A.l_J + (size_t)lt * (LJP * TILE);
#pragma unroll
        for (int g = 0; g < LJP / V; ++g) stv<V>(Jp + (g * TILE + lane) * V, J + g * V);
      }
    }
    cost = warp_sum(cost);
    if (lane == 0) A.cost_tile[tile] = cost;
    if (SCHED == 0) tile += nwarps;
    else if (SCHED == 1) break;
    else { if (lane == 0) tile = atomicAdd(A.counter, 1); tile = __shfl_sync(0xffffffffu, tile, 0); }
  }
}

// ---- practical ceilings: pure streaming write / copy of the same byte volume
template <int V> __global__ void __launch_bounds__(256) write_kernel(double* p, size_t n) {   // n doubles, multiple of V
  const double v[4] = {1.0, 2.0, 3.0, 4.0};
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * V; i < n; i += (size_t)gridDim.x * blockDim.x * V) stv<V>(p + i, v);
}
template <int V> __global__ void __launch_bounds__(256) copy_kernel(double* dst, const double* src, size_t n) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * V; i < n; i += (size_t)gridDim.x * blockDim.x * V) { double v[4]; ldv<V>(src + i, v); stv<V>(dst + i, v); }
}
__global__ void flush_kernel(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

struct Lab {
  int N = 100000, Eo = 0, El = 50000;
  std::vector<int2> oidx, lidx;
  double *pose, *sw, *or_, *oJ, *lr, *lJ, *cost_tile, *flush, *oobs[5], *lobs[5];
  int2 *d_oidx, *d_lidx; int* counter;
  size_t flush_n = (size_t)48 << 20;
  cudaStream_t st; cudaEvent_t e0, e1;
  int nsm = 148;
};

template <int V, int BLK, int MINB, int SCHED>
static void run(Lab& L, int reps, const char* name, long long bytes, bool flush) {
  LabArgs A;
  A.pose = L.pose; A.sw = L.sw; A.o_idx = L.d_oidx; A.l_idx = L.d_lidx; A.o_obs = L.oobs[V]; A.l_obs = L.lobs[V]; A.n_odom = L.Eo; A.n_loop = L.El;
  A.o_r = L.or_; A.o_J = L.oJ; A.l_r = L.lr; A.l_J = L.lJ; A.cost_tile = L.cost_tile; A.counter = L.counter;
  const int T = (L.Eo + 31) / 32 + (L.El + 31) / 32, wpb = BLK / 32;
  int occ = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_lab_kernel<V, BLK, MINB, SCHED>, BLK, 0));
  const int grid = SCHED == 1 ? (T + wpb - 1) / wpb : std::min(L.nsm * occ, (T + wpb - 1) / wpb);
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, sweep_lab_kernel<V, BLK, MINB, SCHED>));
  double sum = 0, mn = 1e30;
  for (int i = 0; i < reps + 3; ++i) {
    if (flush) flush_kernel<<<1184, 256, 0, L.st>>>(L.flush, L.flush_n, (double)i);
    if (SCHED == 2) CK(cudaMemsetAsync(L.counter, 0, sizeof(int), L.st));
    CK(cudaEventRecord(L.e0, L.st));
    sweep_lab_kernel<V, BLK, MINB, SCHED><<<grid, BLK, 0, L.st>>>(A);
    CK(cudaEventRecord(L.e1, L.st));
    CK(cudaEventSynchronize(L.e1));
    float ms; CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
    if (i >= 3) { sum += ms; mn = std::min(mn, (double)ms); }
  }
  CK(cudaGetLastError());
  // checksum of the per-tile costs (all variants must agree)
  std::vector<double> ct(T); CK(cudaMemcpy(ct.data(), L.cost_tile, sizeof(double) * T, cudaMemcpyDeviceToHost));
  double c = 0; for (double x : ct) c += x;
  const double mean = sum / reps;
  printf("%-34s V=%d BLK=%d MINB=%d SCHED=%d regs=%3d occ=%d grid=%5d  mean %7.2f us  min %7.2f us  -> %6.1f GB/s mean, %6.1f GB/s best  cost=%.9e %s\n", name, V, BLK, MINB, SCHED,
         fa.numRegs, occ, grid, mean * 1e3, mn * 1e3, bytes / (mean * 1e-3) / 1e9, bytes / (mn * 1e-3) / 1e9, 0.5 * c, flush ? "flush" : "warm");
  fflush(stdout);
}

template <int V> static void run_write(Lab& L, int reps, long long bytes, int ctas_per_sm) {
  const size_t n = (size_t)bytes / 8 / 4 * 4;
  double sum = 0, mn = 1e30;
  for (int i = 0; i < reps + 3; ++i) {
    flush_kernel<<<1184, 256, 0, L.st>>>(L.flush, L.flush_n, (double)i);
    CK(cudaEventRecord(L.e0, L.st));
    write_kernel<V><<<L.nsm * ctas_per_sm, 256, 0, L.st>>>(L.oJ, n);
    CK(cudaEventRecord(L.e1, L.st)); CK(cudaEventSynchronize(L.e1));
    float ms; CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
    if (i >= 3) { sum += ms; mn = std::min(mn, (double)ms); }
  }
  printf("pure write  V=%d ctas/SM=%d  %lld B  mean %7.2f us  min %7.2f us -> %6.1f GB/s mean, %6.1f best\n", V, ctas_per_sm, bytes, sum / reps * 1e3, mn * 1e3, bytes / (sum / reps * 1e-3) / 1e9, bytes / (mn * 1e-3) / 1e9);
}
template <int V> static void run_copy(Lab& L, int reps, long long bytes, int ctas_per_sm) {
  const size_t n = (size_t)bytes / 2 / 8 / 4 * 4;   // n doubles read + n doubles written = bytes
  double sum = 0, mn = 1e30;
  for (int i = 0; i < reps + 3; ++i) {
    flush_kernel<<<1184, 256, 0, L.st>>>(L.flush, L.flush_n, (double)i);
    CK(cudaEventRecord(L.e0, L.st));
    copy_kernel<V><<<L.nsm * ctas_per_sm, 256, 0, L.st>>>(L.oJ, L.flush, n);
    CK(cudaEventRecord(L.e1, L.st)); CK(cudaEventSynchronize(L.e1));
    float ms; CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
    if (i >= 3) { sum += ms; mn = std::min(mn, (double)ms); }
  }
  printf("copy        V=%d ctas/SM=%d  %lld B  mean %7.2f us  min %7.2f us -> %6.1f GB/s mean, %6.1f best\n", V, ctas_per_sm, bytes, sum / reps * 1e3, mn * 1e3, bytes / (sum / reps * 1e-3) / 1e9, bytes / (mn * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 30;
  Lab L;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&L.nsm, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking)); CK(cudaEventCreate(&L.e0)); CK(cudaEventCreate(&L.e1));
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const int N = L.N;
  for (int u = 1; u < N; ++u) for (int f = 3; f >= 1; --f) if (u - f >= 0) L.oidx.push_back(make_int2(u, u - f));
  L.Eo = (int)L.oidx.size();
  for (int e = 0; e < L.El; ++e) { const int gap = 50 + (int)(rng() % 1951); const int b = (int)(rng() % (N - gap)); L.lidx.push_back(make_int2(b, b + gap)); }
  std::sort(L.lidx.begin(), L.lidx.end(), [](int2 a, int2 b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
  std::vector<double> pose((size_t)N * 8);
  for (int i = 0; i < N; ++i) {
    double q[4] = {0.05 * U(rng), 0.05 * U(rng), 0.3 * U(rng), 1.0}; const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) pose[8 * (size_t)i + k] = q[k] / n;
    pose[8 * (size_t)i + 4] = i + U(rng); pose[8 * (size_t)i + 5] = U(rng); pose[8 * (size_t)i + 6] = U(rng); pose[8 * (size_t)i + 7] = 0;
  }
  auto make_obs = [&](int E, std::vector<double>* out /*[5]*/) {
    const int T = (E + 31) / 32;
    std::vector<double> raw((size_t)E * 8);
    for (int e = 0; e < E; ++e) {
      double q[4] = {0.05 * U(rng), 0.05 * U(rng), 0.05 * U(rng), 1.0}; const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      for (int k = 0; k < 4; ++k) raw[8 * (size_t)e + k] = q[k] / n;
      raw[8 * (size_t)e + 4] = -1.0 + 0.1 * U(rng); raw[8 * (size_t)e + 5] = 0.1 * U(rng); raw[8 * (size_t)e + 6] = 0.1 * U(rng); raw[8 * (size_t)e + 7] = 0.5 + 0.4 * U(rng);
    }
    for (int V : {1, 2, 4}) {
      out[V].assign((size_t)T * 8 * 32, 0.0);
      for (int e = 0; e < E; ++e) for (int k = 0; k < 8; ++k) out[V][(size_t)(e / 32) * 256 + ((k / V) * 32 + (e % 32)) * V + (k % V)] = raw[8 * (size_t)e + k];
    }
  };
  std::vector<double> oo[5], lo[5];
  make_obs(L.Eo, oo); make_obs(L.El, lo);
  const int To = (L.Eo + 31) / 32, Tl = (L.El + 31) / 32;
  auto up = [&](double** d, const std::vector<double>& h) { CK(cudaMalloc((void**)d, sizeof(double) * h.size())); CK(cudaMemcpy(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice)); };
  up(&L.pose, pose);
  for (int V : {1, 2, 4}) { up(&L.oobs[V], oo[V]); up(&L.lobs[V], lo[V]); }
  std::vector<double> sw(L.El, 0.99); up(&L.sw, sw);
  CK(cudaMalloc((void**)&L.d_oidx, sizeof(int2) * L.Eo)); CK(cudaMemcpy(L.d_oidx, L.oidx.data(), sizeof(int2) * L.Eo, cudaMemcpyHostToDevice));
  CK(cudaMalloc((void**)&L.d_lidx, sizeof(int2) * L.El)); CK(cudaMemcpy(L.d_lidx, L.lidx.data(), sizeof(int2) * L.El, cudaMemcpyHostToDevice));
  CK(cudaMalloc((void**)&L.or_, sizeof(double) * (size_t)To * 8 * 32)); CK(cudaMalloc((void**)&L.oJ, sizeof(double) * (size_t)To * 72 * 32 + (128 << 20)));   // + slack: also the target of the pure write/copy kernels
  CK(cudaMalloc((void**)&L.lr, sizeof(double) * (size_t)Tl * 8 * 32)); CK(cudaMalloc((void**)&L.lJ, sizeof(double) * (size_t)Tl * 92 * 32));
  CK(cudaMalloc((void**)&L.cost_tile, sizeof(double) * (To + Tl))); CK(cudaMalloc((void**)&L.counter, sizeof(int)));
  CK(cudaMalloc((void**)&L.flush, sizeof(double) * L.flush_n));
  const long long bytes = 56LL * N + 72LL * L.Eo + 72LL * L.El + 624LL * L.Eo + 784LL * L.El;
  printf("sweep_lab: N=%d Eo=%d El=%d  algorithmic bytes %lld (%.1f B/edge), %d SMs, reps %d\n", N, L.Eo, L.El, bytes, (double)bytes / (L.Eo + L.El), L.nsm, reps);

  for (int c : {2, 4, 8}) { run_write<1>(L, reps, bytes, c); run_write<2>(L, reps, bytes, c); run_write<4>(L, reps, bytes, c); }
  for (int c : {4, 8}) { run_copy<2>(L, reps, bytes, c); run_copy<4>(L, reps, bytes, c); }

  for (int fl = 1; fl >= 0; --fl) {
    const bool f = fl;
    run<1, 256, 1, 0>(L, reps, "baseline persistent strided", bytes, f);
    run<1, 256, 1, 1>(L, reps, "one tile per warp", bytes, f);
    run<1, 256, 1, 2>(L, reps, "atomic tiles", bytes, f);
    run<1, 128, 1, 1>(L, reps, "one tile per warp", bytes, f);
    run<1, 128, 1, 2>(L, reps, "atomic tiles", bytes, f);
    run<2, 256, 1, 0>(L, reps, "128-bit persistent strided", bytes, f);
    run<2, 256, 1, 1>(L, reps, "128-bit one tile per warp", bytes, f);
    run<2, 256, 1, 2>(L, reps, "128-bit atomic tiles", bytes, f);
    run<2, 128, 1, 1>(L, reps, "128-bit one tile per warp", bytes, f);
    run<2, 128, 1, 2>(L, reps, "128-bit atomic tiles", bytes, f);
    run<2, 256, 3, 2>(L, reps, "128-bit atomic tiles", bytes, f);
    run<2, 128, 6, 2>(L, reps, "128-bit atomic tiles", bytes, f);
    run<4, 256, 1, 0>(L, reps, "256-bit persistent strided", bytes, f);
    run<4, 256, 1, 1>(L, reps, "256-bit one tile per warp", bytes, f);
    run<4, 256, 1, 2>(L, reps, "256-bit atomic tiles", bytes, f);
    run<4, 128, 1, 1>(L, reps, "256-bit one tile per warp", bytes, f);
    run<4, 128, 1, 2>(L, reps, "256-bit atomic tiles", bytes, f);
    run<4, 256, 2, 2>(L, reps, "256-bit atomic tiles", bytes, f);
    run<4, 256, 3, 2>(L, reps, "256-bit atomic tiles", bytes, f);
    run<4, 128, 6, 2>(L, reps, "256-bit atomic tiles", bytes, f);
    run<4, 64, 8, 2>(L, reps, "256-bit atomic tiles", bytes, f);
  }
  return 0;
}
