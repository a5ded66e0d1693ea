// Batched evaluation of the reference's alternative (switched-off) edge functors on the device:
// include/pgs_fourdof.h; math in pgs_fourdof.cuh; reference src/CeresResidues.h:252-546.
//
// One lane per edge, one warp per tile of 32 edges, grid-stride over tiles.  The per-edge outputs are row-major
// blocks of 6..91 doubles, i.e. a lane's values are 48..728 B apart in the caller's layout; every warp therefore
// stages its tile in shared memory (odd row pitch: no bank conflicts beyond the two passes a 64-bit access takes)
// and writes it out as one contiguous run of 32 x (rows x cols) doubles, so every store instruction covers whole
// 256-B segments.  Bound: HBM on the Jacobian stores (576 / 728 / 256 B per edge against 72-104 B of inputs), as
// for the live sweep; the atan2 chain adds ~1.5 kflop per edge, still far under the FP64 ridge.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pgs.h"
#include "../../include/pgs_fourdof.h"
#include "pgs_fourdof.cuh"
#include "pgs_solver.h"   // DBuf

namespace pgs {
namespace fourdof {

template <int KIND> struct Shape;
template <> struct Shape<PGS_FOURDOF_ERROR>  { static constexpr int NR = 6, NC = 12; };
template <> struct Shape<PGS_FOURDOF_SWITCH> { static constexpr int NR = 7, NC = 13; };
template <> struct Shape<PGS_FOURDOF_QIN>    { static constexpr int NR = 4, NC = 8; };

struct Args {
  int n_edges;
  const double* __restrict__ rot; const double* __restrict__ t;
  const int* __restrict__ c1; const int* __restrict__ c2;
  const double* __restrict__ obs_rot; const double* __restrict__ obs_t; const double* __restrict__ weight; const double* __restrict__ sw;
  double* __restrict__ r; double* __restrict__ J; double* __restrict__ cost_tile;
};

constexpr int WARPS = 2;   // 2 x 32 x 92 doubles = 46 KB of staging for the widest block: fits the default 48 KB window

__device__ __forceinline__ void load4(const double* __restrict__ p, double* o) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void load3(const double* __restrict__ p, double* o) { o[0] = __ldg(p); o[1] = __ldg(p + 1); o[2] = __ldg(p + 2); }

template <int KIND>
__device__ __forceinline__ void eval_edge(const Args& A, int e, Dual<Shape<KIND>::NC>* res) {
  const int a = __ldg(A.c1 + e), b = __ldg(A.c2 + e);
  double t1[3], t2[3], ot[3];
  load3(A.t + 3 * (size_t)a, t1); load3(A.t + 3 * (size_t)b, t2); load3(A.obs_t + 3 * (size_t)e, ot);
  if constexpr (KIND == PGS_FOURDOF_QIN) {
    double ob[3]; load3(A.obs_rot + 3 * (size_t)e, ob);                         // relative_yaw, pitch_i, roll_i
    qin_four_dof(__ldg(A.rot + 3 * (size_t)a), t1, __ldg(A.rot + 3 * (size_t)b), t2, ot, ob[0], ob[1], ob[2], res);
  } else {
    double q1[4], q2[4], oq[4];
    load4(A.rot + 4 * (size_t)a, q1); load4(A.rot + 4 * (size_t)b, q2); load4(A.obs_rot + 4 * (size_t)e, oq);
    if constexpr (KIND == PGS_FOURDOF_SWITCH) four_dof_error<true, 13>(q1, t1, q2, t2, oq, ot, 1.0, __ldg(A.sw + e), res);
    else four_dof_error<false, 12>(q1, t1, q2, t2, oq, ot, __ldg(A.weight + e), 0.0, res);
  }
}

template <int KIND>
__global__ void __launch_bounds__(32 * WARPS) fourdof_kernel(Args A) {
  constexpr int NR = Shape<KIND>::NR, NC = Shape<KIND>::NC, LD = NR * NC + 1;
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = stage_all + (size_t)warp * 32 * LD;
  const int tiles = (A.n_edges + 31) / 32;
  for (int tile = blockIdx.x * WARPS + warp; tile < tiles; tile += gridDim.x * WARPS) {
    const int e0 = tile * 32, e = e0 + lane;
    const int nvalid = min(32, A.n_edges - e0);
    Dual<NC> res[NR];
    double c = 0.0;
    if (lane < nvalid) {
      eval_edge<KIND>(A, e, res);
#pragma unroll
      for (int i = 0; i < NR; ++i) { c += res[i].a * res[i].a; stage[lane * LD + i] = res[i].a; }
    }
    __syncwarp();
    for (int k = lane; k < nvalid * NR; k += 32) __stcs(A.r + (size_t)e0 * NR + k, stage[(k / NR) * LD + (k % NR)]);
    __syncwarp();
    if (A.J) {
      if (lane < nvalid) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
#pragma unroll
          for (int j = 0; j < NC; ++j) stage[lane * LD + i * NC + j] = res[i].v[j];
        }
      }
      __syncwarp();
      for (int k = lane; k < nvalid * NR * NC; k += 32) __stcs(A.J + (size_t)e0 * NR * NC + k, stage[(k / (NR * NC)) * LD + (k % (NR * NC))]);
      __syncwarp();
    }
    // 1/2 sum r^2 of the tile: xor butterfly, the same order every run
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) A.cost_tile[tile] = 0.5 * c;
  }
}

}  // namespace fourdof
}  // namespace pgs

struct pgs_fourdof_s {
  int dev = 0, sm_count = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  pgs::DBuf<double> rot, t, obs_rot, obs_t, weight, sw, r, J, cost_tile;
  pgs::DBuf<int> c1, c2;
  std::vector<double> h_cost;
  double ms_kernel = 0;
  std::string err;
};

#define FCU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { h->err = std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #x; return e__ == cudaErrorMemoryAllocation ? PGS_ERR_OUT_OF_MEMORY : PGS_ERR_CUDA; } } while (0)

extern "C" {

int pgs_fourdof_create(int32_t device, pgs_fourdof_handle* out) {
  if (!out) return PGS_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return PGS_ERR_CUDA;   // no CPU fallback
  pgs_fourdof_s* h = new pgs_fourdof_s();
  h->dev = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess ||
      cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    delete h; return PGS_ERR_CUDA;
  }
  *out = h;
  return PGS_OK;
}

void pgs_fourdof_destroy(pgs_fourdof_handle h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* pgs_fourdof_last_error(pgs_fourdof_handle h) { return h ? h->err.c_str() : "null handle"; }

int pgs_fourdof_evaluate(pgs_fourdof_handle h, const pgs_fourdof_input* in, double* r, double* J, double* cost) try {
  using namespace pgs::fourdof;
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (!in || in->kind < PGS_FOURDOF_ERROR || in->kind > PGS_FOURDOF_QIN || in->n_nodes < 0 || in->n_edges < 0) { h->err = "pgs_fourdof_evaluate: bad kind or negative size"; return PGS_ERR_INVALID_ARGUMENT; }
  if (cost) *cost = 0.0;
  if (in->n_edges == 0) return PGS_OK;
  if (!in->rot || !in->t || !in->c1 || !in->c2 || !in->obs_rot || !in->obs_t || !r || (in->kind == PGS_FOURDOF_ERROR && !in->weight) ||
      (in->kind == PGS_FOURDOF_SWITCH && !in->sw)) { h->err = "pgs_fourdof_evaluate: null array"; return PGS_ERR_INVALID_ARGUMENT; }
  for (int e = 0; e < in->n_edges; ++e)
    if (in->c1[e] < 0 || in->c1[e] >= in->n_nodes || in->c2[e] < 0 || in->c2[e] >= in->n_nodes) { h->err = "pgs_fourdof_evaluate: edge " + std::to_string(e) + " names a node out of range"; return PGS_ERR_INVALID_ARGUMENT; }
  const bool qin = in->kind == PGS_FOURDOF_QIN;
  const int NR = in->kind == PGS_FOURDOF_ERROR ? 6 : in->kind == PGS_FOURDOF_SWITCH ? 7 : 4;
  const int NC = in->kind == PGS_FOURDOF_ERROR ? 12 : in->kind == PGS_FOURDOF_SWITCH ? 13 : 8;
  const size_t n = (size_t)in->n_nodes, m = (size_t)in->n_edges, rw = qin ? 3 : 4;
  const int tiles = (in->n_edges + 31) / 32;
  FCU(cudaSetDevice(h->dev));
  FCU(h->rot.resize(rw * n)); FCU(h->t.resize(3 * n)); FCU(h->c1.resize(m)); FCU(h->c2.resize(m));
  FCU(h->obs_rot.resize(rw * m)); FCU(h->obs_t.resize(3 * m)); FCU(h->weight.resize(m)); FCU(h->sw.resize(m));
  FCU(h->r.resize(NR * m)); if (J) FCU(h->J.resize((size_t)NR * NC * m)); FCU(h->cost_tile.resize(tiles));
  cudaStream_t s = h->stream;
  FCU(cudaMemcpyAsync(h->rot.p, in->rot, sizeof(double) * rw * n, cudaMemcpyHostToDevice, s));
  FCU(cudaMemcpyAsync(h->t.p, in->t, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
  FCU(cudaMemcpyAsync(h->c1.p, in->c1, sizeof(int) * m, cudaMemcpyHostToDevice, s));
  FCU(cudaMemcpyAsync(h->c2.p, in->c2, sizeof(int) * m, cudaMemcpyHostToDevice, s));
  FCU(cudaMemcpyAsync(h->obs_rot.p, in->obs_rot, sizeof(double) * rw * m, cudaMemcpyHostToDevice, s));
  FCU(cudaMemcpyAsync(h->obs_t.p, in->obs_t, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, s));
  if (in->kind == PGS_FOURDOF_ERROR) FCU(cudaMemcpyAsync(h->weight.p, in->weight, sizeof(double) * m, cudaMemcpyHostToDevice, s));
  if (in->kind == PGS_FOURDOF_SWITCH) FCU(cudaMemcpyAsync(h->sw.p, in->sw, sizeof(double) * m, cudaMemcpyHostToDevice, s));
  Args A;
  A.n_edges = in->n_edges; A.rot = h->rot.p; A.t = h->t.p; A.c1 = h->c1.p; A.c2 = h->c2.p; A.obs_rot = h->obs_rot.p; A.obs_t = h->obs_t.p;
  A.weight = h->weight.p; A.sw = h->sw.p; A.r = h->r.p; A.J = J ? h->J.p : nullptr; A.cost_tile = h->cost_tile.p;
  const int grid = std::max(1, std::min((tiles + WARPS - 1) / WARPS, h->sm_count * 8));   // a multiple of the SM count once the list is long enough
  const size_t smem = sizeof(double) * WARPS * 32 * ((size_t)NR * NC + 1);
  FCU(cudaEventRecord(h->e0, s));
  if (in->kind == PGS_FOURDOF_ERROR) fourdof_kernel<PGS_FOURDOF_ERROR><<<grid, 32 * WARPS, smem, s>>>(A);
  else if (in->kind == PGS_FOURDOF_SWITCH) fourdof_kernel<PGS_FOURDOF_SWITCH><<<grid, 32 * WARPS, smem, s>>>(A);
  else fourdof_kernel<PGS_FOURDOF_QIN><<<grid, 32 * WARPS, smem, s>>>(A);
  FCU(cudaGetLastError());
  FCU(cudaEventRecord(h->e1, s));
  h->h_cost.resize(tiles);
  FCU(cudaMemcpyAsync(r, h->r.p, sizeof(double) * NR * m, cudaMemcpyDeviceToHost, s));
  if (J) FCU(cudaMemcpyAsync(J, h->J.p, sizeof(double) * NR * NC * m, cudaMemcpyDeviceToHost, s));
  FCU(cudaMemcpyAsync(h->h_cost.data(), h->cost_tile.p, sizeof(double) * tiles, cudaMemcpyDeviceToHost, s));
  FCU(cudaStreamSynchronize(s));
  float ms = 0; FCU(cudaEventElapsedTime(&ms, h->e0, h->e1)); h->ms_kernel = ms;
  if (cost) { double c = 0; for (int k = 0; k < tiles; ++k) c += h->h_cost[k]; *cost = c; }   // tile order: reproducible
  return PGS_OK;
} catch (const std::exception& e) {   // nothing may be thrown across the C boundary
  if (h) h->err = std::string("unexpected C++ exception: ") + e.what();
  return PGS_ERR_STATE;
}

int pgs_fourdof_last_timing(pgs_fourdof_handle h, double* ms_kernel) {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (ms_kernel) *ms_kernel = h->ms_kernel;
  return PGS_OK;
}

}  // extern "C"
