// Composer pose assembly on the device (include/pgs_compose.h; reference src/Composer.cpp:24-209).
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/pgs.h"
#include "../../include/pgs_compose.h"
#include "pgs_solver.h"   // DBuf

namespace pgs {

struct Aff { double R[9], t[3]; };   // rigid / affine 3x4

__device__ __forceinline__ Aff aff_load16(const double* __restrict__ T) {
  Aff a;
#pragma unroll
  for (int r = 0; r < 3; ++r) { a.R[3 * r] = T[4 * r]; a.R[3 * r + 1] = T[4 * r + 1]; a.R[3 * r + 2] = T[4 * r + 2]; a.t[r] = T[4 * r + 3]; }
  return a;
}
__device__ __forceinline__ Aff aff_from_qt(const double* __restrict__ q, const double* __restrict__ t) {
  Aff a;
  const double x = q[0], y = q[1], z = q[2], w = q[3];   // raw_xyzw_to_eigenmat (PoseManipUtils.cpp:61-72)
  a.R[0] = 1 - 2 * (y * y + z * z); a.R[1] = 2 * (x * y - z * w);     a.R[2] = 2 * (x * z + y * w);
  a.R[3] = 2 * (x * y + z * w);     a.R[4] = 1 - 2 * (x * x + z * z); a.R[5] = 2 * (y * z - x * w);
  a.R[6] = 2 * (x * z - y * w);     a.R[7] = 2 * (y * z + x * w);     a.R[8] = 1 - 2 * (x * x + y * y);
  a.t[0] = t[0]; a.t[1] = t[1]; a.t[2] = t[2];
  return a;
}
__device__ __forceinline__ Aff aff_mul(const Aff& A, const Aff& B) {
  Aff C;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) C.R[3 * r + c] = A.R[3 * r] * B.R[c] + A.R[3 * r + 1] * B.R[3 + c] + A.R[3 * r + 2] * B.R[6 + c];
    C.t[r] = A.R[3 * r] * B.t[0] + A.R[3 * r + 1] * B.t[1] + A.R[3 * r + 2] * B.t[2] + A.t[r];
  }
  return C;
}
// inverse of a rigid transform; the reference calls Eigen's generic 4x4 inverse on (rigid) poses (Composer.cpp:98,145,163)
__device__ __forceinline__ Aff aff_inv_rigid(const Aff& A) {
  Aff C;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) C.R[3 * r + c] = A.R[3 * c + r];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) C.t[r] = -(C.R[3 * r] * A.t[0] + C.R[3 * r + 1] * A.t[1] + C.R[3 * r + 2] * A.t[2]);
  return C;
}

struct ComposeArgs {
  int n_nodes, n_slam, solved_until, solved_until_world, n_worlds;
  const double* __restrict__ mgr_T; const int* __restrict__ world_id;
  const double* __restrict__ slam_q; const double* __restrict__ slam_t;
  const int* __restrict__ world_end; const int* __restrict__ world_setid; const unsigned char* __restrict__ ws_exists; const double* __restrict__ ws_T_w;
  double* __restrict__ out_T;
};

__device__ __forceinline__ Aff slam_or_mgr(const ComposeArgs& A, int i) {   // Composer.cpp:73-84,157-160
  if (i < A.n_slam) return aff_from_qt(A.slam_q + 4 * (size_t)i, A.slam_t + 3 * (size_t)i);
  return aff_load16(A.mgr_T + 16 * (size_t)i);
}

// assembled pose of a keyframe that is NOT in a dead zone (world >= 0)
__device__ Aff compose_regular(const ComposeArgs& A, int i, int world) {
  const Aff Mi = aff_load16(A.mgr_T + 16 * (size_t)i);
  if (i <= A.solved_until) return slam_or_mgr(A, i);                                  // :69-88
  if (A.solved_until == 0) {                                                           // :128-132, then :172-190
    const int setid = world < A.n_worlds ? A.world_setid[world] : -1;
    if (world != setid && world < A.n_worlds && A.ws_exists[world]) return aff_mul(aff_load16(A.ws_T_w + 16 * (size_t)world), Mi);
    return Mi;
  }
  if (A.solved_until_world == world) {                                                 // :134-136,155-165
    const int last = A.solved_until;
    const Aff w_T_last = slam_or_mgr(A, last);
    const Aff last_M_i = aff_mul(aff_inv_rigid(aff_load16(A.mgr_T + 16 * (size_t)last)), Mi);
    return aff_mul(w_T_last, last_M_i);
  }
  return Mi;                                                                           // :137-139
}

__global__ void __launch_bounds__(128) compose_kernel(ComposeArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_nodes) return;
  const int world = A.world_id[i];
  Aff T;
  if (world >= 0) T = compose_regular(A, i, world);
  else if (i > A.solved_until && A.solved_until == 0) T = aff_load16(A.mgr_T + 16 * (size_t)i);   // :128-132 (no set transform for a dead zone)
  else {
    // dead zone -(k+1): the last assembled pose of world k carried through odometry (:89-101,140-146)
    const int k = -world - 1;
    const int last = (k < A.n_worlds) ? A.world_end[k] : -1;
    if (last < 0 || last >= A.n_nodes) T = aff_load16(A.mgr_T + 16 * (size_t)i);       // the reference asserts / exits here (:137-143)
    else {
      const Aff w_T_last = compose_regular(A, last, k);
      const Aff last_M_i = aff_mul(aff_inv_rigid(aff_load16(A.mgr_T + 16 * (size_t)last)), aff_load16(A.mgr_T + 16 * (size_t)i));
      T = aff_mul(w_T_last, last_M_i);
    }
  }
  double* o = A.out_T + 16 * (size_t)i;
#pragma unroll
  for (int r = 0; r < 3; ++r) { o[4 * r] = T.R[3 * r]; o[4 * r + 1] = T.R[3 * r + 1]; o[4 * r + 2] = T.R[3 * r + 2]; o[4 * r + 3] = T.t[r]; }
  o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

}  // namespace pgs

struct pgs_compose_s {
  int dev = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
  pgs::DBuf<double> mgr_T, slam_q, slam_t, ws_T_w, out_T;
  pgs::DBuf<int> world_id, world_end, world_setid;
  pgs::DBuf<unsigned char> ws_exists;
  double ms_kernel = 0, ms_total = 0;
  std::string err;
};

#define CCU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { h->err = std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #x; return e__ == cudaErrorMemoryAllocation ? PGS_ERR_OUT_OF_MEMORY : PGS_ERR_CUDA; } } while (0)

extern "C" {

int pgs_compose_create(int32_t device, pgs_compose_handle* out) {
  if (!out) return PGS_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return PGS_ERR_CUDA;   // no CPU fallback
  pgs_compose_s* h = new pgs_compose_s();
  h->dev = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess || cudaEventCreate(&h->e2) != cudaSuccess || cudaEventCreate(&h->e3) != cudaSuccess) {
    delete h; return PGS_ERR_CUDA;
  }
  *out = h;
  return PGS_OK;
}

void pgs_compose_destroy(pgs_compose_handle h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  if (h->e2) cudaEventDestroy(h->e2);
  if (h->e3) cudaEventDestroy(h->e3);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* pgs_compose_last_error(pgs_compose_handle h) { return h ? h->err.c_str() : "null handle"; }

int pgs_compose_run(pgs_compose_handle h, const pgs_compose_input* in, double* out_T) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (!in || in->n_nodes < 0 || in->n_slam < 0 || in->n_worlds < 0 || (in->n_nodes > 0 && (!in->mgr_T || !in->world_id || !out_T)) ||
      (in->n_slam > 0 && (!in->slam_q || !in->slam_t)) || (in->n_worlds > 0 && (!in->world_end || !in->world_setid || !in->ws_exists || !in->ws_T_w))) {
    h->err = "pgs_compose_run: null or negative input"; return PGS_ERR_INVALID_ARGUMENT;
  }
  if (in->n_slam > in->n_nodes) { h->err = "pgs_compose_run: more optimised poses than keyframes"; return PGS_ERR_INVALID_ARGUMENT; }
  if (in->n_nodes == 0) return PGS_OK;
  if (in->solved_until < 0 || in->solved_until >= in->n_nodes) { h->err = "pgs_compose_run: solved_until out of range"; return PGS_ERR_INVALID_ARGUMENT; }
  CCU(cudaSetDevice(h->dev));
  const size_t n = (size_t)in->n_nodes, ns = (size_t)in->n_slam, nw = (size_t)in->n_worlds;
  CCU(h->mgr_T.resize(16 * n)); CCU(h->world_id.resize(n)); CCU(h->out_T.resize(16 * n));
  CCU(h->slam_q.resize(4 * ns)); CCU(h->slam_t.resize(3 * ns));
  CCU(h->world_end.resize(nw)); CCU(h->world_setid.resize(nw)); CCU(h->ws_exists.resize(nw)); CCU(h->ws_T_w.resize(16 * nw));
  CCU(cudaEventRecord(h->e0, h->stream));
  CCU(cudaMemcpyAsync(h->mgr_T.p, in->mgr_T, sizeof(double) * 16 * n, cudaMemcpyHostToDevice, h->stream));
  CCU(cudaMemcpyAsync(h->world_id.p, in->world_id, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
  if (ns) {
    CCU(cudaMemcpyAsync(h->slam_q.p, in->slam_q, sizeof(double) * 4 * ns, cudaMemcpyHostToDevice, h->stream));
    CCU(cudaMemcpyAsync(h->slam_t.p, in->slam_t, sizeof(double) * 3 * ns, cudaMemcpyHostToDevice, h->stream));
  }
  if (nw) {
    CCU(cudaMemcpyAsync(h->world_end.p, in->world_end, sizeof(int) * nw, cudaMemcpyHostToDevice, h->stream));
    CCU(cudaMemcpyAsync(h->world_setid.p, in->world_setid, sizeof(int) * nw, cudaMemcpyHostToDevice, h->stream));
    CCU(cudaMemcpyAsync(h->ws_exists.p, in->ws_exists, nw, cudaMemcpyHostToDevice, h->stream));
    CCU(cudaMemcpyAsync(h->ws_T_w.p, in->ws_T_w, sizeof(double) * 16 * nw, cudaMemcpyHostToDevice, h->stream));
  }
  pgs::ComposeArgs A;
  A.n_nodes = in->n_nodes; A.n_slam = in->n_slam; A.solved_until = in->solved_until; A.solved_until_world = in->solved_until_world; A.n_worlds = in->n_worlds;
  A.mgr_T = h->mgr_T.p; A.world_id = h->world_id.p; A.slam_q = h->slam_q.p; A.slam_t = h->slam_t.p;
  A.world_end = h->world_end.p; A.world_setid = h->world_setid.p; A.ws_exists = h->ws_exists.p; A.ws_T_w = h->ws_T_w.p; A.out_T = h->out_T.p;
  CCU(cudaEventRecord(h->e1, h->stream));
  pgs::compose_kernel<<<(in->n_nodes + 127) / 128, 128, 0, h->stream>>>(A);
  CCU(cudaEventRecord(h->e2, h->stream));
  CCU(cudaGetLastError());
  CCU(cudaMemcpyAsync(out_T, h->out_T.p, sizeof(double) * 16 * n, cudaMemcpyDeviceToHost, h->stream));
  CCU(cudaEventRecord(h->e3, h->stream));
  CCU(cudaStreamSynchronize(h->stream));
  float a = 0, b = 0;
  CCU(cudaEventElapsedTime(&a, h->e1, h->e2)); CCU(cudaEventElapsedTime(&b, h->e0, h->e3));
  h->ms_kernel = a; h->ms_total = b;
  return PGS_OK;
} catch (const std::exception& e) {   // nothing may be thrown across the C boundary
  if (h) h->err = std::string("unexpected C++ exception: ") + e.what();
  return PGS_ERR_STATE;
}

int pgs_compose_last_timing(pgs_compose_handle h, double* ms_kernel, double* ms_total) {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (ms_kernel) *ms_kernel = h->ms_kernel;
  if (ms_total) *ms_total = h->ms_total;
  return PGS_OK;
}

}  // extern "C"
