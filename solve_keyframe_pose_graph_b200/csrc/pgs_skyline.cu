// K4 (direct variant): skyline (row-envelope) Cholesky of the reduced pose system, entirely on the device.
//
// Stands in for Ceres' SPARSE_NORMAL_CHOLESKY / CHOLMOD (reference src/PoseGraphSLAM.cpp:1270).  In node order a
// keyframe pose graph is a "thick chain": odometry edges couple i with i-1..i-f, loop edges reach back at most
// a few thousand keyframes, so row i of the Cholesky factor is dense exactly on [min neighbour of i, i] — the
// row envelope — and a skyline factorisation stores no explicit zeros beyond panel alignment.
//
// Layout: scalar row r (6 per node) stores columns [start[r], rowend[r]) contiguously (row-major), start[r] being
// the envelope start rounded down to a panel boundary (PW scalars) and rowend[r] the end of r's own panel, so
// every (row, panel) intersection is a full PW-wide, 16-byte aligned segment.  One extra row n carries b^T:
// factoring it along with the matrix performs the forward substitution for free (row n of L is (L^-1 b)^T).
//
// Right-looking by panels of PW columns with one panel of look-ahead on a second stream:
//   diag    1 CTA    L_dd = chol(A_dd) in shared memory (8-column blocks, 4x4 register tiles), Linv = L_dd^-1 by
//                    block forward substitution (kept for trsm and the backward solve)
//   trsm    |R|/32   X = A[R, panel] * Linv^T for the rows R below the panel whose envelope reaches it
//   update  tiles    A[r, c] -= X[r,:] . X[c,:] for r, c in R, c <= r: 128x64 tiles, 8x4 register micro-tiles,
//                    K = 96 in three cp.async double-buffered chunks.  "next" = the first two column tiles
//                    (they hold every column of panel d+1) runs on the panel stream ahead of diag(d+1);
//                    "rest" runs on the main stream concurrently with diag/trsm of the next panel.
// then a backward sweep (one launch per panel) solves L^T x = y.
//
// Partial factorisation (multi-GPU domain decomposition): only the first n_elim panels are
// eliminated; the trailing rows then hold the Schur complement on the border unknowns and the border part of
// the forward-substituted right-hand side.
#include "pgs_skyline.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/pgs.h"

namespace pgs {

constexpr int PW = 96;            // panel width in scalars = 16 nodes
constexpr int PN = PW / 6;
constexpr int TR = 32;            // trsm rows per CTA
constexpr int LDP = PW + 2;       // padded leading dimension of PW-wide shared tiles: even (16-B rows), LDP/2 odd
constexpr int UM_BULK = 128, UM_NEXT = 32, UN = 64;  // update tiles: rows x cols
constexpr int KC = 32;            // update K chunk
constexpr int LDK = KC + 2;       // 34: even, LDK/2 odd -> conflict-free 128-bit row reads
constexpr int NEXT_TILES = 2;     // column tiles of the look-ahead part of the update (2 * UN >= PW)
constexpr int NEV = 8;            // event ring

struct SkylineFactor {
  int N = 0, n = 0, D = 0;         // nodes, scalars, panels
  int D_elim = 0;                  // panels to eliminate (== D for a full factorisation)
  cudaStream_t stream = nullptr, s1 = nullptr;
  cudaEvent_t ev_trsm[NEV], ev_rest[NEV], ev_fork = nullptr, ev_join = nullptr;
  std::vector<int> h_start;        // per scalar row (n+1 entries, last = rhs row)
  std::vector<long long> h_ptr;    // n+2
  std::vector<int> h_rows_ptr;     // D+1
  std::vector<int> h_lo;           // per panel: smallest envelope start of its rows
  long long nnz = 0;
  int max_rows = 0;
  double* val = nullptr; long long* ptr = nullptr; int* start = nullptr;
  int* rows_ptr = nullptr; int* rows_idx = nullptr;
  double* dinv = nullptr;          // [D][PW*PW] inverse of the diagonal factors
  double* xacc = nullptr;          // [n] backward-solve accumulator
  int* fail = nullptr; int* h_fail = nullptr;
  int* pair_hi = nullptr; int* pair_lo = nullptr; int n_pairs = 0;
};

#define SK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { if (err) *err = std::string("skyline: ") + cudaGetErrorString(e__) + " at " #x; return e__ == cudaErrorMemoryAllocation ? PGS_ERR_OUT_OF_MEMORY : PGS_ERR_CUDA; } } while (0)

void skyline_destroy(SkylineFactor* f) {
  if (!f) return;
  cudaFree(f->val); cudaFree(f->ptr); cudaFree(f->start); cudaFree(f->rows_ptr); cudaFree(f->rows_idx); cudaFree(f->dinv);
  cudaFree(f->xacc); cudaFree(f->fail); cudaFree(f->pair_hi); cudaFree(f->pair_lo);
  if (f->h_fail) cudaFreeHost(f->h_fail);
  for (int i = 0; i < NEV; ++i) { if (f->ev_trsm[i]) cudaEventDestroy(f->ev_trsm[i]); if (f->ev_rest[i]) cudaEventDestroy(f->ev_rest[i]); }
  if (f->ev_fork) cudaEventDestroy(f->ev_fork);
  if (f->ev_join) cudaEventDestroy(f->ev_join);
  if (f->s1) cudaStreamDestroy(f->s1);
  delete f;
}
int64_t skyline_nnz(const SkylineFactor* f) { return f ? f->nnz : 0; }
int skyline_panel_width() { return PW; }

SkylineFactor* skyline_create(int N, int n_pairs, const int* pair_hi, const int* pair_lo, cudaStream_t stream, std::string* err,
                              int n_border_nodes, bool dense) {
  SkylineFactor* f = new SkylineFactor();
  for (int i = 0; i < NEV; ++i) { f->ev_trsm[i] = nullptr; f->ev_rest[i] = nullptr; }
  f->N = N; f->n = 6 * N; f->D = (f->n + PW - 1) / PW; f->stream = stream; f->n_pairs = n_pairs;
  const int n = f->n, D = f->D;
  const int N_int = dense ? 0 : N - n_border_nodes;           // interior nodes come first, border nodes last
  f->D_elim = (!dense && n_border_nodes > 0) ? (6 * N_int) / PW : D;   // the caller pads the interior to a whole number of panels
  // ---- symbolic: envelope start per node = min neighbour, rounded down to a panel boundary
  std::vector<int> nstart(N);
  for (int i = 0; i < N; ++i) nstart[i] = i;
  for (int p = 0; p < n_pairs; ++p) nstart[pair_hi[p]] = std::min(nstart[pair_hi[p]], pair_lo[p]);
  // border rows keep the whole border block (it receives the dense Schur complement)
  for (int i = N_int; i < N; ++i) nstart[i] = std::min(nstart[i], N_int);
  f->h_start.resize(n + 1); f->h_ptr.resize(n + 2);
  long long off = 0;
  for (int i = 0; i < N; ++i) {
    const int s = (nstart[i] / PN) * PW;
    const int rowend = std::min(n, ((6 * i) / PW + 1) * PW);
    for (int k = 0; k < 6; ++k) { const int r = 6 * i + k; f->h_start[r] = s; f->h_ptr[r] = off; off += rowend - s; }
  }
  f->h_start[n] = 0; f->h_ptr[n] = off; off += (long long)D * PW;   // rhs row, padded to whole panels
  f->h_ptr[n + 1] = off; f->nnz = off;
  // ---- rows below each panel whose envelope reaches it (+ the rhs row)
  std::vector<int> cnt(D + 1, 0);
  for (int i = 0; i < N; ++i) { const int d0 = f->h_start[6 * i] / PW, d1 = (6 * i) / PW; for (int d = d0; d < d1; ++d) cnt[d + 1] += 6; }
  // a node's 6 rows may straddle a panel boundary only if PW % 6 != 0 (it is not)
  for (int d = 0; d < D; ++d) cnt[d + 1] += 1;   // rhs row
  f->h_rows_ptr.assign(D + 1, 0);
  for (int d = 0; d < D; ++d) f->h_rows_ptr[d + 1] = f->h_rows_ptr[d] + cnt[d + 1];
  std::vector<int> rows_idx(f->h_rows_ptr[D]), cur(f->h_rows_ptr.begin(), f->h_rows_ptr.end() - 1);
  for (int i = 0; i < N; ++i) { const int d0 = f->h_start[6 * i] / PW, d1 = (6 * i) / PW;
    for (int d = d0; d < d1; ++d) for (int k = 0; k < 6; ++k) rows_idx[cur[d]++] = 6 * i + k; }
  for (int d = 0; d < D; ++d) { rows_idx[cur[d]++] = n; f->max_rows = std::max(f->max_rows, f->h_rows_ptr[d + 1] - f->h_rows_ptr[d]); }
  f->h_lo.resize(D);
  for (int d = 0; d < D; ++d) { int lo = d * PW; for (int i = d * PW; i < std::min(n, (d + 1) * PW); ++i) lo = std::min(lo, f->h_start[i]); f->h_lo[d] = lo; }

  auto bad = [&](cudaError_t e, const char* what) { if (err) *err = std::string("skyline_create: ") + cudaGetErrorString(e) + " (" + what + ", factor needs " + std::to_string((double)f->nnz * 8 / 1e9) + " GB)"; skyline_destroy(f); return (SkylineFactor*)nullptr; };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&f->val, sizeof(double) * (size_t)f->nnz)) != cudaSuccess) return bad(e, "val");
  if ((e = cudaMalloc((void**)&f->ptr, sizeof(long long) * (n + 2))) != cudaSuccess) return bad(e, "ptr");
  if ((e = cudaMalloc((void**)&f->start, sizeof(int) * (n + 1))) != cudaSuccess) return bad(e, "start");
  if ((e = cudaMalloc((void**)&f->rows_ptr, sizeof(int) * (D + 1))) != cudaSuccess) return bad(e, "rows_ptr");
  if ((e = cudaMalloc((void**)&f->rows_idx, sizeof(int) * std::max<size_t>(rows_idx.size(), 1))) != cudaSuccess) return bad(e, "rows_idx");
  if ((e = cudaMalloc((void**)&f->dinv, sizeof(double) * (size_t)std::max(D, 1) * PW * PW)) != cudaSuccess) return bad(e, "dinv");
  if ((e = cudaMalloc((void**)&f->xacc, sizeof(double) * (size_t)std::max(D, 1) * PW)) != cudaSuccess) return bad(e, "xacc");
  if ((e = cudaMalloc((void**)&f->fail, sizeof(int))) != cudaSuccess) return bad(e, "fail");
  if ((e = cudaMallocHost((void**)&f->h_fail, sizeof(int))) != cudaSuccess) return bad(e, "h_fail");
  if ((e = cudaMalloc((void**)&f->pair_hi, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_hi");
  if ((e = cudaMalloc((void**)&f->pair_lo, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_lo");
  { int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);   // the panel stream carries the critical path
    if ((e = cudaStreamCreateWithPriority(&f->s1, cudaStreamNonBlocking, hi_p)) != cudaSuccess) return bad(e, "stream"); }
  for (int i = 0; i < NEV; ++i) {
    if ((e = cudaEventCreateWithFlags(&f->ev_trsm[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
    if ((e = cudaEventCreateWithFlags(&f->ev_rest[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  }
  if ((e = cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  if ((e = cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  cudaMemcpyAsync(f->ptr, f->h_ptr.data(), sizeof(long long) * (n + 2), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(f->start, f->h_start.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(f->rows_ptr, f->h_rows_ptr.data(), sizeof(int) * (D + 1), cudaMemcpyHostToDevice, stream);
  if (!rows_idx.empty()) cudaMemcpyAsync(f->rows_idx, rows_idx.data(), sizeof(int) * rows_idx.size(), cudaMemcpyHostToDevice, stream);
  if (n_pairs) { cudaMemcpyAsync(f->pair_hi, pair_hi, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream);
                 cudaMemcpyAsync(f->pair_lo, pair_lo, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream); }
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bad(e, "upload");
  return f;
}

// ------------------------------------------------------------------------------------------------ kernels
// scatter the block system into the (zeroed) envelope: diagonal blocks (lower triangle), off-diagonal blocks, rhs row
__global__ void sky_scatter_kernel(int N, int n_pairs, const double* __restrict__ Ad, const double* __restrict__ Ao, const double* __restrict__ b,
                                   const int* __restrict__ pair_hi, const int* __restrict__ pair_lo, const long long* __restrict__ ptr,
                                   const int* __restrict__ start, double* __restrict__ val) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 6 * N;
  if (t < 36 * N) {
    const int i = t / 36, a = (t % 36) / 6, c = t % 6;
    if (c <= a) { const int r = 6 * i + a; val[ptr[r] + (6 * i + c - start[r])] = Ad[t]; }
  }
  if (t < 36 * n_pairs) {
    const int p = t / 36, a = (t % 36) / 6, c = t % 6;
    const int r = 6 * pair_hi[p] + a;
    val[ptr[r] + (6 * pair_lo[p] + c - start[r])] = Ao[t];
  }
  if (t < n) val[ptr[n] + t] = b[t];
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NGROUPS>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NGROUPS)); }

// diag: factor the PW x PW diagonal block of panel d in shared memory, store L back, store Linv.
__global__ void __launch_bounds__(256) sky_diag_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                       double* __restrict__ val, double* __restrict__ dinv, int* __restrict__ fail) {
  extern __shared__ __align__(16) double sm_diag[];
  double* L = sm_diag;                 // [PW][LDP]
  double* X = sm_diag + PW * LDP;      // [PW][LDP]
  double* Dinv = X + PW * LDP;         // [PW/8][64] inverses of the 8x8 diagonal blocks
  __shared__ long long rbase[PW];
  __shared__ int bad;
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  if (tid == 0) bad = 0;
  if (tid < PW) { const int r = c0 + tid; rbase[tid] = tid < w ? ptr[r] + (c0 - start[r]) : 0; }
  __syncthreads();
  {
    // all 36 loads of a thread are issued before the first use (one memory latency instead of 36)
    double v[PW * PW / 256];
#pragma unroll
    for (int q = 0; q < PW * PW / 256; ++q) {
      const int e = tid + 256 * q, i = e / PW, j = e % PW;
      v[q] = (i == j && i >= w) ? 1.0 : 0.0;       // identity padding of a short last panel
      if (i < w && j <= i) v[q] = val[rbase[i] + j];
    }
#pragma unroll
    for (int q = 0; q < PW * PW / 256; ++q) {
      const int e = tid + 256 * q, i = e / PW, j = e % PW;
      L[i * LDP + j] = v[q]; X[i * LDP + j] = 0.0;
    }
  }
  __syncthreads();
  // ---- blocked right-looking Cholesky, 8 columns at a time
  for (int jb = 0; jb < PW; jb += 8) {
    const int i = jb + tid;
    double Dg[8][8], arow[8], dinv8[8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) Dg[a][c] = (c <= a) ? L[(jb + a) * LDP + jb + c] : 0.0;
    if (i < PW) {
#pragma unroll
      for (int c = 0; c < 8; ++c) arow[c] = L[i * LDP + jb + c];
    }
    // every thread re-factors the 8x8 diagonal block in registers (no synchronisation inside)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double dd = Dg[c][c];
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < c) dd -= Dg[c][k] * Dg[c][k];
      if (!(dd > 0.0)) { if (tid == 0) bad = 1; dd = 1.0; }
      const double inv = rsqrt(dd);      // 1 ulp; the sqrt + divide pair would put ~600 cycles per column on the critical path
      Dg[c][c] = dd * inv; dinv8[c] = inv;
#pragma unroll
      for (int a = 0; a < 8; ++a) if (a > c) {
        double s = Dg[a][c];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < c) s -= Dg[a][k] * Dg[c][k];
        Dg[a][c] = s * inv;
      }
    }
    __syncthreads();   // everybody has read the diagonal block / its own row before they are overwritten
    if (i < PW) {
      double row[8];
      if (tid < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; ++a) if (a == tid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) if (c <= a) row[c] = Dg[a][c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double s = arow[c];
#pragma unroll
          for (int k = 0; k < 8; ++k) if (k < c) s -= row[k] * Dg[c][k];
          row[c] = s * dinv8[c];
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) L[i * LDP + jb + c] = row[c];
    }
    __syncthreads();
    // trailing update by 4x4 register blocks: A[i][k] -= sum_c L[i][jb+c] L[k][jb+c] on the lower block triangle
    const int t0 = jb + 8, nb = (PW - t0) / 4;
    if (tid < nb * (nb + 1) / 2) {
      int bi = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
      while ((bi + 1) * (bi + 2) / 2 <= tid) ++bi;
      while (bi * (bi + 1) / 2 > tid) --bi;
      const int bj = tid - bi * (bi + 1) / 2;
      const int i0 = t0 + 4 * bi, k0 = t0 + 4 * bj;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        double2 ar[4], bc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { ar[a] = *reinterpret_cast<const double2*>(&L[(i0 + a) * LDP + jb + c]); bc[a] = *reinterpret_cast<const double2*>(&L[(k0 + a) * LDP + jb + c]); }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += ar[a].x * bc[b].x + ar[a].y * bc[b].y;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) L[(i0 + a) * LDP + k0 + b] -= acc[a][b];   // the strict upper part of diagonal blocks is scratch
    }
    __syncthreads();
  }
  // ---- X = L^-1: invert the 8x8 diagonal blocks, then block forward substitution by block distance
  if (tid < PW / 8) {
    const int o = 8 * tid;
    double Xi[8][8], rl[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) rl[c] = 1.0 / L[(o + c) * LDP + o + c];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int a = 0; a < 8; ++a) Xi[a][c] = 0.0;
      Xi[c][c] = rl[c];
#pragma unroll
      for (int a = 0; a < 8; ++a) if (a > c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k >= c && k < a) s += L[(o + a) * LDP + o + k] * Xi[k][c];
        Xi[a][c] = -s * rl[a];
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) { Dinv[tid * 64 + a * 8 + c] = Xi[a][c]; X[(o + a) * LDP + o + c] = Xi[a][c]; }
  }
  __syncthreads();
  constexpr int NB = PW / 8;
  for (int t = 1; t < NB; ++t) {
    const int cnt = (NB - t) * 64;
    // stage 1: S_IJ = sum_{K=J}^{I-1} L_IK X_KJ, written into the (still unused) X_IJ slot
    for (int e = tid; e < cnt; e += blockDim.x) {
      const int J = e / 64, a = (e % 64) / 8, b = e % 8, I = J + t;
      double s = 0.0;
      for (int k = 8 * J; k < 8 * I; ++k) s += L[(8 * I + a) * LDP + k] * X[k * LDP + 8 * J + b];
      X[(8 * I + a) * LDP + 8 * J + b] = s;
    }
    __syncthreads();
    // stage 2: X_IJ = -Dinv_I S_IJ (read everything, then write)
    double out[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = tid + q * 256;
      out[q] = 0.0;
      if (e < cnt) {
        const int J = e / 64, a = (e % 64) / 8, b = e % 8, I = J + t;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c <= a) s += Dinv[I * 64 + a * 8 + c] * X[(8 * I + c) * LDP + 8 * J + b];
        out[q] = -s;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = tid + q * 256;
      if (e < cnt) { const int J = e / 64, a = (e % 64) / 8, b = e % 8, I = J + t; X[(8 * I + a) * LDP + 8 * J + b] = out[q]; }
    }
    __syncthreads();
  }
#pragma unroll 12
  for (int q = 0; q < PW * PW / 256; ++q) {
    const int e = tid + 256 * q, i = e / PW, j = e % PW;
    if (i < w && j <= i) val[rbase[i] + j] = L[i * LDP + j];
    dinv[(size_t)d * PW * PW + e] = (i < w && j < w && j <= i) ? X[i * LDP + j] : 0.0;
  }
  if (tid == 0 && bad) *fail = 1;
}

// trsm: X[r][j] = sum_k A[r][c0+k] * Linv[j][k] for the rows r in R_d, in place.  32 rows x 96 columns per CTA,
// 2 x 6 outputs per thread.
__global__ void __launch_bounds__(256, 2) sky_trsm_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                          const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx,
                                                          const double* __restrict__ dinv, double* __restrict__ val) {
  extern __shared__ __align__(16) double sm[];
  double* Li = sm;                    // [PW][LDP]  Linv
  double* A = sm + PW * LDP;          // [TR][LDP]
  __shared__ long long rbase[TR];
  const int c0 = d * PW, tid = threadIdx.x;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  const int row0 = blockIdx.x * TR;
  if (tid < TR) {
    long long b = -1;
    if (row0 + tid < nr) { const int r = rows_idx[rb + row0 + tid]; b = ptr[r] + (c0 - start[r]); }
    rbase[tid] = b;
  }
  __syncthreads();
  const double* dsrc = dinv + (size_t)d * PW * PW;
  for (int e = tid; e < PW * (PW / 2); e += blockDim.x) {
    const int i = e / (PW / 2), j2 = e % (PW / 2);
    cp_async16(&Li[i * LDP + 2 * j2], dsrc + i * PW + 2 * j2, true);
  }
  for (int e = tid; e < TR * (PW / 2); e += blockDim.x) {
    const int i = e / (PW / 2), j2 = e % (PW / 2);
    const bool ok = rbase[i] >= 0;
    cp_async16(&A[i * LDP + 2 * j2], ok ? (const void*)(val + rbase[i] + 2 * j2) : (const void*)val, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double acc[2][6];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int k = 0; k < PW; k += 2) {
    const double2 a0 = *reinterpret_cast<const double2*>(&A[(2 * ty) * LDP + k]);
    const double2 a1 = *reinterpret_cast<const double2*>(&A[(2 * ty + 1) * LDP + k]);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double2 b = *reinterpret_cast<const double2*>(&Li[(tx + 16 * j) * LDP + k]);
      acc[0][j] += a0.x * b.x + a0.y * b.y;
      acc[1][j] += a1.x * b.x + a1.y * b.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long b = rbase[2 * ty + i];
    if (b < 0) continue;
#pragma unroll
    for (int j = 0; j < 6; ++j) val[b + tx + 16 * j] = acc[i][j];
  }
}

// update: A[r][c] -= X[r,:] . X[c,:] over the tiles (ti, tj) of R_d x R_d that intersect the lower triangle,
// column tiles tj in [tj0, tj0 + gridDim.x).
template <int UM>   // 128: throughput tiles (8x4 per thread) for the bulk; 32: latency tiles (2x4) for the look-ahead columns
__global__ void __launch_bounds__(256, 2) sky_update_kernel(int d, int n, int tj0, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                            const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx, double* __restrict__ val) {
  constexpr int MI = UM / 16;
  const int ti = blockIdx.y, tj = tj0 + blockIdx.x;
  if (tj * UN >= (ti + 1) * UM) return;     // tile entirely above the diagonal
  extern __shared__ __align__(16) double sm[];
  double* As = sm;                           // [2][UM][LDK]
  double* Bs = sm + 2 * UM * LDK;            // [2][UN][LDK]
  __shared__ long long s_abase[UM], s_bbase[UN];   // element offset of (row, c0) in val, -1 = no such row
  __shared__ int s_arow[UM], s_bcol[UN];
  const int c0 = d * PW, tid = threadIdx.x;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  if (tid < UM) {
    const int ir = ti * UM + tid;
    const int r = ir < nr ? rows_idx[rb + ir] : -1;
    s_arow[tid] = r; s_abase[tid] = r >= 0 ? ptr[r] + (c0 - start[r]) : -1;
  } else if (tid < UM + UN) {
    const int t = tid - UM, ic = tj * UN + t;
    int c = ic < nr ? rows_idx[rb + ic] : -1;
    if (c >= n) c = -1;                      // the rhs row is never a column
    s_bcol[t] = c; s_bbase[t] = c >= 0 ? ptr[c] + (c0 - start[c]) : -1;
  }
  __syncthreads();
  auto issue = [&](int chunk, int stage) {
    const int k0 = chunk * KC;
    double* a_dst = As + stage * UM * LDK; double* b_dst = Bs + stage * UN * LDK;
#pragma unroll
    for (int e = tid; e < (UM + UN) * (KC / 2); e += 256) {
      const int row = e / (KC / 2), seg = e % (KC / 2);
      if (row < UM) { const long long b = s_abase[row]; cp_async16(&a_dst[row * LDK + 2 * seg], b >= 0 ? (const void*)(val + b + k0 + 2 * seg) : (const void*)val, b >= 0); }
      else { const int rr = row - UM; const long long b = s_bbase[rr]; cp_async16(&b_dst[rr * LDK + 2 * seg], b >= 0 ? (const void*)(val + b + k0 + 2 * seg) : (const void*)val, b >= 0); }
    }
    cp_async_commit();
  };
  // lane layout inside a warp: 8 (tx) x 4 (ty); warp w covers ty 4*(w>>1).., tx 8*(w&1)..
  const int lane = tid & 31, wid = tid >> 5;
  const int tx = (wid & 1) * 8 + (lane & 7), ty = (wid >> 1) * 4 + (lane >> 3);   // tx 0..15 -> cols tx+16j, ty 0..15 -> rows ty+16i
  // 64-bit shared loads: a warp reads 4 distinct a addresses and 8 distinct b addresses per instruction, all in
  // different banks (row stride 68 words), so every load is a single broadcast wavefront
  double acc[MI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  constexpr int NCH = PW / KC;
  issue(0, 0);
  issue(1, 1);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const double* a_s = As + (ch & 1) * UM * LDK + ty * LDK;
    const double* b_s = Bs + (ch & 1) * UN * LDK + tx * LDK;
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      double a[MI], b[4];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[i] = a_s[(16 * i) * LDK + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = b_s[(16 * j) * LDK + k];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    if (ch + 2 < NCH) { __syncthreads(); issue(ch + 2, ch & 1); }
  }
  // epilogue: read-modify-write in groups of up to 4 rows so up to 16 loads are in flight per thread
  int cidx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cidx[j] = s_bcol[tx + 16 * j];
  constexpr int HR = MI < 4 ? MI : 4;
#pragma unroll
  for (int h = 0; h < MI / HR; ++h) {
    double old[HR][4]; long long base[HR]; int rr[HR];
#pragma unroll
    for (int i = 0; i < HR; ++i) {
      rr[i] = s_arow[ty + 16 * (HR * h + i)];
      base[i] = rr[i] >= 0 ? s_abase[ty + 16 * (HR * h + i)] - c0 : 0;   // offset of (row, column 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = cidx[j]; old[i][j] = (rr[i] >= 0 && c >= 0 && c <= rr[i]) ? val[base[i] + c] : 0.0; }
    }
#pragma unroll
    for (int i = 0; i < HR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = cidx[j]; if (rr[i] >= 0 && c >= 0 && c <= rr[i]) val[base[i] + c] = old[i][j] - acc[HR * h + i][j]; }
  }
}

// backward sweep for panel d (processed D-1 .. 0):  x_d = Linv^T (y_d + acc_d)  [or x_d given, for border panels],
// then push  acc[c] -= sum_{r in panel} L[r][c] x_r  for every column c in [lo, c0) inside the rows' envelopes.
// 256 threads = 32 columns x 8 row groups.
__global__ void __launch_bounds__(256) sky_backward_kernel(int d, int n, int lo, int x_given, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                           const double* __restrict__ val, const double* __restrict__ dinv,
                                                           double* __restrict__ acc, double* __restrict__ x) {
  extern __shared__ __align__(16) double sm[];
  double* Li = sm;                      // [PW][PW+1]
  __shared__ double xs[PW], rhs[PW], red[8][33];
  __shared__ long long rbase[PW];       // ptr[r] - start[r]
  __shared__ int rstart[PW];
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  if (tid < PW) {
    const int r = c0 + tid;
    rbase[tid] = tid < w ? ptr[r] - start[r] : 0; rstart[tid] = tid < w ? start[r] : 0x7fffffff;
    rhs[tid] = tid < w ? val[ptr[n] + c0 + tid] + acc[c0 + tid] : 0.0;
  }
  if (!x_given) {
    const double* dsrc = dinv + (size_t)d * PW * PW;
    for (int e = tid; e < PW * PW; e += blockDim.x) Li[(e / PW) * (PW + 1) + (e % PW)] = dsrc[e];
  }
  __syncthreads();
  if (tid < PW) {
    double s;
    if (x_given) s = tid < w ? x[c0 + tid] : 0.0;
    else {
      double s0 = 0.0, s1 = 0.0;   // x_j = sum_{i>=j} Linv[i][j] rhs_i
      int i = tid;
      for (; i + 1 < w; i += 2) { s0 += Li[i * (PW + 1) + tid] * rhs[i]; s1 += Li[(i + 1) * (PW + 1) + tid] * rhs[i + 1]; }
      if (i < w) s0 += Li[i * (PW + 1) + tid] * rhs[i];
      s = s0 + s1;
      if (blockIdx.x == 0 && tid < w) x[c0 + tid] = s;
    }
    xs[tid] = s;
  }
  __syncthreads();
  const int cx = tid & 31, g = tid >> 5;
  for (int cb = lo + blockIdx.x * 32; cb < c0; cb += gridDim.x * 32) {
    const int c = cb + cx;
    double s = 0.0;
    if (c < c0) {
#pragma unroll
      for (int q = 0; q < PW / 8; ++q) {
        const int i = g * (PW / 8) + q;
        if (c >= rstart[i]) s += val[rbase[i] + c] * xs[i];
      }
    }
    red[g][cx] = s;
    __syncthreads();
    if (g == 0 && c < c0) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][cx];
      acc[c] -= t;
    }
    __syncthreads();
  }
}

static const size_t SM_TRSM = sizeof(double) * (PW * LDP + TR * LDP);
static const size_t SM_UPD = sizeof(double) * (2 * (UM_BULK + UN) * LDK);
static const size_t SM_UPD_NEXT = sizeof(double) * (2 * (UM_NEXT + UN) * LDK);
static const size_t SM_DIAG = sizeof(double) * (2 * PW * LDP + (PW / 8) * 64);
static const size_t SM_BACK = sizeof(double) * (PW * (PW + 1));

static int set_attrs(std::string* err) {
  static bool attr_set = false;
  if (attr_set) return PGS_OK;
  SK(cudaFuncSetAttribute(sky_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_DIAG));
  SK(cudaFuncSetAttribute(sky_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TRSM));
  SK(cudaFuncSetAttribute(sky_update_kernel<UM_BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD));
  SK(cudaFuncSetAttribute(sky_update_kernel<UM_NEXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD_NEXT));
  SK(cudaFuncSetAttribute(sky_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BACK));
  attr_set = true;
  return PGS_OK;
}

static int skyline_begin(SkylineFactor* f, std::string* err) {
  if (int rc = set_attrs(err)) return rc;
  SK(cudaMemsetAsync(f->val, 0, sizeof(double) * (size_t)f->nnz, f->stream));
  SK(cudaMemsetAsync(f->xacc, 0, sizeof(double) * (size_t)std::max(f->D, 1) * PW, f->stream));
  SK(cudaMemsetAsync(f->fail, 0, sizeof(int), f->stream));
  return PGS_OK;
}

// dense (border) system given as a packed lower triangle + rhs; every row of a dense factor starts at column 0
__global__ void sky_load_packed_kernel(int n, const long long* __restrict__ ptr, const double* __restrict__ S, const double* __restrict__ rhs,
                                       double* __restrict__ val) {
  const long long tot = (long long)n * (n + 1) / 2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= e) ++i;
    while ((long long)i * (i + 1) / 2 > e) --i;
    const int j = (int)(e - (long long)i * (i + 1) / 2);
    val[ptr[i] + j] = S[e];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) val[ptr[n] + i] = rhs[i];
}
int skyline_load_packed(SkylineFactor* f, const double* S_packed, const double* rhs, std::string* err) {
  if (int rc = skyline_begin(f, err)) return rc;
  sky_load_packed_kernel<<<592, 256, 0, f->stream>>>(f->n, f->ptr, S_packed, rhs, f->val);
  SK(cudaGetLastError());
  return PGS_OK;
}

// Numeric factorisation of the first D_elim panels (all of them for a single-GPU solve).
int skyline_factor(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, std::string* err) {
  if (int rc = skyline_begin(f, err)) return rc;
  const int tot = std::max(36 * std::max(f->N, f->n_pairs), f->n);
  sky_scatter_kernel<<<(tot + 255) / 256, 256, 0, f->stream>>>(f->N, f->n_pairs, Ad, Ao, b, f->pair_hi, f->pair_lo, f->ptr, f->start, f->val);
  return skyline_factor_numeric(f, err);
}

int skyline_factor_numeric(SkylineFactor* f, std::string* err) {
  cudaStream_t s0 = f->stream, s1 = f->s1;
  const int n = f->n;
  SK(cudaEventRecord(f->ev_fork, s0));
  SK(cudaStreamWaitEvent(s1, f->ev_fork, 0));
  // panel stream s1:  [wait rest(d-1)] next(d) ... diag(d) trsm(d) -> ev_trsm[d]
  // main  stream s0:  [wait ev_trsm[d]] rest(d) -> ev_rest[d]
  for (int d = 0; d < f->D_elim; ++d) {
    const int nr = f->h_rows_ptr[d + 1] - f->h_rows_ptr[d];
    const int Tr = (nr + UM_BULK - 1) / UM_BULK, Trn = (nr + UM_NEXT - 1) / UM_NEXT, Tc = (nr + UN - 1) / UN;
    sky_diag_kernel<<<1, 256, SM_DIAG, s1>>>(d, n, f->ptr, f->start, f->val, f->dinv, f->fail);
    sky_trsm_kernel<<<(nr + TR - 1) / TR, 256, SM_TRSM, s1>>>(d, n, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->dinv, f->val);
    SK(cudaEventRecord(f->ev_trsm[d % NEV], s1));
    // rest(d) on the main stream, concurrent with next(d) / diag(d+1) / trsm(d+1)
    if (Tc > NEXT_TILES) {
      SK(cudaStreamWaitEvent(s0, f->ev_trsm[d % NEV], 0));
      sky_update_kernel<UM_BULK><<<dim3(Tc - NEXT_TILES, Tr), 256, SM_UPD, s0>>>(d, n, NEXT_TILES, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->val);
    }
    SK(cudaEventRecord(f->ev_rest[d % NEV], s0));
    // next(d): columns of panel d+1 (and a little beyond); must follow rest(d-1), which may touch the same columns
    if (d > 0) SK(cudaStreamWaitEvent(s1, f->ev_rest[(d - 1) % NEV], 0));
    sky_update_kernel<UM_NEXT><<<dim3(std::min(Tc, NEXT_TILES), Trn), 256, SM_UPD_NEXT, s1>>>(d, n, 0, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->val);
  }
  // join: the main stream continues after both streams are done
  SK(cudaEventRecord(f->ev_join, s1));
  SK(cudaStreamWaitEvent(s0, f->ev_join, 0));
  SK(cudaGetLastError());
  return PGS_OK;
}

// Backward substitution L^T x = y.  Panels >= D_elim (border, multi-GPU) take x as given in y[] beforehand.
int skyline_backward(SkylineFactor* f, double* y, std::string* err) {
  cudaStream_t st = f->stream;
  const int n = f->n;
  for (int d = f->D - 1; d >= 0; --d) {
    const int cols = d * PW - f->h_lo[d];
    const int grid = std::max(1, std::min(592, (cols + 31) / 32));
    const int given = d >= f->D_elim ? 1 : 0;
    sky_backward_kernel<<<grid, 256, SM_BACK, st>>>(d, n, f->h_lo[d], given, f->ptr, f->start, f->val, f->dinv, f->xacc, y);
  }
  SK(cudaGetLastError());
  return PGS_OK;
}

int skyline_check(SkylineFactor* f, std::string* err) {
  SK(cudaMemcpyAsync(f->h_fail, f->fail, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  SK(cudaStreamSynchronize(f->stream));
  if (*f->h_fail) { if (err) *err = "skyline Cholesky: non-positive pivot"; return PGS_ERR_LINEAR_SOLVER; }
  return PGS_OK;
}

int skyline_factor_solve(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, double* y, std::string* err) {
  if (int rc = skyline_factor(f, Ad, Ao, b, err)) return rc;
  if (int rc = skyline_backward(f, y, err)) return rc;
  return skyline_check(f, err);
}

// ---- border access for the multi-GPU Schur scheme: the trailing (border) rows after a partial factorisation
// out[(i*(i+1))/2 + j] = S[i][j] (packed lower triangle over the nb = n - 6*N_int border scalars), rhs[i] = forward-substituted b.
__global__ void sky_border_get_kernel(int n, int nint, const long long* __restrict__ ptr, const int* __restrict__ start, const double* __restrict__ val,
                                      double* __restrict__ S, double* __restrict__ rhs) {
  const int nb = n - nint;
  const long long tot = (long long)nb * (nb + 1) / 2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= e) ++i;
    while ((long long)i * (i + 1) / 2 > e) --i;
    const int j = (int)(e - (long long)i * (i + 1) / 2);
    const int r = nint + i;
    S[e] = val[ptr[r] + (nint + j - start[r])];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) rhs[i] = val[ptr[n] + nint + i];
}
int skyline_border_get(SkylineFactor* f, double* S_packed, double* rhs, std::string* err) {
  const int nint = f->D_elim * PW;
  sky_border_get_kernel<<<592, 256, 0, f->stream>>>(f->n, nint, f->ptr, f->start, f->val, S_packed, rhs);
  SK(cudaGetLastError());
  return PGS_OK;
}
int skyline_interior_scalars(const SkylineFactor* f) { return f->D_elim * PW; }
const int* skyline_fail_flag(const SkylineFactor* f) { return f->fail; }

}  // namespace pgs
