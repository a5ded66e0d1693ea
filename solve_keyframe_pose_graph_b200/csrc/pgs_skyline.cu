// K4 (direct variant): skyline (row-envelope) Cholesky of the reduced pose system, entirely on the device.
//
// Stands in for Ceres' SPARSE_NORMAL_CHOLESKY / CHOLMOD (reference src/PoseGraphSLAM.cpp:1270).  In node order a
// keyframe pose graph is a "thick chain": odometry edges couple i with i-1..i-f, loop edges reach back at most
// a few thousand keyframes, so row i of the Cholesky factor is dense exactly on [min neighbour of i, i] — the
// row envelope — and a skyline factorisation stores no explicit zeros beyond panel alignment.
//
// Layout: scalar row r (6 per node) stores columns [start[r], rowend[r]) contiguously (row-major), start[r] being
// the envelope start rounded down to a panel boundary (PW scalars) and rowend[r] the end of r's own panel, so
// every (row, panel) intersection is a full PW-wide segment.  One extra row n carries b^T: factoring it along
// with the matrix performs the forward substitution for free (row n of L is (L^-1 b)^T).
//
// Right-looking by panels of PW columns; per panel three launches on one stream, no host synchronisation:
//   diag   1 CTA   : L_dd = chol(A_dd) in shared memory, Linv = L_dd^-1 (kept for the solves)
//   trsm   |R|/32  : X = A[R, panel] * Linv^T for the rows R below the panel whose envelope reaches it
//   update tiles   : A[r, c] -= X[r,:] . X[c,:] for r, c in R, c <= r (64x64 tiles, 4x4 register micro-tiles)
// then a backward sweep (one launch per panel) solves L^T x = y.
#include "pgs_skyline.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/pgs.h"

namespace pgs {

constexpr int PW = 96;            // panel width in scalars = 16 nodes
constexpr int PN = PW / 6;
constexpr int UT = 64;            // update tile
constexpr int TR = 32;            // trsm rows per CTA

struct SkylineFactor {
  int N = 0, n = 0, D = 0;         // nodes, scalars, panels
  cudaStream_t stream = nullptr;
  std::vector<int> h_start;        // per scalar row (n+1 entries, last = rhs row)
  std::vector<long long> h_ptr;    // n+2
  std::vector<int> h_rows_ptr;     // D+1
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
  delete f;
}
int64_t skyline_nnz(const SkylineFactor* f) { return f ? f->nnz : 0; }

SkylineFactor* skyline_create(int N, int n_pairs, const int* pair_hi, const int* pair_lo, cudaStream_t stream, std::string* err) {
  SkylineFactor* f = new SkylineFactor();
  f->N = N; f->n = 6 * N; f->D = (f->n + PW - 1) / PW; f->stream = stream; f->n_pairs = n_pairs;
  const int n = f->n, D = f->D;
  // ---- symbolic: envelope start per node = min neighbour, rounded down to a panel boundary
  std::vector<int> nstart(N);
  for (int i = 0; i < N; ++i) nstart[i] = i;
  for (int p = 0; p < n_pairs; ++p) nstart[pair_hi[p]] = std::min(nstart[pair_hi[p]], pair_lo[p]);
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

  auto bad = [&](cudaError_t e, const char* what) { if (err) *err = std::string("skyline_create: ") + cudaGetErrorString(e) + " (" + what + ", factor needs " + std::to_string((double)f->nnz * 8 / 1e9) + " GB)"; skyline_destroy(f); return (SkylineFactor*)nullptr; };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&f->val, sizeof(double) * (size_t)f->nnz)) != cudaSuccess) return bad(e, "val");
  if ((e = cudaMalloc((void**)&f->ptr, sizeof(long long) * (n + 2))) != cudaSuccess) return bad(e, "ptr");
  if ((e = cudaMalloc((void**)&f->start, sizeof(int) * (n + 1))) != cudaSuccess) return bad(e, "start");
  if ((e = cudaMalloc((void**)&f->rows_ptr, sizeof(int) * (D + 1))) != cudaSuccess) return bad(e, "rows_ptr");
  if ((e = cudaMalloc((void**)&f->rows_idx, sizeof(int) * std::max<size_t>(rows_idx.size(), 1))) != cudaSuccess) return bad(e, "rows_idx");
  if ((e = cudaMalloc((void**)&f->dinv, sizeof(double) * (size_t)D * PW * PW)) != cudaSuccess) return bad(e, "dinv");
  if ((e = cudaMalloc((void**)&f->xacc, sizeof(double) * (size_t)D * PW)) != cudaSuccess) return bad(e, "xacc");
  if ((e = cudaMalloc((void**)&f->fail, sizeof(int))) != cudaSuccess) return bad(e, "fail");
  if ((e = cudaMallocHost((void**)&f->h_fail, sizeof(int))) != cudaSuccess) return bad(e, "h_fail");
  if ((e = cudaMalloc((void**)&f->pair_hi, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_hi");
  if ((e = cudaMalloc((void**)&f->pair_lo, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_lo");
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

// diag: factor the PW x PW diagonal block of panel d in shared memory, store L back, store Linv.
__global__ void __launch_bounds__(256) sky_diag_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                       double* __restrict__ val, double* __restrict__ dinv, int* __restrict__ fail) {
  extern __shared__ double sm_diag[];
  double (*L)[PW + 1] = reinterpret_cast<double (*)[PW + 1]>(sm_diag);
  double (*X)[PW + 1] = reinterpret_cast<double (*)[PW + 1]>(sm_diag + PW * (PW + 1));
  __shared__ int bad;
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  if (tid == 0) bad = 0;
  for (int e = tid; e < PW * PW; e += blockDim.x) {
    const int i = e / PW, j = e % PW;
    double v = 0.0;
    if (i < w && j <= i) { const int r = c0 + i; v = val[ptr[r] + (c0 + j - start[r])]; }
    L[i][j] = v;
  }
  __syncthreads();
  // blocked right-looking Cholesky, 8 columns at a time: every thread owning a row re-factors the 8x8 diagonal
  // block in registers (no synchronisation), then solves its own row against it.
  for (int jb = 0; jb < w; jb += 8) {
    const int bw = min(8, w - jb);
    const int i = jb + tid;
    double Dg[8][8], arow[8];
    if (i < w) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) Dg[a][c] = (a < bw && c <= a) ? L[jb + a][jb + c] : (a == c ? 1.0 : 0.0);
#pragma unroll
      for (int c = 0; c < 8; ++c) arow[c] = (c < bw) ? L[i][jb + c] : 0.0;
    }
    __syncthreads();   // everybody has read the diagonal block before its owners overwrite it
    if (i < w) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double dd = Dg[c][c];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < c) dd -= Dg[c][k] * Dg[c][k];
        if (!(dd > 0.0)) { bad = 1; dd = 1.0; }
        dd = sqrt(dd); Dg[c][c] = dd;
        const double inv = 1.0 / dd;
#pragma unroll
        for (int a = 0; a < 8; ++a) if (a > c) {
          double s = Dg[a][c];
#pragma unroll
          for (int k = 0; k < 8; ++k) if (k < c) s -= Dg[a][k] * Dg[c][k];
          Dg[a][c] = s * inv;
        }
      }
      double row[8];
      if (tid < bw) {
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
          row[c] = s / Dg[c][c];
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) if (c < bw) L[i][jb + c] = row[c];
    }
    __syncthreads();
    // trailing update: A[i][k] -= sum_c L[i][jb+c] L[k][jb+c] for jb+bw <= k <= i
    const int t0 = jb + bw, m = w - t0;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int ii = e / m, kk = e % m;
      if (kk <= ii) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c < bw) s += L[t0 + ii][jb + c] * L[t0 + kk][jb + c];
        L[t0 + ii][t0 + kk] -= s;
      }
    }
    __syncthreads();
  }
  // X = L^-1 (lower triangular): thread pair-of-4 per column, forward substitution
  for (int e = tid; e < PW * PW; e += blockDim.x) X[e / PW][e % PW] = 0.0;
  __syncthreads();
  {
    const int j = tid >> 2, part = tid & 3;   // 4 lanes cooperate on one column; groups in a warp run different trip counts
    const unsigned gmask = 0xFu << ((tid & 31) & ~3);
    for (int col = j; col < w; col += (blockDim.x >> 2)) {
      for (int i = col; i < w; ++i) {
        double s = 0.0;
        for (int k = col + part; k < i; k += 4) s += L[i][k] * X[k][col];
        s += __shfl_xor_sync(gmask, s, 1); s += __shfl_xor_sync(gmask, s, 2);
        if (part == 0) X[i][col] = ((i == col ? 1.0 : 0.0) - s) / L[i][i];
        __syncwarp(gmask);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < PW * PW; e += blockDim.x) {
    const int i = e / PW, j = e % PW;
    if (i < w && j <= i) { const int r = c0 + i; val[ptr[r] + (c0 + j - start[r])] = L[i][j]; }
    dinv[(size_t)d * PW * PW + e] = (i < w && j < w) ? X[i][j] : 0.0;
  }
  if (tid == 0 && bad) *fail = 1;
}

// trsm: X[r][j] = sum_{k<=j} A[r][c0+k] * Linv[j][k] for the rows r in R_d, in place.
__global__ void __launch_bounds__(256) sky_trsm_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                       const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx,
                                                       const double* __restrict__ dinv, double* __restrict__ val) {
  extern __shared__ double sm[];
  double* Li = sm;                    // [PW][PW+1]  Linv
  double* A = sm + PW * (PW + 1);     // [TR][PW+1]
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  const int row0 = blockIdx.x * TR;
  for (int e = tid; e < PW * PW; e += blockDim.x) Li[(e / PW) * (PW + 1) + (e % PW)] = dinv[(size_t)d * PW * PW + e];
  for (int e = tid; e < TR * PW; e += blockDim.x) {
    const int i = e / PW, k = e % PW;
    double v = 0.0;
    if (row0 + i < nr && k < w) { const int r = rows_idx[rb + row0 + i]; v = val[ptr[r] + (c0 + k - start[r])]; }
    A[i * (PW + 1) + k] = v;
  }
  __syncthreads();
  for (int e = tid; e < TR * PW; e += blockDim.x) {
    const int i = e / PW, j = e % PW;
    if (row0 + i < nr && j < w) {
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for (; k + 1 <= j; k += 2) { s0 += A[i * (PW + 1) + k] * Li[j * (PW + 1) + k]; s1 += A[i * (PW + 1) + k + 1] * Li[j * (PW + 1) + k + 1]; }
      if (k <= j) s0 += A[i * (PW + 1) + k] * Li[j * (PW + 1) + k];
      const int r = rows_idx[rb + row0 + i];
      val[ptr[r] + (c0 + j - start[r])] = s0 + s1;
    }
  }
}

// update: A[r][c] -= X[r,:] . X[c,:] over lower-triangular 64x64 tiles of R_d x R_d.
__global__ void __launch_bounds__(256) sky_update_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                         const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx, double* __restrict__ val) {
  extern __shared__ double sm[];
  constexpr int LD = UT + 2;          // k-major tiles, padded, 16-B aligned rows
  double* Xr = sm;                    // [PW][LD]
  double* Xc = sm + PW * LD;          // [PW][LD]
  __shared__ int s_row[UT], s_col[UT];
  __shared__ long long s_rbase[UT];   // ptr[r] - start[r]
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  // linear block id -> (ti >= tj)
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while (ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  if (tid < UT) {
    const int ir = ti * UT + tid, ic = tj * UT + tid;
    const int r = ir < nr ? rows_idx[rb + ir] : -1, c = ic < nr ? rows_idx[rb + ic] : -1;
    s_row[tid] = r; s_col[tid] = c;
    s_rbase[tid] = r >= 0 ? ptr[r] - start[r] : 0;
  }
  __syncthreads();
  for (int e = tid; e < UT * PW; e += blockDim.x) {
    const int i = e / PW, k = e % PW;
    double vr = 0.0, vc = 0.0;
    if (k < w) {
      const int r = s_row[i], c = s_col[i];
      if (r >= 0) vr = val[s_rbase[i] + c0 + k];
      if (c >= 0 && c < n) vc = val[ptr[c] + (c0 + k - start[c])];
    }
    Xr[k * LD + i] = vr; Xc[k * LD + i] = vc;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int k = 0; k < PW; ++k) {
    const double2 a01 = *reinterpret_cast<const double2*>(&Xr[k * LD + ty * 4]);
    const double2 a23 = *reinterpret_cast<const double2*>(&Xr[k * LD + ty * 4 + 2]);
    const double2 b01 = *reinterpret_cast<const double2*>(&Xc[k * LD + tx * 4]);
    const double2 b23 = *reinterpret_cast<const double2*>(&Xc[k * LD + tx * 4 + 2]);
    const double a[4] = {a01.x, a01.y, a23.x, a23.y}, bb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = s_row[ty * 4 + i];
    if (r < 0) continue;
    const long long base = s_rbase[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = s_col[tx * 4 + j];
      if (c >= 0 && c < n && c <= r) val[base + c] -= acc[i][j];
    }
  }
}

// backward sweep for panel d (processed D-1 .. 0):  x_d = Linv^T (y_d + acc_d), then push
//   acc[c] -= sum_{r in panel} L[r][c] x_r   for every column c < c0 inside the rows' envelopes.
__global__ void __launch_bounds__(256) sky_backward_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                           const double* __restrict__ val, const double* __restrict__ dinv,
                                                           double* __restrict__ acc, double* __restrict__ x) {
  __shared__ double xs[PW];
  __shared__ double rhs[PW];
  __shared__ int smin;
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  if (tid < PW) rhs[tid] = tid < w ? val[ptr[n] + c0 + tid] + acc[c0 + tid] : 0.0;
  if (tid == 0) { int m = c0; for (int i = 0; i < w; ++i) m = min(m, start[c0 + i]); smin = m; }
  __syncthreads();
  if (tid < PW) {
    double s = 0.0;   // x_j = sum_{i>=j} Linv[i][j] rhs_i
    for (int i = tid; i < w; ++i) s += dinv[(size_t)d * PW * PW + i * PW + tid] * rhs[i];
    xs[tid] = s;
    if (blockIdx.x == 0 && tid < w) x[c0 + tid] = s;
  }
  __syncthreads();
  const int lo = smin;
  for (int c = lo + blockIdx.x * blockDim.x + tid; c < c0; c += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < w; ++i) { const int r = c0 + i; const int st = start[r]; if (c >= st) s += val[ptr[r] + (c - st)] * xs[i]; }
    acc[c] -= s;
  }
}

int skyline_factor_solve(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, double* y, std::string* err) {
  cudaStream_t st = f->stream;
  const int n = f->n, D = f->D;
  static bool attr_set = false;
  const size_t sm_trsm = sizeof(double) * (PW * (PW + 1) + TR * (PW + 1));
  const size_t sm_upd = sizeof(double) * (2 * PW * (UT + 2));
  const size_t sm_diag = sizeof(double) * (2 * PW * (PW + 1));
  if (!attr_set) {
    SK(cudaFuncSetAttribute(sky_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_diag));
    SK(cudaFuncSetAttribute(sky_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_trsm));
    SK(cudaFuncSetAttribute(sky_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_upd));
    attr_set = true;
  }
  SK(cudaMemsetAsync(f->val, 0, sizeof(double) * (size_t)f->nnz, st));
  SK(cudaMemsetAsync(f->xacc, 0, sizeof(double) * (size_t)D * PW, st));
  SK(cudaMemsetAsync(f->fail, 0, sizeof(int), st));
  const int tot = std::max(36 * std::max(f->N, f->n_pairs), n);
  sky_scatter_kernel<<<(tot + 255) / 256, 256, 0, st>>>(f->N, f->n_pairs, Ad, Ao, b, f->pair_hi, f->pair_lo, f->ptr, f->start, f->val);
  for (int d = 0; d < D; ++d) {
    const int nr = f->h_rows_ptr[d + 1] - f->h_rows_ptr[d];
    sky_diag_kernel<<<1, 256, sm_diag, st>>>(d, n, f->ptr, f->start, f->val, f->dinv, f->fail);
    sky_trsm_kernel<<<(nr + TR - 1) / TR, 256, sm_trsm, st>>>(d, n, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->dinv, f->val);
    const int T = (nr + UT - 1) / UT;
    sky_update_kernel<<<T * (T + 1) / 2, 256, sm_upd, st>>>(d, n, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->val);
  }
  for (int d = D - 1; d >= 0; --d) {
    // columns to push into: from the smallest envelope start of the panel's rows up to c0
    int lo = d * PW;
    for (int i = d * PW; i < std::min(n, (d + 1) * PW); ++i) lo = std::min(lo, f->h_start[i]);
    const int cols = d * PW - lo;
    const int grid = std::max(1, std::min(296, (cols + 255) / 256));
    sky_backward_kernel<<<grid, 256, 0, st>>>(d, n, f->ptr, f->start, f->val, f->dinv, f->xacc, y);
  }
  SK(cudaGetLastError());
  SK(cudaMemcpyAsync(f->h_fail, f->fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  if (*f->h_fail) { if (err) *err = "skyline Cholesky: non-positive pivot"; return PGS_ERR_LINEAR_SOLVER; }
  return PGS_OK;
}

}  // namespace pgs
