// K4 (iterative variant): block-Jacobi preconditioned conjugate gradients on the reduced system
//   A y = b,  A = block-sparse SPD over N nodes (6x6 blocks): Ad[N][36] diagonal, Ao[P][36] = block (hi,lo)
// Everything stays on the device; scalars (alpha, beta, residual norms) live in a small device array and
// are produced by a deterministic "last block reduces the per-block partials in fixed order" pattern, so
// one PCG iteration is three launches and no host synchronisation.
#pragma once
#include "pgs_kernels.cuh"

namespace pgs {

// scalar slots
enum { S_RZ0 = 0, S_RZ1 = 1, S_PAP = 2, S_RR = 3, S_BB = 4, S_NSLOTS = 8 };

struct PcgArgs {
  int N, n_pairs;
  const double* __restrict__ Ad; const double* __restrict__ Ao; const double* __restrict__ Minv;
  const int2* __restrict__ pair; const int* __restrict__ adj_ptr; const int* __restrict__ adj_item;  // (pair<<1 | node_is_hi)
  const double* __restrict__ b;
  double* __restrict__ x; double* __restrict__ r; double* __restrict__ rn; double* __restrict__ z; double* __restrict__ p; double* __restrict__ Ap;
  double* __restrict__ partial;   // [3][grid]
  double* __restrict__ scal;      // [S_NSLOTS]
  unsigned int* __restrict__ counter;
};

// Last-arriving block sums partial[0..nblk) in index order and stores it; returns true in that block's thread 0.
__device__ __forceinline__ void finalize_partials(const double* partial, int nblk, double* out, unsigned int* counter, double* sm) {
  __shared__ bool is_last;
  __threadfence();
  if (threadIdx.x == 0) { const unsigned int t = atomicInc(counter, (unsigned int)nblk - 1); is_last = (t == (unsigned int)nblk - 1); }
  __syncthreads();
  if (is_last) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += ((volatile const double*)partial)[i];
    const double t = block_sum(s, sm);
    if (threadIdx.x == 0) *out = t;
  }
}

// 6x6 SPD inverse via Cholesky, one thread per node.
__global__ void __launch_bounds__(128) block_inverse_kernel(int N, const double* __restrict__ Ad, double* __restrict__ Minv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double L[36], X[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) L[k] = Ad[36 * (size_t)i + k];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double d = L[c * 7];
#pragma unroll
    for (int k = 0; k < c; ++k) d -= L[c * 6 + k] * L[c * 6 + k];
    d = sqrt(fmax(d, 1e-300));
    L[c * 7] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      double s = L[r * 6 + c];
#pragma unroll
      for (int k = 0; k < c; ++k) s -= L[r * 6 + k] * L[c * 6 + k];
      L[r * 6 + c] = s * inv;
    }
  }
  // X = L^-1 (lower), then Minv = X^T X
#pragma unroll
  for (int c = 0; c < 6; ++c) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r < c) { X[r * 6 + c] = 0.0; continue; }
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; ++k) if (k >= c) s -= L[r * 6 + k] * X[k * 6 + c];
      X[r * 6 + c] = s / L[r * 7];
    }
  }
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += X[k * 6 + a] * X[k * 6 + c];
      Minv[36 * (size_t)i + a * 6 + c] = s;
    }
}

// x = 0, r = b, z = Minv r, p = z; scal[S_RZ0] = r.z, scal[S_RR] = r.r, scal[S_BB] = b.b
__global__ void __launch_bounds__(192) pcg_init_kernel(PcgArgs A) {
  __shared__ double sm[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, rr = 0.0;
  if (t < 6 * A.N) {
    const int i = t / 6, row = t % 6;
    double zz = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) zz += A.Minv[36 * (size_t)i + row * 6 + c] * A.b[6 * (size_t)i + c];
    const double bi = A.b[t];
    A.x[t] = 0.0; A.r[t] = bi; A.z[t] = zz; A.p[t] = zz;
    rz = bi * zz; rr = bi * bi;
  }
  const double s0 = block_sum(rz, sm), s1 = block_sum(rr, sm);
  if (threadIdx.x == 0) { A.partial[blockIdx.x] = s0; A.partial[gridDim.x + blockIdx.x] = s1; }
  __shared__ bool is_last;
  __threadfence();
  if (threadIdx.x == 0) { const unsigned int c = atomicInc(A.counter, gridDim.x - 1); is_last = (c == gridDim.x - 1); }
  __syncthreads();
  if (is_last) {
    double a = 0.0, b2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { a += ((volatile double*)A.partial)[i]; b2 += ((volatile double*)A.partial)[gridDim.x + i]; }
    const double ta = block_sum(a, sm), tb = block_sum(b2, sm);
    if (threadIdx.x == 0) { A.scal[S_RZ0] = ta; A.scal[S_RR] = tb; A.scal[S_BB] = tb; }
  }
}

// Ap = A p (gather over the node's adjacency, one thread per scalar row); scal[S_PAP] = p.Ap
__global__ void __launch_bounds__(192) pcg_spmv_kernel(PcgArgs A) {
  __shared__ double sm[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double pap = 0.0;
  if (t < 6 * A.N) {
    const int i = t / 6, row = t % 6;
    double acc = 0.0;
    const double* D = A.Ad + 36 * (size_t)i + row * 6;
    const double* pi = A.p + 6 * (size_t)i;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc += D[c] * pi[c];
    for (int q = A.adj_ptr[i]; q < A.adj_ptr[i + 1]; ++q) {
      const int code = __ldg(A.adj_item + q);
      const int pr = code >> 1;
      const int2 hl = A.pair[pr];
      const double* B = A.Ao + 36 * (size_t)pr;
      if (code & 1) {  // this node is the row (hi) node: y += B x_lo
        const double* xv = A.p + 6 * (size_t)hl.y;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += B[row * 6 + c] * xv[c];
      } else {         // this node is the column (lo) node: y += B^T x_hi
        const double* xv = A.p + 6 * (size_t)hl.x;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += B[c * 6 + row] * xv[c];
      }
    }
    A.Ap[t] = acc;
    pap = acc * pi[row];
  }
  const double s = block_sum(pap, sm);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
  finalize_partials(A.partial, gridDim.x, A.scal + S_PAP, A.counter, sm);
}

// alpha = rz/pAp; x += alpha p; rn = r - alpha Ap; z = Minv rn; scal[rz_new] = rn.z, scal[S_RR] = rn.rn
// (the new residual goes to a separate buffer because the 5 sibling rows of a node read r; the host swaps r/rn)
__global__ void __launch_bounds__(192) pcg_update_kernel(PcgArgs A, int rz_old_slot, int rz_new_slot) {
  __shared__ double sm[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double pap = A.scal[S_PAP];
  const double alpha = pap > 0.0 ? A.scal[rz_old_slot] / pap : 0.0;
  double rz = 0.0, rr = 0.0;
  if (t < 6 * A.N) {
    const int i = t / 6, row = t % 6;
    double rn[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) rn[c] = A.r[6 * (size_t)i + c] - alpha * A.Ap[6 * (size_t)i + c];
    double zz = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) zz += A.Minv[36 * (size_t)i + row * 6 + c] * rn[c];
    A.x[t] += alpha * A.p[t];
    A.z[t] = zz;
    A.rn[t] = rn[row];
    rz = rn[row] * zz; rr = rn[row] * rn[row];
  }
  const double s0 = block_sum(rz, sm), s1 = block_sum(rr, sm);
  if (threadIdx.x == 0) { A.partial[blockIdx.x] = s0; A.partial[gridDim.x + blockIdx.x] = s1; }
  __shared__ bool is_last;
  __threadfence();
  if (threadIdx.x == 0) { const unsigned int c = atomicInc(A.counter, gridDim.x - 1); is_last = (c == gridDim.x - 1); }
  __syncthreads();
  if (is_last) {
    double a = 0.0, b2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { a += ((volatile double*)A.partial)[i]; b2 += ((volatile double*)A.partial)[gridDim.x + i]; }
    const double ta = block_sum(a, sm), tb = block_sum(b2, sm);
    if (threadIdx.x == 0) { A.scal[rz_new_slot] = ta; A.scal[S_RR] = tb; }
  }
}

// beta = rz_new/rz_old; p = z + beta p
__global__ void __launch_bounds__(256) pcg_direction_kernel(PcgArgs A, int rz_old_slot, int rz_new_slot) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double o = A.scal[rz_old_slot];
  const double beta = o > 0.0 ? A.scal[rz_new_slot] / o : 0.0;
  if (t < 6 * A.N) A.p[t] = A.z[t] + beta * A.p[t];
}

}  // namespace pgs
