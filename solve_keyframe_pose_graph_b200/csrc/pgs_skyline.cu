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
// Right-looking by panels of PW columns on three streams; the panel chain runs one panel ahead of the bulk update:
//   chain stream   C(d):     1 CTA.  Applies the one update the bulk has not yet given the diagonal block — the
//                            rank-96 product of X[d, d-1] from trsm(d-1) — then L_dd = chol(A_dd) and Linv = L_dd^-1,
//                            every matrix-shaped piece on FP64 tensor cores (DMMA).  Waits for trsm(d-1), rest(d-2).
//                  trsm(d):  X = A[R, panel] * Linv^T for the rows R below the panel whose envelope reaches it.
//                            Behind C(d) on the same stream (chain mode 1); waits for next(d-1).
//   panel stream   next(d):  the update tiles that hold the columns of panel d+1 (what trsm(d+1) needs), except the
//                            diagonal block of panel d+1: C(d+1) does that one itself, which is what lets it start
//                            right after trsm(d) instead of after the update.  Waits for trsm(d), rest(d-1).
//   main stream    rest(d):  every other tile of A[r, c] -= X[r,:] . X[c,:], r, c in R, c <= r.  128x64 DMMA tiles in a
//                            persistent kernel fed by bulk copies of the packed panel.  Waits for trsm(d).
// so C(d+1), next(d) and rest(d) overlap, and a panel costs max(C + trsm + two launch gaps, rest) (c3, one chain: 38.8 us).
// Then a backward sweep (one launch per panel, programmatic dependent launches) solves L^T x = y.
//
// Partial factorisation (multi-GPU domain decomposition): only the first n_elim panels are
// eliminated; the trailing rows then hold the Schur complement on the border unknowns and the border part of
// the forward-substituted right-hand side.
#include "pgs_skyline.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <vector>

#include "../../include/pgs.h"

namespace pgs {

constexpr int PW = 96;            // panel width in scalars = 16 nodes
constexpr int PN = PW / 6;
constexpr int TR = 32;            // trsm rows per CTA
constexpr int LDT = PW + 4;       // 100 doubles = 200 words = 8 mod 32: conflict-free DMMA fragment loads (8 rows x 4 k per half-warp pair)
constexpr int UM = 128, UN = 64;  // update tiles: rows x cols
constexpr int KC = 32;            // update K chunk
constexpr int LDK = KC + 4;       // 36 doubles = 72 words = 8 mod 32: conflict-free DMMA fragment loads, 16-B aligned rows
constexpr int NEV = 8;            // event ring
constexpr int XP_RING = 4;        // packed panel copies in flight: trsm(d) cannot run before rest(d-2) is done

struct SkylineFactor {
  int N = 0, n = 0, D = 0;         // nodes, scalars, panels
  int D_elim = 0;                  // panels to eliminate (== D for a full factorisation)
  int share = 1;                   // factorisations that run on this GPU at the same time (chains): the persistent update takes 1/share of its SMs
  cudaStream_t stream = nullptr, s1 = nullptr, s2 = nullptr;   // main (rest), chain (C), panel (trsm + next)
  cudaEvent_t ev_trsm[NEV], ev_rest[NEV], ev_c[NEV], ev_next[NEV], ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  unsigned int* sched = nullptr;   // [2] CTAs-done counter of the backward sweep (self-resetting)
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
  cudaGraphExec_t bw_graph = nullptr; double* bw_y = nullptr; bool bw_graph_failed = false;   // the backward sweep as a replayable graph (one tiny launch per panel)
  // Packed copy of the current panel's X = A[R, panel] Linv^T for the trailing update (ring over panels): layout
  // [K chunk][list position][LDK] with the shared-memory padding already in place, so that a tile's operand chunk — 128 or 64
  // consecutive list positions — is ONE contiguous block a single bulk copy can fetch; rinfo = per list position the row index
  // (-1 past the end of the list) and the offset of (row, column 0) in val.
  double* xp = nullptr; long long* rinfo = nullptr; long long xp_stride = 0, rinfo_stride = 0; int rpad = 0;
  int* node_src = nullptr; int* pair_src = nullptr;   // optional: factor node -> row of Ad / b (-1 = none), factor pair -> row of Ao
  long long tail = 0;              // extra doubles behind the envelope (travel with it in the border all-reduce)
};

#ifdef SKY_TIMELINE   // tools/timeline_lab.py: first-CTA-in / last-CTA-out wall-clock stamps (%globaltimer, ns) of every factorisation
                      // kernel for panels [TL_D0, TL_D0 + TL_ND) of up to four factors that run at the same time
constexpr int TL_D0 = 1000, TL_ND = 24, TL_KINDS = 5;
__device__ unsigned long long g_tl[4][TL_KINDS][TL_ND][2];
__device__ const double* g_tl_val[4];
__device__ __forceinline__ void tl_stamp(const double* val, int kind, int d, int which) {
  if (threadIdx.x != 0 || d < TL_D0 || d >= TL_D0 + TL_ND) return;
  int ch = -1;
  for (int i = 0; i < 4; ++i) if (g_tl_val[i] == val) ch = i;
  if (ch < 0) return;
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (which == 0) atomicMin(&g_tl[ch][kind][d - TL_D0][0], t); else atomicMax(&g_tl[ch][kind][d - TL_D0][1], t);
}
#define TL_IN(val, kind, d) tl_stamp(val, kind, d, 0)
#define TL_OUT(val, kind, d) tl_stamp(val, kind, d, 1)
#else
#define TL_IN(val, kind, d) do { } while (0)
#define TL_OUT(val, kind, d) do { } while (0)
#endif

#define SK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { if (err) *err = std::string("skyline: ") + cudaGetErrorString(e__) + " at " #x; return e__ == cudaErrorMemoryAllocation ? PGS_ERR_OUT_OF_MEMORY : PGS_ERR_CUDA; } } while (0)

void skyline_destroy(SkylineFactor* f) {
  if (!f) return;
  cudaFree(f->val); cudaFree(f->ptr); cudaFree(f->start); cudaFree(f->rows_ptr); cudaFree(f->rows_idx); cudaFree(f->dinv);
  cudaFree(f->xacc); cudaFree(f->fail); cudaFree(f->pair_hi); cudaFree(f->pair_lo); cudaFree(f->sched); cudaFree(f->node_src); cudaFree(f->pair_src); cudaFree(f->xp); cudaFree(f->rinfo);
  if (f->bw_graph) cudaGraphExecDestroy(f->bw_graph);
  if (f->h_fail) cudaFreeHost(f->h_fail);
  for (int i = 0; i < NEV; ++i) { if (f->ev_trsm[i]) cudaEventDestroy(f->ev_trsm[i]); if (f->ev_rest[i]) cudaEventDestroy(f->ev_rest[i]); if (f->ev_c[i]) cudaEventDestroy(f->ev_c[i]); if (f->ev_next[i]) cudaEventDestroy(f->ev_next[i]); }
  if (f->ev_fork) cudaEventDestroy(f->ev_fork);
  if (f->ev_join) cudaEventDestroy(f->ev_join);
  if (f->ev_join2) cudaEventDestroy(f->ev_join2);
  if (f->s1) cudaStreamDestroy(f->s1);
  if (f->s2) cudaStreamDestroy(f->s2);
  delete f;
}
int64_t skyline_nnz(const SkylineFactor* f) { return f ? f->nnz : 0; }
void skyline_set_share(SkylineFactor* f, int share) { if (f) f->share = share < 1 ? 1 : share; }
int skyline_panel_width() { return PW; }

SkylineFactor* skyline_create(int N, int n_pairs, const int* pair_hi, const int* pair_lo, cudaStream_t stream, std::string* err,
                              int n_border_nodes, bool dense, const int* node_src, const int* pair_src, long long tail) {
  SkylineFactor* f = new SkylineFactor();
  for (int i = 0; i < NEV; ++i) { f->ev_trsm[i] = nullptr; f->ev_rest[i] = nullptr; f->ev_c[i] = nullptr; f->ev_next[i] = nullptr; }
  f->N = N; f->n = 6 * N; f->D = (f->n + PW - 1) / PW; f->stream = stream; f->n_pairs = n_pairs; f->tail = tail;
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
  if ((e = cudaMalloc((void**)&f->val, sizeof(double) * (size_t)(f->nnz + f->tail))) != cudaSuccess) return bad(e, "val");
  if (node_src) { if ((e = cudaMalloc((void**)&f->node_src, sizeof(int) * std::max(N, 1))) != cudaSuccess) return bad(e, "node_src");
                  cudaMemcpyAsync(f->node_src, node_src, sizeof(int) * N, cudaMemcpyHostToDevice, stream); }
  if (pair_src) { if ((e = cudaMalloc((void**)&f->pair_src, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_src");
                  cudaMemcpyAsync(f->pair_src, pair_src, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream); }
  if ((e = cudaMalloc((void**)&f->ptr, sizeof(long long) * (n + 2))) != cudaSuccess) return bad(e, "ptr");
  if ((e = cudaMalloc((void**)&f->start, sizeof(int) * (n + 1))) != cudaSuccess) return bad(e, "start");
  if ((e = cudaMalloc((void**)&f->rows_ptr, sizeof(int) * (D + 1))) != cudaSuccess) return bad(e, "rows_ptr");
  if ((e = cudaMalloc((void**)&f->rows_idx, sizeof(int) * std::max<size_t>(rows_idx.size(), 1))) != cudaSuccess) return bad(e, "rows_idx");
  if ((e = cudaMalloc((void**)&f->dinv, sizeof(double) * (size_t)std::max(D, 1) * PW * PW)) != cudaSuccess) return bad(e, "dinv");
  if ((e = cudaMemsetAsync(f->dinv, 0, sizeof(double) * (size_t)std::max(D, 1) * PW * PW, stream)) != cudaSuccess) return bad(e, "dinv");   // the kernels write the lower triangles only
  if ((e = cudaMalloc((void**)&f->xacc, sizeof(double) * (size_t)std::max(D, 1) * PW)) != cudaSuccess) return bad(e, "xacc");
  if ((e = cudaMalloc((void**)&f->fail, sizeof(int))) != cudaSuccess) return bad(e, "fail");
  f->rpad = ((f->max_rows + UM - 1) / UM) * UM;
  f->xp_stride = (long long)(PW / KC) * f->rpad * LDK; f->rinfo_stride = 2LL * f->rpad;
  if ((e = cudaMalloc((void**)&f->xp, sizeof(double) * (size_t)std::max<long long>(XP_RING * f->xp_stride, 1))) != cudaSuccess) return bad(e, "xp");
  if ((e = cudaMalloc((void**)&f->rinfo, sizeof(long long) * (size_t)std::max<long long>(XP_RING * f->rinfo_stride, 1))) != cudaSuccess) return bad(e, "rinfo");
  if ((e = cudaMalloc((void**)&f->sched, 4 * sizeof(unsigned int))) != cudaSuccess) return bad(e, "sched");
  if ((e = cudaMemsetAsync(f->sched, 0, 4 * sizeof(unsigned int), stream)) != cudaSuccess) return bad(e, "sched");
  if ((e = cudaMallocHost((void**)&f->h_fail, sizeof(int))) != cudaSuccess) return bad(e, "h_fail");
  if ((e = cudaMalloc((void**)&f->pair_hi, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_hi");
  if ((e = cudaMalloc((void**)&f->pair_lo, sizeof(int) * std::max(n_pairs, 1))) != cudaSuccess) return bad(e, "pair_lo");
  { int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);   // the panel stream carries the critical path
    if ((e = cudaStreamCreateWithPriority(&f->s1, cudaStreamNonBlocking, hi_p)) != cudaSuccess) return bad(e, "stream");
    if ((e = cudaStreamCreateWithPriority(&f->s2, cudaStreamNonBlocking, hi_p)) != cudaSuccess) return bad(e, "stream"); }
  for (int i = 0; i < NEV; ++i) {
    if ((e = cudaEventCreateWithFlags(&f->ev_trsm[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
    if ((e = cudaEventCreateWithFlags(&f->ev_rest[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
    if ((e = cudaEventCreateWithFlags(&f->ev_c[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
    if ((e = cudaEventCreateWithFlags(&f->ev_next[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  }
  if ((e = cudaEventCreateWithFlags(&f->ev_join2, cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  if ((e = cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  if ((e = cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bad(e, "event");
  cudaMemcpyAsync(f->ptr, f->h_ptr.data(), sizeof(long long) * (n + 2), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(f->start, f->h_start.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(f->rows_ptr, f->h_rows_ptr.data(), sizeof(int) * (D + 1), cudaMemcpyHostToDevice, stream);
  if (!rows_idx.empty()) cudaMemcpyAsync(f->rows_idx, rows_idx.data(), sizeof(int) * rows_idx.size(), cudaMemcpyHostToDevice, stream);
  if (n_pairs) { cudaMemcpyAsync(f->pair_hi, pair_hi, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream);
                 cudaMemcpyAsync(f->pair_lo, pair_lo, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream); }
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bad(e, "upload");
#ifdef SKY_TIMELINE
  { static int n_reg = 0;
    if (n_reg == 0) { static unsigned long long init[4][TL_KINDS][TL_ND][2]; for (auto& a : init) for (auto& b : a) for (auto& c : b) { c[0] = ~0ull; c[1] = 0; } cudaMemcpyToSymbol(g_tl, init, sizeof(init)); }
    if (n_reg < 4) { const double* v = f->val; cudaMemcpyToSymbol(g_tl_val, &v, sizeof(v), sizeof(v) * n_reg); ++n_reg; } }
#endif
  return f;
}

// ------------------------------------------------------------------------------------------------ kernels
// scatter the block system into the (zeroed) envelope: diagonal blocks (lower triangle), off-diagonal blocks, rhs row
// node_src / pair_src (may be null = identity): the factor holds a sub-system — node i takes its diagonal block and
// right-hand side from row node_src[i] of Ad / b (-1: none, the entries stay zero), pair p its block from row pair_src[p] of Ao.
__global__ void sky_scatter_kernel(int N, int n_pairs, const double* __restrict__ Ad, const double* __restrict__ Ao, const double* __restrict__ b,
                                   const int* __restrict__ pair_hi, const int* __restrict__ pair_lo, const int* __restrict__ node_src,
                                   const int* __restrict__ pair_src, const long long* __restrict__ ptr,
                                   const int* __restrict__ start, double* __restrict__ val) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 6 * N;
  if (t < 36 * N) {
    const int i = t / 36, a = (t % 36) / 6, c = t % 6;
    const int src = node_src ? node_src[i] : i;
    if (c <= a && src >= 0) { const int r = 6 * i + a; val[ptr[r] + (6 * i + c - start[r])] = Ad[36 * (size_t)src + 6 * a + c]; }
  }
  if (t < 36 * n_pairs) {
    const int p = t / 36, a = (t % 36) / 6, c = t % 6;
    const int r = 6 * pair_hi[p] + a;
    const int src = pair_src ? pair_src[p] : p;
    val[ptr[r] + (6 * pair_lo[p] + c - start[r])] = Ao[36 * (size_t)src + 6 * a + c];
  }
  if (t < n) { const int src = node_src ? node_src[t / 6] : t / 6; if (src >= 0) val[ptr[n] + t] = b[6 * (size_t)src + t % 6]; }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
// programmatic dependent launch: wait = the launch before this one on the stream is complete and its writes are visible;
// launch = the launch after this one may be scheduled (it runs up to its own wait)
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NGROUPS>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NGROUPS)); }

// FP64 tensor-core step: D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l>>2][l&3], B[l&3][l>>2] and
// D[l>>2][2*(l&3) + {0,1}].  On B200 DMMA sustains 37 TFLOP/s against 31-34 for DFMA (tools/fp64_lab.cu) and
// needs one shared-memory load per 2 (128x64 tile) / 1 (32x64 tile) mma instead of 12 per 32 FMAs.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// diag: factor the PW x PW diagonal block of panel d in shared memory and store X = L^-1.
// One CTA on the critical path of the whole factorisation, so every matrix-shaped piece runs on DMMA and the
// inverse is built alongside the factor.  Twelve 8-column steps of two barrier phases each:
//   phase A  row threads (tid < rows left): re-factor the 8x8 diagonal block in registers (redundantly, no
//            synchronisation inside) and solve their own row of the block column;
//            thread 255: the same factor, then W = its inverse -> Dinv[I] and the diagonal block of X;
//            free warps: finish X row-block I-1 = -W_{I-1} * S (S from the previous phase B).
//   phase B  8x8 DMMA blocks of the trailing update  A[i,k] -= L[i,I] L[k,I]^T  and of
//            S = L[I, 0:jb] * X[0:jb, 0:jb] for row-block I, dealt round-robin to the 8 warps.
// Fragment layout of dmma884: g = lane >> 2, t = lane & 3; A[g][t], B[t][g], C[g][2t..2t+1].
__device__ __forceinline__ void diag_block_of(int b, int& bi, int& bk) {
  bi = (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
  while (bi * (bi + 1) / 2 > b) --bi;
  bk = b - bi * (bi + 1) / 2;
}
#ifdef SKY_DIAG_CLOCKS   // tools/diag_lab.cu: cycle stamps of thread 0 after every barrier phase
__device__ long long g_diag_clk[64];
#define DIAG_STAMP(i) do { if (threadIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#define DIAG_STAMP_U(i) do { if (threadIdx.x == 128) g_diag_clk[i] = clock64(); } while (0)
#else
#define DIAG_STAMP(i) do { } while (0)
#define DIAG_STAMP_U(i) do { } while (0)
#endif
__global__ void __launch_bounds__(256) sky_diag_kernel(int d, int n, int prev, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                       double* __restrict__ val, double* __restrict__ dinv, int* __restrict__ fail) {
  constexpr int NB = PW / 8, LDQ = LDT;
  DIAG_STAMP(0);
  extern __shared__ __align__(16) double sm_diag[];
  double* L = sm_diag;                 // [PW][LDQ]
  double* X = sm_diag + PW * LDQ;      // [PW][LDQ]  only the lower block triangle is ever written or read
  double* Dinv = X + PW * LDQ;         // [NB][64] inverses of the 8x8 diagonal factors (full 8x8, upper part zero)
  double* Sb = Dinv + NB * 64;         // [8][LDQ]  S of the current row-block
  __shared__ long long rbase[PW];
  __shared__ int bad;
  __shared__ unsigned char blk_i[NB * (NB + 1) / 2], blk_k[NB * (NB + 1) / 2];   // b -> (bi, bk), row-major lower block triangle
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) bad = 0;
  if (tid < NB * (NB + 1) / 2) { int bi, bk; diag_block_of(tid, bi, bk); blk_i[tid] = (unsigned char)bi; blk_k[tid] = (unsigned char)bk; }
  __shared__ int rprev[PW];            // 1 = the row's envelope reaches panel d-1
  if (tid < PW) {
    const int r = c0 + tid;
    const int st = tid < w ? start[r] : 0;
    rbase[tid] = tid < w ? ptr[r] + (c0 - st) : 0;
    rprev[tid] = (prev && tid < w && st <= c0 - PW) ? 1 : 0;
  }
  __syncthreads();
  // warp wid loads rows wid, wid+8, ...; a lane the 16-byte pairs lane and lane+32 (48 pairs per row)
#pragma unroll 4
  for (int i = wid; i < PW; i += 8) {
    const double* src = val + rbase[i];
    const bool row_ok = i < w;
    {
      const bool ok = row_ok && 2 * lane <= i;               // lower triangle by pairs, the rest zero-filled
      cp_async16(&L[i * LDQ + 2 * lane], ok ? (const void*)(src + 2 * lane) : (const void*)val, ok);
    }
    if (lane < PW / 2 - 32) {
      const int j2 = lane + 32;
      const bool ok = row_ok && 2 * j2 <= i;
      cp_async16(&L[i * LDQ + 2 * j2], ok ? (const void*)(src + 2 * j2) : (const void*)val, ok);
    }
    if (prev) {
      // X[d, d-1] (written by trsm(d-1)) staged in the X buffer, which the factorisation does not touch before step 0
      const bool ok = rprev[i] != 0;
      cp_async16(&X[i * LDQ + 2 * lane], ok ? (const void*)(src - PW + 2 * lane) : (const void*)val, ok);
      if (lane < PW / 2 - 32) cp_async16(&X[i * LDQ + 2 * (lane + 32)], ok ? (const void*)(src - PW + 2 * (lane + 32)) : (const void*)val, ok);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (prev) {
    // A_dd -= X X^T on the lower block triangle: 78 blocks of 8x8, two per warp in flight, K = 96
    constexpr int NBLK = NB * (NB + 1) / 2;
    for (int b = 2 * wid; b < NBLK; b += 16) {
      const bool two = b + 1 < NBLK;
      const int bsec = two ? b + 1 : b;
      const int bi0 = blk_i[b], bk0 = blk_k[b], bi1 = blk_i[bsec], bk1 = blk_k[bsec];
      double2* cp0 = reinterpret_cast<double2*>(&L[(8 * bi0 + g) * LDQ + 8 * bk0 + 2 * t]);
      double2* cp1 = reinterpret_cast<double2*>(&L[(8 * bi1 + g) * LDQ + 8 * bk1 + 2 * t]);
      double2 u = *cp0, v = *cp1;
      const double* a0 = &X[(8 * bi0 + g) * LDQ + t]; const double* b0 = &X[(8 * bk0 + g) * LDQ + t];
      const double* a1 = &X[(8 * bi1 + g) * LDQ + t]; const double* b1 = &X[(8 * bk1 + g) * LDQ + t];
#pragma unroll 6
      for (int k = 0; k < PW; k += 4) {
        dmma884(u.x, u.y, -a0[k], b0[k]);
        dmma884(v.x, v.y, -a1[k], b1[k]);
      }
      *cp0 = u;
      if (two) *cp1 = v;
    }
    __syncthreads();
  }
  if (tid >= w && tid < PW) L[tid * LDQ + tid] = 1.0;         // identity padding of a short last panel
  __syncthreads();
  DIAG_STAMP(1);
  double Dg[8][8], dinv8[8];
  for (int I = 0; I < NB; ++I) {
    const int jb = 8 * I, nrow = PW - jb;
    const bool row_thread = tid < nrow, inv_thread = tid == 255;
    // ---------------- phase A
    if (row_thread || inv_thread) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) Dg[a][c] = (c <= a) ? L[(jb + a) * LDQ + jb + c] : 0.0;
      // Right-looking (outer-product) form, everything in registers: as soon as column c is scaled, its rank-1 update
      // goes into the columns to its right — independent FMAs — so the next pivot is one FMA behind the previous one
      // instead of a dot product of growing length (the dependent chain per column is rsqrt + mul + fma).  The thread's
      // own row of the block column rides along: row[b] -= row[c] * Dg[b][c], the same operations a forward
      // substitution would do, without its serial dot products.
      double arow[8];
      if (row_thread) {
#pragma unroll
        for (int c = 0; c < 8; ++c) arow[c] = L[(jb + tid) * LDQ + jb + c];
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) arow[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double dd = Dg[c][c];
        if (!(dd > 0.0)) { if (tid == 0) bad = 1; dd = 1.0; }
        const double inv = rsqrt(dd);      // 1 ulp; the sqrt + divide pair would put ~600 cycles per column on the critical path
        Dg[c][c] = dd * inv; dinv8[c] = inv;
        arow[c] *= inv;
#pragma unroll
        for (int a = 0; a < 8; ++a) if (a > c) Dg[a][c] *= inv;
#pragma unroll
        for (int b = 0; b < 8; ++b) if (b > c) {
          arow[b] -= arow[c] * Dg[b][c];
#pragma unroll
          for (int a = 0; a < 8; ++a) if (a >= b) Dg[a][b] -= Dg[a][c] * Dg[b][c];
        }
      }
      if (row_thread && tid >= 8) {        // rows below the diagonal block: x = arow * Dg^-T, computed above
#pragma unroll
        for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(&L[(jb + tid) * LDQ + jb + c]) = make_double2(arow[c], arow[c + 1]);
      }
      if (inv_thread) {                    // W = Dg^-1 (lower triangular) -> Dinv[I] and the diagonal block of X
        double Xi[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int a = 0; a < 8; ++a) Xi[a][c] = 0.0;
          Xi[c][c] = dinv8[c];
#pragma unroll
          for (int a = 0; a < 8; ++a) if (a > c) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k >= c && k < a) sacc += Dg[a][k] * Xi[k][c];
            Xi[a][c] = -sacc * dinv8[a];
          }
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            *reinterpret_cast<double2*>(&Dinv[I * 64 + a * 8 + c]) = make_double2(Xi[a][c], Xi[a][c + 1]);
            *reinterpret_cast<double2*>(&X[(jb + a) * LDQ + jb + c]) = make_double2(Xi[a][c], Xi[a][c + 1]);
          }
      }
    }
    {
      // X[I-1, 8J..8J+7] = -W_{I-1} * S[:, 8J..8J+7] on the warps that hold no row thread (warp 7 is the inverse warp)
      const int first_free = (nrow + 31) >> 5, nfw = 7 - first_free, P = I - 1;
      if (I >= 1 && wid >= first_free && wid < 7) {
        const double* Wp = Dinv + P * 64;
        for (int J = wid - first_free; J < P; J += nfw) {
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int sk = 0; sk < 2; ++sk) dmma884(d0, d1, Wp[g * 8 + 4 * sk + t], Sb[(4 * sk + t) * LDQ + 8 * J + g]);
          *reinterpret_cast<double2*>(&X[(8 * P + g) * LDQ + 8 * J + 2 * t]) = make_double2(-d0, -d1);
        }
      }
    }
    __syncthreads();
    DIAG_STAMP(2 + 2 * I);
    // ---------------- phase B
    if (tid < 8) {                         // factored diagonal block (nobody reads these rows during phase B)
#pragma unroll
      for (int a = 0; a < 8; ++a) if (a == tid) {
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c <= a) L[(jb + a) * LDQ + jb + c] = Dg[a][c];
      }
    }
    __syncwarp();
    {
      // trailing update, 8x8 blocks (bi >= bk) of the rows below the block column
      const int t0 = jb + 8, m = (PW - t0) / 8, nblk = m * (m + 1) / 2;
      for (int b = wid; b < nblk; b += 24) {   // three blocks per round so their loads and DMMAs overlap
        double2* cp[3]; double2 c[3]; const double* ap[3]; const double* bp[3]; bool on[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int bb = b + 8 * u;
          on[u] = bb < nblk;
          const int bq = on[u] ? bb : b;
          const int bi = blk_i[bq], bk = blk_k[bq];
          const int i0 = t0 + 8 * bi, k0 = t0 + 8 * bk;
          cp[u] = reinterpret_cast<double2*>(&L[(i0 + g) * LDQ + k0 + 2 * t]);
          ap[u] = &L[(i0 + g) * LDQ + jb + t]; bp[u] = &L[(k0 + g) * LDQ + jb + t];
          c[u] = *cp[u];
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) { dmma884(c[u].x, c[u].y, -ap[u][0], bp[u][0]); dmma884(c[u].x, c[u].y, -ap[u][4], bp[u][4]); }
#pragma unroll
        for (int u = 0; u < 3; ++u) if (on[u]) *cp[u] = c[u];   // diagonal blocks also get their (unused) upper half
      }
      // S[:, 8J..8J+7] = sum_{k = 8J}^{jb-1} L[jb+., k] X[k, 8J+.]; two interleaved accumulators per block
      for (int J = 7 - wid; J < I; J += 8) {
        double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
        const double* ap = &L[(jb + g) * LDQ + t];
        const double* bp = &X[t * LDQ + 8 * J + g];
        int k = 8 * J;
        for (; k + 8 <= jb; k += 8) {
          dmma884(e0, e1, ap[k], bp[k * LDQ]);
          dmma884(f0, f1, ap[k + 4], bp[(k + 4) * LDQ]);
        }
        *reinterpret_cast<double2*>(&Sb[g * LDQ + 8 * J + 2 * t]) = make_double2(e0 + f0, e1 + f1);
      }
    }
    __syncthreads();
    DIAG_STAMP(3 + 2 * I);
  }
  {                                        // last row-block: X[NB-1, 0:8(NB-1)] = -W_{NB-1} * S
    const double* Wp = Dinv + (NB - 1) * 64;
    for (int J = wid; J < NB - 1; J += 8) {
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int sk = 0; sk < 2; ++sk) dmma884(d0, d1, Wp[g * 8 + 4 * sk + t], Sb[(4 * sk + t) * LDQ + 8 * J + g]);
      *reinterpret_cast<double2*>(&X[(8 * (NB - 1) + g) * LDQ + 8 * J + 2 * t]) = make_double2(-d0, -d1);
    }
  }
  __syncthreads();
  DIAG_STAMP(2 + 2 * NB);
  // Only Linv goes back to memory.  L_dd itself has no reader: trsm and the backward sweep multiply by Linv, and the rows
  // below the panel hold X = A Linv^T, which IS their part of L.
  double* dout = dinv + (size_t)d * PW * PW;
#pragma unroll 4
  for (int i = wid; i < PW; i += 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * (lane + 32 * h);                        // j even: (j, j+1) both in the lower triangle unless j == i
      if (j >= PW) break;
      double2 x = make_double2(0.0, 0.0);
      if (i < w && j <= i) { x = *reinterpret_cast<const double2*>(&X[i * LDQ + j]); if (j == i) x.y = 0.0; }
      *reinterpret_cast<double2*>(dout + i * PW + j) = x;
    }
  }
  if (tid == 0 && bad) *fail = 1;
  DIAG_STAMP(3 + 2 * NB);
}

// diag, pipelined (the default): the same factorisation with the two barrier phases of a step overlapped.  512 threads:
//   panel warps 0-3   step I: wait until block column I has its last update; every row thread re-factors the 8x8
//                     diagonal block in registers and scales its own row; the block's own eight rows compute the rows
//                     of W = (8x8 factor)^-1 instead -> Dinv[I] and the diagonal block of X.  Then on to step I+1 as soon
//                     as the update warps have done the ONE block column it needs.
//   update warps 4-15 step I: wait for block column I; first the "urgent" blocks — block column I+1 of the trailing
//                     matrix, what the panel warps are waiting for — then, under the panel warps' next factorisation,
//                     X row-block I-1 = -W_{I-1} S, S = L[I, 0:8I] X[0:8I, 0:8I] for row-block I, and the rest of the
//                     rank-8 trailing update.  A global 8x8 block always belongs to the same warp, so the updates of
//                     one block stay in program order from step to step.
// Named barriers: 1 = "block column I is scaled" (panel arrives, update warps wait), 2 = "block column I+1 is updated"
// (update warps arrive, panel waits), 3 = among the update warps (S complete / X row-block complete).
constexpr int DG2_THREADS = 512, DG2_UW = 12;
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__global__ void __launch_bounds__(DG2_THREADS) sky_diag2_kernel(int d, int n, int prev, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                                 double* __restrict__ val, double* __restrict__ dinv, int* __restrict__ fail) {
  constexpr int NB = PW / 8, LDQ = LDT, NW = DG2_THREADS / 32;
  DIAG_STAMP(0);
  TL_IN(val, 0, d);
  extern __shared__ __align__(16) double sm_diag[];
  double* L = sm_diag;                 // [PW][LDQ]
  double* X = sm_diag + PW * LDQ;      // [PW][LDQ]  only the lower block triangle is ever written or read
  double* Dinv = X + PW * LDQ;         // [NB][64] inverses of the 8x8 diagonal factors (full 8x8, upper part zero)
  double* Sb = Dinv + NB * 64;         // [8][LDQ]  S of the current row-block
  __shared__ long long rbase[PW];
  __shared__ int bad;
  __shared__ int rprev[PW];            // 1 = the row's envelope reaches panel d-1
  const int c0 = d * PW, w = min(PW, n - c0), tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) bad = 0;
  if (tid < PW) {
    const int r = c0 + tid;
    const int st = tid < w ? start[r] : 0;
    rbase[tid] = tid < w ? ptr[r] + (c0 - st) : 0;
    rprev[tid] = (prev && tid < w && st <= c0 - PW) ? 1 : 0;
  }
  __syncthreads();
#pragma unroll 2
  for (int i = wid; i < PW; i += NW) {
    const double* src = val + rbase[i];
    const bool row_ok = i < w;
    {
      const bool ok = row_ok && 2 * lane <= i;               // lower triangle by pairs, the rest zero-filled
      cp_async16(&L[i * LDQ + 2 * lane], ok ? (const void*)(src + 2 * lane) : (const void*)val, ok);
    }
    if (lane < PW / 2 - 32) {
      const int j2 = lane + 32;
      const bool ok = row_ok && 2 * j2 <= i;
      cp_async16(&L[i * LDQ + 2 * j2], ok ? (const void*)(src + 2 * j2) : (const void*)val, ok);
    }
  }
  if (prev) {
#pragma unroll 2
    for (int i = wid; i < PW; i += NW) {
      const double* src = val + rbase[i];
      const bool ok = rprev[i] != 0;
      cp_async16(&X[i * LDQ + 2 * lane], ok ? (const void*)(src - PW + 2 * lane) : (const void*)val, ok);
      if (lane < PW / 2 - 32) cp_async16(&X[i * LDQ + 2 * (lane + 32)], ok ? (const void*)(src - PW + 2 * (lane + 32)) : (const void*)val, ok);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (prev) {
    // A_dd -= X X^T on the lower block triangle: 78 blocks of 8x8, two per warp in flight, K = 96
    constexpr int NBLK = NB * (NB + 1) / 2;
    for (int b = 2 * wid; b < NBLK; b += 2 * NW) {
      const bool two = b + 1 < NBLK;
      const int bsec = two ? b + 1 : b;
      int bi0, bk0, bi1, bk1; diag_block_of(b, bi0, bk0); diag_block_of(bsec, bi1, bk1);
      double2* cp0 = reinterpret_cast<double2*>(&L[(8 * bi0 + g) * LDQ + 8 * bk0 + 2 * t]);
      double2* cp1 = reinterpret_cast<double2*>(&L[(8 * bi1 + g) * LDQ + 8 * bk1 + 2 * t]);
      double2 u = *cp0, v = *cp1;
      const double* a0 = &X[(8 * bi0 + g) * LDQ + t]; const double* b0 = &X[(8 * bk0 + g) * LDQ + t];
      const double* a1 = &X[(8 * bi1 + g) * LDQ + t]; const double* b1 = &X[(8 * bk1 + g) * LDQ + t];
#pragma unroll 6
      for (int k = 0; k < PW; k += 4) {
        dmma884(u.x, u.y, -a0[k], b0[k]);
        dmma884(v.x, v.y, -a1[k], b1[k]);
      }
      *cp0 = u;
      if (two) *cp1 = v;
    }
    __syncthreads();
  }
  if (tid >= w && tid < PW) L[tid * LDQ + tid] = 1.0;         // identity padding of a short last panel
  __syncthreads();
  DIAG_STAMP(1);
  if (wid < 4) {
    // ------------------------------------------------------------------ panel warps
    double Dg[8][8];
    for (int I = 0; I < NB; ++I) {
      const int jb = 8 * I, nrow = PW - jb;
      if (I > 0) bar_sync_n(2, DG2_THREADS);               // block column I carries every update of the steps before
      if (tid < nrow) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) Dg[a][c] = (c <= a) ? L[(jb + a) * LDQ + jb + c] : 0.0;
        double arow[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) arow[c] = L[(jb + tid) * LDQ + jb + c];
        double dinv8[8];
        // right-looking, in registers: scale column c, push its rank-1 update to the right; the thread's own row rides along
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double dd = Dg[c][c];
          if (!(dd > 0.0)) { if (tid == 0) bad = 1; dd = 1.0; }
          const double inv = rsqrt(dd);
          Dg[c][c] = dd * inv; dinv8[c] = inv;
          arow[c] *= inv;
#pragma unroll
          for (int a = 0; a < 8; ++a) if (a > c) Dg[a][c] *= inv;
#pragma unroll
          for (int b = 0; b < 8; ++b) if (b > c) {
            arow[b] -= arow[c] * Dg[b][c];
#pragma unroll
            for (int a = 0; a < 8; ++a) if (a >= b) Dg[a][b] -= Dg[a][c] * Dg[b][c];
          }
        }
        if (tid >= 8) {                      // rows below the diagonal block: x = arow * Dg^-T
#pragma unroll
          for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(&L[(jb + tid) * LDQ + jb + c]) = make_double2(arow[c], arow[c + 1]);
        } else {
          // row `tid` of W = Dg^-1: w Dg = e_tid, solved from the right (entries left of the diagonal only)
          double wr[8];
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            double sacc = (j == tid) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k > j) sacc -= wr[k] * Dg[k][j];
            wr[j] = (j <= tid) ? sacc * dinv8[j] : 0.0;
          }
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            *reinterpret_cast<double2*>(&Dinv[I * 64 + tid * 8 + c]) = make_double2(wr[c], wr[c + 1]);
            *reinterpret_cast<double2*>(&X[(jb + tid) * LDQ + jb + c]) = make_double2(wr[c], wr[c + 1]);
          }
        }
      }
      __threadfence_block();
      bar_arrive_n(1, DG2_THREADS);                          // block column I is scaled, W_I is there
      DIAG_STAMP(2 + 2 * I);
    }
  } else {
    // ------------------------------------------------------------------ update warps
    const int uw = wid - 4;
    auto owner_row = [&](int Cb) { const int r = uw - Cb; return r < 0 ? r + DG2_UW : r; };   // block (Rb, Cb) belongs to warp (Rb + Cb) mod 12: one block row per block column and warp, no division
    auto update_block = [&](int Rb, int Cb, int jb) {
      double2* cp = reinterpret_cast<double2*>(&L[(8 * Rb + g) * LDQ + 8 * Cb + 2 * t]);
      double2 c = *cp;
      const double* ap = &L[(8 * Rb + g) * LDQ + jb + t]; const double* bp = &L[(8 * Cb + g) * LDQ + jb + t];
      dmma884(c.x, c.y, -ap[0], bp[0]); dmma884(c.x, c.y, -ap[4], bp[4]);
      *cp = c;                                               // diagonal blocks also get their (unused) upper half
    };
    for (int I = 0; I < NB; ++I) {
      const int jb = 8 * I;
      bar_sync_n(1, DG2_THREADS);                            // block column I is scaled
      if (I + 1 < NB) {
        const int Rb = owner_row(I + 1);                     // urgent: block column I+1, one block per warp at most
        if (Rb >= I + 1 && Rb < NB) update_block(Rb, I + 1, jb);
        __threadfence_block();
        bar_arrive_n(2, DG2_THREADS);
      }
      // X row-block I-1 = -W_{I-1} * S  (S of row-block I-1 was completed by all update warps in the step before)
      if (I >= 1) {
        bar_sync_n(3, DG2_UW * 32);
        const int P = I - 1;
        if (uw < P) {
          const int J = uw;
          const double* Wp = Dinv + P * 64;
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int sk = 0; sk < 2; ++sk) dmma884(d0, d1, Wp[g * 8 + 4 * sk + t], Sb[(4 * sk + t) * LDQ + 8 * J + g]);
          *reinterpret_cast<double2*>(&X[(8 * P + g) * LDQ + 8 * J + 2 * t]) = make_double2(-d0, -d1);
        }
        bar_sync_n(3, DG2_UW * 32);                          // X row-block I-1 complete; S may be overwritten
      }
      // S[:, 8J..8J+7] = sum_{k = 8J}^{jb-1} L[jb+., k] X[k, 8J+.] for row-block I; two interleaved accumulators
      if (uw < I) {
        const int J = uw;
        double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
        const double* ap = &L[(jb + g) * LDQ + t];
        const double* bp = &X[t * LDQ + 8 * J + g];
        for (int k = 8 * J; k + 8 <= jb; k += 8) {
          dmma884(e0, e1, ap[k], bp[k * LDQ]);
          dmma884(f0, f1, ap[k + 4], bp[(k + 4) * LDQ]);
        }
        *reinterpret_cast<double2*>(&Sb[g * LDQ + 8 * J + 2 * t]) = make_double2(e0 + f0, e1 + f1);
      }
      // the rest of the trailing update: block columns I+2 .. (one block per column at most; four columns at a time with
      // predicated dummies was measured slower: 30.1 vs 25.7 us, tools/diag_lab.cu)
      for (int Cb = I + 2; Cb < NB; ++Cb) {
        const int Rb = owner_row(Cb);
        if (Rb >= Cb && Rb < NB) update_block(Rb, Cb, jb);
      }
      DIAG_STAMP_U(3 + 2 * I);
    }
    // last row-block: X[NB-1, 0:8(NB-1)] = -W_{NB-1} * S
    bar_sync_n(3, DG2_UW * 32);
    if (uw < NB - 1) {
      const int J = uw;
      const double* Wp = Dinv + (NB - 1) * 64;
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int sk = 0; sk < 2; ++sk) dmma884(d0, d1, Wp[g * 8 + 4 * sk + t], Sb[(4 * sk + t) * LDQ + 8 * J + g]);
      *reinterpret_cast<double2*>(&X[(8 * (NB - 1) + g) * LDQ + 8 * J + 2 * t]) = make_double2(-d0, -d1);
    }
  }
  __syncthreads();
  DIAG_STAMP(2 + 2 * NB);
  // Only Linv goes back to memory (L_dd itself has no reader), and only its lower triangle: the array is zeroed when the
  // factor is created and nothing else is ever written above the diagonal or into the padding rows of a short last panel.
  double* dout = dinv + (size_t)d * PW * PW;
#pragma unroll 2
  for (int i = wid; i < w; i += NW) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * (lane + 32 * h);                        // j even: (j, j+1) both in the lower triangle unless j == i
      if (j > i) break;
      double2 x = *reinterpret_cast<const double2*>(&X[i * LDQ + j]);
      if (j == i) x.y = 0.0;
      *reinterpret_cast<double2*>(dout + i * PW + j) = x;
    }
  }
  if (tid == 0 && bad) *fail = 1;
  DIAG_STAMP(3 + 2 * NB);
  TL_OUT(val, 0, d);
}

// trsm: X[r][j] = sum_k A[r][c0+k] * Linv[j][k] for the rows r in R_d, in place.  32 rows x 96 columns per CTA;
// warp (wm, wn) of the 2 x 4 warp grid owns rows 16 wm.., columns 24 wn.. as 2 x 3 DMMA tiles.  Linv is lower
// triangular, so output columns j0..j0+7 only need k <= j0+7.
__global__ void __launch_bounds__(256, 2) sky_trsm_kernel(int d, int n, const long long* __restrict__ ptr, const int* __restrict__ start,
                                                          const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx,
                                                          const double* __restrict__ dinv, double* __restrict__ val,
                                                          double* __restrict__ xp, long long* __restrict__ rinfo, int rpad) {
  extern __shared__ __align__(16) double sm[];
  double* Li = sm;                    // [PW][LDT]  Linv
  double* A = sm + PW * LDT;          // [TR][LDT]
  __shared__ long long rbase[TR];
  const int c0 = d * PW, tid = threadIdx.x;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  const int row0 = blockIdx.x * TR;
  TL_IN(val, 1, d);
  if (tid < TR) {
    long long b = -1; int r = -1;
    if (row0 + tid < nr) { r = rows_idx[rb + row0 + tid]; b = ptr[r] + (c0 - start[r]); }
    rbase[tid] = b;
    // row table of the packed copy: (row index or -1, offset of (row, column 0)); CTAs past the end of the list only pad
    rinfo[2 * (size_t)(row0 + tid)] = r; rinfo[2 * (size_t)(row0 + tid) + 1] = b >= 0 ? b - c0 : 0;
  }
  __syncthreads();
  if (row0 >= nr) {   // padding rows of the packed copy (the list is padded to whole 128-row tiles): zeros
    for (int e = tid; e < TR * (PW / 2); e += blockDim.x) {
      const int i = e / (PW / 2), c = 2 * (e % (PW / 2));
      *reinterpret_cast<double2*>(xp + ((size_t)(c / KC) * rpad + row0 + i) * LDK + c % KC) = make_double2(0.0, 0.0);
    }
    return;
  }
  const double* dsrc = dinv + (size_t)d * PW * PW;
  for (int e = tid; e < TR * (PW / 2); e += blockDim.x) {
    const int i = e / (PW / 2), j2 = e % (PW / 2);
    const bool ok = rbase[i] >= 0;
    cp_async16(&A[i * LDT + 2 * j2], ok ? (const void*)(val + rbase[i] + 2 * j2) : (const void*)val, ok);
  }
  for (int e = tid; e < PW * (PW / 2); e += blockDim.x) {
    const int i = e / (PW / 2), j2 = e % (PW / 2);
    cp_async16(&Li[i * LDT + 2 * j2], dsrc + i * PW + 2 * j2, 2 * j2 <= i);   // above the diagonal: zero-filled, nothing read
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = wid & 1, wn = wid >> 1;
  double acc[2][3][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const double* a_s = A + (16 * wm + g) * LDT + t;
  const double* b_s = Li + (24 * wn + g) * LDT + t;
  const int kmax = 24 * wn + 24;      // columns of this warp need k < kmax
  for (int k = 0; k < kmax; k += 4) {
    double a[2], b[3];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = a_s[(8 * i) * LDT + k];
#pragma unroll
    for (int j = 0; j < 3; ++j) b[j] = b_s[(8 * j) * LDT + k];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rl = 16 * wm + 8 * i + g;
    const long long b = rbase[rl];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = 24 * wn + 8 * j + 2 * t;
      const double2 v = b >= 0 ? make_double2(acc[i][j][0], acc[i][j][1]) : make_double2(0.0, 0.0);   // rows past the end of the list: A was zero-filled
      if (b >= 0) *reinterpret_cast<double2*>(val + b + c) = v;
      *reinterpret_cast<double2*>(xp + ((size_t)(c / KC) * rpad + row0 + rl) * LDK + c % KC) = v;
    }
  }
  TL_OUT(val, 1, d);
}

// update: A[r][c] -= X[r,:] . X[c,:] over the tiles (ti, tj) of R_d x R_d that intersect the lower triangle, except
// the elements with r < skip_below (the diagonal block of panel d+1, which C(d+1) updates itself).
// Row tile ti owns the column tiles tj = 0 .. 2 ti + 1 (UM == 2 UN).  Two launches per panel:
//   part 0 ("next")  the column tiles tj < 2 — they hold every column of panel d+1, which trsm(d+1) is waiting for;
//                    tile index T = 2 ti + tj;
//   part 1 ("rest")  tj >= 2, T = ti (ti - 1) + (tj - 2), on the low-priority stream.
// One 128x64 tile per 256-thread CTA (108 KB of shared memory, two CTAs per SM): a tile is ~5k cycles of loads
// (L2 -> SM bandwidth bound) followed by ~12k cycles of DMMA that one CTA can keep saturated, and two CTAs scheduled
// independently drift out of phase so that one loads while the other multiplies.  (512-thread CTAs of two tiles in
// lock-step were measured 20 % slower, tools/upd_lab.cu.)
// 8 warps as 4 x 2, every warp a grid of 4 x 4 8x8 DMMA tiles, K = 96 in three cp.async chunks.
// predicated 128/64-bit global accesses: no branch, so a thread's sixteen tile loads are all in flight at once
__device__ __forceinline__ void ldg128_if(double& x, double& y, const double* p, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q ld.global.v2.f64 {%0,%1}, [%2];\n}" : "+d"(x), "+d"(y) : "l"(p), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void ldg64_if(double& x, const double* p, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.f64 %0, [%1];\n}" : "+d"(x) : "l"(p), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void stg128_if(double* p, double x, double y, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q st.global.v2.f64 [%2], {%0,%1};\n}" ::"d"(x), "d"(y), "l"(p), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void stg64_if(double* p, double x, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q st.global.f64 [%1], %0;\n}" ::"d"(x), "l"(p), "r"((int)pred) : "memory");
}
#ifdef SKY_UPD_CLOCKS   // tools/upd_lab.cu: cycle stamps of group 0 / thread 0 of the first CTAs of panel SKY_UPD_CLOCKS
__device__ long long g_upd_clk[2][64][10];
#define UPD_STAMP(i) do { if (d == SKY_UPD_CLOCKS && threadIdx.x == 0 && blockIdx.x < 64) g_upd_clk[PART][blockIdx.x][i] = clock64(); } while (0)
#else
#define UPD_STAMP(i) do { } while (0)
#endif
template <int PART>
__global__ void __launch_bounds__(256, 2) sky_update_kernel(int d, int n, int skip_below, int Tr, int Tc,
                                                            const long long* __restrict__ ptr, const int* __restrict__ start,
                                                            const int* __restrict__ rows_ptr, const int* __restrict__ rows_idx, double* __restrict__ val) {
  constexpr int WARPS_M = 4, WARPS_N = 2;
  constexpr int WTM = UM / WARPS_M, WTN = UN / WARPS_N, FM = WTM / 8, FN = WTN / 8;
  static_assert(UM == 2 * UN, "triangular tile enumeration assumes UM == 2 UN");
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  double* As = sm;                                 // [2][UM][LDK]
  double* Bs = As + 2 * UM * LDK;                  // [2][UN][LDK]
  __shared__ long long s_abase[UM], s_bbase[UN];   // element offset of (row, c0) in val, -1 = no such row
  __shared__ int s_arow[UM], s_bcol[UN];
  const int c0 = d * PW;
  const int rb = rows_ptr[d], nr = rows_ptr[d + 1] - rb;
  const int lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = wid % WARPS_M, wn = wid / WARPS_M;
  constexpr int NCH = PW / KC;
  {
    const int T = blockIdx.x;
    int ti, tj;
    if (PART == 0) { ti = T >> 1; tj = T & 1; }
    else {
      ti = (int)((1.0f + sqrtf(1.0f + 4.0f * (float)T)) * 0.5f);
      while (ti * (ti - 1) > T) --ti;
      while ((ti + 1) * ti <= T) ++ti;
      tj = 2 + (T - ti * (ti - 1));
    }
    if (ti >= Tr || tj >= Tc) return;        // the last row tile can run past the last column tile
    UPD_STAMP(0);
    if (tid < UM) {
      const int ir = ti * UM + tid;
      const int r = ir < nr ? rows_idx[rb + ir] : -1;
      s_arow[tid] = r; s_abase[tid] = r >= 0 ? ptr[r] + (c0 - start[r]) : -1;
    } else if (tid < UM + UN) {
      const int q = tid - UM, ic = tj * UN + q;
      int c = ic < nr ? rows_idx[rb + ic] : -1;
      if (c >= n) c = -1;                    // the rhs row is never a column
      s_bcol[q] = c; s_bbase[q] = c >= 0 ? ptr[c] + (c0 - start[c]) : -1;
    }
    __syncthreads();
    UPD_STAMP(1);
    auto issue = [&](int chunk, int stage) {
      const int k0 = chunk * KC;
      double* a_dst = As + stage * UM * LDK; double* b_dst = Bs + stage * UN * LDK;
#pragma unroll
      for (int e = tid; e < (UM + UN) * (KC / 2); e += 256) {
        const int row = e / (KC / 2), seg = e % (KC / 2);
        if (row < UM) { const long long bo = s_abase[row]; cp_async16(&a_dst[row * LDK + 2 * seg], bo >= 0 ? (const void*)(val + bo + k0 + 2 * seg) : (const void*)val, bo >= 0); }
        else { const int rr = row - UM; const long long bo = s_bbase[rr]; cp_async16(&b_dst[rr * LDK + 2 * seg], bo >= 0 ? (const void*)(val + bo + k0 + 2 * seg) : (const void*)val, bo >= 0); }
      }
      cp_async_commit();
    };
    issue(0, 0);
    issue(1, 1);
    // The accumulators start at A_old and the A fragments are negated, so the tile leaves the loop as
    // A_old - X X^T and the epilogue is a plain store: the read half of the read-modify-write is in flight together
    // with the first operand chunk (loaded straight into the accumulator registers) instead of stalling the
    // epilogue.  A lane owns two adjacent list positions (even, odd) of a row; list positions of a node's six
    // scalars are consecutive and start even, so the pair is always (c, c + 1), 16-byte aligned; on the diagonal
    // (c == r) only the first of the two exists.
    double acc[FM][FN][2];
    int cidx[FN], ridx[FM];
    long long rowoff[FM];
#pragma unroll
    for (int j = 0; j < FN; ++j) cidx[j] = s_bcol[wn * WTN + 8 * j + 2 * t];
#pragma unroll
    for (int i = 0; i < FM; ++i) {
      const int rl = wm * WTM + 8 * i + g;
      const int r = s_arow[rl];
      ridx[i] = r >= skip_below ? r : -1;                      // -1: no such row, or a row C(d+1) owns
      rowoff[i] = r >= 0 ? s_abase[rl] - c0 : 0;              // offset of (row, column 0)
    }
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        const int c = cidx[j], r = ridx[i];
        // c <= r covers the pair (c, c + 1) below the diagonal and the diagonal element c == r; there the second slot is
        // (r, r + 1), which exists in the row's storage (c is even, a node's six scalars never straddle a panel) and is
        // never stored back.  ONE load per fragment: a second, differently predicated load into the same registers
        // would have to wait for the first (scoreboard), serialising sixteen round trips to L2 per tile.
        const bool any = c >= 0 && c <= r;
        const double* p = val + (any ? rowoff[i] + c : 0);
        acc[i][j][0] = 0.0; acc[i][j][1] = 0.0;
        ldg128_if(acc[i][j][0], acc[i][j][1], p, any);
      }
    UPD_STAMP(2);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (ch + 1 < NCH) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      UPD_STAMP(3 + 2 * ch);
      const double* a_s = As + (ch & 1) * UM * LDK + (wm * WTM + g) * LDK + t;
      const double* b_s = Bs + (ch & 1) * UN * LDK + (wn * WTN + g) * LDK + t;
#pragma unroll
      for (int k = 0; k < KC; k += 4) {
        double a[FM], bf[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) a[i] = -a_s[(8 * i) * LDK + k];
#pragma unroll
        for (int j = 0; j < FN; ++j) bf[j] = b_s[(8 * j) * LDK + k];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
      }
      UPD_STAMP(4 + 2 * ch);
      if (ch + 2 < NCH) { __syncthreads(); issue(ch + 2, ch & 1); }
    }
    // epilogue: indices re-read from shared memory so that nothing but the accumulators lives across the MMA loop
#pragma unroll
    for (int i = 0; i < FM; ++i) {
      const int rl = wm * WTM + 8 * i + g;
      const int r0 = s_arow[rl];
      const int r = r0 >= skip_below ? r0 : -1;
      const long long ro = r0 >= 0 ? s_abase[rl] - c0 : 0;
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        const int c = s_bcol[wn * WTN + 8 * j + 2 * t];
        const bool pair = c >= 0 && c + 1 <= r, diag = c >= 0 && c == r;
        double* p = val + ((pair || diag) ? ro + c : 0);
        stg128_if(p, acc[i][j][0], acc[i][j][1], pair);
        stg64_if(p, acc[i][j][0], diag);
      }
    }
    UPD_STAMP(9);
  }
}

// ------------------------------------------------------------------------------------------------ update, persistent pipeline
// The same tiles as sky_update_kernel, as a persistent software pipeline (one 512-thread CTA per SM, 166 KB of shared
// memory): a CTA walks its tiles T = blockIdx.x, blockIdx.x + gridDim.x, ... and treats their K chunks as ONE stream
// through a 3-stage ring.  The operands come from the packed copy of X that trsm leaves behind ([K chunk][list position]
// [LDK], shared-memory padding included): a tile's chunk is two contiguous blocks, 128 and 64 list positions, fetched
// by TWO bulk asynchronous copies (cp.async.bulk, 36 KB and 18 KB, completion counted in bytes on the stage's
// mbarrier).  As soon as every warp is done with chunk ch of tile t, chunk ch of tile t+1 is requested into the same
// stage, so the operands of the next tile arrive while this one is being multiplied.  (One bulk copy per ROW and chunk
// straight from the envelope — 576 copies of 256 B per tile — kept the tile at twice its multiply time: measured
// 12.3 us against the 6.3 us the same loop takes from resident shared memory, tools/mma_lab.cu.)
// A_old is not preloaded into the accumulators: its loads go into registers of their own, one fragment per k-step of the
// middle chunk, land during the multiply, and the epilogue stores A_old - X X^T.  Sixteen warps, each a 32 x 16 piece
// of the 128 x 64 tile, synchronised only through the ring (no block-wide barrier inside the loop).
constexpr int WS_THREADS = 512;
constexpr int WS_NS = PW / KC;                 // ring stages == K chunks of a tile
static_assert(WS_NS == 3, "stage index == chunk index");
constexpr int WS_STAGE = (UM + UN) * LDK;      // doubles per stage

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void upd_tile_of(int part, int T, int& ti, int& tj) {
  if (part == 0) { ti = T >> 1; tj = T & 1; return; }
  ti = (int)((1.0f + sqrtf(1.0f + 4.0f * (float)T)) * 0.5f);
  while (ti * (ti - 1) > T) --ti;
  while ((ti + 1) * ti <= T) ++ti;
  tj = 2 + (T - ti * (ti - 1));
}

#ifdef SKY_WS_CLOCKS   // tools/ws_lab.cu: cycle stamps of thread 0 of the first CTAs of panel SKY_WS_CLOCKS (rest kernel)
__device__ long long g_ws_clk[64][48];
#define WS_STAMP(i) do { if (PART == 1 && d == SKY_WS_CLOCKS && threadIdx.x == 0 && blockIdx.x < 64 && (i) < 48) g_ws_clk[blockIdx.x][i] = clock64(); } while (0)
#else
#define WS_STAMP(i) do { } while (0)
#endif
template <int PART>
__global__ void __launch_bounds__(WS_THREADS, 1) sky_update_ws_kernel(int d, int n, int skip_below, int Tr, int Tc, int n_tiles, int rpad,
                                                                      const double* __restrict__ xp, const long long* __restrict__ rinfo,
                                                                      double* __restrict__ val) {
  constexpr int WARPS_M = 4, WARPS_N = 4, NWARPS = WARPS_M * WARPS_N;
  constexpr int WTM = UM / WARPS_M, WTN = UN / WARPS_N, FM = WTM / 8, FN = WTN / 8;
  static_assert(FM * FN == KC / 4, "one A_old fragment load per k-step of a chunk");
  constexpr unsigned A_BYTES = UM * LDK * sizeof(double), B_BYTES = UN * LDK * sizeof(double), TAB_BYTES = 2 * (UM + UN) * sizeof(long long);
  extern __shared__ __align__(128) unsigned char ws_smem[];
  double* ring = reinterpret_cast<double*>(ws_smem);                                  // [WS_NS][UM + UN][LDK]
  long long* tab = reinterpret_cast<long long*>(ring + WS_NS * WS_STAGE);             // [UM + UN][2] row table of the current tile
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tab + 2 * (UM + UN));   // [WS_NS]
  int* done = reinterpret_cast<int*>(full + WS_NS);                                   // [WS_NS] warps finished with the stage
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = wid % WARPS_M, wn = wid / WARPS_M;
  WS_STAMP(0);
  TL_IN(val, 2 + PART, d);
  if (tid == 0) {
    for (int s = 0; s < WS_NS; ++s) { mbar_init(full + s, 1); done[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // chunk ch of tile (ti, tj) -> stage ch: two contiguous blocks of the packed copy; with chunk 0 also the tile's row table
  // (row index, offset) for its 128 + 64 list positions.  Called by ONE lane.  No proxy fence: the warps' reads of the
  // stage have returned their data (the multiplies that consumed it have been issued) before the counter that elects
  // this lane is incremented.
  auto issue = [&](int ti, int tj, int ch) {
    if (tj >= Tc) tj = Tc - 1;                         // the last row tile can run past the last column tile: nothing of it is stored
    mbar_expect_tx(full + ch, A_BYTES + B_BYTES + (ch == 0 ? TAB_BYTES : 0u));
    bulk_g2s(ring + ch * WS_STAGE, xp + ((size_t)ch * rpad + (size_t)ti * UM) * LDK, A_BYTES, full + ch);
    bulk_g2s(ring + ch * WS_STAGE + UM * LDK, xp + ((size_t)ch * rpad + (size_t)tj * UN) * LDK, B_BYTES, full + ch);
    if (ch == 0) {
      bulk_g2s(tab, rinfo + 2 * (size_t)ti * UM, 2 * UM * sizeof(long long), full + ch);
      bulk_g2s(tab + 2 * UM, rinfo + 2 * (size_t)tj * UN, 2 * UN * sizeof(long long), full + ch);
    }
  };
  int T = blockIdx.x;
  if (T >= n_tiles) return;
  int ti, tj; upd_tile_of(PART, T, ti, tj);
  if (tid == 0) {
#pragma unroll
    for (int ch = 0; ch < WS_NS; ++ch) issue(ti, tj, ch);
  }
  // No block-wide barrier below: a warp that is done with a stage says so on a counter and moves on; the warp that
  // arrives last requests the stage's next contents.  Warps drift apart, so that the global loads of A_old and the
  // stores of one warp run under the multiplies of the others instead of stopping all sixteen at once (measured in
  // lock-step: ~4k cycles of loads + ~3k of stores around 12k cycles of DMMA per tile, tools/ws_lab.cu).
  for (int it = 0; T < n_tiles; T += gridDim.x, ++it) {
    const unsigned par = (unsigned)(it & 1);
    const int Tn = T + gridDim.x;
    const bool more = Tn < n_tiles;
    const bool live = tj < Tc;
    int tin = 0, tjn = 0;
    if (more) upd_tile_of(PART, Tn, tin, tjn);
    int cidx[FN], ridx[FM];
    long long rowoff[FM];
    double cold[FM][FN][2], acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) { cold[i][j][0] = 0.0; cold[i][j][1] = 0.0; acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    WS_STAMP(1 + 12 * it);
#pragma unroll
    for (int ch = 0; ch < WS_NS; ++ch) {
      mbar_wait(full + ch, par);
      if (ch == 0) {
        // rows / columns of this thread's fragments from the tile's row table (it came with chunk 0; the next tile's
        // table can only replace it after every warp has passed this point and finished the chunk)
#pragma unroll
        for (int j = 0; j < FN; ++j) {
          const long long c = live ? tab[2 * (UM + wn * WTN + 8 * j + 2 * t)] : -1;
          cidx[j] = c >= n ? -1 : (int)c;               // the rhs row is never a column
        }
#pragma unroll
        for (int i = 0; i < FM; ++i) {
          const int rl = wm * WTM + 8 * i + g;
          const int r = (int)tab[2 * rl];
          ridx[i] = r >= skip_below ? r : -1;           // -1: no such row, or a row C(d+1) owns
          rowoff[i] = tab[2 * rl + 1];
        }
      }
      WS_STAMP(2 + 12 * it + 3 * ch);
      const double* a_s = ring + ch * WS_STAGE + (wm * WTM + g) * LDK + t;
      const double* b_s = ring + ch * WS_STAGE + (UM + wn * WTN + g) * LDK + t;
#pragma unroll
      for (int k = 0; k < KC; k += 4) {
        double a[FM], bf[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) a[i] = a_s[(8 * i) * LDK + k];
#pragma unroll
        for (int j = 0; j < FN; ++j) bf[j] = b_s[(8 * j) * LDK + k];
        if (ch == 1) {
          // A_old, one fragment per k-step of the middle chunk (the row table has arrived by now, the values are not
          // needed before the epilogue).  A lane owns two adjacent list positions (even, odd) of a row — always columns
          // (c, c + 1), 16-byte aligned; c <= r covers the pair below the diagonal and the diagonal element c == r, where
          // the second slot (r, r + 1) exists in the row's storage and is never stored back.
          const int fi = (k / 4) / FN, fj = (k / 4) % FN;
          const int c = cidx[fj], r = ridx[fi];
          const bool any = c >= 0 && c <= r;
          ldg128_if(cold[fi][fj][0], cold[fi][fj][1], val + (any ? rowoff[fi] + c : 0), any);
        }
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
      }
      WS_STAMP(3 + 12 * it + 3 * ch);
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        // (no fence: the warps publish nothing through shared memory, and their reads of the stage have completed —
        // a fence here would also wait for this lane's A_old loads, which are meant to stay in flight)
        last = atomicAdd(done + ch, 1) == NWARPS - 1;
        if (last) { done[ch] = 0; if (more) issue(tin, tjn, ch); }
      }
      WS_STAMP(4 + 12 * it + 3 * ch);
    }
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        const int c = cidx[j], r = ridx[i];
        const bool pair = c >= 0 && c + 1 <= r, diag = c >= 0 && c == r;
        double* p = val + ((pair || diag) ? rowoff[i] + c : 0);
        stg128_if(p, cold[i][j][0] - acc[i][j][0], cold[i][j][1] - acc[i][j][1], pair);
        stg64_if(p, cold[i][j][0] - acc[i][j][0], diag);
      }
    WS_STAMP(11 + 12 * it);
    ti = tin; tj = tjn;
  }
  TL_OUT(val, 2 + PART, d);
}

// backward sweep, one launch per panel d = D-1 .. 0 (plus one leading launch that only computes x of the last panel):
//   push      acc[c] -= sum_{r in panel d} L[r][c] x_r  for every column c in [lo, c0) inside the rows' envelopes
//             (256 threads = 32 columns x 8 row groups per CTA);
//   next x    CTA 0 owns the 96 columns of panel d-1: it pushes into them first — which completes that panel's
//             accumulator, every later panel having pushed in earlier launches — and goes straight on to
//             x_{d-1} = Linv_{d-1}^T (y_{d-1} + acc_{d-1}) while the other CTAs push into the columns further left.
//             No counter, no fence: every accumulator entry is touched by one CTA per launch.
// Border panels of a partial factorisation (d >= D_elim) take x as given: they push, nobody computes them.
// The sweep is a chain of thousands of launches, each waiting for the one before and each a few microseconds long, so
// what it costs is latency.  Launches are therefore programmatic dependent launches: everything a launch needs that no
// earlier launch of the sweep writes — the factor entries, Linv, the right-hand side, the row tables — is fetched BEFORE
// griddepcontrol.wait, i.e. while the launch before is still running; only x of panel d and the accumulators are read
// after it.  launch_dependents follows the wait, so at most two launches of a sweep are resident at a time.
__global__ void __launch_bounds__(256) sky_backward_kernel(int d, int n, int lo, int do_push, int next_d, long long rhs_off, const long long* __restrict__ ptr,
                                                           const int* __restrict__ start, const double* __restrict__ val, const double* __restrict__ dinv,
                                                           double* __restrict__ acc, double* __restrict__ x) {
  constexpr int BWB = 4;                // blocks of 32 columns per CTA and round (BW_COLS = 32 BWB columns)
  __shared__ double xs[PW], red[BWB][8][33];
  __shared__ long long rbase[PW];       // ptr[r] - start[r]
  __shared__ int rstart[PW];
  const int tid = threadIdx.x;
  const int c0 = d * PW;
  const int w = do_push ? min(PW, n - c0) : 0;
  if (do_push && tid < PW) {
    const int r = c0 + tid;
    rbase[tid] = tid < w ? ptr[r] - start[r] : 0; rstart[tid] = tid < w ? start[r] : 0x7fffffff;
  }
  __syncthreads();
  if (blockIdx.x != 0) {
    // ---- CTAs 1..: the columns left of panel d-1, 128 per CTA and round: thread (column cx of each of the four blocks,
    // row group g of 12 rows) has its 48 factor entries in flight at once; one reduction over the row groups per round
    const int cx = tid & 31, g = tid >> 5;
    const int cend = max(lo, c0 - PW);
    const int cb0 = lo + (blockIdx.x - 1) * (32 * BWB), cstep = (gridDim.x - 1) * (32 * BWB);
    double v[BWB][PW / 8];
    auto fetch = [&](int cb) {
#pragma unroll
      for (int k = 0; k < BWB; ++k) {
        const int c = cb + 32 * k + cx;
#pragma unroll
        for (int q = 0; q < PW / 8; ++q) { const int i = g * (PW / 8) + q; v[k][q] = (c < cend && c >= rstart[i]) ? val[rbase[i] + c] : 0.0; }
      }
    };
    if (cb0 < cend) fetch(cb0);                       // the first (mostly the only) round: before the wait
    grid_dep_wait();
    grid_dep_launch();
    if (tid < PW) xs[tid] = tid < w ? x[c0 + tid] : 0.0;
    const int ck = tid >> 5, cc = tid & 31;           // threads 0..127 own column 32 ck + cc of the round
    double a_old = 0.0;
    if (tid < 32 * BWB && cb0 + tid < cend) a_old = acc[cb0 + tid];
    __syncthreads();
    for (int cb = cb0; cb < cend; cb += cstep) {
      if (cb != cb0) { fetch(cb); if (tid < 32 * BWB && cb + tid < cend) a_old = acc[cb + tid]; }
#pragma unroll
      for (int k = 0; k < BWB; ++k) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < PW / 8; ++q) s += v[k][q] * xs[g * (PW / 8) + q];
        red[k][g][cx] = s;
      }
      __syncthreads();
      if (tid < 32 * BWB && cb + tid < cend) {
        double tt = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) tt += red[ck][q][cc];
        acc[cb + tid] = a_old - tt;
      }
      __syncthreads();
    }
    return;
  }
  // ---- CTA 0: finish panel d-1.  Everything that does not depend on the earlier launches is fetched first: the
  // right-hand side, this thread's share of Linv_{d-1} and of the factor block (panel d rows, panel d-1 columns).
  __shared__ double half[2][PW], part[8][PW];
  const int cn0 = next_d * PW, wn = next_d >= 0 ? min(PW, n - cn0) : 0;
  const int ln = tid & 31, wq = tid >> 5;
  double rhs_pref = 0.0, li[PW / 8][3];
  if (next_d >= 0) {
    if (tid < wn) rhs_pref = val[rhs_off + cn0 + tid];
    const double* Li = dinv + (size_t)next_d * PW * PW;
#pragma unroll
    for (int q = 0; q < PW / 8; ++q) { const int i = wq + 8 * q; li[q][0] = Li[i * PW + ln]; li[q][1] = Li[i * PW + 32 + ln]; li[q][2] = Li[i * PW + 64 + ln]; }
  }
  // push of panel d into the 96 columns of panel d-1 in one round trip: thread (column, row half), all 48 loads in flight
  const int cl = tid % PW, h = tid / PW, cpush = c0 - PW + cl;
  const bool pusher = tid < 2 * PW && do_push && cpush >= lo;
  double v[PW / 2];
  if (pusher) {
#pragma unroll
    for (int q = 0; q < PW / 2; ++q) { const int i = h * (PW / 2) + q; v[q] = cpush >= rstart[i] ? val[rbase[i] + cpush] : 0.0; }
  }
  grid_dep_wait();
  grid_dep_launch();
  if (tid < PW) xs[tid] = tid < w ? x[c0 + tid] : 0.0;
  if (next_d >= 0 && tid < wn) rhs_pref += acc[cn0 + tid];     // the accumulator as the earlier launches left it
  __syncthreads();
  if (tid < 2 * PW) {
    double s = 0.0;
    if (pusher) {
#pragma unroll
      for (int q = 0; q < PW / 2; ++q) s += v[q] * xs[h * (PW / 2) + q];
    }
    half[h][cl] = s;
  }
  __syncthreads();
  if (tid < PW) {
    const double tt = half[0][tid] + half[1][tid];
    const int c = c0 - PW + tid;
    if (do_push && c >= lo && c >= 0) acc[c] -= tt;
    xs[tid] = (next_d >= 0 && tid < wn) ? rhs_pref - ((next_d == d - 1) ? tt : 0.0) : 0.0;   // rhs of panel d-1 (xs reused)
  }
  if (next_d < 0) return;
  __syncthreads();
  // x_j = sum_{i>=j} Linv[i][j] rhs_i: warp wq takes rows i = wq, wq+8, ..., a lane three columns.  Linv is stored with
  // explicit zeros above the diagonal, so no triangle test is needed.
  double p0 = 0.0, p1 = 0.0, p2 = 0.0;
#pragma unroll
  for (int q = 0; q < PW / 8; ++q) { const double ri = xs[wq + 8 * q]; p0 += li[q][0] * ri; p1 += li[q][1] * ri; p2 += li[q][2] * ri; }
  part[wq][ln] = p0; part[wq][32 + ln] = p1; part[wq][64 + ln] = p2;
  __syncthreads();
  if (tid < wn) {
    double sacc = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) sacc += part[q][tid];
    x[cn0 + tid] = sacc;
  }
}

static const size_t SM_TRSM = sizeof(double) * (PW * LDT + TR * LDT);
static const size_t SM_UPD = sizeof(double) * (2 * (UM + UN) * LDK);   // two stages: 108 KB, two CTAs per SM
static const size_t SM_DIAG = sizeof(double) * (2 * PW * LDT + (PW / 8) * 64 + 8 * LDT);
static const size_t SM_UPD_WS = sizeof(double) * (size_t)WS_NS * WS_STAGE + 2 * (UM + UN) * sizeof(long long) + WS_NS * sizeof(unsigned long long) + WS_NS * sizeof(int) + 4;

static int g_rest_ctas = 132;     // grid of the persistent update kernel
static int g_diag_mode = 1;       // 0: two barrier phases per step (sky_diag_kernel), 1: panel / update warps pipelined (sky_diag2_kernel)
static int g_chain_mode = 1;      // 1: C(d) and trsm(d) alternate on the chain stream, 0: trsm(d) on the panel stream (PGS_CHAIN_MODE)
static int g_backward_pdl = 1;    // backward sweep as programmatic dependent launches (PGS_BACKWARD_PDL=0: plain launches)
static int g_update_mode = 1;     // 0: one tile per CTA (sky_update_kernel), 1: warp-specialised persistent pipeline for rest(d), 2: for next(d) too
static int set_attrs(std::string* err) {
  // per device (the attributes live in the context) and under a lock (chains are enqueued from several host threads)
  static std::mutex mu; static std::set<int> ready;
  std::lock_guard<std::mutex> lk(mu);
  int cur = 0; cudaGetDevice(&cur);
  if (ready.count(cur)) return PGS_OK;
  SK(cudaFuncSetAttribute(sky_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_DIAG));
  SK(cudaFuncSetAttribute(sky_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_DIAG));
  SK(cudaFuncSetAttribute(sky_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TRSM));
  SK(cudaFuncSetAttribute(sky_update_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD));
  SK(cudaFuncSetAttribute(sky_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD));
  SK(cudaFuncSetAttribute(sky_update_ws_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD_WS));
  SK(cudaFuncSetAttribute(sky_update_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_UPD_WS));
  { int dev = 0, nsm = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // the persistent update CTAs fill a whole SM each; a few SMs stay free for the kernels of the panel chain
    const char* e = getenv("PGS_REST_SMS"); g_rest_ctas = e ? atoi(e) : nsm - 16; if (g_rest_ctas < 1) g_rest_ctas = 1;
    const char* m = getenv("PGS_UPDATE_MODE"); g_update_mode = m ? atoi(m) : 2;
    const char* dm = getenv("PGS_DIAG_MODE"); g_diag_mode = dm ? atoi(dm) : 1;
    const char* bp = getenv("PGS_BACKWARD_PDL"); g_backward_pdl = bp ? atoi(bp) : 1;
    const char* cm = getenv("PGS_CHAIN_MODE"); g_chain_mode = cm ? atoi(cm) : 1; }
  ready.insert(cur);
  return PGS_OK;
}

static int skyline_begin(SkylineFactor* f, std::string* err) {
  if (int rc = set_attrs(err)) return rc;
  SK(cudaMemsetAsync(f->val, 0, sizeof(double) * (size_t)(f->nnz + f->tail), f->stream));
  SK(cudaMemsetAsync(f->xacc, 0, sizeof(double) * (size_t)std::max(f->D, 1) * PW, f->stream));
  SK(cudaMemsetAsync(f->fail, 0, sizeof(int), f->stream));
  return PGS_OK;
}

int skyline_begin_border(SkylineFactor* f, std::string* err) { return skyline_begin(f, err); }
double* skyline_values(SkylineFactor* f) { return f->val; }
long long skyline_values_count(const SkylineFactor* f) { return f->nnz + f->tail; }
double* skyline_tail(SkylineFactor* f) { return f->val + f->nnz; }

// Numeric factorisation of the first D_elim panels (all of them for a single-GPU solve).
int skyline_factor(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, std::string* err) {
  if (int rc = skyline_begin(f, err)) return rc;
  const int tot = std::max(36 * std::max(f->N, f->n_pairs), f->n);
  sky_scatter_kernel<<<(tot + 255) / 256, 256, 0, f->stream>>>(f->N, f->n_pairs, Ad, Ao, b, f->pair_hi, f->pair_lo, f->node_src, f->pair_src, f->ptr, f->start, f->val);
  return skyline_factor_numeric(f, err);
}

// (A CUDA-graph replay of this loop was measured and is slower than direct launches: c3 5.6 s vs 4.2 s per 10-iteration
// solve — the three-stream overlap and the chain stream's priority do not survive the capture as well.)
int skyline_factor_numeric(SkylineFactor* f, std::string* err) {
  cudaStream_t s0 = f->stream, s1 = f->s1, s2 = f->s2;
  const int n = f->n;
  SK(cudaEventRecord(f->ev_fork, s0));
  SK(cudaStreamWaitEvent(s1, f->ev_fork, 0));
  SK(cudaStreamWaitEvent(s2, f->ev_fork, 0));
  // Chain mode 0 (PGS_CHAIN_MODE=0, the earlier arrangement):
  //   chain stream s1:  [wait trsm(d-1), rest(d-2)] C(d)    -> ev_c[d]
  //   panel stream s2:  [wait C(d)]                 trsm(d) -> ev_trsm[d]   [wait rest(d-1)] next(d)
  //   main  stream s0:  [wait trsm(d)]              rest(d) -> ev_rest[d]
  // next(d-1) precedes trsm(d) on s2, rest(d-1) precedes rest(d) on s0.
  // The persistent update CTAs fill a whole SM each; g_rest_ctas SMs are theirs, split between the factorisations that
  // share the GPU, the others stay free for the kernels of the panel chains (diag, trsm, next) at all times.
  const int rest_ctas = std::max(1, g_rest_ctas / f->share);
  // Chain mode 1 (the default): C(d) and trsm(d) alternate on the chain stream, so the critical path C(d) -> trsm(d) ->
  // C(d+1) has no cross-stream event hop of its own, and next(d) has the panel stream to itself:
  //   chain stream s1:  [wait rest(d-2)] C(d)   [wait next(d-1)] trsm(d) -> ev_trsm[d]
  //   panel stream s2:  [wait trsm(d), rest(d-1)] next(d) -> ev_next[d]
  //   main  stream s0:  [wait trsm(d)] rest(d) -> ev_rest[d]
  // Measured 38.8 us per panel against 41.0 for one chain on config 3 (profiles/r02_timeline_c3_one_chain*.txt).  The two
  // are plain launches: with a wait for another stream's event between them a programmatic dependent launch does not start
  // early (tools/pdl_lab.cu), and waiting on the device instead, with event records left between the launches, stalled
  // the events (profiles/r02_backward_pdl_and_chain_labs.txt).
  const bool chain1 = g_chain_mode == 1 && g_diag_mode == 1;
  for (int d = 0; d < f->D_elim; ++d) {
    const int nr = f->h_rows_ptr[d + 1] - f->h_rows_ptr[d];
    const int Tr = (nr + UM - 1) / UM, Tc = (nr + UN - 1) / UN;
    double* xp = f->xp + (size_t)(d % XP_RING) * f->xp_stride; long long* rinfo = f->rinfo + (size_t)(d % XP_RING) * f->rinfo_stride;
    if (chain1) {
      if (d > 1) SK(cudaStreamWaitEvent(s1, f->ev_rest[(d - 2) % NEV], 0));
      sky_diag2_kernel<<<1, DG2_THREADS, SM_DIAG, s1>>>(d, n, d > 0 ? 1 : 0, f->ptr, f->start, f->val, f->dinv, f->fail);
      if (d > 0) SK(cudaStreamWaitEvent(s1, f->ev_next[(d - 1) % NEV], 0));
      sky_trsm_kernel<<<Tr * (UM / TR), 256, SM_TRSM, s1>>>(d, n, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->dinv, f->val, xp, rinfo, f->rpad);
      SK(cudaEventRecord(f->ev_trsm[d % NEV], s1));
      SK(cudaStreamWaitEvent(s2, f->ev_trsm[d % NEV], 0));
    } else {
      if (d > 0) SK(cudaStreamWaitEvent(s1, f->ev_trsm[(d - 1) % NEV], 0));
      if (d > 1) SK(cudaStreamWaitEvent(s1, f->ev_rest[(d - 2) % NEV], 0));   // rest(d-2) holds part of panel d-2's update of A_dd
      if (g_diag_mode == 1) sky_diag2_kernel<<<1, DG2_THREADS, SM_DIAG, s1>>>(d, n, d > 0 ? 1 : 0, f->ptr, f->start, f->val, f->dinv, f->fail);
      else sky_diag_kernel<<<1, 256, SM_DIAG, s1>>>(d, n, d > 0 ? 1 : 0, f->ptr, f->start, f->val, f->dinv, f->fail);
      SK(cudaEventRecord(f->ev_c[d % NEV], s1));
      SK(cudaStreamWaitEvent(s2, f->ev_c[d % NEV], 0));
      sky_trsm_kernel<<<Tr * (UM / TR), 256, SM_TRSM, s2>>>(d, n, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->dinv, f->val, xp, rinfo, f->rpad);
      SK(cudaEventRecord(f->ev_trsm[d % NEV], s2));
    }
    // C(d+1) applies panel d's update to its own diagonal block; past the eliminated part nobody does, so next(d) keeps it.
    // The rhs row (index n) is always live, also when the last panel is short.
    const int skip_below = (d + 1 < f->D_elim) ? std::min((d + 2) * PW, n) : 0;
    if (d > 0) SK(cudaStreamWaitEvent(s2, f->ev_rest[(d - 1) % NEV], 0));
    if (g_update_mode >= 2) sky_update_ws_kernel<0><<<std::min(2 * Tr, g_rest_ctas), WS_THREADS, SM_UPD_WS, s2>>>(d, n, skip_below, Tr, Tc, 2 * Tr, f->rpad, xp, rinfo, f->val);
    else sky_update_kernel<0><<<2 * Tr, 256, SM_UPD, s2>>>(d, n, skip_below, Tr, Tc, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->val);
    if (chain1) SK(cudaEventRecord(f->ev_next[d % NEV], s2));
    SK(cudaStreamWaitEvent(s0, f->ev_trsm[d % NEV], 0));
    if (Tr > 1) {
      const int nt = Tr * (Tr - 1);
      if (g_update_mode >= 1) sky_update_ws_kernel<1><<<std::min(nt, rest_ctas), WS_THREADS, SM_UPD_WS, s0>>>(d, n, skip_below, Tr, Tc, nt, f->rpad, xp, rinfo, f->val);
      else sky_update_kernel<1><<<nt, 256, SM_UPD, s0>>>(d, n, skip_below, Tr, Tc, f->ptr, f->start, f->rows_ptr, f->rows_idx, f->val);
    }
    SK(cudaEventRecord(f->ev_rest[d % NEV], s0));
  }
  // join: the main stream continues after all three are done
  SK(cudaEventRecord(f->ev_join, s1));
  SK(cudaStreamWaitEvent(s0, f->ev_join, 0));
  SK(cudaEventRecord(f->ev_join2, s2));
  SK(cudaStreamWaitEvent(s0, f->ev_join2, 0));
  SK(cudaGetLastError());
  return PGS_OK;
}

// Backward substitution L^T x = y.  Panels >= D_elim (border, multi-GPU) take x as given in y[] beforehand.
// Every launch but the first is a programmatic dependent launch of the one before (see the kernel); the first one is an
// ordinary launch, so nothing of the sweep starts before the work queued ahead of it on the stream is complete.
static int skyline_backward_launches(SkylineFactor* f, double* y, std::string* err, bool allow_pdl = true) {
  cudaStream_t st = f->stream;
  const int n = f->n, D = f->D;
  const long long rhs_off = f->h_ptr[n];
  bool first = true;
  auto launch = [&](int grid, int d, int lo, int do_push, int next_d) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    const bool pdl = allow_pdl && g_backward_pdl && !first;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    first = false;
    return cudaLaunchKernelEx(&cfg, sky_backward_kernel, d, n, lo, do_push, next_d, rhs_off, (const long long*)f->ptr, (const int*)f->start,
                              (const double*)f->val, (const double*)f->dinv, f->xacc, y);
  };
  // x of the last panel (unless it is a given border panel)
  if (D - 1 < f->D_elim) SK(launch(1, D, 0, 0, D - 1));
  for (int d = D - 1; d >= 0; --d) {
    const int cols = d * PW - f->h_lo[d];
    const int next_d = (d - 1 >= 0 && d - 1 < f->D_elim) ? d - 1 : -1;
    if (cols <= 0 && next_d < 0) continue;
    const int left = std::max(0, cols - PW);                      // columns left of panel d-1, shared by CTAs 1.. (128 per CTA and round)
    const int grid = 1 + std::min(147, (left + 127) / 128);
    SK(launch(grid, d, f->h_lo[d], cols > 0 ? 1 : 0, next_d));
  }
  SK(cudaGetLastError());
  return PGS_OK;
}
// One small kernel per panel, each waiting for the one before.  The sequence never changes for a given factor, so it can
// be captured once into a CUDA graph and replayed (PGS_BACKWARD_GRAPH=1).  Measured on B200 this is NOT faster than plain
// launches (config 3: 1049 vs 984 ms per three LM iterations, config 2: 52.7 vs 50.7 ms): the sweep is bound by the
// kernel-to-kernel dependency latency on the device, not by the host's launch rate, so plain launches are the default.
int skyline_backward(SkylineFactor* f, double* y, std::string* err) {
  if (f->D == 0) return PGS_OK;
  static const bool use_graph = [] { const char* e = getenv("PGS_BACKWARD_GRAPH"); return e && atoi(e) != 0; }();
  if (!use_graph || f->bw_graph_failed || f->D < 64) return skyline_backward_launches(f, y, err);
  if (f->bw_graph && f->bw_y != y) { cudaGraphExecDestroy(f->bw_graph); f->bw_graph = nullptr; }
  if (!f->bw_graph) {
    cudaGraph_t g = nullptr;
    bool ok = cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const int rc = skyline_backward_launches(f, y, err, false);
      ok = cudaStreamEndCapture(f->stream, &g) == cudaSuccess && rc == PGS_OK && g;
    }
    if (ok) ok = cudaGraphInstantiate(&f->bw_graph, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (!ok) { cudaGetLastError(); f->bw_graph = nullptr; f->bw_graph_failed = true; return skyline_backward_launches(f, y, err); }
    f->bw_y = y;
  }
  SK(cudaGraphLaunch(f->bw_graph, f->stream));
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

// ---- border access for the Schur scheme: after a partial factorisation the trailing rows of a chain factor hold
// its contribution to the border system (dense over the nbc border scalars it holds) and to the border right-hand side.
// They are ADDED into the border factor `dst`, whose node order is the global border order: chain border node j sits at
// border node bmap[j] (ascending, so lower triangles map onto lower triangles).
__global__ void sky_border_accumulate_kernel(int n, int nint, const long long* __restrict__ ptr, const int* __restrict__ start, const double* __restrict__ val,
                                             const int* __restrict__ bmap, int dn, const long long* __restrict__ dptr, const int* __restrict__ dstart,
                                             double* __restrict__ dval) {
  const int nb = n - nint;
  const long long tot = (long long)nb * (nb + 1) / 2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= e) ++i;
    while ((long long)i * (i + 1) / 2 > e) --i;
    const int j = (int)(e - (long long)i * (i + 1) / 2);
    const int r = nint + i;
    const int R = 6 * bmap[i / 6] + i % 6, Cc = 6 * bmap[j / 6] + j % 6;
    dval[dptr[R] + (Cc - dstart[R])] += val[ptr[r] + (nint + j - start[r])];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) dval[dptr[dn] + 6 * bmap[i / 6] + i % 6] += val[ptr[n] + nint + i];
}
int skyline_border_accumulate(SkylineFactor* f, SkylineFactor* dst, const int* bmap_dev, cudaStream_t st, std::string* err) {
  const int nint = f->D_elim * PW;
  if (f->n == nint) return PGS_OK;
  sky_border_accumulate_kernel<<<592, 256, 0, st>>>(f->n, nint, f->ptr, f->start, f->val, bmap_dev, dst->n, dst->ptr, dst->start, dst->val);
  SK(cudaGetLastError());
  return PGS_OK;
}
__global__ void sky_add_diagonal_kernel(int n, const long long* __restrict__ ptr, const int* __restrict__ start, const double* __restrict__ add, double* __restrict__ val) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) val[ptr[r] + (r - start[r])] += add[r];
}
int skyline_add_diagonal(SkylineFactor* f, const double* add, std::string* err) {
  if (f->n) sky_add_diagonal_kernel<<<(f->n + 255) / 256, 256, 0, f->stream>>>(f->n, f->ptr, f->start, add, f->val);
  SK(cudaGetLastError());
  return PGS_OK;
}
#ifdef SKY_TIMELINE
extern "C" int pgs_debug_timeline(unsigned long long* out, int* d0, int* nd, int* kinds) {
  *d0 = TL_D0; *nd = TL_ND; *kinds = TL_KINDS;
  return cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 4 * TL_KINDS * TL_ND * 2) == cudaSuccess ? 0 : 1;
}
#endif
int skyline_interior_scalars(const SkylineFactor* f) { return f->D_elim * PW; }
const int* skyline_fail_flag(const SkylineFactor* f) { return f->fail; }

}  // namespace pgs
