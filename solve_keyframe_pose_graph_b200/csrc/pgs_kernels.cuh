// Device code of the B200 pose-graph hot path (sm_100a, fp64).
//
// K1  sweep_kernel          residuals + tangent Jacobian blocks for every odometry edge
//                           (SixDOFError, reference src/CeresResidues.h:19-90), every switchable loop
//                           edge (SixDOFErrorWithSwitchingConstraints, :145-222) and every node
//                           regulariser (NodePoseRegularization, :96-141) in ONE launch, + cost.
// K2  assemble_*            block-sparse J^T J (6x6 blocks) and J^T r, switch coupling per loop edge.
// K3  build_system_*        Jacobi scaling, clamped LM diagonal, analytic Schur elimination of the
//                           scalar switch unknowns -> reduced SPD system over the 6N pose unknowns.
// K4  pcg_* / skyline_*     linear solve on device (pgs_linear.cuh).
// K5  retract / norms       x+ = Plus(x, delta) (ceres::EigenQuaternionParameterization), step norms,
//                           model cost change.
//
// Data layout in HBM:
//   pose        double[N][8]            (qx,qy,qz,qw, tx,ty,tz, pad) — 64-B records, two 32-B sectors
//   edge index  int2[E]                 (c1,c2), edges sorted by (c1,c2) so a warp touches a short
//                                       window of consecutive nodes
//   edge consts double[tiles][8][32]    warp-tiled SoA: (qo4, to3, w); lane = edge within a tile of 32
//   residuals   double[tiles][R][32]    R = 6 (odom) / 7 (loop)
//   Jacobians   double[tiles][P][32]    P = 72 (odom: side*36+row*6+col) / 91 (loop: side*42+row*6+col, 84+row = switch col)
// Every global load/store of a plane by a warp is one fully-used 256-B segment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgs {

constexpr int TILE = 32;
constexpr int OD_R = 6, OD_J = 72, LP_R = 7, LP_J = 91, OBS = 8;

struct SweepArgs {
  const double* __restrict__ pose;     // [N][8]
  const double* __restrict__ sw;       // [El] switches (sorted loop order)
  const int2* __restrict__ o_idx; const double* __restrict__ o_obs; int n_odom;
  const int2* __restrict__ l_idx; const double* __restrict__ l_obs; int n_loop;
  const int* __restrict__ r_node; const double* __restrict__ r_anchor; int n_reg;   // anchor [K][8] = qf4,tf3,w
  double* __restrict__ o_r; double* __restrict__ o_J;
  double* __restrict__ l_r; double* __restrict__ l_J;
  double* __restrict__ g_r; double* __restrict__ g_J;      // regulariser [K][6], [K][36] (AoS, a handful)
  double* __restrict__ cost_tile;                          // [tiles] sum of squared residuals per tile of 32 edges
  unsigned int* __restrict__ sched;                        // [2] {next tile, blocks done}; zero between launches
  double* __restrict__ cost_out;                           // 0.5 * sum of cost_tile, written by the last block to finish
  // A launch may cover only part of the tiles (the end-to-end step sweeps the edges whose keyframes have arrived while the
  // rest of the poses is still on the bus): tiles [o_t0, o_t1) of the odometry edges, [l_t0, l_t1) of the loop edges and
  // [r_t0, r_t1) of the regularisers; reduce != 0 on the launch that completes the sweep (sums ALL per-tile partials).
  int o_t0, o_t1, l_t0, l_t1, r_t0, r_t1, reduce;
};

// ------------------------------------------------------------------ small math
struct Q4 { double x, y, z, w; };
struct P7 { Q4 q; double tx, ty, tz; double fixed; };   // fixed: 1.0 for a constant parameter block (8th slot of the pose record), else 0.0

__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ void qtoR(const Q4& q, double R[9]) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
__device__ __forceinline__ P7 load_pose(const double* __restrict__ pose, int i) {
  const double2* p = reinterpret_cast<const double2*>(pose + 8 * (size_t)i);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  P7 r; r.q.x = a.x; r.q.y = a.y; r.q.z = b.x; r.q.w = b.y; r.tx = c.x; r.ty = c.y; r.tz = d.x; r.fixed = d.y;
  return r;
}
// M = (L(A) Rm(b))[0:3,0:3]: derivative of vec(A (x) dq (x) b) w.r.t. the half-angle increment of dq.
__device__ __forceinline__ void quat_M(const Q4& A, const Q4& b, double M[9]) {
  const double G0 = A.w, G1 = -A.z, G2 = A.y, G3 = A.z, G4 = A.w, G5 = -A.x, G6 = -A.y, G7 = A.x, G8 = A.w;
  // [b_v]x G
  const double S0 = -b.z * G3 + b.y * G6, S1 = -b.z * G4 + b.y * G7, S2 = -b.z * G5 + b.y * G8;
  const double S3 = b.z * G0 - b.x * G6, S4 = b.z * G1 - b.x * G7, S5 = b.z * G2 - b.x * G8;
  const double S6 = -b.y * G0 + b.x * G3, S7 = -b.y * G1 + b.x * G4, S8 = -b.y * G2 + b.x * G5;
  M[0] = -b.x * A.x + b.w * G0 - S0; M[1] = -b.x * A.y + b.w * G1 - S1; M[2] = -b.x * A.z + b.w * G2 - S2;
  M[3] = -b.y * A.x + b.w * G3 - S3; M[4] = -b.y * A.y + b.w * G4 - S4; M[5] = -b.y * A.z + b.w * G5 - S5;
  M[6] = -b.z * A.x + b.w * G6 - S6; M[7] = -b.z * A.y + b.w * G7 - S7; M[8] = -b.z * A.z + b.w * G8 - S8;
}

// e = [R2^T (R1 t_o - t2 + t1) ; 2 vec(q2* (x) q1 (x) q_o)] and the four 3x3 blocks its tangent
// Jacobian is made of: Rt = R2^T, Ba = R2^T [R1 t_o]x, Bv = R2^T [v]x, M.
template <bool JAC>
__device__ __forceinline__ void sixdof_core(const P7& p1, const P7& p2, const Q4& qo, double ox, double oy, double oz,
                                            double e[6], double Rt[9], double Ba[9], double Bv[9], double M[9]) {
  double R1[9], R2[9];
  qtoR(p1.q, R1); qtoR(p2.q, R2);
  const double a0 = R1[0] * ox + R1[1] * oy + R1[2] * oz;
  const double a1 = R1[3] * ox + R1[4] * oy + R1[5] * oz;
  const double a2 = R1[6] * ox + R1[7] * oy + R1[8] * oz;
  const double v0 = a0 - p2.tx + p1.tx, v1 = a1 - p2.ty + p1.ty, v2 = a2 - p2.tz + p1.tz;
  e[0] = R2[0] * v0 + R2[3] * v1 + R2[6] * v2;
  e[1] = R2[1] * v0 + R2[4] * v1 + R2[7] * v2;
  e[2] = R2[2] * v0 + R2[5] * v1 + R2[8] * v2;
  const Q4 b = qmul(p1.q, qo);
  const Q4 A{-p2.q.x, -p2.q.y, -p2.q.z, p2.q.w};
  const Q4 dq = qmul(A, b);
  e[3] = 2.0 * dq.x; e[4] = 2.0 * dq.y; e[5] = 2.0 * dq.z;
  if (JAC) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double r0 = R2[i], r1 = R2[3 + i], r2 = R2[6 + i];   // row i of R2^T
      Rt[3 * i] = r0; Rt[3 * i + 1] = r1; Rt[3 * i + 2] = r2;
      Ba[3 * i] = r1 * a2 - r2 * a1; Ba[3 * i + 1] = r2 * a0 - r0 * a2; Ba[3 * i + 2] = r0 * a1 - r1 * a0;
      Bv[3 * i] = r1 * v2 - r2 * v1; Bv[3 * i + 1] = r2 * v0 - r0 * v2; Bv[3 * i + 2] = r0 * v1 - r1 * v0;
    }
    quat_M(A, b, M);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Streaming store: the Jacobian planes are written once and not re-read by this kernel.
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

// ------------------------------------------------------------------ K1
// MODE 0: residuals + Jacobians (ceres Evaluate with jacobians)   MODE 1: cost only (candidate point)
// Tiles of 32 edges are handed to the warps of a persistent grid by an atomic counter: the sweep is a 40-50 us,
// store-bound kernel, and a static round-robin leaves the last partial round on a quarter of the SMs (measured
// 7 % slower, tools/sweep_lab.cu).  Every tile writes its own cost partial, so the summation order — and with it
// the cost, bit for bit — does not depend on which warp happened to take which tile.
#ifndef PGS_SWEEP_MINB
#define PGS_SWEEP_MINB 2
#endif
template <int MODE>
__global__ void __launch_bounds__(256, PGS_SWEEP_MINB) sweep_kernel(SweepArgs A) {
  const int lane = threadIdx.x & 31;
  const int To = (A.n_odom + TILE - 1) / TILE, Tl = (A.n_loop + TILE - 1) / TILE, Tr = (A.n_reg + TILE - 1) / TILE;
  const int T = To + Tl + Tr;

  // (Drawing the ticket for the NEXT tile before processing the current one — to hide the atomic's round trip — was
  // measured on the same box: 51.2 us against 48.0 us per sweep.  The plain loop stays.)
  const int no = A.o_t1 - A.o_t0, nl = A.l_t1 - A.l_t0, nt = no + nl + (A.r_t1 - A.r_t0);
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = (int)atomicAdd(A.sched, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= nt) break;
    tile = tile < no ? A.o_t0 + tile : (tile < no + nl ? To + A.l_t0 + (tile - no) : To + Tl + A.r_t0 + (tile - no - nl));   // ticket -> tile of this launch's ranges
    double cost = 0.0;
    if (tile < To) {
      // ---- odometry edges: r = w e, J = w Je
      const int e = tile * TILE + lane;
      if (e < A.n_odom) {
        const int2 ij = __ldg(A.o_idx + e);
        const double* ob = A.o_obs + (size_t)tile * (OBS * TILE) + lane;
        const Q4 qo{__ldg(ob), __ldg(ob + TILE), __ldg(ob + 2 * TILE), __ldg(ob + 3 * TILE)};
        const double ox = __ldg(ob + 4 * TILE), oy = __ldg(ob + 5 * TILE), oz = __ldg(ob + 6 * TILE), w = __ldg(ob + 7 * TILE);
        const P7 p1 = load_pose(A.pose, ij.x), p2 = load_pose(A.pose, ij.y);
        double ev[6], Rt[9], Ba[9], Bv[9], M[9];
        sixdof_core<MODE == 0>(p1, p2, qo, ox, oy, oz, ev, Rt, Ba, Bv, M);
#pragma unroll
        for (int i = 0; i < 6; ++i) { ev[i] *= w; cost += ev[i] * ev[i]; }
        if (MODE == 0) {
          double* r = A.o_r + (size_t)tile * (OD_R * TILE) + lane;
#pragma unroll
          for (int i = 0; i < 6; ++i) st_stream(r + i * TILE, ev[i]);
          double* J = A.o_J + (size_t)tile * (OD_J * TILE) + lane;
          // a constant parameter block (ceres SetParameterBlockConstant, reference PoseGraphSLAM.cpp:150-151) gets zero columns
          const double wa = w * (1.0 - p1.fixed), wb = w * (1.0 - p2.fixed), wa2 = 2.0 * wa, wb2 = 2.0 * wb;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              st_stream(J + (i * 6 + j) * TILE, -wa2 * Ba[3 * i + j]);            // d e_t / d th1
              st_stream(J + (i * 6 + 3 + j) * TILE, wa * Rt[3 * i + j]);          // d e_t / d t1
              st_stream(J + ((3 + i) * 6 + j) * TILE, wa2 * M[3 * i + j]);        // d e_r / d th1
              st_stream(J + ((3 + i) * 6 + 3 + j) * TILE, 0.0);
              st_stream(J + (36 + i * 6 + j) * TILE, wb2 * Bv[3 * i + j]);        // d e_t / d th2
              st_stream(J + (36 + i * 6 + 3 + j) * TILE, -wb * Rt[3 * i + j]);    // d e_t / d t2
              st_stream(J + (36 + (3 + i) * 6 + j) * TILE, -wb2 * M[3 * i + j]);  // d e_r / d th2
              st_stream(J + (36 + (3 + i) * 6 + 3 + j) * TILE, 0.0);
            }
          }
        }
      }
    } else if (tile < To + Tl) {
      // ---- switchable loop edges: r = s [e ; 1-s], pose cols = s Je, switch col = [e ; 1-2s]
      const int lt = tile - To;
      const int e = lt * TILE + lane;
      if (e < A.n_loop) {
        const int2 ij = __ldg(A.l_idx + e);
        const double* ob = A.l_obs + (size_t)lt * (OBS * TILE) + lane;
        const Q4 qo{__ldg(ob), __ldg(ob + TILE), __ldg(ob + 2 * TILE), __ldg(ob + 3 * TILE)};
        const double ox = __ldg(ob + 4 * TILE), oy = __ldg(ob + 5 * TILE), oz = __ldg(ob + 6 * TILE);
        const double s = __ldg(A.sw + e);
        const P7 p1 = load_pose(A.pose, ij.x), p2 = load_pose(A.pose, ij.y);
        double ev[6], Rt[9], Ba[9], Bv[9], M[9];
        sixdof_core<MODE == 0>(p1, p2, qo, ox, oy, oz, ev, Rt, Ba, Bv, M);
        const double r6 = s * (1.0 - s);
        cost += r6 * r6;
#pragma unroll
        for (int i = 0; i < 6; ++i) { const double ri = s * ev[i]; cost += ri * ri; }
        if (MODE == 0) {
          double* r = A.l_r + (size_t)lt * (LP_R * TILE) + lane;
#pragma unroll
          for (int i = 0; i < 6; ++i) st_stream(r + i * TILE, s * ev[i]);
          st_stream(r + 6 * TILE, r6);
          double* J = A.l_J + (size_t)lt * (LP_J * TILE) + lane;
          const double sa = s * (1.0 - p1.fixed), sb = s * (1.0 - p2.fixed), sa2 = 2.0 * sa, sb2 = 2.0 * sb;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              st_stream(J + (i * 6 + j) * TILE, -sa2 * Ba[3 * i + j]);
              st_stream(J + (i * 6 + 3 + j) * TILE, sa * Rt[3 * i + j]);
              st_stream(J + ((3 + i) * 6 + j) * TILE, sa2 * M[3 * i + j]);
              st_stream(J + ((3 + i) * 6 + 3 + j) * TILE, 0.0);
              st_stream(J + (42 + i * 6 + j) * TILE, sb2 * Bv[3 * i + j]);
              st_stream(J + (42 + i * 6 + 3 + j) * TILE, -sb * Rt[3 * i + j]);
              st_stream(J + (42 + (3 + i) * 6 + j) * TILE, -sb2 * M[3 * i + j]);
              st_stream(J + (42 + (3 + i) * 6 + 3 + j) * TILE, 0.0);
            }
          }
#pragma unroll
          for (int j = 0; j < 6; ++j) { st_stream(J + (36 + j) * TILE, 0.0); st_stream(J + (78 + j) * TILE, 0.0); }  // row 6 of both sides
#pragma unroll
          for (int i = 0; i < 6; ++i) st_stream(J + (84 + i) * TILE, ev[i]);
          st_stream(J + 90 * TILE, 1.0 - 2.0 * s);
        }
      }
    } else {
      // ---- node regularisers: r = w [Rf^T (t - tf) ; 2 sgn vec(qf* (x) q)]
      const int k = (tile - To - Tl) * TILE + lane;
      if (k < A.n_reg) {
        const int node = __ldg(A.r_node + k);
        const double* an = A.r_anchor + 8 * (size_t)k;
        const Q4 qf{an[0], an[1], an[2], an[3]};
        const double w = an[7];
        const P7 p = load_pose(A.pose, node);
        double Rf[9]; qtoR(qf, Rf);
        const double d0 = p.tx - an[4], d1 = p.ty - an[5], d2 = p.tz - an[6];
        const Q4 Aq{-qf.x, -qf.y, -qf.z, qf.w};
        const Q4 d = qmul(Aq, p.q);
        // sign Eigen's Quaternion(Matrix3) would give the relative rotation
        const double n2 = d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        double sgn;
        if (4.0 * d.w * d.w - n2 > 0.0) sgn = d.w >= 0.0 ? 1.0 : -1.0;
        else {
          const double m0 = d.x * d.x, m1 = d.y * d.y, m2 = d.z * d.z;
          double c = d.x, mi = m0;
          if (m1 > mi) { c = d.y; mi = m1; }
          if (m2 > mi) { c = d.z; }
          sgn = c >= 0.0 ? 1.0 : -1.0;
        }
        double rv[6];
        rv[0] = w * (Rf[0] * d0 + Rf[3] * d1 + Rf[6] * d2);
        rv[1] = w * (Rf[1] * d0 + Rf[4] * d1 + Rf[7] * d2);
        rv[2] = w * (Rf[2] * d0 + Rf[5] * d1 + Rf[8] * d2);
        rv[3] = 2.0 * w * sgn * d.x; rv[4] = 2.0 * w * sgn * d.y; rv[5] = 2.0 * w * sgn * d.z;
#pragma unroll
        for (int i = 0; i < 6; ++i) cost += rv[i] * rv[i];
        if (MODE == 0) {
          double M[9]; quat_M(Aq, p.q, M);
#pragma unroll
          for (int i = 0; i < 6; ++i) A.g_r[6 * k + i] = rv[i];
          double* J = A.g_J + 36 * (size_t)k;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double wf = w * (1.0 - p.fixed);
              J[i * 6 + j] = 0.0; J[i * 6 + 3 + j] = wf * Rf[3 * j + i];
              J[(3 + i) * 6 + j] = 2.0 * wf * sgn * M[3 * i + j]; J[(3 + i) * 6 + 3 + j] = 0.0;
            }
        }
      }
    }
    cost = warp_sum(cost);
    if (lane == 0) A.cost_tile[tile] = cost;
  }
  // The last block to leave sums the per-tile partials in a fixed order (thread i takes tiles i, i+256, ...; then
  // the usual shuffle tree) and re-arms the scheduler: no second launch for the cost, and the same bits every run.
  __shared__ double red[8];
  __shared__ int is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(A.sched + 1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  if (!A.reduce) { if (threadIdx.x == 0) { A.sched[0] = 0u; A.sched[1] = 0u; } return; }   // a partial launch only re-arms the scheduler
  __threadfence();
  double s = 0.0;
  for (int i0 = threadIdx.x; i0 < T; i0 += 8 * blockDim.x) {   // eight loads in flight, added in index order
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < T ? __ldcg(A.cost_tile + i) : 0.0; }
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  s = warp_sum(s);
  if (lane == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *A.cost_out = 0.5 * t;
    A.sched[0] = 0u; A.sched[1] = 0u;
  }
}

// out[slot] = scale * sum(partial[0..n))  — single block, fixed order (deterministic).
__global__ void reduce_sum_kernel(const double* __restrict__ partial, int n, double scale, double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < (blockDim.x >> 5); ++i) t += sm[i]; *out = scale * t; }
}
__global__ void reduce_max_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmax(s, partial[i]);
  s = warp_max(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < (blockDim.x >> 5); ++i) t = fmax(t, sm[i]); *out = t; }
}

// block-level helper: every thread contributes v; thread 0 gets the block sum (fixed order)
__device__ __forceinline__ double block_sum(double v, double* sm /*[32]*/) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < ((blockDim.x + 31) >> 5); ++i) t += sm[i];
  __syncthreads();
  return t;
}
__device__ __forceinline__ double block_max(double v, double* sm) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < ((blockDim.x + 31) >> 5); ++i) t = fmax(t, sm[i]);
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------ accessors into the tiled J/r
__device__ __forceinline__ double jo(const double* __restrict__ J, int e, int k) { return __ldg(J + ((size_t)(e >> 5) * OD_J + k) * TILE + (e & 31)); }
__device__ __forceinline__ double jl(const double* __restrict__ J, int e, int k) { return __ldg(J + ((size_t)(e >> 5) * LP_J + k) * TILE + (e & 31)); }
__device__ __forceinline__ double ro(const double* __restrict__ r, int e, int k) { return __ldg(r + ((size_t)(e >> 5) * OD_R + k) * TILE + (e & 31)); }
__device__ __forceinline__ double rl(const double* __restrict__ r, int e, int k) { return __ldg(r + ((size_t)(e >> 5) * LP_R + k) * TILE + (e & 31)); }

struct AsmArgs {
  int N, n_odom, n_loop, n_reg, n_pairs;
  const double* __restrict__ o_r; const double* __restrict__ o_J;
  const double* __restrict__ l_r; const double* __restrict__ l_J;
  const double* __restrict__ g_r; const double* __restrict__ g_J;
  const int* __restrict__ inc_ptr; const int* __restrict__ inc_item;     // per node: (edge<<3 | kind<<1 | side); kind 0 odom, 1 loop, 2 reg
  const int* __restrict__ pe_ptr; const int* __restrict__ pe_item;       // per pair: (edge<<2 | kind<<1 | c1_is_hi)
  double* __restrict__ Hd;   // [N][36]
  double* __restrict__ g;    // [N][6]
  double* __restrict__ Ho;   // [P][36]  block (row hi, col lo)
  double* __restrict__ lv;   // [El][12] J_p^T j_s
  double* __restrict__ lh;   // [El]     j_s^T j_s
  double* __restrict__ lg;   // [El]     j_s^T r
};

// ------------------------------------------------------------------ K2a: diagonal blocks + gradient, one thread per node
__global__ void __launch_bounds__(128) assemble_diag_kernel(AsmArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.N) return;
  double H[21], g[6];
#pragma unroll
  for (int k = 0; k < 21; ++k) H[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  const int b = A.inc_ptr[i], e_ = A.inc_ptr[i + 1];
  for (int p = b; p < e_; ++p) {
    const int code = __ldg(A.inc_item + p);
    const int e = code >> 3, kind = (code >> 1) & 3, side = code & 1;
    const int rows = kind == 1 ? 6 : 6;   // row 6 of a loop block is zero in the pose columns
    for (int rrow = 0; rrow < rows; ++rrow) {
      double jr[6], rv;
      if (kind == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) jr[c] = jo(A.o_J, e, side * 36 + rrow * 6 + c);
        rv = ro(A.o_r, e, rrow);
      } else if (kind == 1) {
#pragma unroll
        for (int c = 0; c < 6; ++c) jr[c] = jl(A.l_J, e, side * 42 + rrow * 6 + c);
        rv = rl(A.l_r, e, rrow);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) jr[c] = A.g_J[36 * (size_t)e + rrow * 6 + c];
        rv = A.g_r[6 * (size_t)e + rrow];
      }
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        g[a] += jr[a] * rv;
#pragma unroll
        for (int c = 0; c <= a; ++c) H[k++] += jr[a] * jr[c];
      }
    }
  }
  double* Hd = A.Hd + 36 * (size_t)i;
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = 0; c <= a; ++c) { Hd[a * 6 + c] = H[k]; Hd[c * 6 + a] = H[k]; ++k; }
#pragma unroll
  for (int a = 0; a < 6; ++a) A.g[6 * (size_t)i + a] = g[a];
}

// ------------------------------------------------------------------ K2b: off-diagonal blocks, one thread per node pair
__global__ void __launch_bounds__(128) assemble_offdiag_kernel(AsmArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.n_pairs) return;
  double B[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) B[k] = 0.0;
  for (int q = A.pe_ptr[p]; q < A.pe_ptr[p + 1]; ++q) {
    const int code = __ldg(A.pe_item + q);
    const int e = code >> 2, kind = (code >> 1) & 1, c1hi = code & 1;
    const int hi_side = c1hi ? 0 : 1, lo_side = 1 - hi_side;
    for (int rrow = 0; rrow < 6; ++rrow) {
      double jh[6], jw[6];
      if (kind == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { jh[c] = jo(A.o_J, e, hi_side * 36 + rrow * 6 + c); jw[c] = jo(A.o_J, e, lo_side * 36 + rrow * 6 + c); }
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) { jh[c] = jl(A.l_J, e, hi_side * 42 + rrow * 6 + c); jw[c] = jl(A.l_J, e, lo_side * 42 + rrow * 6 + c); }
      }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 6; ++c) B[a * 6 + c] += jh[a] * jw[c];
    }
  }
  double* Ho = A.Ho + 36 * (size_t)p;
#pragma unroll
  for (int k = 0; k < 36; ++k) Ho[k] = B[k];
}

// ------------------------------------------------------------------ K2c: switch coupling, one thread per loop edge
__global__ void __launch_bounds__(128) assemble_switch_kernel(AsmArgs A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_loop) return;
  double js[7], h = 0.0, gs = 0.0;
#pragma unroll
  for (int i = 0; i < 7; ++i) { js[i] = jl(A.l_J, e, 84 + i); h += js[i] * js[i]; gs += js[i] * rl(A.l_r, e, i); }
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    const int side = c / 6, col = c % 6;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += jl(A.l_J, e, side * 42 + i * 6 + col) * js[i];
    A.lv[12 * (size_t)e + c] = s;
  }
  A.lh[e] = h; A.lg[e] = gs;
}

// ------------------------------------------------------------------ K3: scaling, LM diagonal, switch elimination
struct SysArgs {
  int N, n_loop, n_pairs;
  int first_border;   // nodes >= first_border are border unknowns of a domain decomposition (== N when there are none)
  double inv_radius;
  const int2* __restrict__ l_idx;
  const double* __restrict__ Hd; const double* __restrict__ g; const double* __restrict__ Ho;
  const double* __restrict__ lv; const double* __restrict__ lh; const double* __restrict__ lg;
  const double* __restrict__ scale_p; const double* __restrict__ scale_s;   // [6N], [El]
  const double* __restrict__ diag_p; const double* __restrict__ diag_s;     // clamped squared column norms of the scaled J
  const int* __restrict__ inc_ptr; const int* __restrict__ inc_item;
  const int* __restrict__ pe_ptr; const int* __restrict__ pe_item;
  const int2* __restrict__ pair;     // (hi, lo)
  double* __restrict__ lvt;  // [El][12] scaled v
  double* __restrict__ lw;   // [El]  1 / (scaled hss + D_s^2)
  double* __restrict__ lgt;  // [El]  scaled g_s
  double* __restrict__ Ad; double* __restrict__ Ao; double* __restrict__ b;
};

__global__ void __launch_bounds__(128) system_switch_kernel(SysArgs A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_loop) return;
  const int2 ij = A.l_idx[e];
  const double ss = A.scale_s[e];
  const double h = ss * ss * A.lh[e] + A.diag_s[e] * A.inv_radius;
  A.lw[e] = 1.0 / h;
  A.lgt[e] = ss * A.lg[e];
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    const int node = c < 6 ? ij.x : ij.y;
    A.lvt[12 * (size_t)e + c] = A.lv[12 * (size_t)e + c] * A.scale_p[6 * (size_t)node + (c % 6)] * ss;
  }
}

__global__ void __launch_bounds__(128) system_diag_kernel(SysArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.N) return;
  const int b0 = A.inc_ptr[i], b1 = A.inc_ptr[i + 1];
  double* Ad = A.Ad + 36 * (size_t)i;
  double* rhs = A.b + 6 * (size_t)i;
  if (b0 == b1) {  // node in no residual block: Ceres drops the parameter block; keep the system SPD
    // (a border node without local blocks contributes nothing to the summed border system)
    const double one = i >= A.first_border ? 0.0 : 1.0;
#pragma unroll
    for (int k = 0; k < 36; ++k) Ad[k] = (k % 7 == 0) ? one : 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) rhs[k] = 0.0;
    return;
  }
  double s[6], M[36], r[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) { s[a] = A.scale_p[6 * (size_t)i + a]; r[a] = s[a] * A.g[6 * (size_t)i + a]; }
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = 0; c < 6; ++c) M[a * 6 + c] = s[a] * A.Hd[36 * (size_t)i + a * 6 + c] * s[c];
#pragma unroll
  for (int a = 0; a < 6; ++a) M[a * 7] += A.diag_p[6 * (size_t)i + a] * A.inv_radius;
  for (int p = b0; p < b1; ++p) {
    const int code = __ldg(A.inc_item + p);
    if (((code >> 1) & 3) != 1) continue;
    const int e = code >> 3, side = code & 1;
    const double w = A.lw[e], gt = A.lgt[e];
    double v[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) v[a] = A.lvt[12 * (size_t)e + side * 6 + a];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      r[a] -= w * gt * v[a];
#pragma unroll
      for (int c = 0; c < 6; ++c) M[a * 6 + c] -= w * v[a] * v[c];
    }
  }
#pragma unroll
  for (int k = 0; k < 36; ++k) Ad[k] = M[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) rhs[k] = r[k];
}

__global__ void __launch_bounds__(128) system_offdiag_kernel(SysArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.n_pairs) return;
  const int2 hl = A.pair[p];
  double sh[6], sl[6], M[36];
#pragma unroll
  for (int a = 0; a < 6; ++a) { sh[a] = A.scale_p[6 * (size_t)hl.x + a]; sl[a] = A.scale_p[6 * (size_t)hl.y + a]; }
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = 0; c < 6; ++c) M[a * 6 + c] = sh[a] * A.Ho[36 * (size_t)p + a * 6 + c] * sl[c];
  for (int q = A.pe_ptr[p]; q < A.pe_ptr[p + 1]; ++q) {
    const int code = __ldg(A.pe_item + q);
    if (((code >> 1) & 1) != 1) continue;
    const int e = code >> 2, c1hi = code & 1;
    const double w = A.lw[e];
    const double* vh = A.lvt + 12 * (size_t)e + (c1hi ? 0 : 6);
    const double* vl = A.lvt + 12 * (size_t)e + (c1hi ? 6 : 0);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int c = 0; c < 6; ++c) M[a * 6 + c] -= w * vh[a] * vl[c];
  }
#pragma unroll
  for (int k = 0; k < 36; ++k) A.Ao[36 * (size_t)p + k] = M[k];
}

// switch back-substitution and sign flip: y_s = w (g~_s - v~^T y_p); step = -y; delta = step * scale
__global__ void __launch_bounds__(128) finish_step_kernel(int N, int n_loop, const int2* __restrict__ l_idx, const double* __restrict__ y,
                                                          const double* __restrict__ lvt, const double* __restrict__ lw, const double* __restrict__ lgt,
                                                          const double* __restrict__ scale_p, const double* __restrict__ scale_s,
                                                          double* __restrict__ delta_p, double* __restrict__ delta_s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * N) delta_p[i] = -y[i] * scale_p[i];
  if (i < n_loop) {
    const int2 ij = l_idx[i];
    double s = lgt[i];
#pragma unroll
    for (int c = 0; c < 12; ++c) s -= lvt[12 * (size_t)i + c] * y[6 * (size_t)(c < 6 ? ij.x : ij.y) + (c % 6)];
    delta_s[i] = -(s * lw[i]) * scale_s[i];
  }
}

// ------------------------------------------------------------------ model cost change  -(J d)^T (r + J d / 2)
struct MccArgs {
  int n_odom, n_loop, n_reg;
  const int2* __restrict__ o_idx; const int2* __restrict__ l_idx; const int* __restrict__ r_node;
  const double* __restrict__ o_r; const double* __restrict__ o_J; const double* __restrict__ l_r; const double* __restrict__ l_J;
  const double* __restrict__ g_r; const double* __restrict__ g_J;
  const double* __restrict__ dp; const double* __restrict__ ds;
  double* __restrict__ partial;
};
__global__ void __launch_bounds__(256) model_cost_kernel(MccArgs A) {
  __shared__ double sm[32];
  double acc = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < A.n_odom; e += stride) {
    const int2 ij = A.o_idx[e];
    double d[12];
#pragma unroll
    for (int c = 0; c < 6; ++c) { d[c] = A.dp[6 * (size_t)ij.x + c]; d[6 + c] = A.dp[6 * (size_t)ij.y + c]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double m = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) m += jo(A.o_J, e, i * 6 + c) * d[c] + jo(A.o_J, e, 36 + i * 6 + c) * d[6 + c];
      acc += m * (ro(A.o_r, e, i) + 0.5 * m);
    }
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < A.n_loop; e += stride) {
    const int2 ij = A.l_idx[e];
    double d[12];
#pragma unroll
    for (int c = 0; c < 6; ++c) { d[c] = A.dp[6 * (size_t)ij.x + c]; d[6 + c] = A.dp[6 * (size_t)ij.y + c]; }
    const double dsw = A.ds[e];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      double m = jl(A.l_J, e, 84 + i) * dsw;
#pragma unroll
      for (int c = 0; c < 6; ++c) m += jl(A.l_J, e, i * 6 + c) * d[c] + jl(A.l_J, e, 42 + i * 6 + c) * d[6 + c];
      acc += m * (rl(A.l_r, e, i) + 0.5 * m);
    }
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < A.n_reg; k += stride) {
    const int node = A.r_node[k];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double m = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) m += A.g_J[36 * (size_t)k + i * 6 + c] * A.dp[6 * (size_t)node + c];
      acc += m * (A.g_r[6 * (size_t)k + i] + 0.5 * m);
    }
  }
  const double t = block_sum(acc, sm);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = t;
}

// ------------------------------------------------------------------ K5: retraction + norms
// cand = Plus(x, sign*d) per used node / switch.  Partials: [0]=sum (x-cand)^2, [1]=sum x^2, [2]=max |x-cand|
// Nodes with node_counted == 0 are retracted but left out of the norms (a border node is counted by one rank only).
__global__ void __launch_bounds__(256) retract_kernel(int N, const char* __restrict__ node_counted, int n_loop, const char* __restrict__ node_used, const double* __restrict__ pose,
                                                      const double* __restrict__ sw, const double* __restrict__ dp, const double* __restrict__ ds,
                                                      double sign, double* __restrict__ cpose, double* __restrict__ csw,
                                                      double* __restrict__ p_diff2, double* __restrict__ p_x2, double* __restrict__ p_max) {
  __shared__ double sm[32];
  double d2 = 0.0, x2 = 0.0, mx = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const double* x = pose + 8 * (size_t)i; double* c = cpose + 8 * (size_t)i;
    if (!node_used[i]) { for (int k = 0; k < 8; ++k) c[k] = x[k]; continue; }   // unused or constant blocks stay as they are
    const double a0 = sign * dp[6 * (size_t)i], a1 = sign * dp[6 * (size_t)i + 1], a2 = sign * dp[6 * (size_t)i + 2];
    const double n = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    double o[7];
    if (n > 0.0) {
      const double sn = sin(n) / n;
      const Q4 dq{sn * a0, sn * a1, sn * a2, cos(n)};
      const Q4 r = qmul(dq, Q4{x[0], x[1], x[2], x[3]});
      o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
    } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
    o[4] = x[4] + sign * dp[6 * (size_t)i + 3]; o[5] = x[5] + sign * dp[6 * (size_t)i + 4]; o[6] = x[6] + sign * dp[6 * (size_t)i + 5];
    const bool counted = node_counted[i] != 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) { const double df = x[k] - o[k]; if (counted) { d2 += df * df; x2 += x[k] * x[k]; mx = fmax(mx, fabs(df)); } c[k] = o[k]; }
    c[7] = x[7];
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_loop; e += stride) {
    const double s = sw[e], c = s + sign * ds[e];
    csw[e] = c; const double df = s - c; d2 += df * df; x2 += s * s; mx = fmax(mx, fabs(df));
  }
  const double t0 = block_sum(d2, sm), t1 = block_sum(x2, sm), t2 = block_max(mx, sm);
  if (threadIdx.x == 0) { p_diff2[blockIdx.x] = t0; p_x2[blockIdx.x] = t1; p_max[blockIdx.x] = t2; }
}

// gradient of the switches into a dense vector (for |Plus(x,-g)-x| and pgs_gradient)
__global__ void scatter_lg_kernel(int n_loop, const double* __restrict__ lg, double* __restrict__ gs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_loop) gs[e] = lg[e];
}

// Jacobi scaling (once, at iteration 0) and the clamped LM diagonal (whenever !reuse_diagonal):
//   scale = 1/(1+sqrt(colnorm2)) ; diag = clamp(colnorm2 * scale^2, lo, hi)  with colnorm2 = diag(J^T J)
// Border unknowns of a domain decomposition (i >= 6*first_border) stay unscaled and undamped here: their column
// norms are only partial sums; scaling and damping are applied to the summed border system.
__global__ void scaling_kernel(int N, int first_border, int n_loop, const double* __restrict__ Hd, const double* __restrict__ lh, int compute_scale, int jacobi,
                               double lo, double hi, double* __restrict__ scale_p, double* __restrict__ scale_s,
                               double* __restrict__ diag_p, double* __restrict__ diag_s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * N) {
    if (i >= 6 * first_border) { scale_p[i] = 1.0; diag_p[i] = 0.0; }
    else {
      const double n2 = Hd[36 * (size_t)(i / 6) + (i % 6) * 7];
      if (compute_scale) scale_p[i] = jacobi ? 1.0 / (1.0 + sqrt(n2)) : 1.0;
      const double s = scale_p[i];
      diag_p[i] = fmin(fmax(n2 * s * s, lo), hi);
    }
  }
  if (i < n_loop) {
    const double n2 = lh[i];
    if (compute_scale) scale_s[i] = jacobi ? 1.0 / (1.0 + sqrt(n2)) : 1.0;
    const double s = scale_s[i];
    diag_s[i] = fmin(fmax(n2 * s * s, lo), hi);
  }
}

// ------------------------------------------------------------------ Summary::fixed_cost
// 1/2 sum |r|^2 over the residual blocks whose parameter blocks are all constant (odometry blocks between two constant
// keyframes, regularisers on a constant keyframe): Ceres' preprocessor takes them out of the reduced program and reports
// their cost as Summary::fixed_cost.  One block, fixed summation order.
__global__ void fixed_cost_kernel(int n_fo, const int* __restrict__ fo, const double* __restrict__ o_r, int n_fr, const int* __restrict__ fr,
                                  const double* __restrict__ g_r, double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_fo; i += blockDim.x) { const int e = fo[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double v = ro(o_r, e, k); s += v * v; } }
  for (int i = threadIdx.x; i < n_fr; i += blockDim.x) { const int k0 = fr[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double v = g_r[6 * (size_t)k0 + k]; s += v * v; } }
  const double t = block_sum(s, sm);
  if (threadIdx.x == 0) *out = 0.5 * t;
}

// ------------------------------------------------------------------ re-layout for the C-ABI (parity / debugging path)
// tiled SoA (sorted edge order) -> caller order, row-major r[E][R], J[E][R][C] with the two sides interleaved per row
__global__ void export_odom_kernel(int n, const int* __restrict__ perm, const double* __restrict__ r, const double* __restrict__ J,
                                   double* __restrict__ r_out, double* __restrict__ J_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int o = perm[e];
  if (r_out) for (int i = 0; i < 6; ++i) r_out[6 * (size_t)o + i] = ro(r, e, i);
  if (J_out) for (int i = 0; i < 6; ++i) for (int c = 0; c < 6; ++c) {
    J_out[72 * (size_t)o + 12 * i + c] = jo(J, e, i * 6 + c);
    J_out[72 * (size_t)o + 12 * i + 6 + c] = jo(J, e, 36 + i * 6 + c);
  }
}
__global__ void export_loop_kernel(int n, const int* __restrict__ perm, const double* __restrict__ r, const double* __restrict__ J,
                                   double* __restrict__ r_out, double* __restrict__ J_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int o = perm[e];
  if (r_out) for (int i = 0; i < 7; ++i) r_out[7 * (size_t)o + i] = rl(r, e, i);
  if (J_out) for (int i = 0; i < 7; ++i) {
    for (int c = 0; c < 6; ++c) {
      J_out[91 * (size_t)o + 13 * i + c] = jl(J, e, i * 6 + c);
      J_out[91 * (size_t)o + 13 * i + 6 + c] = jl(J, e, 42 + i * 6 + c);
    }
    J_out[91 * (size_t)o + 13 * i + 12] = jl(J, e, 84 + i);
  }
}

}  // namespace pgs
