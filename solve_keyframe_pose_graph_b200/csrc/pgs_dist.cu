// Multi-GPU solve: node-range sharding as a domain decomposition.
//
// Outer solver (the handle the caller holds): keeps the full problem on the host, partitions it
// (host/partition.h), builds this rank's LOCAL problem — interior nodes in natural order, padding up to a
// whole skyline panel, then ALL border nodes — in an inner Solver, and gathers the result.
// Inner solver: the ordinary LM loop of pgs_solver.cu on the local problem, plus the collectives below.
//
// Per linear solve, on the locally scaled system A' = S'HS' + D_int (S' = interior Jacobi scale, border
// unknowns unscaled and undamped, switches eliminated per edge):
//   1. skyline_factor eliminates the interior panels; the trailing rows then hold this rank's contribution
//      to the border Schur complement S'_bb and to the border right-hand side;
//   2. ONE all-reduce sums [S'_bb packed | rhs_b | diag(J^T J)_b] over the ranks (the "border blocks");
//   3. every rank adds the LM diagonal of the border unknowns, clamp(diag s_b^2)/(radius s_b^2) with
//      s_b = 1/(1+sqrt(diag at iteration 0)) — identical to scaling+damping the full system (T-congruence) —
//      factors the dense border system redundantly and back-substitutes its own interior.
// Small all-reduces carry the border gradient + cost after each evaluation and five scalars per step.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "host/partition.h"
#include "pgs_skyline.h"
#include "pgs_solver.h"

namespace pgs {

#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_fail(e__, #x); } while (0)
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// buf[0..nb6) = g[border], buf[nb6] = local cost
__global__ void border_pack_grad_kernel(int nb6, const double* __restrict__ g_border, const double* __restrict__ cost, double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) buf[i] = g_border[i];
  if (i == 0) buf[nb6] = *cost;
}
// gfull = g with the border part replaced by the summed one; cost slot <- summed cost
__global__ void border_unpack_grad_kernel(int n6, int first6, const double* __restrict__ g, const double* __restrict__ buf, double* __restrict__ gfull,
                                          double* __restrict__ cost) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n6) gfull[i] = i >= first6 ? buf[i - first6] : g[i];
  if (i == 0) *cost = buf[n6 - first6];
}
// diagH[i] = (J^T J)_ii of border scalar i (this rank's partial sum)
__global__ void border_pack_diag_kernel(int nb6, int first_border, const double* __restrict__ Hd, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) out[i] = Hd[36 * (size_t)(first_border + i / 6) + (i % 6) * 7];
}
// Jacobi scale (once) and clamped LM diagonal (unless reused) of the border unknowns from the SUMMED diag(J^T J);
// adds diag_b / (radius s_b^2) to the diagonal of the summed, locally unscaled border system.
__global__ void border_damp_kernel(int nb6, const double* __restrict__ diagH, int compute_scale, int jacobi, int reuse_diag, double lo, double hi,
                                   double inv_radius, double* __restrict__ sb, double* __restrict__ diagb, double* __restrict__ S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb6) return;
  const double n2 = diagH[i];
  if (compute_scale) sb[i] = jacobi ? 1.0 / (1.0 + sqrt(n2)) : 1.0;
  const double s = sb[i];
  if (!reuse_diag) diagb[i] = fmin(fmax(n2 * s * s, lo), hi);
  // a border node that appears in no residual block at all keeps a unit pivot (Ceres drops such blocks)
  const double add = n2 > 0.0 ? diagb[i] * inv_radius / (s * s) : 1.0;
  S[(long long)i * (i + 1) / 2 + i] += add;
}
__global__ void border_fail_kernel(const int* __restrict__ f0, const int* __restrict__ f1, double* __restrict__ out) {
  *out = (double)((f0 ? *f0 : 0) + (f1 ? *f1 : 0));
}

// ------------------------------------------------------------------------------------------------ inner solver
int Solver::border_gradient_exchange() {
  const int n6 = 6 * N, b6 = nb6(), first6 = n6 - b6;
  CU(d_xbuf.resize((size_t)b6 + 1)); CU(d_gfull.resize((size_t)std::max(n6, 1)));
  border_pack_grad_kernel<<<cdiv(b6 + 1, 256), 256, 0, stream>>>(b6, d_g.p + first6, d_scal.p + L_COST, d_xbuf.p);
  if (int rc = comm->allreduce_sum(d_xbuf.p, (size_t)b6 + 1, stream, &err)) return rc;
  border_unpack_grad_kernel<<<cdiv(std::max(n6, 1), 256), 256, 0, stream>>>(n6, first6, d_g.p, d_xbuf.p, d_gfull.p, d_scal.p + L_COST);
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::border_solve() {
  const int b6 = nb6(), nbn = b6 / 6;
  const size_t tri = (size_t)b6 * (b6 + 1) / 2;
  CU(d_xbuf.resize(tri + 2 * (size_t)b6)); CU(d_sb.resize(b6)); CU(d_diagb.resize(b6)); CU(d_zb.resize(b6));
  double* S = d_xbuf.p; double* rhs = S + tri; double* diagH = rhs + b6;
  if (int rc = skyline_border_get(sky, S, rhs, &err)) return rc;
  border_pack_diag_kernel<<<cdiv(b6, 256), 256, 0, stream>>>(b6, first_border, d_Hd.p, diagH);
  if (int rc = comm->allreduce_sum(d_xbuf.p, tri + 2 * (size_t)b6, stream, &err)) return rc;
  dstats.border_buffer_bytes = (int64_t)((tri + 2 * (size_t)b6) * sizeof(double));
  border_damp_kernel<<<cdiv(b6, 256), 256, 0, stream>>>(b6, diagH, border_scale_ready ? 0 : 1, opt.jacobi_scaling, cur_reuse_diag ? 1 : 0, opt.min_lm_diagonal,
                                                       opt.max_lm_diagonal, 1.0 / cur_radius, d_sb.p, d_diagb.p, S);
  border_scale_ready = true;
  CU(cudaGetLastError());
  if (!sky_border) {
    sky_border = skyline_create(nbn, 0, nullptr, nullptr, stream, &err, 0, /*dense=*/true);
    if (!sky_border) return PGS_ERR_OUT_OF_MEMORY;
  }
  if (int rc = skyline_load_packed(sky_border, S, rhs, &err)) return rc;
  if (int rc = skyline_factor_numeric(sky_border, &err)) return rc;
  if (int rc = skyline_backward(sky_border, d_zb.p, &err)) return rc;
  // border solution becomes the "given" tail of y for the interior back-substitution
  CU(cudaMemcpyAsync(d_y.p + (size_t)6 * first_border, d_zb.p, sizeof(double) * b6, cudaMemcpyDeviceToDevice, stream));
  return PGS_OK;
}

int Solver::dist_fail_flag() {
  border_fail_kernel<<<1, 1, 0, stream>>>(sky ? skyline_fail_flag(sky) : nullptr, sky_border ? skyline_fail_flag(sky_border) : nullptr, d_scal.p + L_FAIL);
  CU(cudaGetLastError());
  return PGS_OK;
}

// ------------------------------------------------------------------------------------------------ outer solver
int Solver::dist_init(int rank, int world, const void* id128) {
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(PGS_ERR_INVALID_ARGUMENT, "dist_init: bad rank/world/id");
  CU(cudaSetDevice(dev));
  comm_owned.reset(new Comm());
  if (int rc = comm_owned->init(rank, world, id128, &err)) { comm_owned.reset(); return rc; }
  inner.reset();
  structure_dirty = true;
  return PGS_OK;
}

int Solver::dist_stats(pgs_dist_stats* out) {
  if (!comm_owned) return fail(PGS_ERR_STATE, "dist_stats: pgs_dist_init was not called");
  *out = dstats;
  out->rank = comm_owned->rank; out->world = comm_owned->world;
  if (inner) out->border_buffer_bytes = inner->dstats.border_buffer_bytes;
  out->n_collectives = comm_owned->n_collectives; out->bytes_reduced = comm_owned->bytes_reduced;
  return PGS_OK;
}

int Solver::solve_dist(pgs_summary* sum, pgs_iteration* iters, int cap) {
  Comm* C = comm_owned.get();
  const int rank = C->rank, world = C->world;
  const int Eo = (int)o_c1.size(), El = (int)l_a.size(), K = (int)r_node.size();
  const int PN = skyline_panel_width() / 6;
  if (structure_dirty || !inner) {
    Partition P;
    make_partition(N, world, Eo, o_c1.data(), o_c2.data(), El, l_a.data(), l_b.data(), K, r_node.data(), &P);
    // local numbering: interior (natural order) | padding to a whole panel | all border nodes
    std::vector<int> glob2loc(N, -1);
    loc2glob.clear();
    for (int i = P.cut[rank]; i < P.cut[rank + 1]; ++i) if (P.node_owner[i] == rank) { glob2loc[i] = (int)loc2glob.size(); loc2glob.push_back(i); }
    const int n_int = (int)loc2glob.size();
    while (loc2glob.size() % PN) loc2glob.push_back(-1);
    const int fb = (int)loc2glob.size();
    for (int b : P.border) { glob2loc[b] = (int)loc2glob.size(); loc2glob.push_back(b); }
    const int Nl = (int)loc2glob.size();
    inner.reset(new Solver(opt));
    if (int rc = inner->init()) { err = inner->err; inner.reset(); return rc; }
    inner->comm = C; inner->first_border = P.border.empty() ? -1 : fb; inner->count_border = (rank == 0);
    std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl);
    for (int l = 0; l < Nl; ++l) {
      const int g = loc2glob[l];
      if (g < 0) { q[4 * (size_t)l] = q[4 * (size_t)l + 1] = q[4 * (size_t)l + 2] = 0.0; q[4 * (size_t)l + 3] = 1.0; t[3 * (size_t)l] = t[3 * (size_t)l + 1] = t[3 * (size_t)l + 2] = 0.0; }
      else { std::memcpy(&q[4 * (size_t)l], &h_q[4 * (size_t)g], 32); std::memcpy(&t[3 * (size_t)l], &h_t[3 * (size_t)g], 24); }
    }
    int rc = inner->set_nodes(Nl, q.data(), t.data(), false);
    // owned residual blocks, re-indexed
    std::vector<int> c1, c2; std::vector<double> eq, et, ew;
    for (int e = 0; e < Eo && rc == PGS_OK; ++e) if (P.odom_owner[e] == rank) {
      if (glob2loc[o_c1[e]] < 0 || glob2loc[o_c2[e]] < 0) return fail(PGS_ERR_STATE, "solve_dist: an owned edge has a node outside interior + border");
      c1.push_back(glob2loc[o_c1[e]]); c2.push_back(glob2loc[o_c2[e]]);
      eq.insert(eq.end(), &o_q[4 * (size_t)e], &o_q[4 * (size_t)e] + 4); et.insert(et.end(), &o_t[3 * (size_t)e], &o_t[3 * (size_t)e] + 3); ew.push_back(o_w[e]);
    }
    dstats.n_odom_owned = (int)c1.size();
    if (rc == PGS_OK && !c1.empty()) rc = inner->add_odom((int)c1.size(), c1.data(), c2.data(), eq.data(), et.data(), ew.data());
    c1.clear(); c2.clear(); eq.clear(); et.clear(); ew.clear(); loop2glob.clear();
    for (int e = 0; e < El; ++e) if (P.loop_owner[e] == rank) {
      if (glob2loc[l_a[e]] < 0 || glob2loc[l_b[e]] < 0) return fail(PGS_ERR_STATE, "solve_dist: an owned loop edge has a node outside interior + border");
      c1.push_back(glob2loc[l_a[e]]); c2.push_back(glob2loc[l_b[e]]); loop2glob.push_back(e);
      eq.insert(eq.end(), &l_q[4 * (size_t)e], &l_q[4 * (size_t)e] + 4); et.insert(et.end(), &l_t[3 * (size_t)e], &l_t[3 * (size_t)e] + 3); ew.push_back(l_w[e]);
    }
    dstats.n_loop_owned = (int)c1.size();
    if (rc == PGS_OK && !c1.empty()) rc = inner->add_loop((int)c1.size(), c1.data(), c2.data(), eq.data(), et.data(), ew.data());
    std::vector<int> rn; eq.clear(); et.clear(); ew.clear();
    for (int k = 0; k < K; ++k) if (P.reg_owner[k] == rank) {
      rn.push_back(glob2loc[r_node[k]]);
      eq.insert(eq.end(), &r_q[4 * (size_t)k], &r_q[4 * (size_t)k] + 4); et.insert(et.end(), &r_t[3 * (size_t)k], &r_t[3 * (size_t)k] + 3); ew.push_back(r_w[k]);
    }
    dstats.n_reg_owned = (int)rn.size();
    if (rc == PGS_OK) rc = inner->set_regs((int)rn.size(), rn.data(), eq.data(), et.data(), ew.data());
    if (rc != PGS_OK) { err = "solve_dist: building the local problem failed: " + inner->err; inner.reset(); return rc; }
    dstats.n_interior_nodes = n_int; dstats.n_border_nodes = (int)P.border.size();
    structure_dirty = false;
  } else {
    // same structure, new initial guesses
    const int Nl = (int)loc2glob.size();
    std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl);
    for (int l = 0; l < Nl; ++l) {
      const int g = loc2glob[l];
      if (g < 0) { q[4 * (size_t)l + 3] = 1.0; continue; }
      std::memcpy(&q[4 * (size_t)l], &h_q[4 * (size_t)g], 32); std::memcpy(&t[3 * (size_t)l], &h_t[3 * (size_t)g], 24);
    }
    if (int rc = inner->update_nodes(0, Nl, q.data(), t.data())) { err = inner->err; return rc; }
  }
  if (!loop2glob.empty()) {
    std::vector<double> s(loop2glob.size());
    for (size_t l = 0; l < s.size(); ++l) s[l] = h_sw[loop2glob[l]];
    if (int rc = inner->set_switches(0, (int)s.size(), s.data())) { err = inner->err; return rc; }
  }
  C->n_collectives = 0; C->bytes_reduced = 0;
  if (int rc = inner->solve(sum, iters, cap)) { err = inner->err; return rc; }

  // gather: every rank contributes its interior poses and owned switches (rank 0 also the border), summed into a
  // zero-initialised global vector -> every rank ends up with the complete solution
  const int Nl = (int)loc2glob.size();
  std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl), s(loop2glob.size());
  if (int rc = inner->get_poses(0, Nl, q.data(), t.data())) { err = inner->err; return rc; }
  if (!s.empty()) if (int rc = inner->get_switches(0, (int)s.size(), s.data())) { err = inner->err; return rc; }
  const size_t tot = 7 * (size_t)N + (size_t)El;
  std::vector<double> g(tot, 0.0);
  const int fb = inner->first_border >= 0 ? inner->first_border : Nl;
  for (int l = 0; l < Nl; ++l) {
    const int gi = loc2glob[l];
    if (gi < 0 || (l >= fb && rank != 0)) continue;
    std::memcpy(&g[4 * (size_t)gi], &q[4 * (size_t)l], 32); std::memcpy(&g[4 * (size_t)N + 3 * (size_t)gi], &t[3 * (size_t)l], 24);
  }
  for (size_t l = 0; l < s.size(); ++l) g[7 * (size_t)N + loop2glob[l]] = s[l];
  DBuf<double> dg;
  CU(dg.resize(tot));
  CU(cudaMemcpyAsync(dg.p, g.data(), sizeof(double) * tot, cudaMemcpyHostToDevice, stream));
  if (int rc = C->allreduce_sum(dg.p, tot, stream, &err)) return rc;
  CU(cudaMemcpyAsync(g.data(), dg.p, sizeof(double) * tot, cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  std::memcpy(h_q.data(), g.data(), sizeof(double) * 4 * (size_t)N);
  std::memcpy(h_t.data(), g.data() + 4 * (size_t)N, sizeof(double) * 3 * (size_t)N);
  if (El) std::memcpy(h_sw.data(), g.data() + 7 * (size_t)N, sizeof(double) * (size_t)El);
  host_params_newer = true; device_params_newer = false;
  return PGS_OK;
}

}  // namespace pgs
