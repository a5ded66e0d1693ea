// Sharded linear solve: node-range sharding as a domain decomposition, on one GPU
// (two chains burning from both ends of the keyframe chain) and across GPUs (one process per GPU, NCCL).
//
// Outer solver (the handle the caller holds): keeps the full problem on the host, plans it (host/partition.h),
// builds this rank's LOCAL problem — the interiors of its chains in elimination order, each padded to whole
// skyline panels, then the border nodes its chains touch — in an inner Solver, and gathers the result.
// Inner solver: the ordinary LM loop of pgs_solver.cu on the local problem, plus what is below.
//
// Per linear solve, on the locally scaled system A' = S'HS' + D_int (S' = interior Jacobi scale, border
// unknowns unscaled and undamped, switches eliminated per edge):
//   1. every chain is eliminated by its own skyline factor, all chains of a rank concurrently on their own
//      streams; the trailing rows of a chain factor then hold its contribution to the border Schur complement
//      and to the border right-hand side;
//   2. the contributions are added into the border system, a skyline over ALL border nodes of the graph whose
//      envelope is block tridiagonal in the separators (a chain only couples the separators at its two ends);
//      across GPUs ONE all-reduce sums [border envelope | rhs | diag(J^T J)_b] over the ranks (the "border blocks");
//   3. every rank adds the LM diagonal of the border unknowns, clamp(diag s_b^2)/(radius s_b^2) with
//      s_b = 1/(1+sqrt(diag at iteration 0)) — identical to scaling+damping the full system (T-congruence) —
//      factors the border system (redundantly across ranks: it is small) and back-substitutes its chains.
// Small all-reduces carry the border gradient + cost after each evaluation and a few scalars per step.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>

#include "host/partition.h"
#include "pgs_skyline.h"
#include "pgs_solver.h"

namespace pgs {

#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_fail(e__, #x); } while (0)
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// buf (zeroed, global border order) <- g of the local border nodes; buf[6 nG] = local cost
__global__ void border_pack_grad_kernel(int nb6, int first6, const int* __restrict__ gpos, const double* __restrict__ g, const double* __restrict__ cost,
                                        int ng6, double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) buf[6 * gpos[i / 6] + i % 6] = g[first6 + i];
  if (i == 0) buf[ng6] = *cost;
}
// gfull = g with the border part replaced by the summed one; cost slot <- summed cost
__global__ void border_unpack_grad_kernel(int n6, int first6, const int* __restrict__ gpos, const double* __restrict__ g, const double* __restrict__ buf,
                                          int ng6, double* __restrict__ gfull, double* __restrict__ cost) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n6) gfull[i] = i >= first6 ? buf[6 * gpos[(i - first6) / 6] + (i - first6) % 6] : g[i];
  if (i == 0) *cost = buf[ng6];
}
// out (zeroed, global border order)[.] = (J^T J)_ii of the local border scalars (this rank's partial sum)
__global__ void border_pack_diag_kernel(int nb6, int first_border, const int* __restrict__ gpos, const double* __restrict__ Hd, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) out[6 * gpos[i / 6] + i % 6] = Hd[36 * (size_t)(first_border + i / 6) + (i % 6) * 7];
}
// Jacobi scale (once) and clamped LM diagonal (unless reused) of the border unknowns from the SUMMED diag(J^T J);
// damp[i] = diag_b / (radius s_b^2) is what gets added to the diagonal of the summed, locally unscaled border system.
__global__ void border_damp_kernel(int ng6, const double* __restrict__ diagH, int compute_scale, int jacobi, int reuse_diag, double lo, double hi,
                                   double inv_radius, double* __restrict__ sb, double* __restrict__ diagb, double* __restrict__ damp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng6) return;
  const double n2 = diagH[i];
  if (compute_scale) sb[i] = jacobi ? 1.0 / (1.0 + sqrt(n2)) : 1.0;
  const double s = sb[i];
  if (!reuse_diag) diagb[i] = fmin(fmax(n2 * s * s, lo), hi);
  // a border node that appears in no residual block at all (or only as a constant block) keeps a unit pivot (Ceres drops such blocks)
  damp[i] = n2 > 0.0 ? diagb[i] * inv_radius / (s * s) : 1.0;
}
__global__ void border_fail_kernel(int n, const int* const* __restrict__ flags, double* __restrict__ out) {
  int s = 0;
  for (int i = 0; i < n; ++i) s += *flags[i];
  *out = (double)s;
}
// y of the local border nodes <- the border solution (global order); optionally the same into a chain's tail
__global__ void border_take_kernel(int nb6, const int* __restrict__ gpos, const double* __restrict__ zb, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) y[i] = zb[6 * gpos[i / 6] + i % 6];
}

// ------------------------------------------------------------------------------------------------ inner solver
int Solver::border_gradient_exchange() {
  const int n6 = 6 * N, b6 = nb6(), first6 = n6 - b6, ng6 = 6 * n_gborder;
  CU(d_gfull.resize((size_t)std::max(n6, 1)));
  if (!comm) {   // one rank: the local border gradient is already the whole sum
    CU(cudaMemcpyAsync(d_gfull.p, d_g.p, sizeof(double) * n6, cudaMemcpyDeviceToDevice, stream));
    return PGS_OK;
  }
  CU(d_xbuf.resize((size_t)ng6 + 1));
  CU(cudaMemsetAsync(d_xbuf.p, 0, sizeof(double) * ((size_t)ng6 + 1), stream));
  border_pack_grad_kernel<<<cdiv(std::max(b6, 1), 256), 256, 0, stream>>>(b6, first6, d_border_gpos.p, d_g.p, d_scal.p + L_COST, ng6, d_xbuf.p);
  if (int rc = comm->allreduce_sum(d_xbuf.p, (size_t)ng6 + 1, stream, &err)) return rc;
  border_unpack_grad_kernel<<<cdiv(std::max(n6, 1), 256), 256, 0, stream>>>(n6, first6, d_border_gpos.p, d_g.p, d_xbuf.p, ng6, d_gfull.p, d_scal.p + L_COST);
  CU(cudaGetLastError());
  return PGS_OK;
}

void Solver::release_chains() {
  for (ChainState& c : cstate) {
    if (c.f) skyline_destroy(c.f);
    if (c.done) cudaEventDestroy(c.done);
    if (c.st) cudaStreamDestroy(c.st);
  }
  cstate.clear();
  if (sky_border) { skyline_destroy(sky_border); sky_border = nullptr; }
}

// Chain factors over (chain interior + the border nodes it touches) and the border factor over all border nodes.
int Solver::prepare_chains() {
  if (!cstate.empty()) return PGS_OK;
  CU(cudaSetDevice(dev));
  const int fb = first_border >= 0 ? first_border : N, nbl = N - fb;
  if (!ev_fork) CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  std::vector<int> chain_of(N, -1);               // interior node -> chain
  for (size_t c = 0; c < chains.size(); ++c) for (int j = 0; j < chains[c].len; ++j) chain_of[chains[c].off + j] = (int)c;
  std::vector<std::vector<char>> member(chains.size(), std::vector<char>(std::max(nbl, 1), 0));
  std::vector<int> diag_chain(std::max(nbl, 1), -1);   // the chain that carries a border node's own diagonal block and rhs
  for (size_t c = 0; c < chains.size(); ++c)
    for (int b : chains[c].border) { member[c][b - fb] = 1; if (diag_chain[b - fb] < 0) diag_chain[b - fb] = (int)c; }
  for (int b = 0; b < nbl; ++b) if (diag_chain[b] < 0) return fail(PGS_ERR_STATE, "prepare_chains: a local border node belongs to no chain");
  // pairs -> chains
  std::vector<std::vector<int>> pairs_of(chains.size());
  for (int p = 0; p < n_pairs; ++p) {
    const int hi = h_pair_hi[p], lo = h_pair_lo[p];
    int c = lo < fb ? chain_of[lo] : (hi < fb ? chain_of[hi] : -1);
    if (c < 0) { for (size_t k = 0; k < chains.size() && c < 0; ++k) if (member[k][hi - fb] && member[k][lo - fb]) c = (int)k; }
    if (c < 0) return fail(PGS_ERR_STATE, "prepare_chains: a block couples nodes no chain holds together");
    if ((hi >= fb && !member[c][hi - fb]) || (lo >= fb && !member[c][lo - fb])) return fail(PGS_ERR_STATE, "prepare_chains: a chain's block reaches a border node outside its factor");
    if ((hi < fb && chain_of[hi] != c) || (lo < fb && chain_of[lo] != c)) return fail(PGS_ERR_STATE, "prepare_chains: a block couples two chain interiors");
    pairs_of[c].push_back(p);
  }
  int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
  cstate.resize(chains.size());
  factor_nnz = 0;
  std::vector<int> fidx(N, -1);
  for (size_t c = 0; c < chains.size(); ++c) {
    const ChainSpec& ch = chains[c];
    ChainState& cs = cstate[c];
    cs.nb = (int)ch.border.size(); cs.nf = ch.len + cs.nb;
    std::vector<int> node_src(std::max(cs.nf, 1), -1), bmap(std::max(cs.nb, 1), 0);
    for (int j = 0; j < ch.len; ++j) { fidx[ch.off + j] = j; node_src[j] = ch.off + j; }
    for (int k = 0; k < cs.nb; ++k) { const int b = ch.border[k]; fidx[b] = ch.len + k; node_src[ch.len + k] = diag_chain[b - fb] == (int)c ? b : -1; bmap[k] = border_gpos[b - fb]; }
    std::vector<int> phi(pairs_of[c].size()), plo(pairs_of[c].size());
    for (size_t k = 0; k < pairs_of[c].size(); ++k) { const int p = pairs_of[c][k]; phi[k] = fidx[h_pair_hi[p]]; plo[k] = fidx[h_pair_lo[p]]; }
    CU(cudaStreamCreateWithFlags(&cs.st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&cs.done, cudaEventDisableTiming));
    cs.f = skyline_create(cs.nf, (int)phi.size(), phi.data(), plo.data(), cs.st, &err, cs.nb, false, node_src.data(), pairs_of[c].data());
    if (!cs.f) return PGS_ERR_OUT_OF_MEMORY;
    skyline_set_share(cs.f, (int)chains.size());
    factor_nnz += skyline_nnz(cs.f);
    CU(cs.y.resize((size_t)std::max(6 * cs.nf, 1), true));
    CU(cs.bmap.upload(bmap, cs.st));
    CU(cudaStreamSynchronize(cs.st));
    for (int j = 0; j < ch.len; ++j) fidx[ch.off + j] = -1;
    for (int b : ch.border) fidx[b] = -1;
  }
  if (n_gborder > 0) {
    std::vector<int> phi, plo;
    for (int b = 0; b < n_gborder; ++b) if (gborder_env[b] < b) { phi.push_back(b); plo.push_back(gborder_env[b]); }
    sky_border = skyline_create(n_gborder, (int)phi.size(), phi.data(), plo.data(), stream, &err, 0, false, nullptr, nullptr, 6LL * n_gborder);
    if (!sky_border) return PGS_ERR_OUT_OF_MEMORY;
    factor_nnz += skyline_nnz(sky_border);
    dstats.border_buffer_bytes = (int64_t)(skyline_values_count(sky_border) * sizeof(double));
    CU(d_sb.resize(6 * (size_t)n_gborder)); CU(d_diagb.resize(6 * (size_t)n_gborder)); CU(d_zb.resize(6 * (size_t)n_gborder, true)); CU(d_dampb.resize(6 * (size_t)n_gborder));
  }
  dstats.factor_nnz = factor_nnz;
  return PGS_OK;
}

int Solver::solve_chains() {
  if (int rc = prepare_chains()) return rc;
  const int fb = first_border >= 0 ? first_border : N, b6 = nb6(), ng6 = 6 * n_gborder;
  // 1. the chains, concurrently.  Pivot failures are not checked here: the flags travel with the scalar all-reduce
  //    of the LM step so that every rank takes the same branch.
  CU(cudaEventRecord(ev_fork, stream));
  {
    // A chain is ~4 launches and ~8 event operations per panel: one host thread per chain, or the second chain's first
    // kernel would be enqueued only after all of the first chain's and the chains would hardly overlap on the device.
    std::vector<int> rcs(cstate.size(), PGS_OK); std::vector<std::string> errs(cstate.size());
    auto enqueue = [&](size_t c) {
      ChainState& cs = cstate[c];
      cudaSetDevice(dev);
      cudaError_t e = cudaStreamWaitEvent(cs.st, ev_fork, 0);
      if (e != cudaSuccess) { rcs[c] = PGS_ERR_CUDA; errs[c] = cudaGetErrorString(e); return; }
      rcs[c] = skyline_factor(cs.f, d_Ad.p, d_Ao.p, d_b.p, &errs[c]);
      if (rcs[c] == PGS_OK && cudaEventRecord(cs.done, cs.st) != cudaSuccess) { rcs[c] = PGS_ERR_CUDA; errs[c] = "cudaEventRecord"; }
    };
    std::vector<std::thread> th;
    for (size_t c = 1; c < cstate.size(); ++c) th.emplace_back(enqueue, c);
    enqueue(0);
    for (std::thread& t : th) t.join();
    for (size_t c = 0; c < cstate.size(); ++c) if (rcs[c] != PGS_OK) { err = errs[c]; return rcs[c]; }
  }
  if (sky_border) if (int rc = skyline_begin_border(sky_border, &err)) return rc;
  for (ChainState& cs : cstate) CU(cudaStreamWaitEvent(stream, cs.done, 0));
  if (!ev_ph[0]) for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&ev_ph[i]));
  CU(cudaEventRecord(ev_ph[0], stream));   // this rank's chains are eliminated
  if (sky_border) {
    // 2. border system = sum of the chains' Schur complements (+ the other ranks')
    for (ChainState& cs : cstate) if (cs.nb) if (int rc = skyline_border_accumulate(cs.f, sky_border, cs.bmap.p, stream, &err)) return rc;
    double* diagH = skyline_tail(sky_border);
    if (b6) border_pack_diag_kernel<<<cdiv(b6, 256), 256, 0, stream>>>(b6, fb, d_border_gpos.p, d_Hd.p, diagH);
    CU(cudaEventRecord(ev_ph[1], stream));
    if (comm) if (int rc = comm->allreduce_sum(skyline_values(sky_border), (size_t)skyline_values_count(sky_border), stream, &err)) return rc;
    CU(cudaEventRecord(ev_ph[2], stream));   // ... and every other rank's: the wait for the slowest rank is in [1, 2]
    // 3. damping of the border unknowns, border factorisation and solve
    border_damp_kernel<<<cdiv(ng6, 256), 256, 0, stream>>>(ng6, diagH, border_scale_ready ? 0 : 1, opt.jacobi_scaling, cur_reuse_diag ? 1 : 0, opt.min_lm_diagonal,
                                                          opt.max_lm_diagonal, 1.0 / cur_radius, d_sb.p, d_diagb.p, d_dampb.p);
    border_scale_ready = true;
    CU(cudaGetLastError());
    if (int rc = skyline_add_diagonal(sky_border, d_dampb.p, &err)) return rc;
    if (int rc = skyline_factor_numeric(sky_border, &err)) return rc;
    if (int rc = skyline_backward(sky_border, d_zb.p, &err)) return rc;
    if (b6) border_take_kernel<<<cdiv(b6, 256), 256, 0, stream>>>(b6, d_border_gpos.p, d_zb.p, d_y.p + 6 * (size_t)fb);
    CU(cudaGetLastError());
  }
  // 4. back-substitution of the chains, concurrently again
  CU(cudaEventRecord(ev_ph[3], stream));
  ph_pending = sky_border != nullptr;
  CU(cudaEventRecord(ev_fork, stream));
  {
    std::vector<int> rcs(cstate.size(), PGS_OK); std::vector<std::string> errs(cstate.size());
    auto enqueue = [&](size_t c) {
      ChainState& cs = cstate[c];
      const int len6 = 6 * chains[c].len;
      cudaSetDevice(dev);
      if (cudaStreamWaitEvent(cs.st, ev_fork, 0) != cudaSuccess) { rcs[c] = PGS_ERR_CUDA; errs[c] = "cudaStreamWaitEvent"; return; }
      if (cs.nb) border_take_kernel<<<cdiv(6 * cs.nb, 256), 256, 0, cs.st>>>(6 * cs.nb, cs.bmap.p, d_zb.p, cs.y.p + len6);
      rcs[c] = skyline_backward(cs.f, cs.y.p, &errs[c]);
      if (rcs[c] != PGS_OK) return;
      if (len6 && cudaMemcpyAsync(d_y.p + 6 * (size_t)chains[c].off, cs.y.p, sizeof(double) * len6, cudaMemcpyDeviceToDevice, cs.st) != cudaSuccess) { rcs[c] = PGS_ERR_CUDA; errs[c] = "cudaMemcpyAsync"; return; }
      if (cudaEventRecord(cs.done, cs.st) != cudaSuccess) { rcs[c] = PGS_ERR_CUDA; errs[c] = "cudaEventRecord"; }
    };
    std::vector<std::thread> th;
    for (size_t c = 1; c < cstate.size(); ++c) th.emplace_back(enqueue, c);
    enqueue(0);
    for (std::thread& t : th) t.join();
    for (size_t c = 0; c < cstate.size(); ++c) if (rcs[c] != PGS_OK) { err = errs[c]; return rcs[c]; }
  }
  for (ChainState& cs : cstate) CU(cudaStreamWaitEvent(stream, cs.done, 0));
  return PGS_OK;
}

// Phase times of the last solve_chains (call after the stream has been synchronised past it): local elimination is
// billed to ms_linear_solve's caller, the border exchange (all-reduce incl. the wait for the slowest rank) to ms_comm.
void Solver::collect_chain_times(float ms_from_start_to_eliminated) {
  if (!ph_pending) return;
  ph_pending = false;
  float a = 0, b = 0;
  if (cudaEventElapsedTime(&a, ev_ph[1], ev_ph[2]) == cudaSuccess) ms_exchange += a;
  if (cudaEventElapsedTime(&b, ev_ph[2], ev_ph[3]) == cudaSuccess) ms_border += b;
  ms_eliminate += ms_from_start_to_eliminated;
}

int Solver::dist_fail_flag() {
  std::vector<const int*> flags;
  if (sky) flags.push_back(skyline_fail_flag(sky));
  if (sky_border) flags.push_back(skyline_fail_flag(sky_border));
  for (ChainState& cs : cstate) if (cs.f) flags.push_back(skyline_fail_flag(cs.f));
  if (d_flag_ptrs.n < flags.size() || !flag_ptrs_ready) {
    CU(d_flag_ptrs.resize(std::max<size_t>(flags.size(), 1)));
    CU(cudaMemcpyAsync(d_flag_ptrs.p, flags.data(), sizeof(const int*) * flags.size(), cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));   // `flags` is a local
    flag_ptrs_ready = true;
  }
  border_fail_kernel<<<1, 1, 0, stream>>>((int)flags.size(), d_flag_ptrs.p, d_scal.p + L_FAIL);
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
  structure_dirty = true; inner_dirty = true;
  return PGS_OK;
}
int Solver::dist_init_local(int rank, int world, const char* group) {
  if (world < 1 || rank < 0 || rank >= world || !group) return fail(PGS_ERR_INVALID_ARGUMENT, "dist_init_local: bad rank/world/group");
  CU(cudaSetDevice(dev));
  comm_owned.reset(new Comm());
  if (int rc = comm_owned->init_local(rank, world, group, &err)) { comm_owned.reset(); return rc; }
  inner.reset();
  structure_dirty = true; inner_dirty = true;
  return PGS_OK;
}

int Solver::dist_stats(pgs_dist_stats* out) {
  if (!comm_owned && !inner) return fail(PGS_ERR_STATE, "dist_stats: no sharded solve has run on this handle");
  *out = dstats;
  out->rank = comm_owned ? comm_owned->rank : 0; out->world = comm_owned ? comm_owned->world : 1;
  if (inner) { out->border_buffer_bytes = inner->dstats.border_buffer_bytes; out->factor_nnz = inner->dstats.factor_nnz; out->ms_comm = inner->ms_comm + inner->ms_exchange;
               out->ms_eliminate = inner->ms_eliminate; out->ms_border = inner->ms_border; out->ms_exchange = inner->ms_exchange; }
  out->n_collectives = comm_owned ? comm_owned->n_collectives : 0; out->bytes_reduced = comm_owned ? comm_owned->bytes_reduced : 0;
  return PGS_OK;
}

// Two chains meeting in the middle do the same work as one, on two dependency chains instead of one: while one chain
// sits in its diagonal block (one CTA, ~33 us per panel) the other one's panel solve and trailing update have the GPU.
// Measured on BASELINE config 3 (front ~2700 rows): 327 -> 267 ms per factorisation; config 2 (thin front): 31 -> 51 LM it/s.
bool Solver::want_chains() const {
  if (is_inner || comm_owned || opt.linear_solver != PGS_SKYLINE_CHOLESKY) return false;
  if (opt.chains >= 2) return true;
  return opt.chains == 0 && N >= 4096;
}

int Solver::solve_dist(pgs_summary* sum, pgs_iteration* iters, int cap) {
  Comm* C = comm_owned.get();
  const int rank = C ? C->rank : 0, world = C ? C->world : 1;
  HostLap lap;
  if (int rc = sync_params_to_host()) return rc;
  const int Eo = (int)o_c1.size(), El = (int)l_a.size(), K = (int)r_node.size();
  const int PN = skyline_panel_width() / 6;
  int build_rc = PGS_OK;
  if (inner_dirty || !inner) {
    Partition P;
    make_partition(N, world, Eo, o_c1.data(), o_c2.data(), El, l_a.data(), l_b.data(), K, r_node.data(), &P, opt.chains);
    const int R = (int)P.ranges.size();
    if (world == 1 && (R < 2 || P.border.empty() || (opt.chains == 0 && (int)P.border.size() > N / 4))) {
      // nothing to gain: one natural-order chain on this handle (no edge crosses the cut, or the separator is a
      // large part of the graph — loop closures that span most of the trajectory)
      inner.reset(); plain_chain = true; inner_dirty = false;
      return PGS_PLAIN_CHAIN;
    }
    plain_chain = false;
    // local numbering: the interiors of this rank's chains in elimination order, each padded to a whole panel | local border nodes
    std::vector<int> glob2loc(N, -1);
    loc2glob.clear();
    std::vector<ChainSpec> specs;
    std::vector<int> my_border;   // global border positions this rank's chains touch
    for (int c = 0; c < R; ++c) if (P.ranges[c].rank == rank) {
      ChainSpec cs; cs.off = (int)loc2glob.size();
      const PlanRange& rg = P.ranges[c];
      if (!rg.down) { for (int i = rg.lo; i < rg.hi; ++i) if (P.node_chain[i] == c) { glob2loc[i] = (int)loc2glob.size(); loc2glob.push_back(i); } }
      else { for (int i = rg.hi - 1; i >= rg.lo; --i) if (P.node_chain[i] == c) { glob2loc[i] = (int)loc2glob.size(); loc2glob.push_back(i); } }
      while (loc2glob.size() % PN) loc2glob.push_back(-1);
      cs.len = (int)loc2glob.size() - cs.off;
      specs.push_back(cs);
      my_border.insert(my_border.end(), P.chain_border[c].begin(), P.chain_border[c].end());
    }
    std::sort(my_border.begin(), my_border.end());
    my_border.erase(std::unique(my_border.begin(), my_border.end()), my_border.end());
    const int n_int = (int)loc2glob.size();
    const int fb = n_int;
    for (int bp : my_border) { glob2loc[P.border[bp]] = (int)loc2glob.size(); loc2glob.push_back(P.border[bp]); }
    const int Nl = (int)loc2glob.size();
    { int k = 0; for (int c = 0; c < R; ++c) if (P.ranges[c].rank == rank) { for (int bp : P.chain_border[c]) specs[k].border.push_back(glob2loc[P.border[bp]]); ++k; } }
    // who counts (and publishes) a border node: the lowest rank that holds it
    std::vector<int> lowest(P.border.size(), world);
    for (int c = 0; c < R; ++c) for (int bp : P.chain_border[c]) lowest[bp] = std::min(lowest[bp], P.ranges[c].rank);
    // Ceres' reduced program drops parameter blocks that no residual block uses — of the WHOLE problem
    std::vector<char> used(N, 0);
    for (int e = 0; e < Eo; ++e) { used[o_c1[e]] = 1; used[o_c2[e]] = 1; }
    for (int e = 0; e < El; ++e) { used[l_a[e]] = 1; used[l_b[e]] = 1; }
    for (int k = 0; k < K; ++k) used[r_node[k]] = 1;

    pgs_options io = opt; io.chains = 1;
    inner.reset(new Solver(io));
    inner->is_inner = true;
    int rc = inner->init();
    if (rc == PGS_OK) {
      inner->comm = C; inner->first_border = fb; inner->chains = specs;
      inner->n_gborder = (int)P.border.size(); inner->gborder_env = P.border_env;
      inner->border_gpos = my_border;
      inner->border_counted.resize(my_border.size());
      for (size_t k = 0; k < my_border.size(); ++k) inner->border_counted[k] = lowest[my_border[k]] == rank;
      inner->forced_used.assign(Nl, 0);
      for (int l = fb; l < Nl; ++l) inner->forced_used[l] = used[loc2glob[l]];
      std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl);
      for (int l = 0; l < Nl; ++l) {
        const int g = loc2glob[l];
        if (g < 0) { q[4 * (size_t)l] = q[4 * (size_t)l + 1] = q[4 * (size_t)l + 2] = 0.0; q[4 * (size_t)l + 3] = 1.0; t[3 * (size_t)l] = t[3 * (size_t)l + 1] = t[3 * (size_t)l + 2] = 0.0; }
        else { std::memcpy(&q[4 * (size_t)l], &h_q[4 * (size_t)g], 32); std::memcpy(&t[3 * (size_t)l], &h_t[3 * (size_t)g], 24); }
      }
      rc = inner->set_nodes(Nl, q.data(), t.data(), false);
      for (int l = 0; l < Nl && rc == PGS_OK; ++l) { const int g = loc2glob[l]; if (g >= 0 && g < (int)h_node_const.size() && h_node_const[g]) rc = inner->set_constant(l, 1, 1); }
    }
    // owned residual blocks, re-indexed
    std::vector<int> c1, c2; std::vector<double> eq, et, ew;
    for (int e = 0; e < Eo && rc == PGS_OK; ++e) if (P.odom_owner[e] == rank) {
      if (glob2loc[o_c1[e]] < 0 || glob2loc[o_c2[e]] < 0) { rc = fail(PGS_ERR_STATE, "solve_dist: an owned edge has a node outside interior + border"); break; }
      c1.push_back(glob2loc[o_c1[e]]); c2.push_back(glob2loc[o_c2[e]]);
      eq.insert(eq.end(), &o_q[4 * (size_t)e], &o_q[4 * (size_t)e] + 4); et.insert(et.end(), &o_t[3 * (size_t)e], &o_t[3 * (size_t)e] + 3); ew.push_back(o_w[e]);
    }
    dstats.n_odom_owned = (int)c1.size();
    if (rc == PGS_OK && !c1.empty()) rc = inner->add_odom((int)c1.size(), c1.data(), c2.data(), eq.data(), et.data(), ew.data());
    c1.clear(); c2.clear(); eq.clear(); et.clear(); ew.clear(); loop2glob.clear();
    for (int e = 0; e < El && rc == PGS_OK; ++e) if (P.loop_owner[e] == rank) {
      if (glob2loc[l_a[e]] < 0 || glob2loc[l_b[e]] < 0) { rc = fail(PGS_ERR_STATE, "solve_dist: an owned loop edge has a node outside interior + border"); break; }
      c1.push_back(glob2loc[l_a[e]]); c2.push_back(glob2loc[l_b[e]]); loop2glob.push_back(e);
      eq.insert(eq.end(), &l_q[4 * (size_t)e], &l_q[4 * (size_t)e] + 4); et.insert(et.end(), &l_t[3 * (size_t)e], &l_t[3 * (size_t)e] + 3); ew.push_back(l_w[e]);
    }
    dstats.n_loop_owned = (int)c1.size();
    if (rc == PGS_OK && !c1.empty()) rc = inner->add_loop((int)c1.size(), c1.data(), c2.data(), eq.data(), et.data(), ew.data());
    std::vector<int> rn; eq.clear(); et.clear(); ew.clear();
    for (int k = 0; k < K && rc == PGS_OK; ++k) if (P.reg_owner[k] == rank) {
      rn.push_back(glob2loc[r_node[k]]);
      eq.insert(eq.end(), &r_q[4 * (size_t)k], &r_q[4 * (size_t)k] + 4); et.insert(et.end(), &r_t[3 * (size_t)k], &r_t[3 * (size_t)k] + 3); ew.push_back(r_w[k]);
    }
    dstats.n_reg_owned = (int)rn.size();
    if (rc == PGS_OK) rc = inner->set_regs((int)rn.size(), rn.data(), eq.data(), et.data(), ew.data());
    if (rc != PGS_OK && inner && !inner->err.empty()) err = "solve_dist: building the local problem failed: " + inner->err;
    dstats.n_interior_nodes = n_int; dstats.n_border_nodes = (int)P.border.size();
    dstats.n_chains = (int)specs.size(); dstats.n_local_border_nodes = (int)my_border.size();
    build_rc = rc;
    if (rc == PGS_OK) inner_dirty = false;
  } else {
    // same structure, new initial guesses
    const int Nl = (int)loc2glob.size();
    std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl);
    for (int l = 0; l < Nl; ++l) {
      const int g = loc2glob[l];
      if (g < 0) { q[4 * (size_t)l + 3] = 1.0; continue; }
      std::memcpy(&q[4 * (size_t)l], &h_q[4 * (size_t)g], 32); std::memcpy(&t[3 * (size_t)l], &h_t[3 * (size_t)g], 24);
    }
    build_rc = inner->update_nodes(0, Nl, q.data(), t.data());
    if (build_rc) err = inner->err;
  }
  if (build_rc == PGS_OK && !loop2glob.empty()) {
    std::vector<double> s(loop2glob.size());
    for (size_t l = 0; l < s.size(); ++l) s[l] = h_sw[loop2glob[l]];
    build_rc = inner->set_switches(0, (int)s.size(), s.data());
    if (build_rc) err = inner->err;
  }
  // a rank whose local problem could not be built must not leave the others waiting in the first collective
  if (C) { const int a = C->agree(build_rc, stream, &err); if (a) { inner.reset(); inner_dirty = true; return build_rc ? build_rc : a; } }
  else if (build_rc) { inner.reset(); inner_dirty = true; return build_rc; }
  if (C) { C->n_collectives = 0; C->bytes_reduced = 0; }
  lap.lap("plan + local problem");
  if (int rc = inner->solve(sum, iters, cap)) { err = inner->err; return rc; }
  lap.lap("inner solve (all of it)");
  if (sum && world == 1) sum->factor_flops = est_flops;   // natural-order estimate: two chains meeting in the middle do the same work

  // gather: every rank contributes its interior poses, the border poses it counts and its owned switches
  const int Nl = (int)loc2glob.size();
  std::vector<double> q(4 * (size_t)Nl), t(3 * (size_t)Nl), s(loop2glob.size());
  if (int rc = inner->get_poses(0, Nl, q.data(), t.data())) { err = inner->err; return rc; }
  if (!s.empty()) if (int rc = inner->get_switches(0, (int)s.size(), s.data())) { err = inner->err; return rc; }
  const int fb = inner->first_border >= 0 ? inner->first_border : Nl;
  if (world == 1) {
    for (int l = 0; l < Nl; ++l) {
      const int gi = loc2glob[l];
      if (gi < 0) continue;
      std::memcpy(&h_q[4 * (size_t)gi], &q[4 * (size_t)l], 32); std::memcpy(&h_t[3 * (size_t)gi], &t[3 * (size_t)l], 24);
    }
    for (size_t l = 0; l < s.size(); ++l) h_sw[loop2glob[l]] = s[l];
    host_params_newer = true; device_params_newer = false;
    lap.lap("read-back");
    return PGS_OK;
  }
  // summed into a zero-initialised global vector -> every rank ends up with the complete solution
  const size_t tot = 7 * (size_t)N + (size_t)El;
  std::vector<double> g(tot, 0.0);
  for (int l = 0; l < Nl; ++l) {
    const int gi = loc2glob[l];
    if (gi < 0 || (l >= fb && !inner->border_counted[l - fb])) continue;
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
