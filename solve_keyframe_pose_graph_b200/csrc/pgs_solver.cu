// Host driver: structure build, kernel launches, Ceres-faithful LM loop.  See pgs_solver.h.
#include "pgs_solver.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>

#include "pgs_kernels.cuh"
#include "pgs_pcg.cuh"
#include "pgs_skyline.h"

namespace pgs {

#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_fail(e__, #x); } while (0)


// ---- small pack/unpack kernels between the C-ABI's SoA (q[4N], t[3N], s[El] caller order) and the device layout
// the 8th slot of a pose record carries the constant-block flag (1.0 = constant), which the sweep turns into zero Jacobian columns
// (grid-stride: q and t may be pinned HOST memory read over the bus by a small grid, see evaluate_from_host)
__global__ void pack_pose_kernel(int first, int n, const double* __restrict__ q, const double* __restrict__ t, const char* __restrict__ fixed, double* __restrict__ pose) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 q01 = *reinterpret_cast<const double2*>(q + 4 * (size_t)i), q23 = *reinterpret_cast<const double2*>(q + 4 * (size_t)i + 2);
    const double t0 = t[3 * (size_t)i], t1 = t[3 * (size_t)i + 1], t2 = t[3 * (size_t)i + 2];
    double* p = pose + 8 * (size_t)(first + i);
    *reinterpret_cast<double2*>(p) = q01; *reinterpret_cast<double2*>(p + 2) = q23;
    *reinterpret_cast<double2*>(p + 4) = make_double2(t0, t1); *reinterpret_cast<double2*>(p + 6) = make_double2(t2, fixed[first + i] ? 1.0 : 0.0);
  }
}
__global__ void unpack_pose_kernel(int first, int n, const double* __restrict__ pose, double* __restrict__ q, double* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = pose + 8 * (size_t)(first + i);
  q[4 * (size_t)i] = p[0]; q[4 * (size_t)i + 1] = p[1]; q[4 * (size_t)i + 2] = p[2]; q[4 * (size_t)i + 3] = p[3];
  t[3 * (size_t)i] = p[4]; t[3 * (size_t)i + 1] = p[5]; t[3 * (size_t)i + 2] = p[6];
}
// sorted[e] = caller[perm[e]]
__global__ void gather_kernel(int n, const int* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) dst[e] = src[perm[e]];
}
// caller[perm[e]] = sorted[e]
__global__ void scatter_kernel(int n, const int* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) dst[perm[e]] = src[e];
}
__global__ void flush_kernel(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// reads n doubles; the store never happens (v is never NaN) but keeps the loads alive
__global__ void drain_kernel(double* p, size_t n) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i];
  if (s != s) p[0] = s;
}
__global__ void __launch_bounds__(256) stream_write_kernel(double2* p, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, make_double2(1.0, 2.0));
}

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

Solver::Solver(const pgs_options& o) : opt(o) {}

Solver::~Solver() {
  inner.reset();
  release_chains();
  if (ev_fork) cudaEventDestroy(ev_fork);
  for (cudaEvent_t e : ev_ph) if (e) cudaEventDestroy(e);
  if (sky) skyline_destroy(sky);
  if (h_scal) cudaFreeHost(h_scal);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ev_t0) cudaEventDestroy(ev_t0);
  if (ev_t1) cudaEventDestroy(ev_t1);
  for (cudaEvent_t e : ev_chunk) if (e) cudaEventDestroy(e);
  if (ev_copy_go) cudaEventDestroy(ev_copy_go);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (stream) cudaStreamDestroy(stream);
}

int Solver::cuda_fail(cudaError_t e, const char* what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  err = buf;
  return e == cudaErrorMemoryAllocation ? PGS_ERR_OUT_OF_MEMORY : PGS_ERR_CUDA;
}

int Solver::init() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(PGS_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e) + " (libpgs has no CPU fallback)");
  if (opt.device < 0 || opt.device >= count) return fail(PGS_ERR_INVALID_ARGUMENT, "pgs_options.device out of range");
  dev = opt.device;
  CU(cudaSetDevice(dev));
  CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&ev0)); CU(cudaEventCreate(&ev1)); CU(cudaEventCreate(&ev_t0)); CU(cudaEventCreate(&ev_t1));
  CU(cudaMallocHost((void**)&h_scal, sizeof(double) * L_NSCAL));
  CU(d_scal.resize(L_NSCAL, true));
  CU(d_partial.resize(4 * MAX_GRID, true));
  CU(d_counter.resize(4, true));
  int nsm = 148, occ = 1;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel<0>, 256, 0);
  sweep_grid = std::min(MAX_GRID, nsm * std::max(occ, 1));
  return PGS_OK;
}

void Solver::tic() { cudaEventRecord(ev0, stream); }
double Solver::toc() { cudaEventRecord(ev1, stream); cudaEventSynchronize(ev1); float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); return ms; }

// ------------------------------------------------------------------------------------------------ problem construction
int Solver::set_nodes(int n, const double* q, const double* t, bool append) {
  if (n < 0 || (n > 0 && (!q || !t))) return fail(PGS_ERR_INVALID_ARGUMENT, "set_nodes: null input");
  if (!append && n < N) {
    // replacing the node set by a smaller one must not leave residual blocks pointing past its end
    auto past = [n](const std::vector<int>& v) { for (int x : v) if (x >= n) return true; return false; };
    if (past(o_c1) || past(o_c2) || past(l_a) || past(l_b) || past(r_node))
      return fail(PGS_ERR_STATE, "set_nodes: existing residual blocks reference nodes beyond the new node count");
  }
  if (int rc = sync_params_to_host()) return rc;
  if (!append) { h_q.clear(); h_t.clear(); h_node_const.clear(); }
  h_q.insert(h_q.end(), q, q + 4 * (size_t)n);
  h_t.insert(h_t.end(), t, t + 3 * (size_t)n);
  N = (int)(h_t.size() / 3);
  structure_dirty = true; inner_dirty = true; host_params_newer = true;
  return PGS_OK;
}
int Solver::update_nodes(int first, int n, const double* q, const double* t) {
  if (first < 0 || n < 0 || (int64_t)first + n > N) return fail(PGS_ERR_INVALID_ARGUMENT, "update_nodes: range out of bounds");
  if (n > 0 && (!q || !t)) return fail(PGS_ERR_INVALID_ARGUMENT, "update_nodes: null input");
  if (int rc = sync_params_to_host()) return rc;
  std::memcpy(&h_q[4 * (size_t)first], q, sizeof(double) * 4 * n);
  std::memcpy(&h_t[3 * (size_t)first], t, sizeof(double) * 3 * n);
  host_params_newer = true;
  return PGS_OK;
}
int Solver::get_poses(int first, int n, double* q, double* t) {
  if (first < 0 || n < 0 || (int64_t)first + n > N) return fail(PGS_ERR_INVALID_ARGUMENT, "get_poses: range out of bounds");
  if (int rc = sync_params_to_host()) return rc;
  if (q) std::memcpy(q, &h_q[4 * (size_t)first], sizeof(double) * 4 * n);
  if (t) std::memcpy(t, &h_t[3 * (size_t)first], sizeof(double) * 3 * n);
  return PGS_OK;
}
int Solver::set_constant(int first, int n, int constant) {
  if (first < 0 || n < 0 || (int64_t)first + n > N) return fail(PGS_ERR_INVALID_ARGUMENT, "set_constant_nodes: range out of bounds");
  if (int rc = sync_params_to_host()) return rc;
  if ((int)h_node_const.size() < N) h_node_const.resize(N, 0);
  for (int i = first; i < first + n; ++i) h_node_const[i] = constant ? 1 : 0;
  structure_dirty = true; inner_dirty = true; host_params_newer = true;   // node_used and the pose records' flag slot are rebuilt
  return PGS_OK;
}
int Solver::set_switches(int first, int n, const double* s) {
  if (first < 0 || n < 0 || (int64_t)first + n > (int64_t)h_sw.size()) return fail(PGS_ERR_INVALID_ARGUMENT, "set_switches: range out of bounds");
  if (n > 0 && !s) return fail(PGS_ERR_INVALID_ARGUMENT, "set_switches: null input");
  if (int rc = sync_params_to_host()) return rc;
  std::memcpy(&h_sw[first], s, sizeof(double) * n);
  host_params_newer = true;
  return PGS_OK;
}
int Solver::get_switches(int first, int n, double* s) {
  if (first < 0 || n < 0 || (int64_t)first + n > (int64_t)h_sw.size()) return fail(PGS_ERR_INVALID_ARGUMENT, "get_switches: range out of bounds");
  if (n > 0 && !s) return fail(PGS_ERR_INVALID_ARGUMENT, "get_switches: null output");
  if (int rc = sync_params_to_host()) return rc;
  std::memcpy(s, &h_sw[first], sizeof(double) * n);
  return PGS_OK;
}
int Solver::add_odom(int m, const int* c1, const int* c2, const double* q, const double* t, const double* w) {
  if (m < 0 || (m > 0 && (!c1 || !c2 || !q || !t || !w))) return fail(PGS_ERR_INVALID_ARGUMENT, "add_odom_edges: null input");
  for (int i = 0; i < m; ++i)
    if (c1[i] < 0 || c2[i] < 0 || c1[i] >= N || c2[i] >= N || c1[i] == c2[i]) return fail(PGS_ERR_INVALID_ARGUMENT, "add_odom_edges: node index out of range or c1 == c2");
  o_c1.insert(o_c1.end(), c1, c1 + m); o_c2.insert(o_c2.end(), c2, c2 + m);
  o_q.insert(o_q.end(), q, q + 4 * (size_t)m); o_t.insert(o_t.end(), t, t + 3 * (size_t)m); o_w.insert(o_w.end(), w, w + m);
  structure_dirty = true; inner_dirty = true;
  return PGS_OK;
}
int Solver::add_loop(int m, const int* a, const int* b, const double* q, const double* t, const double* w) {
  if (m < 0 || (m > 0 && (!a || !b || !q || !t))) return fail(PGS_ERR_INVALID_ARGUMENT, "add_loop_edges: null input");
  for (int i = 0; i < m; ++i)
    if (a[i] < 0 || b[i] < 0 || a[i] >= N || b[i] >= N || a[i] == b[i]) return fail(PGS_ERR_INVALID_ARGUMENT, "add_loop_edges: node index out of range or a == b");
  if (int rc = sync_params_to_host()) return rc;
  l_a.insert(l_a.end(), a, a + m); l_b.insert(l_b.end(), b, b + m);
  l_q.insert(l_q.end(), q, q + 4 * (size_t)m); l_t.insert(l_t.end(), t, t + 3 * (size_t)m);
  for (int i = 0; i < m; ++i) { l_w.push_back(w ? w[i] : 1.0); h_sw.push_back(opt.switch_init); }
  structure_dirty = true; inner_dirty = true; host_params_newer = true;
  return PGS_OK;
}
int Solver::set_regs(int k, const int* node, const double* q, const double* t, const double* w) {
  if (k < 0 || (k > 0 && (!node || !q || !t || !w))) return fail(PGS_ERR_INVALID_ARGUMENT, "set_regularizers: null input");
  for (int i = 0; i < k; ++i) if (node[i] < 0 || node[i] >= N) return fail(PGS_ERR_INVALID_ARGUMENT, "set_regularizers: node index out of range");
  r_node.assign(node, node + k); r_q.assign(q, q + 4 * (size_t)k); r_t.assign(t, t + 3 * (size_t)k); r_w.assign(w, w + k);
  structure_dirty = true; inner_dirty = true;   // incidence lists contain the regulariser blocks
  return PGS_OK;
}
void Solver::sizes(pgs_sizes* s) {
  if (structure_dirty && !comm_owned) finalize();   // a multi-GPU handle never builds the full problem on one device
  s->n_nodes = N; s->n_odom = (int)o_c1.size(); s->n_loop = (int)l_a.size(); s->n_reg = (int)r_node.size(); s->n_pairs = n_pairs;
}
int64_t Solver::sweep_bytes() const {
  const int64_t Eo = (int64_t)o_c1.size(), El = (int64_t)l_a.size(), K = (int64_t)r_node.size();
  return 56 * (int64_t)N + 72 * Eo + 72 * El + 68 * K + 624 * Eo + 784 * El + 336 * K;
}

// ------------------------------------------------------------------------------------------------ structure
int Solver::finalize() {
  if (!structure_dirty) return PGS_OK;
  CU(cudaSetDevice(dev));
  const int Eo = (int)o_c1.size(), El = (int)l_a.size(), K = (int)r_node.size();
  // 1. sort edges by (c1, c2): a warp of 32 consecutive edges then touches a short window of nodes
  perm_o.resize(Eo); std::iota(perm_o.begin(), perm_o.end(), 0);
  std::stable_sort(perm_o.begin(), perm_o.end(), [&](int x, int y) { return o_c1[x] != o_c1[y] ? o_c1[x] < o_c1[y] : o_c2[x] < o_c2[y]; });
  perm_l.resize(El); std::iota(perm_l.begin(), perm_l.end(), 0);
  // loop edge (a,b) binds parameters (c1,c2) = (b,a)  [reference PoseGraphSLAM.cpp:1553-1554]
  std::stable_sort(perm_l.begin(), perm_l.end(), [&](int x, int y) { return l_b[x] != l_b[y] ? l_b[x] < l_b[y] : l_a[x] < l_a[y]; });
  inv_perm_l.resize(El);
  for (int e = 0; e < El; ++e) inv_perm_l[perm_l[e]] = e;

  const int To = cdiv(Eo, TILE), Tl = cdiv(El, TILE);
  std::vector<int2> oidx(std::max(Eo, 1)), lidx(std::max(El, 1));
  std::vector<double> oobs((size_t)To * OBS * TILE, 0.0), lobs((size_t)Tl * OBS * TILE, 0.0);
  for (int e = 0; e < Eo; ++e) {
    const int o = perm_o[e];
    oidx[e] = make_int2(o_c1[o], o_c2[o]);
    double* ob = &oobs[(size_t)(e / TILE) * OBS * TILE + (e % TILE)];
    for (int k = 0; k < 4; ++k) ob[k * TILE] = o_q[4 * (size_t)o + k];
    for (int k = 0; k < 3; ++k) ob[(4 + k) * TILE] = o_t[3 * (size_t)o + k];
    ob[7 * TILE] = o_w[o];
  }
  for (int e = 0; e < El; ++e) {
    const int o = perm_l[e];
    lidx[e] = make_int2(l_b[o], l_a[o]);
    double* ob = &lobs[(size_t)(e / TILE) * OBS * TILE + (e % TILE)];
    for (int k = 0; k < 4; ++k) ob[k * TILE] = l_q[4 * (size_t)o + k];
    for (int k = 0; k < 3; ++k) ob[(4 + k) * TILE] = l_t[3 * (size_t)o + k];
    ob[7 * TILE] = l_w[o];
  }
  std::vector<double> ranchor((size_t)std::max(K, 1) * 8, 0.0);
  for (int k = 0; k < K; ++k) {
    for (int c = 0; c < 4; ++c) ranchor[8 * (size_t)k + c] = r_q[4 * (size_t)k + c];
    for (int c = 0; c < 3; ++c) ranchor[8 * (size_t)k + 4 + c] = r_t[3 * (size_t)k + c];
    ranchor[8 * (size_t)k + 7] = r_w[k];
  }
  // per tile of 32 edges: the largest keyframe index any tile up to and including it touches (end-to-end step: a prefix
  // of the tiles can be swept as soon as the keyframes below a bound have arrived)
  o_pm.assign(To, 0); l_pm.assign(Tl, 0); r_pm.assign(cdiv(K, TILE), 0);
  for (int e = 0; e < Eo; ++e) o_pm[e / TILE] = std::max(o_pm[e / TILE], std::max(oidx[e].x, oidx[e].y));
  for (int e = 0; e < El; ++e) l_pm[e / TILE] = std::max(l_pm[e / TILE], std::max(lidx[e].x, lidx[e].y));
  for (int k = 0; k < K; ++k) r_pm[k / TILE] = std::max(r_pm[k / TILE], r_node[k]);
  for (size_t i = 1; i < o_pm.size(); ++i) o_pm[i] = std::max(o_pm[i], o_pm[i - 1]);
  for (size_t i = 1; i < l_pm.size(); ++i) l_pm[i] = std::max(l_pm[i], l_pm[i - 1]);
  for (size_t i = 1; i < r_pm.size(); ++i) r_pm[i] = std::max(r_pm[i], r_pm[i - 1]);
  // 2. node incidence lists (edge<<3 | kind<<1 | side)
  std::vector<int> inc_ptr(N + 1, 0);
  for (int e = 0; e < Eo; ++e) { ++inc_ptr[oidx[e].x + 1]; ++inc_ptr[oidx[e].y + 1]; }
  for (int e = 0; e < El; ++e) { ++inc_ptr[lidx[e].x + 1]; ++inc_ptr[lidx[e].y + 1]; }
  for (int k = 0; k < K; ++k) ++inc_ptr[r_node[k] + 1];
  for (int i = 0; i < N; ++i) inc_ptr[i + 1] += inc_ptr[i];
  std::vector<int> inc_item(std::max(inc_ptr[N], 1)), cur(inc_ptr.begin(), inc_ptr.end() - 1);
  for (int e = 0; e < Eo; ++e) { inc_item[cur[oidx[e].x]++] = (e << 3) | 0; inc_item[cur[oidx[e].y]++] = (e << 3) | 1; }
  for (int e = 0; e < El; ++e) { inc_item[cur[lidx[e].x]++] = (e << 3) | 2; inc_item[cur[lidx[e].y]++] = (e << 3) | 3; }
  for (int k = 0; k < K; ++k) inc_item[cur[r_node[k]]++] = (k << 3) | 4;
  h_node_used.assign(std::max(N, 1), 0);
  h_node_const.resize(std::max(N, 1), 0);
  // a block is updated (and counted in the step / gradient norms) if some residual block uses it and it is not constant:
  // Ceres' reduced program drops unused and constant parameter blocks alike
  for (int i = 0; i < N; ++i) h_node_used[i] = (inc_ptr[i + 1] > inc_ptr[i] || (i < (int)forced_used.size() && forced_used[i])) && !h_node_const[i];
  {
    std::vector<int> fo, fr;
    for (int e = 0; e < Eo; ++e) if (h_node_const[oidx[e].x] && h_node_const[oidx[e].y]) fo.push_back(e);
    for (int k = 0; k < K; ++k) if (h_node_const[r_node[k]]) fr.push_back(k);
    n_fixed_o = (int)fo.size(); n_fixed_r = (int)fr.size();
    if (fo.empty()) fo.push_back(0);
    if (fr.empty()) fr.push_back(0);
    CU(d_fixed_o.upload(fo, stream)); CU(d_fixed_r.upload(fr, stream));
  }
  // 3. distinct node pairs (hi, lo) and their edges (edge<<2 | kind<<1 | c1_is_hi)
  std::vector<std::pair<uint64_t, int>> keyed; keyed.reserve((size_t)Eo + El);
  auto key_of = [](int c1, int c2) { const uint64_t hi = (uint64_t)std::max(c1, c2), lo = (uint64_t)std::min(c1, c2); return (hi << 32) | lo; };
  for (int e = 0; e < Eo; ++e) keyed.push_back({key_of(oidx[e].x, oidx[e].y), (e << 2) | 0 | (oidx[e].x > oidx[e].y ? 1 : 0)});
  for (int e = 0; e < El; ++e) keyed.push_back({key_of(lidx[e].x, lidx[e].y), (e << 2) | 2 | (lidx[e].x > lidx[e].y ? 1 : 0)});
  std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a.first < b.first; });
  h_pair_hi.clear(); h_pair_lo.clear();
  std::vector<int> pe_ptr(1, 0), pe_item(std::max(keyed.size(), (size_t)1));
  for (size_t i = 0; i < keyed.size(); ++i) {
    if (i == 0 || keyed[i].first != keyed[i - 1].first) {
      if (i) pe_ptr.push_back((int)i);
      h_pair_hi.push_back((int)(keyed[i].first >> 32)); h_pair_lo.push_back((int)(keyed[i].first & 0xffffffffu));
    }
    pe_item[i] = keyed[i].second;
  }
  pe_ptr.push_back((int)keyed.size());
  n_pairs = (int)h_pair_hi.size();
  if (keyed.empty()) { pe_ptr.assign(2, 0); }
  std::vector<int2> pairs(std::max(n_pairs, 1));
  for (int p = 0; p < n_pairs; ++p) pairs[p] = make_int2(h_pair_hi[p], h_pair_lo[p]);
  // 4. node adjacency over pairs for the SpMV (pair<<1 | node_is_hi)
  std::vector<int> adj_ptr(N + 1, 0);
  for (int p = 0; p < n_pairs; ++p) { ++adj_ptr[h_pair_hi[p] + 1]; ++adj_ptr[h_pair_lo[p] + 1]; }
  for (int i = 0; i < N; ++i) adj_ptr[i + 1] += adj_ptr[i];
  std::vector<int> adj_item(std::max(adj_ptr[N], 1)), cur2(adj_ptr.begin(), adj_ptr.end() - 1);
  for (int p = 0; p < n_pairs; ++p) { adj_item[cur2[h_pair_hi[p]]++] = (p << 1) | 1; adj_item[cur2[h_pair_lo[p]]++] = (p << 1) | 0; }

  // 5. upload
  CU(d_oidx.upload(oidx, stream)); CU(d_lidx.upload(lidx, stream)); CU(d_oobs.upload(oobs, stream)); CU(d_lobs.upload(lobs, stream));
  CU(d_ranchor.upload(ranchor, stream));
  { std::vector<int> rn(r_node); if (rn.empty()) rn.push_back(0); CU(d_rnode.upload(rn, stream)); }
  { std::vector<int> po(perm_o), pl(perm_l); if (po.empty()) po.push_back(0); if (pl.empty()) pl.push_back(0); CU(d_perm_o.upload(po, stream)); CU(d_perm_l.upload(pl, stream)); }
  CU(d_inc_ptr.upload(inc_ptr, stream)); CU(d_inc_item.upload(inc_item, stream));
  CU(d_pe_ptr.upload(pe_ptr, stream)); CU(d_pe_item.upload(pe_item, stream)); CU(d_pair.upload(pairs, stream));
  CU(d_adj_ptr.upload(adj_ptr, stream)); CU(d_adj_item.upload(adj_item, stream));
  CU(d_node_used.upload(h_node_used, stream)); CU(d_node_const.upload(h_node_const, stream));
  // 6. outputs and work arrays
  CU(d_pose.resize((size_t)N * 8, true)); CU(d_cpose.resize((size_t)N * 8, true));
  CU(d_sw.resize(std::max(El, 1), true)); CU(d_csw.resize(std::max(El, 1), true));
  CU(d_stage_q.resize((size_t)std::max(N, 1) * 4)); CU(d_stage_t.resize((size_t)std::max(N, 1) * 3)); CU(d_stage_s.resize(std::max(El, 1)));
  CU(d_or.resize((size_t)std::max(To, 1) * OD_R * TILE, true)); CU(d_oJ.resize((size_t)std::max(To, 1) * OD_J * TILE, true));
  CU(d_lr.resize((size_t)std::max(Tl, 1) * LP_R * TILE, true)); CU(d_lJ.resize((size_t)std::max(Tl, 1) * LP_J * TILE, true));
  CU(d_gr.resize((size_t)std::max(K, 1) * 6, true)); CU(d_gJ.resize((size_t)std::max(K, 1) * 36, true));
  CU(d_cost_tile.resize((size_t)To + Tl + cdiv(K, TILE) + 1, true));
  CU(d_counter.resize(4, true));   // tile scheduler of the sweep: must be zero between launches
  CU(d_Hd.resize((size_t)std::max(N, 1) * 36, true)); CU(d_g.resize((size_t)std::max(N, 1) * 6, true)); CU(d_Ho.resize((size_t)std::max(n_pairs, 1) * 36, true));
  CU(d_lv.resize((size_t)std::max(El, 1) * 12, true)); CU(d_lh.resize(std::max(El, 1), true)); CU(d_lg.resize(std::max(El, 1), true));
  CU(d_lvt.resize((size_t)std::max(El, 1) * 12, true)); CU(d_lw.resize(std::max(El, 1), true)); CU(d_lgt.resize(std::max(El, 1), true));
  CU(d_scale_p.resize((size_t)std::max(N, 1) * 6, true)); CU(d_scale_s.resize(std::max(El, 1), true));
  CU(d_diag_p.resize((size_t)std::max(N, 1) * 6, true)); CU(d_diag_s.resize(std::max(El, 1), true));
  CU(d_Ad.resize((size_t)std::max(N, 1) * 36, true)); CU(d_Ao.resize((size_t)std::max(n_pairs, 1) * 36, true)); CU(d_b.resize((size_t)std::max(N, 1) * 6, true));
  CU(d_y.resize((size_t)std::max(N, 1) * 6, true)); CU(d_dp.resize((size_t)std::max(N, 1) * 6, true)); CU(d_ds.resize(std::max(El, 1), true));
  if (sky) { skyline_destroy(sky); sky = nullptr; }
  release_chains();
  flag_ptrs_ready = false;
  {
    // which nodes this rank counts in the step / gradient / x norms: everything but the border nodes another rank counts
    std::vector<char> counted(std::max(N, 1), 1);
    const int fb = first_border >= 0 ? first_border : N;
    for (int i = fb; i < N; ++i) counted[i] = (i - fb) < (int)border_counted.size() ? border_counted[i - fb] : 1;
    CU(d_node_counted.upload(counted, stream));
    std::vector<int> gp(border_gpos); if (gp.empty()) gp.push_back(0);
    CU(d_border_gpos.upload(gp, stream));
  }
  CU(cudaStreamSynchronize(stream));
  structure_dirty = false; host_params_newer = true; device_params_newer = false;
  return PGS_OK;
}

int Solver::sync_params_to_device() {
  if (int rc = finalize()) return rc;
  if (!host_params_newer) return PGS_OK;
  const int El = (int)l_a.size();
  if (N) {
    CU(cudaMemcpyAsync(d_stage_q.p, h_q.data(), sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_stage_t.p, h_t.data(), sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, stream));
    pack_pose_kernel<<<cdiv(N, 256), 256, 0, stream>>>(0, N, d_stage_q.p, d_stage_t.p, d_node_const.p, d_pose.p);
  }
  if (El) {
    CU(cudaMemcpyAsync(d_stage_s.p, h_sw.data(), sizeof(double) * (size_t)El, cudaMemcpyHostToDevice, stream));
    gather_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_stage_s.p, d_sw.p);
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(stream));
  host_params_newer = false; device_params_newer = false;
  return PGS_OK;
}

int Solver::sync_params_to_host() {
  if (!device_params_newer) return PGS_OK;
  CU(cudaSetDevice(dev));
  const int El = (int)l_a.size();
  if (N) {
    unpack_pose_kernel<<<cdiv(N, 256), 256, 0, stream>>>(0, N, d_pose.p, d_stage_q.p, d_stage_t.p);
    CU(cudaMemcpyAsync(h_q.data(), d_stage_q.p, sizeof(double) * 4 * (size_t)N, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(h_t.data(), d_stage_t.p, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToHost, stream));
  }
  if (El) {
    scatter_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_sw.p, d_stage_s.p);
    CU(cudaMemcpyAsync(h_sw.data(), d_stage_s.p, sizeof(double) * (size_t)El, cudaMemcpyDeviceToHost, stream));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(stream));
  device_params_newer = false;
  return PGS_OK;
}

int Solver::read_scalars(int n) {
  CU(cudaMemcpyAsync(h_scal, d_scal.p, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  return PGS_OK;
}

// ------------------------------------------------------------------------------------------------ kernels
int Solver::launch_sweep(int mode, const double* pose, const double* sw, double* cost_out_dev, cudaEvent_t after_kernel, const int* ranges, int reduce) {
  SweepArgs A;
  A.pose = pose; A.sw = sw;
  A.o_idx = d_oidx.p; A.o_obs = d_oobs.p; A.n_odom = (int)o_c1.size();
  A.l_idx = d_lidx.p; A.l_obs = d_lobs.p; A.n_loop = (int)l_a.size();
  A.r_node = d_rnode.p; A.r_anchor = d_ranchor.p; A.n_reg = (int)r_node.size();
  A.o_r = d_or.p; A.o_J = d_oJ.p; A.l_r = d_lr.p; A.l_J = d_lJ.p; A.g_r = d_gr.p; A.g_J = d_gJ.p;
  A.cost_tile = d_cost_tile.p; A.sched = d_counter.p + 2;   // slots 0-1 belong to the PCG kernels
  A.o_t0 = 0; A.o_t1 = cdiv(A.n_odom, TILE); A.l_t0 = 0; A.l_t1 = cdiv(A.n_loop, TILE); A.r_t0 = 0; A.r_t1 = cdiv(A.n_reg, TILE); A.reduce = reduce;
  if (ranges) { A.o_t0 = ranges[0]; A.o_t1 = ranges[1]; A.l_t0 = ranges[2]; A.l_t1 = ranges[3]; A.r_t0 = ranges[4]; A.r_t1 = ranges[5]; }
  const int tiles = (A.o_t1 - A.o_t0) + (A.l_t1 - A.l_t0) + (A.r_t1 - A.r_t0);
  if (tiles == 0 && !reduce) return PGS_OK;
  const int grid = std::max(1, std::min(sweep_grid, cdiv(tiles, 8)));
  A.cost_out = cost_out_dev;
  if (mode == 0) sweep_kernel<0><<<grid, 256, 0, stream>>>(A); else sweep_kernel<1><<<grid, 256, 0, stream>>>(A);
  if (after_kernel) cudaEventRecord(after_kernel, stream);
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::run_assemble() {
  AsmArgs A;
  A.N = N; A.n_odom = (int)o_c1.size(); A.n_loop = (int)l_a.size(); A.n_reg = (int)r_node.size(); A.n_pairs = n_pairs;
  A.o_r = d_or.p; A.o_J = d_oJ.p; A.l_r = d_lr.p; A.l_J = d_lJ.p; A.g_r = d_gr.p; A.g_J = d_gJ.p;
  A.inc_ptr = d_inc_ptr.p; A.inc_item = d_inc_item.p; A.pe_ptr = d_pe_ptr.p; A.pe_item = d_pe_item.p;
  A.Hd = d_Hd.p; A.g = d_g.p; A.Ho = d_Ho.p; A.lv = d_lv.p; A.lh = d_lh.p; A.lg = d_lg.p;
  if (N) assemble_diag_kernel<<<cdiv(N, 128), 128, 0, stream>>>(A);
  if (n_pairs) assemble_offdiag_kernel<<<cdiv(n_pairs, 128), 128, 0, stream>>>(A);
  if (A.n_loop) assemble_switch_kernel<<<cdiv(A.n_loop, 128), 128, 0, stream>>>(A);
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::compute_scaling(bool compute_scale) {
  const int El = (int)l_a.size();
  const int n = std::max(6 * N, El);
  if (n) scaling_kernel<<<cdiv(n, 256), 256, 0, stream>>>(N, first_border >= 0 ? first_border : N, El, d_Hd.p, d_lh.p, compute_scale ? 1 : 0, opt.jacobi_scaling, opt.min_lm_diagonal,
                                                         opt.max_lm_diagonal, d_scale_p.p, d_scale_s.p, d_diag_p.p, d_diag_s.p);
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::build_system(double radius) {
  SysArgs A;
  A.N = N; A.n_loop = (int)l_a.size(); A.n_pairs = n_pairs; A.inv_radius = 1.0 / radius;
  A.first_border = first_border >= 0 ? first_border : N;
  A.l_idx = d_lidx.p; A.Hd = d_Hd.p; A.g = d_g.p; A.Ho = d_Ho.p; A.lv = d_lv.p; A.lh = d_lh.p; A.lg = d_lg.p;
  A.scale_p = d_scale_p.p; A.scale_s = d_scale_s.p; A.diag_p = d_diag_p.p; A.diag_s = d_diag_s.p;
  A.inc_ptr = d_inc_ptr.p; A.inc_item = d_inc_item.p; A.pe_ptr = d_pe_ptr.p; A.pe_item = d_pe_item.p; A.pair = d_pair.p;
  A.lvt = d_lvt.p; A.lw = d_lw.p; A.lgt = d_lgt.p; A.Ad = d_Ad.p; A.Ao = d_Ao.p; A.b = d_b.p;
  if (A.n_loop) system_switch_kernel<<<cdiv(A.n_loop, 128), 128, 0, stream>>>(A);
  if (N) system_diag_kernel<<<cdiv(N, 128), 128, 0, stream>>>(A);
  if (n_pairs) system_offdiag_kernel<<<cdiv(n_pairs, 128), 128, 0, stream>>>(A);
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::solve_pcg(int* iters) {
  const size_t n6 = (size_t)N * 6;
  CU(d_Minv.resize((size_t)N * 36)); CU(d_px.resize(n6)); CU(d_pr.resize(n6)); CU(d_prn.resize(n6)); CU(d_pz.resize(n6)); CU(d_pp.resize(n6)); CU(d_pAp.resize(n6));
  PcgArgs A;
  A.N = N; A.n_pairs = n_pairs; A.Ad = d_Ad.p; A.Ao = d_Ao.p; A.Minv = d_Minv.p; A.pair = d_pair.p; A.adj_ptr = d_adj_ptr.p; A.adj_item = d_adj_item.p;
  A.b = d_b.p; A.x = d_y.p; A.r = d_pr.p; A.rn = d_prn.p; A.z = d_pz.p; A.p = d_pp.p; A.Ap = d_pAp.p;
  A.partial = d_partial.p; A.scal = d_scal.p; A.counter = d_counter.p;
  const int grid = cdiv(6LL * N, 192);
  if (grid > MAX_GRID * 2) { CU(d_partial.resize((size_t)grid * 2 + 16, true)); A.partial = d_partial.p; }
  block_inverse_kernel<<<cdiv(N, 128), 128, 0, stream>>>(N, d_Ad.p, d_Minv.p);
  pcg_init_kernel<<<grid, 192, 0, stream>>>(A);
  int old_slot = S_RZ0, new_slot = S_RZ1, it = 0;
  const int check_every = 8;
  const double tol2 = opt.pcg_tolerance * opt.pcg_tolerance;
  while (it < opt.pcg_max_iterations) {
    for (int k = 0; k < check_every && it < opt.pcg_max_iterations; ++k, ++it) {
      pcg_spmv_kernel<<<grid, 192, 0, stream>>>(A);
      pcg_update_kernel<<<grid, 192, 0, stream>>>(A, old_slot, new_slot);
      std::swap(A.r, A.rn);
      pcg_direction_kernel<<<cdiv(6LL * N, 256), 256, 0, stream>>>(A, old_slot, new_slot);
      std::swap(old_slot, new_slot);
    }
    CU(cudaGetLastError());
    if (int rc = read_scalars(S_NSLOTS)) return rc;
    const double rr = h_scal[S_RR], bb = h_scal[S_BB];
    if (!(rr == rr)) return fail(PGS_ERR_LINEAR_SOLVER, "PCG produced NaN");
    if (rr <= tol2 * bb || bb == 0.0) break;
  }
  if (iters) *iters = it;
  return PGS_OK;
}

int Solver::prepare_linear() {
  if (use_pcg()) return PGS_OK;
  if (!chains.empty()) return prepare_chains();
  if (!sky) {
    sky = skyline_create(N, n_pairs, h_pair_hi.data(), h_pair_lo.data(), stream, &err);
    if (!sky) return PGS_ERR_OUT_OF_MEMORY;
    factor_nnz = skyline_nnz(sky);
  }
  return PGS_OK;
}

int Solver::solve_skyline() {
  if (int rc = prepare_linear()) return rc;
  if (!chains.empty()) return solve_chains();
  return skyline_factor_solve(sky, d_Ad.p, d_Ao.p, d_b.p, d_y.p, &err);
}

int Solver::solve_linear(int* iters) {
  if (iters) *iters = 0;
  if (use_pcg()) return solve_pcg(iters);
  return solve_skyline();
}

int Solver::get_backward_errors(double* out, int cap, int* n) {
  const std::vector<double>& v = inner ? inner->backward_error : backward_error;
  if (n) *n = (int)v.size();
  for (int i = 0; out && i < cap && i < (int)v.size(); ++i) out[i] = v[i];
  return PGS_OK;
}

// Envelope size and factorisation cost of the natural-order skyline, straight from the edge lists (host only):
// row i spans [min neighbour, i]; nnz = sum of the row widths; flops = sum over the columns of (rows that reach the
// column)^2 — the rank-1 update every eliminated column applies to the rows below it.
void Solver::estimate_skyline(double* bytes, double* flops) const {
  std::vector<int> nstart(std::max(N, 1));
  for (int i = 0; i < N; ++i) nstart[i] = i;
  auto edge = [&](int i, int j) { const int hi = std::max(i, j), lo = std::min(i, j); nstart[hi] = std::min(nstart[hi], lo); };
  for (size_t e = 0; e < o_c1.size(); ++e) edge(o_c1[e], o_c2[e]);
  for (size_t e = 0; e < l_a.size(); ++e) edge(l_a[e], l_b[e]);
  const int PW = skyline_panel_width(), PN = PW / 6;
  const int D = (6 * N + PW - 1) / PW;
  std::vector<int> diff(D + 1, 0);                 // rows (in nodes) that reach panel d from below
  double nnz = 0.0;
  for (int i = 0; i < N; ++i) {
    const int d0 = nstart[i] / PN, d1 = i / PN;
    nnz += 6.0 * (double)((d1 + 1 - d0) * PW);
    if (d0 < d1) { ++diff[d0]; --diff[d1]; }
  }
  double fl = 0.0; int reach = 0;
  for (int d = 0; d < D; ++d) { reach += diff[d]; const double R = 6.0 * reach + PW / 2.0; fl += PW * R * R; }
  *bytes = nnz * sizeof(double); *flops = fl;
}

// The skyline factor is dense inside the row envelope: loop closures that reach far back make it large.  Past the
// memory or flop budget the iterative solver takes over for this problem (pgs_summary.linear_solver_used says so).
int Solver::choose_linear_solver() {
  pcg_fallback = false; est_flops = 0.0;
  if (opt.linear_solver != PGS_SKYLINE_CHOLESKY || is_inner) return PGS_OK;
  double bytes = 0.0;
  estimate_skyline(&bytes, &est_flops);
  if (comm_owned) return PGS_OK;   // sharded across GPUs: every rank holds a part; no fallback there
  double limit = opt.max_factor_bytes;
  if (limit <= 0.0) { size_t fr = 0, tot = 0; CU(cudaMemGetInfo(&fr, &tot)); limit = 0.8 * (double)fr; }
  if (bytes > limit || (opt.max_factor_flops > 0.0 && est_flops > opt.max_factor_flops)) pcg_fallback = true;
  return PGS_OK;
}

// r = b - A y of the reduced block system (Ad, Ao) and den = |A| |y| + |b|, row by row; one thread per scalar row
__global__ void lin_residual_kernel(int N, const double* __restrict__ Ad, const double* __restrict__ Ao, const int2* __restrict__ pair,
                                    const int* __restrict__ adj_ptr, const int* __restrict__ adj_item, const double* __restrict__ b,
                                    const double* __restrict__ y, double* __restrict__ r, double* __restrict__ den) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * N) return;
  const int i = t / 6, a = t % 6;
  double s = b[t], m = fabs(b[t]);
#pragma unroll
  for (int c = 0; c < 6; ++c) { const double p = Ad[36 * (size_t)i + 6 * a + c] * y[6 * (size_t)i + c]; s -= p; m += fabs(p); }
  for (int q = adj_ptr[i]; q < adj_ptr[i + 1]; ++q) {
    const int code = adj_item[q], pp = code >> 1;
    const int2 hl = pair[pp];
    if (code & 1) {   // this node is the pair's hi: block (row hi, col lo)
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double p = Ao[36 * (size_t)pp + 6 * a + c] * y[6 * (size_t)hl.y + c]; s -= p; m += fabs(p); }
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double p = Ao[36 * (size_t)pp + 6 * c + a] * y[6 * (size_t)hl.x + c]; s -= p; m += fabs(p); }
    }
  }
  r[t] = s; den[t] = m;
}
// partial[blk] = max_i |r_i| / den_i over [0, n) (rows with den == 0 — unused blocks — are skipped)
__global__ void maxratio_kernel(int n, const double* __restrict__ r, const double* __restrict__ den, double* __restrict__ partial) {
  __shared__ double sm[32];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (den[i] > 0.0) m = fmax(m, fabs(r[i]) / den[i]);
  const double t = block_max(m, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// buf[gpos] <- r of the local border rows, buf[ng6 + gpos] <- den of the local border rows
__global__ void border_pack_res_kernel(int nb6, int first6, const int* __restrict__ gpos, const double* __restrict__ r, const double* __restrict__ den, int ng6,
                                       double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb6) { const int g = 6 * gpos[i / 6] + i % 6; buf[g] = r[first6 + i]; buf[ng6 + g] = den[first6 + i]; }
}
// the summed border rows lack the damping term: r_b -= damp * z_b, den_b += damp * |z_b|
__global__ void border_fix_res_kernel(int ng6, const double* __restrict__ damp, const double* __restrict__ zb, double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ng6) { buf[i] -= damp[i] * zb[i]; buf[ng6 + i] += damp[i] * fabs(zb[i]); }
}

// Backward error of the linear solve just done, componentwise (Oettli-Prager): max_i |b - A y|_i / (|A| |y| + |b|)_i on
// the reduced pose system (scaled, damped, switches eliminated) — the size of the smallest relative perturbation of A
// and b for which y is the exact solution.  Independent of the conditioning and of how small the gradient b has become.
// Sharded: interior rows are local; border rows (numerator and denominator) are summed over the ranks first.
int Solver::linear_residual(double* rel) {
  const int n6 = 6 * N;
  *rel = 0.0;
  if (n6 == 0) return PGS_OK;
  CU(d_pr.resize((size_t)n6)); CU(d_prn.resize((size_t)n6));
  lin_residual_kernel<<<cdiv(n6, 128), 128, 0, stream>>>(N, d_Ad.p, d_Ao.p, d_pair.p, d_adj_ptr.p, d_adj_item.p, d_b.p, d_y.p, d_pr.p, d_prn.p);
  const int fb = (!chains.empty() && first_border >= 0) ? first_border : N, int6 = 6 * fb, b6 = n6 - int6, ng6 = 6 * n_gborder;
  const int grid = std::max(1, std::min(MAX_GRID, cdiv(std::max(int6, 1), 256)));
  double* P = d_partial.p + 2 * MAX_GRID;
  maxratio_kernel<<<grid, 256, 0, stream>>>(int6, d_pr.p, d_prn.p, P);
  reduce_max_kernel<<<1, 256, 0, stream>>>(P, grid, d_scal.p + L_RES);
  if (comm) if (int rc = comm->allreduce_max(d_scal.p + L_RES, 1, stream, &err)) return rc;
  const bool border = !chains.empty() && ng6 > 0;
  if (border) {
    CU(d_xbuf.resize(2 * (size_t)ng6));
    CU(cudaMemsetAsync(d_xbuf.p, 0, sizeof(double) * 2 * (size_t)ng6, stream));
    if (b6) border_pack_res_kernel<<<cdiv(b6, 256), 256, 0, stream>>>(b6, int6, d_border_gpos.p, d_pr.p, d_prn.p, ng6, d_xbuf.p);
    if (comm) if (int rc = comm->allreduce_sum(d_xbuf.p, 2 * (size_t)ng6, stream, &err)) return rc;
    border_fix_res_kernel<<<cdiv(ng6, 256), 256, 0, stream>>>(ng6, d_dampb.p, d_zb.p, d_xbuf.p);
    const int g2 = std::max(1, std::min(MAX_GRID, cdiv(ng6, 256)));
    maxratio_kernel<<<g2, 256, 0, stream>>>(ng6, d_xbuf.p, d_xbuf.p + ng6, P);
    reduce_max_kernel<<<1, 256, 0, stream>>>(P, g2, d_scal.p + L_RES + 1);
  }
  CU(cudaGetLastError());
  if (int rc = read_scalars(L_NSCAL)) return rc;
  *rel = std::max(h_scal[L_RES], border ? h_scal[L_RES + 1] : 0.0);
  return PGS_OK;
}

// ------------------------------------------------------------------------------------------------ public ops
int Solver::evaluate(double* cost, double* r_o, double* J_o, double* r_l, double* J_l, double* r_r, double* J_r) {
  CU(cudaSetDevice(dev));
  if (int rc = sync_params_to_device()) return rc;
  if (int rc = launch_sweep(0, d_pose.p, d_sw.p, d_scal.p + L_COST)) return rc;
  if (int rc = read_scalars(L_NSCAL)) return rc;
  if (cost) *cost = h_scal[L_COST];
  const int Eo = (int)o_c1.size(), El = (int)l_a.size(), K = (int)r_node.size();
  if ((r_o || J_o) && Eo) {
    DBuf<double> tr, tJ;
    if (r_o) CU(tr.resize((size_t)Eo * 6));
    if (J_o) CU(tJ.resize((size_t)Eo * 72));
    export_odom_kernel<<<cdiv(Eo, 128), 128, 0, stream>>>(Eo, d_perm_o.p, d_or.p, d_oJ.p, r_o ? tr.p : nullptr, J_o ? tJ.p : nullptr);
    if (r_o) CU(cudaMemcpyAsync(r_o, tr.p, sizeof(double) * 6 * (size_t)Eo, cudaMemcpyDeviceToHost, stream));
    if (J_o) CU(cudaMemcpyAsync(J_o, tJ.p, sizeof(double) * 72 * (size_t)Eo, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
  }
  if ((r_l || J_l) && El) {
    DBuf<double> tr, tJ;
    if (r_l) CU(tr.resize((size_t)El * 7));
    if (J_l) CU(tJ.resize((size_t)El * 91));
    export_loop_kernel<<<cdiv(El, 128), 128, 0, stream>>>(El, d_perm_l.p, d_lr.p, d_lJ.p, r_l ? tr.p : nullptr, J_l ? tJ.p : nullptr);
    if (r_l) CU(cudaMemcpyAsync(r_l, tr.p, sizeof(double) * 7 * (size_t)El, cudaMemcpyDeviceToHost, stream));
    if (J_l) CU(cudaMemcpyAsync(J_l, tJ.p, sizeof(double) * 91 * (size_t)El, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
  }
  if (r_r && K) CU(cudaMemcpy(r_r, d_gr.p, sizeof(double) * 6 * (size_t)K, cudaMemcpyDeviceToHost));
  if (J_r && K) CU(cudaMemcpy(J_r, d_gJ.p, sizeof(double) * 36 * (size_t)K, cudaMemcpyDeviceToHost));
  return PGS_OK;
}

int Solver::gradient(double* g_pose, double* g_switch) {
  CU(cudaSetDevice(dev));
  if (int rc = sync_params_to_device()) return rc;
  if (int rc = launch_sweep(0, d_pose.p, d_sw.p, d_scal.p + L_COST)) return rc;
  if (int rc = run_assemble()) return rc;
  const int El = (int)l_a.size();
  if (g_pose && N) CU(cudaMemcpyAsync(g_pose, d_g.p, sizeof(double) * 6 * (size_t)N, cudaMemcpyDeviceToHost, stream));
  if (g_switch && El) {
    scatter_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_lg.p, d_stage_s.p);
    CU(cudaMemcpyAsync(g_switch, d_stage_s.p, sizeof(double) * (size_t)El, cudaMemcpyDeviceToHost, stream));
  }
  CU(cudaStreamSynchronize(stream));
  return PGS_OK;
}

int Solver::assemble(double* diag, int* pair_hi, int* pair_lo, double* offdiag, double* loop_v, double* loop_hss) {
  CU(cudaSetDevice(dev));
  if (int rc = sync_params_to_device()) return rc;
  if (int rc = launch_sweep(0, d_pose.p, d_sw.p, d_scal.p + L_COST)) return rc;
  if (int rc = run_assemble()) return rc;
  CU(cudaStreamSynchronize(stream));
  const int El = (int)l_a.size();
  if (diag && N) CU(cudaMemcpy(diag, d_Hd.p, sizeof(double) * 36 * (size_t)N, cudaMemcpyDeviceToHost));
  if (pair_hi) std::memcpy(pair_hi, h_pair_hi.data(), sizeof(int) * n_pairs);
  if (pair_lo) std::memcpy(pair_lo, h_pair_lo.data(), sizeof(int) * n_pairs);
  if (offdiag && n_pairs) CU(cudaMemcpy(offdiag, d_Ho.p, sizeof(double) * 36 * (size_t)n_pairs, cudaMemcpyDeviceToHost));
  if (El && (loop_v || loop_hss)) {
    std::vector<double> v((size_t)El * 12), hs(El);
    CU(cudaMemcpy(v.data(), d_lv.p, sizeof(double) * 12 * (size_t)El, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hs.data(), d_lh.p, sizeof(double) * (size_t)El, cudaMemcpyDeviceToHost));
    for (int e = 0; e < El; ++e) {   // sorted -> caller order; v is stored (c1 side, c2 side) = (b, a)
      const int o = perm_l[e];
      if (loop_v) std::memcpy(loop_v + 12 * (size_t)o, &v[12 * (size_t)e], sizeof(double) * 12);
      if (loop_hss) loop_hss[o] = hs[e];
    }
  }
  return PGS_OK;
}

int Solver::linear_step(double radius, double* delta_pose, double* delta_switch, double* mcc, int* lin_iters) {
  CU(cudaSetDevice(dev));
  if (int rc = sync_params_to_device()) return rc;
  if (int rc = launch_sweep(0, d_pose.p, d_sw.p, d_scal.p + L_COST)) return rc;
  if (int rc = run_assemble()) return rc;
  if (int rc = compute_scaling(true)) return rc;
  if (int rc = build_system(radius)) return rc;
  if (int rc = solve_linear(lin_iters)) return rc;
  const int El = (int)l_a.size();
  const int n = std::max(6 * N, El);
  finish_step_kernel<<<cdiv(n, 128), 128, 0, stream>>>(N, El, d_lidx.p, d_y.p, d_lvt.p, d_lw.p, d_lgt.p, d_scale_p.p, d_scale_s.p, d_dp.p, d_ds.p);
  MccArgs M;
  M.n_odom = (int)o_c1.size(); M.n_loop = El; M.n_reg = (int)r_node.size(); M.o_idx = d_oidx.p; M.l_idx = d_lidx.p; M.r_node = d_rnode.p;
  M.o_r = d_or.p; M.o_J = d_oJ.p; M.l_r = d_lr.p; M.l_J = d_lJ.p; M.g_r = d_gr.p; M.g_J = d_gJ.p; M.dp = d_dp.p; M.ds = d_ds.p; M.partial = d_partial.p;
  const int mg = std::max(1, std::min(1024, cdiv(std::max(M.n_odom, M.n_loop), 256)));
  model_cost_kernel<<<mg, 256, 0, stream>>>(M);
  reduce_sum_kernel<<<1, 256, 0, stream>>>(d_partial.p, mg, -1.0, d_scal.p + L_MCC);
  CU(cudaGetLastError());
  if (int rc = read_scalars(L_NSCAL)) return rc;
  if (mcc) *mcc = h_scal[L_MCC];
  if (delta_pose && N) CU(cudaMemcpy(delta_pose, d_dp.p, sizeof(double) * 6 * (size_t)N, cudaMemcpyDeviceToHost));
  if (delta_switch && El) {
    scatter_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_ds.p, d_stage_s.p);
    CU(cudaMemcpyAsync(delta_switch, d_stage_s.p, sizeof(double) * (size_t)El, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
  }
  return PGS_OK;
}

static constexpr size_t FLUSH_DOUBLES = (size_t)48 << 20;   // 384 MiB > 126 MB L2

int Solver::flush_l2_now() {
  CU(d_flush.resize(FLUSH_DOUBLES));
  flush_kernel<<<1184, 256, 0, stream>>>(d_flush.p, FLUSH_DOUBLES, 1.0);
  drain_kernel<<<1184, 256, 0, stream>>>(d_flush.p, (size_t)32 << 20);   // 256 MiB re-read: the dirty tail of the write is evicted
  CU(cudaGetLastError());
  return PGS_OK;
}

int Solver::time_sweep(int mode, int reps, int flush_l2, double* ms, double* ms_kernel, int64_t* launches) {
  CU(cudaSetDevice(dev));
  if (int rc = sync_params_to_device()) return rc;
  if (reps < 1) reps = 1;
  cudaEvent_t evk; CU(cudaEventCreate(&evk));
  double total = 0.0, total_k = 0.0;
  // every repetition is timed on its own: [ev0] sweep kernel (cost reduction fused in) [evk][ev1]; the optional L2
  // flush runs between repetitions, outside the timed span
  for (int i = 0; i < reps; ++i) {
    if (flush_l2) if (int rc = flush_l2_now()) return rc;
    CU(cudaEventRecord(ev0, stream));
    if (int rc = launch_sweep(mode, d_pose.p, d_sw.p, d_scal.p + L_COST, evk)) return rc;
    CU(cudaEventRecord(ev1, stream));
    CU(cudaEventSynchronize(ev1));
    float t = 0, tk = 0; CU(cudaEventElapsedTime(&t, ev0, ev1)); CU(cudaEventElapsedTime(&tk, ev0, evk));
    total += t; total_k += tk;
  }
  cudaEventDestroy(evk);
  if (ms) *ms = total / reps;
  if (ms_kernel) *ms_kernel = total_k / reps;
  if (launches) *launches = 1LL * reps;   // one launch per repetition: the sweep's last block also reduces the cost
  return PGS_OK;
}

int Solver::time_stream_write(int64_t bytes, int reps, int flush_l2, double* ms) {
  CU(cudaSetDevice(dev));
  if (bytes <= 0 || bytes > ((int64_t)16 << 30)) return fail(PGS_ERR_INVALID_ARGUMENT, "time_stream_write: bytes must be in (0, 16 GiB]");
  if (reps < 1) reps = 1;
  CU(d_flush.resize(std::max(FLUSH_DOUBLES, (size_t)(bytes + 7) / 8)));
  double total = 0.0;
  for (int i = 0; i < reps; ++i) {
    if (flush_l2) if (int rc = flush_l2_now()) return rc;
    CU(cudaEventRecord(ev0, stream));
    stream_write_kernel<<<sweep_grid, 256, 0, stream>>>(reinterpret_cast<double2*>(d_flush.p), (size_t)bytes / 16);
    CU(cudaEventRecord(ev1, stream));
    CU(cudaEventSynchronize(ev1));
    float t = 0; CU(cudaEventElapsedTime(&t, ev0, ev1));
    total += t;
  }
  CU(cudaGetLastError());
  if (ms) *ms = total / reps;
  return PGS_OK;
}

// End-to-end step: poses and switches arrive from host memory, the mode-J sweep runs, the cost goes back.  The upload is
// the long pole (6 MB over PCIe against a 50 us sweep), so it is cut into chunks of keyframes on a copy stream, and
// after every chunk the compute stream packs those poses and sweeps the tiles whose keyframes are all there (edges are
// sorted by keyframe, so that is a growing prefix of every tile list); only the last chunk's share of the sweep and
// the 8-byte read-back are left when the bus falls silent.
int Solver::evaluate_from_host(const double* q, const double* t, const double* s, double* cost) {
  CU(cudaSetDevice(dev));
  if (int rc = finalize()) return rc;
  if (host_params_newer) if (int rc = sync_params_to_device()) return rc;
  const int El = (int)l_a.size();
  if ((q || t) && N && (!q || !t)) return fail(PGS_ERR_INVALID_ARGUMENT, "evaluate_from_host: q and t must be given together");
  if (!copy_stream) {
    CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_copy_go, cudaEventDisableTiming));
  }
  const int To = (int)o_pm.size(), Tl = (int)l_pm.size(), Tr = (int)r_pm.size();
  // The cost goes straight into the pinned read-back buffer (device-visible under unified addressing): the CTA that reduces the
  // tile partials stores it over the bus, and the stream synchronisation below is all the host waits for — no 8-byte copy.
  double* cost_dst = h_scal + L_COST;
  const char* e2e_env = getenv("PGS_E2E_MODE");                                   // 0 chunked copies, 1 chunked reads of pinned memory, 2 one shot
  const int e2e_mode = e2e_env ? atoi(e2e_env) : 0;
  if (!(q && t) || N < 4096 || e2e_mode == 2) {
    // nothing to overlap: one shot
    if (q && N) {
      CU(cudaMemcpyAsync(d_stage_q.p, q, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, stream));
      CU(cudaMemcpyAsync(d_stage_t.p, t, sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, stream));
      pack_pose_kernel<<<cdiv(N, 256), 256, 0, stream>>>(0, N, d_stage_q.p, d_stage_t.p, d_node_const.p, d_pose.p);
    }
    if (s && El) {
      CU(cudaMemcpyAsync(d_stage_s.p, s, sizeof(double) * (size_t)El, cudaMemcpyHostToDevice, stream));
      gather_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_stage_s.p, d_sw.p);
    }
    if (int rc = launch_sweep(0, d_pose.p, d_sw.p, cost_dst)) return rc;
  } else {
    constexpr int MAXCHUNK = (int)(sizeof(ev_chunk) / sizeof(ev_chunk[0]));
    // Two chunks, the first 70 % of the keyframes: every copy costs a few microseconds of set-up on the copy engine, so few
    // chunks beat many (4 equal chunks 188 us, 3: 182, 2: 176), and with the sweep ~2.4x faster than the bus the share of the
    // first chunk is swept just as the second one lands (first chunk 50 / 60 / 70 / 80 %: 177 / 175 / 171 / 179 us;
    // profiles/r02_e2e_lab_v2.txt).  PGS_E2E_CHUNKS / PGS_E2E_SPLIT override.
    static const int NCHUNK = [] { const char* e = getenv("PGS_E2E_CHUNKS"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > MAXCHUNK ? MAXCHUNK : v); }();
    // Pinned (registered) host memory is visible to the device: the pack kernel then reads q and t straight over the bus,
    // chunk by chunk, with no staging copy and none of the per-copy set-up cost that nine small cudaMemcpyAsync calls have.
    // Pageable memory goes through cudaMemcpyAsync as before.
    bool direct = false;
    if (e2e_mode == 1) {
      cudaPointerAttributes aq{}, at{};
      direct = cudaPointerGetAttributes(&aq, q) == cudaSuccess && cudaPointerGetAttributes(&at, t) == cudaSuccess && aq.type == cudaMemoryTypeHost && at.type == cudaMemoryTypeHost &&
               aq.devicePointer && at.devicePointer;
      cudaGetLastError();
      if (direct) { q = (const double*)aq.devicePointer; t = (const double*)at.devicePointer; }
    }
    CU(cudaEventRecord(ev_copy_go, stream));                 // the staging buffers are free once earlier work on the compute stream is done
    CU(cudaStreamWaitEvent(copy_stream, ev_copy_go, 0));
    if (s && El) {
      CU(cudaMemcpyAsync(d_stage_s.p, s, sizeof(double) * (size_t)El, cudaMemcpyHostToDevice, copy_stream));
    }
    int done_o = 0, done_l = 0, done_r = 0;
    for (int k = 0; k < NCHUNK; ++k) {
      int n0 = (int)((long long)N * k / NCHUNK), n1 = (int)((long long)N * (k + 1) / NCHUNK);
      if (NCHUNK == 2) { static const double f = [] { const char* e = getenv("PGS_E2E_SPLIT"); const double v = e ? atof(e) : 0.7; return (v > 0.05 && v < 0.95) ? v : 0.7; }();
                         const int cut = (int)(f * N); if (k == 0) n1 = cut; else n0 = cut; }
      if (direct) {
        // few CTAs: the bus is the bound, and the SMs are wanted by the sweep of the previous chunk
        pack_pose_kernel<<<std::min(cdiv(n1 - n0, 256), 64), 256, 0, copy_stream>>>(n0, n1 - n0, q + 4 * (size_t)n0, t + 3 * (size_t)n0, d_node_const.p, d_pose.p);
      } else {
        CU(cudaMemcpyAsync(d_stage_q.p + 4 * (size_t)n0, q + 4 * (size_t)n0, sizeof(double) * 4 * (size_t)(n1 - n0), cudaMemcpyHostToDevice, copy_stream));
        CU(cudaMemcpyAsync(d_stage_t.p + 3 * (size_t)n0, t + 3 * (size_t)n0, sizeof(double) * 3 * (size_t)(n1 - n0), cudaMemcpyHostToDevice, copy_stream));
      }
      CU(cudaEventRecord(ev_chunk[k], copy_stream));
      CU(cudaStreamWaitEvent(stream, ev_chunk[k], 0));
      if (k == 0 && s && El) gather_kernel<<<cdiv(El, 256), 256, 0, stream>>>(El, d_perm_l.p, d_stage_s.p, d_sw.p);
      if (!direct) pack_pose_kernel<<<cdiv(n1 - n0, 256), 256, 0, stream>>>(n0, n1 - n0, d_stage_q.p + 4 * (size_t)n0, d_stage_t.p + 3 * (size_t)n0, d_node_const.p, d_pose.p);
      // tiles whose keyframes are all below n1
      const bool last = k == NCHUNK - 1;
      const int up_o = last ? To : (int)(std::upper_bound(o_pm.begin(), o_pm.end(), n1 - 1) - o_pm.begin());
      const int up_l = last ? Tl : (int)(std::upper_bound(l_pm.begin(), l_pm.end(), n1 - 1) - l_pm.begin());
      const int up_r = last ? Tr : (int)(std::upper_bound(r_pm.begin(), r_pm.end(), n1 - 1) - r_pm.begin());
      const int ranges[6] = {done_o, up_o, done_l, up_l, done_r, up_r};
      if (int rc = launch_sweep(0, d_pose.p, d_sw.p, cost_dst, nullptr, ranges, last ? 1 : 0)) return rc;
      done_o = up_o; done_l = up_l; done_r = up_r;
    }
  }
  CU(cudaStreamSynchronize(stream));
  if (cost) *cost = h_scal[L_COST];
  device_params_newer = true;   // host mirrors are refreshed lazily by the getters
  return PGS_OK;
}

// ------------------------------------------------------------------------------------------------ the LM loop
// Follows Ceres 1.12-1.14 TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy.
int Solver::solve(pgs_summary* sum, pgs_iteration* iters, int cap) {
  CU(cudaSetDevice(dev));
  if (comm_owned && opt.linear_solver != PGS_SKYLINE_CHOLESKY) return fail(PGS_ERR_INVALID_ARGUMENT, "multi-GPU solve needs the skyline Cholesky solver");
  if (!is_inner) if (int rc = choose_linear_solver()) return rc;
  if (comm_owned || (want_chains() && !pcg_fallback)) {   // outer solver of a sharded run (several GPUs, or two chains on this one)
    if (inner_dirty) plain_chain = false;
    if (!plain_chain) { const int rc = solve_dist(sum, iters, cap); if (rc != PGS_PLAIN_CHAIN) return rc; }
  }
  HostLap lap;
  if (int rc = sync_params_to_device()) return rc;
  lap.lap(is_inner ? "inner: structure + upload" : "structure + upload");
  const bool sharded = !chains.empty();
  border_scale_ready = false;
  ms_sweep = ms_asm = ms_lin = ms_comm = ms_eliminate = ms_exchange = ms_border = 0.0;
  backward_error.clear();
  {
    // the factors are allocated up front: a rank that cannot hold its share says so before anyone waits for it
    int prc = prepare_linear();
    if (comm) { const int a = comm->agree(prc, stream, &err); if (a) return prc ? prc : a; }
    else if (prc) return prc;
  }
  lap.lap("factor symbolic + allocation");
  CU(cudaEventRecord(ev_t0, stream));
  const int El = (int)l_a.size(), Eo = (int)o_c1.size(), K = (int)r_node.size();
  const int rgrid = std::max(1, std::min(1024, cdiv(std::max(N, El), 256)));
  const int mg = std::max(1, std::min(1024, cdiv(std::max(Eo, El), 256)));
  double* P0 = d_partial.p; double* P1 = d_partial.p + MAX_GRID; double* P2 = d_partial.p + 2 * MAX_GRID;

  std::vector<pgs_iteration> rows;
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid = 0, n_succ = 0, n_unsucc = 0, lin_total = 0;
  double x_cost = 0, grad_max = 0, grad_norm = 0, fixed_cost = 0;
  int termination = PGS_NO_CONVERGENCE;
  int rc = PGS_OK;

  auto eval_grad_jac = [&](int iteration) -> int {
    tic();
    if (int r = launch_sweep(0, d_pose.p, d_sw.p, d_scal.p + L_COST)) return r;
    ms_sweep += toc();
    tic();
    if (int r = run_assemble()) return r;
    if (iteration == 0) if (int r = compute_scaling(true)) return r;
    ms_asm += toc();
    tic();
    if (sharded) if (int r = border_gradient_exchange()) return r;   // summed border gradient -> d_gfull, summed cost -> L_COST
    if (comm) ms_comm += toc(); else ms_asm += toc();
    tic();
    // |Plus(x, -g) - x| in the ambient space (Ceres' projected-gradient norms)
    retract_kernel<<<rgrid, 256, 0, stream>>>(N, d_node_counted.p, El, d_node_used.p, d_pose.p, d_sw.p, sharded ? d_gfull.p : d_g.p, d_lg.p, -1.0, d_cpose.p, d_csw.p, P0, P1, P2);
    reduce_sum_kernel<<<1, 256, 0, stream>>>(P0, rgrid, 1.0, d_scal.p + L_DIFF2);
    reduce_max_kernel<<<1, 256, 0, stream>>>(P2, rgrid, d_scal.p + L_MAX);
    if (cudaGetLastError() != cudaSuccess) return cuda_fail(cudaPeekAtLastError(), "eval_grad_jac");
    if (comm) {
      if (int r = comm->allreduce_sum(d_scal.p + L_DIFF2, 1, stream, &err)) return r;
      if (int r = comm->allreduce_max(d_scal.p + L_MAX, 1, stream, &err)) return r;
    }
    if (iteration == 0) {
      // Summary::fixed_cost: the blocks that bind constant keyframes only do not move; evaluated once, as Ceres does
      fixed_cost_kernel<<<1, 256, 0, stream>>>(n_fixed_o, d_fixed_o.p, d_or.p, n_fixed_r, d_fixed_r.p, d_gr.p, d_scal.p + L_FIXED);
      if (comm) if (int r = comm->allreduce_sum(d_scal.p + L_FIXED, 1, stream, &err)) return r;
    }
    ms_asm += toc();
    if (int r = read_scalars(L_NSCAL)) return r;
    if (iteration == 0) fixed_cost = h_scal[L_FIXED];
    x_cost = h_scal[L_COST] - fixed_cost; grad_norm = std::sqrt(h_scal[L_DIFF2]); grad_max = h_scal[L_MAX];   // the minimiser sees the reduced program's cost
    return PGS_OK;
  };

  if ((rc = eval_grad_jac(0))) return rc;
  const double initial_cost = x_cost + fixed_cost;
  pgs_iteration it{}; it.iteration = 0; it.cost = x_cost; it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm;
  it.step_is_valid = 1; it.step_is_successful = 1;

  while (true) {
    if (it.step_is_successful) ++n_succ; else ++n_unsucc;
    it.trust_region_radius = radius;
    rows.push_back(it);
    if (it.iteration >= opt.max_num_iterations) { termination = PGS_NO_CONVERGENCE; break; }
    if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { termination = PGS_CONVERGENCE; break; }
    if (radius <= opt.min_trust_region_radius) { termination = PGS_CONVERGENCE; break; }
    pgs_iteration prev = it; it = pgs_iteration{}; it.iteration = prev.iteration + 1;
    it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;

    // ---- ComputeTrustRegionStep
    tic();
    if (!reuse_diagonal) if ((rc = compute_scaling(false))) return rc;
    if ((rc = build_system(radius))) return rc;
    ms_asm += toc();
    tic();
    int lin_it = 0;
    cur_radius = radius; cur_reuse_diag = reuse_diagonal;
    const int lrc = solve_linear(&lin_it);
    ms_lin += toc();
    if (ph_pending) { float e = 0; cudaEventElapsedTime(&e, ev0, ev_ph[0]); collect_chain_times(e); }
    if (lrc != PGS_OK && lrc != PGS_ERR_LINEAR_SOLVER) return lrc;
    if (lrc == PGS_OK && opt.check_linear_solves) { double rel = 0.0; if ((rc = linear_residual(&rel))) return rc; backward_error.push_back(rel); }
    it.linear_solver_iterations = lin_it; lin_total += lin_it;
    reuse_diagonal = true;
    bool step_ok = (lrc == PGS_OK);
    double mcc = 0.0, cand_cost = 0.0, diff2 = 0.0, x2 = 0.0;
    if (step_ok) {
      const int n = std::max(6 * N, El);
      finish_step_kernel<<<cdiv(n, 128), 128, 0, stream>>>(N, El, d_lidx.p, d_y.p, d_lvt.p, d_lw.p, d_lgt.p, d_scale_p.p, d_scale_s.p, d_dp.p, d_ds.p);
      MccArgs M;
      M.n_odom = Eo; M.n_loop = El; M.n_reg = K; M.o_idx = d_oidx.p; M.l_idx = d_lidx.p; M.r_node = d_rnode.p;
      M.o_r = d_or.p; M.o_J = d_oJ.p; M.l_r = d_lr.p; M.l_J = d_lJ.p; M.g_r = d_gr.p; M.g_J = d_gJ.p; M.dp = d_dp.p; M.ds = d_ds.p; M.partial = P0;
      model_cost_kernel<<<mg, 256, 0, stream>>>(M);
      reduce_sum_kernel<<<1, 256, 0, stream>>>(P0, mg, -1.0, d_scal.p + L_MCC);
      // candidate point + its cost, queued speculatively so one read-back serves all the tests
      retract_kernel<<<rgrid, 256, 0, stream>>>(N, d_node_counted.p, El, d_node_used.p, d_pose.p, d_sw.p, d_dp.p, d_ds.p, 1.0, d_cpose.p, d_csw.p, P0, P1, P2);
      reduce_sum_kernel<<<1, 256, 0, stream>>>(P0, rgrid, 1.0, d_scal.p + L_DIFF2);
      reduce_sum_kernel<<<1, 256, 0, stream>>>(P1, rgrid, 1.0, d_scal.p + L_X2);
      tic();
      if ((rc = launch_sweep(1, d_cpose.p, d_csw.p, d_scal.p + L_CCOST))) return rc;
      ms_sweep += toc();
      if (sharded) if ((rc = dist_fail_flag())) return rc;
      // model change, candidate cost, step norms and the pivot flags, summed over the ranks
      if (comm) if ((rc = comm->allreduce_sum(d_scal.p + L_MCC, L_FAIL - L_MCC + 1, stream, &err))) return rc;
      if ((rc = read_scalars(L_NSCAL))) return rc;
      mcc = h_scal[L_MCC]; cand_cost = h_scal[L_CCOST] - fixed_cost; diff2 = h_scal[L_DIFF2]; x2 = h_scal[L_X2];
      step_ok = std::isfinite(mcc) && std::isfinite(diff2) && mcc > 0.0;
      if (sharded && h_scal[L_FAIL] != 0.0) step_ok = false;   // some chain (of some rank) hit a non-positive pivot
    }
    it.step_is_valid = step_ok ? 1 : 0;
    if (!step_ok) {
      // ---- HandleInvalidStep
      if (++invalid >= opt.max_num_consecutive_invalid_steps) { termination = PGS_FAILURE; it.cost = x_cost; it.trust_region_radius = radius; rows.push_back(it); break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      it.cost = x_cost; it.step_is_successful = 0;
      continue;
    }
    invalid = 0;
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    // ---- ParameterToleranceReached
    it.step_norm = std::sqrt(diff2);
    const double x_norm = std::sqrt(x2);
    if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { termination = PGS_CONVERGENCE; it.cost = x_cost; rows.push_back(it); break; }
    // ---- FunctionToleranceReached
    it.cost_change = x_cost - cand_cost;
    if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) { termination = PGS_CONVERGENCE; it.cost = x_cost; rows.push_back(it); break; }
    // ---- IsStepSuccessful
    it.relative_decrease = (x_cost - cand_cost) / mcc;
    if (it.relative_decrease > opt.min_relative_decrease) {
      std::swap(d_pose.p, d_cpose.p); std::swap(d_sw.p, d_csw.p);
      if ((rc = eval_grad_jac(it.iteration))) return rc;
      it.cost = x_cost; it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm; it.step_is_successful = 1;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
    } else {
      it.step_is_successful = 0; it.cost = cand_cost;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  device_params_newer = true;
  CU(cudaEventRecord(ev_t1, stream)); CU(cudaEventSynchronize(ev_t1));
  float total_ms = 0; CU(cudaEventElapsedTime(&total_ms, ev_t0, ev_t1));
  if (sum) {
    sum->initial_cost = initial_cost; sum->final_cost = x_cost + fixed_cost; sum->termination = termination;
    sum->num_successful_steps = n_succ; sum->num_unsuccessful_steps = n_unsucc; sum->num_iterations = (int)rows.size();
    sum->linear_solver_iterations = lin_total; sum->ms_sweep = ms_sweep; sum->ms_assemble = ms_asm; sum->ms_linear_solve = ms_lin; sum->ms_total = total_ms;
    sum->factor_nnz = use_pcg() ? 0 : factor_nnz;
    sum->linear_solver_used = use_pcg() ? PGS_BLOCK_PCG : PGS_SKYLINE_CHOLESKY;
    sum->n_chains = use_pcg() ? 0 : std::max<int>(1, (int)chains.size());
    sum->factor_flops = use_pcg() ? 0.0 : est_flops;
    sum->max_linear_backward_error = -1.0;
    for (double e : backward_error) sum->max_linear_backward_error = std::max(sum->max_linear_backward_error, e);
    sum->fixed_cost = fixed_cost;
    sum->ms_comm = ms_comm + ms_exchange;
  }
  for (int i = 0; iters && i < (int)rows.size() && i < cap; ++i) iters[i] = rows[i];
  return PGS_OK;
}

}  // namespace pgs
