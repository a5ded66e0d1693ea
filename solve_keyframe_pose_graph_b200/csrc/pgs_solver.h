// Host-side driver of the device pose-graph solver (C++ behind the C-ABI of include/pgs.h).
//
// Owns the problem (parameter blocks + residual blocks as the reference's ceres::Problem does,
// reference src/PoseGraphSLAM.cpp:1340-1367,1550-1556,1629-1633,1847-1849), the device buffers and
// the trust-region Levenberg-Marquardt loop that replaces ceres::Solve (PoseGraphSLAM.cpp:1903).
// The loop restates Ceres 1.12-1.14's TrustRegionMinimizer + LevenbergMarquardtStrategy semantics
// (SURVEY §3.4); only scalars cross PCIe per iteration.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <memory>
#include "../../include/pgs.h"
#include "host/host_lap.h"
#include "pgs_comm.h"

namespace pgs {



template <class T>
struct DBuf {
  T* p = nullptr; size_t n = 0;
  ~DBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t resize(size_t m, bool zero = false) {
    if (m <= n && p) { if (zero) return cudaMemset(p, 0, sizeof(T) * (m ? m : 1)); return cudaSuccess; }
    release();
    cudaError_t e = cudaMalloc((void**)&p, sizeof(T) * (m ? m : 1));
    if (e != cudaSuccess) { p = nullptr; return e; }
    n = m;
    if (zero) return cudaMemset(p, 0, sizeof(T) * (m ? m : 1));
    return cudaSuccess;
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t s) {
    cudaError_t e = resize(h.size()); if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s);
  }
};

struct SkylineFactor;  // pgs_skyline.cu

static constexpr int MAX_GRID = 2048;
// slots of the small device/pinned scalar array.  L_MCC..L_FAIL are contiguous: the multi-GPU loop sums them
// over the ranks with one small all-reduce
enum { PGS_PLAIN_CHAIN = 1 };   // internal: solve_dist found nothing to split, the caller solves on its own handle
enum { L_COST = 8, L_MCC = 9, L_DIFF2 = 10, L_X2 = 11, L_CCOST = 12, L_FAIL = 13, L_MAX = 14, L_FIXED = 15, L_RES = 16, L_NSCAL = 32 };

class Solver {
 public:
  explicit Solver(const pgs_options& o);
  ~Solver();
  int init();

  // ---- problem construction (host side; marks the structure dirty)
  int set_nodes(int n, const double* q, const double* t, bool append);
  int update_nodes(int first, int n, const double* q, const double* t);
  int get_poses(int first, int n, double* q, double* t);
  int set_constant(int first, int n, int constant);   // ceres::Problem::SetParameterBlockConstant / Variable on node blocks
  int set_switches(int first, int n, const double* s);
  int get_switches(int first, int n, double* s);
  int add_odom(int m, const int* c1, const int* c2, const double* q, const double* t, const double* w);
  int add_loop(int m, const int* a, const int* b, const double* q, const double* t, const double* w);
  int set_regs(int k, const int* node, const double* q, const double* t, const double* w);

  // ---- evaluation / assembly / solve
  int evaluate(double* cost, double* r_o, double* J_o, double* r_l, double* J_l, double* r_r, double* J_r);
  int gradient(double* g_pose, double* g_switch);
  int assemble(double* diag, int* pair_hi, int* pair_lo, double* offdiag, double* loop_v, double* loop_hss);
  int linear_step(double radius, double* delta_pose, double* delta_switch, double* mcc, int* lin_iters);
  int solve(pgs_summary* sum, pgs_iteration* iters, int cap);
  int time_sweep(int mode, int reps, int flush_l2, double* ms, double* ms_kernel, int64_t* launches);
  int time_stream_write(int64_t bytes, int reps, int flush_l2, double* ms);
  int evaluate_from_host(const double* q, const double* t, const double* s, double* cost);
  int64_t sweep_bytes() const;
  void sizes(pgs_sizes* s);

  // ---- multi-GPU
  // On the handle the caller sees: dist_init() attaches a communicator; solve() then partitions the graph by node
  // range, loads this rank's interior + the border into an inner Solver and runs the sharded LM in it.
  int dist_init(int rank, int world, const void* nccl_id128);
  int dist_stats(pgs_dist_stats* out);
  int get_backward_errors(double* out, int cap, int* n);
  int dist_init_local(int rank, int world, const char* group);   // the same over the in-process transport (pgs_comm.h)
  // On the inner (per-rank) solver: set by the outer one before the first solve.
  struct ChainSpec { int off = 0, len = 0; std::vector<int> border; };   // interior = local nodes [off, off+len) in elimination order (padded to whole panels); border = local indices, ascending
  Comm* comm = nullptr;          // not owned; non-null adds the collectives
  bool is_inner = false;
  int first_border = -1;         // local index of the first border node (they come last); -1 = no border
  std::vector<ChainSpec> chains; // the chains this rank eliminates
  std::vector<int> border_gpos;  // per local border node: its position in the global border order
  std::vector<char> border_counted;   // per local border node: this rank counts it in the norms and publishes its pose
  std::vector<char> forced_used; // per local node: used by some residual block of SOME rank (Ceres' reduced program is global)
  int n_gborder = 0;             // border nodes of the whole graph
  std::vector<int> gborder_env;  // per global border position: first position of its row envelope in the border system

  std::string err;
  pgs_options opt;

 private:
  int fail(int code, const std::string& msg) { err = msg; return code; }
  int cuda_fail(cudaError_t e, const char* what);
  int finalize();                        // sort edges, build incidence/pair/adjacency structures, upload
  int sync_params_to_device();           // host q,t,sw -> device pose / sw (if dirty)
  int sync_params_to_host();             // device -> host mirrors (if device is newer)
  int launch_sweep(int mode, const double* pose, const double* sw, double* cost_out_dev, cudaEvent_t after_kernel = nullptr,
                   const int* tile_ranges = nullptr, int reduce = 1);   // tile_ranges: {o0, o1, l0, l1, r0, r1} for a partial sweep
  int run_assemble();
  int compute_scaling(bool compute_scale);
  int build_system(double radius);
  int solve_linear(int* iters);          // Ad/Ao/b -> y  (PCG or skyline Cholesky)
  int solve_pcg(int* iters);
  int solve_skyline();
  int solve_dist(pgs_summary* sum, pgs_iteration* iters, int cap);   // outer: partition, inner solve, gather
  int solve_chains();                    // inner: concurrent chain eliminations, border system, back-substitution
  int prepare_linear();                  // allocate the factor(s) before the LM loop (so that a failure is known up front)
  int prepare_chains();
  int border_gradient_exchange();        // inner: all-reduce of the border gradient + cost
  int dist_fail_flag();                  // inner: pivot flags of all factors -> d_scal[L_FAIL]
  bool want_chains() const;
  bool use_pcg() const { return opt.linear_solver == PGS_BLOCK_PCG || pcg_fallback; }
  void estimate_skyline(double* bytes, double* flops) const;
  int choose_linear_solver();              // single GPU: split the elimination into two chains burning from both ends
  int nb6() const { return first_border >= 0 ? 6 * (N - first_border) : 0; }
  int linear_residual(double* rel);      // ||b - A y|| / ||b|| of the reduced system just solved (backward error of the step)
  int flush_l2_now();                    // evict everything from L2 and leave no dirty lines behind
  int read_scalars(int n);               // d_scal -> h_scal (pinned), synchronises the stream

  int N = 0;
  // host problem, caller order
  std::vector<double> h_q, h_t, h_sw;
  std::vector<int> o_c1, o_c2; std::vector<double> o_q, o_t, o_w;
  std::vector<int> l_a, l_b; std::vector<double> l_q, l_t, l_w;
  std::vector<int> r_node; std::vector<double> r_q, r_t, r_w;
  bool structure_dirty = true, regs_dirty = true, host_params_newer = true, device_params_newer = false;
  bool inner_dirty = true;               // outer: the inner solver of a sharded solve has to be rebuilt
  bool pcg_fallback = false;             // the skyline estimate exceeded the budget: this problem is solved with PGS_BLOCK_PCG
  double est_flops = 0.0;
  bool plain_chain = false;              // outer, one GPU: the plan found nothing to split (no crossing edge / separator too large)

  // sorted structure (host)
  std::vector<int> perm_o, perm_l;       // sorted index -> caller index
  std::vector<int> inv_perm_l;
  int n_pairs = 0;
  std::vector<int> h_pair_hi, h_pair_lo;
  std::vector<char> h_node_used, h_node_const;
  std::vector<int> o_pm, l_pm, r_pm;     // per tile: largest keyframe index touched by the tiles up to it (prefix maximum)
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t ev_copy_go = nullptr;

  // device
  int dev = 0; cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  DBuf<double> d_pose, d_cpose, d_sw, d_csw, d_stage_q, d_stage_t, d_stage_s;
  DBuf<int2> d_oidx, d_lidx, d_pair;
  DBuf<double> d_oobs, d_lobs, d_ranchor;
  DBuf<int> d_rnode, d_perm_o, d_perm_l;
  DBuf<double> d_or, d_oJ, d_lr, d_lJ, d_gr, d_gJ;
  DBuf<int> d_inc_ptr, d_inc_item, d_pe_ptr, d_pe_item, d_adj_ptr, d_adj_item;
  DBuf<char> d_node_used, d_node_const;
  DBuf<double> d_Hd, d_g, d_Ho, d_lv, d_lh, d_lg, d_lvt, d_lw, d_lgt;
  DBuf<double> d_scale_p, d_scale_s, d_diag_p, d_diag_s;
  DBuf<double> d_Ad, d_Ao, d_b, d_y, d_dp, d_ds;
  DBuf<double> d_Minv, d_px, d_pr, d_prn, d_pz, d_pp, d_pAp;
  DBuf<double> d_partial, d_scal, d_flush, d_cost_tile;
  DBuf<unsigned int> d_counter;
  double* h_scal = nullptr;              // pinned
  double* h_pin_q = nullptr; double* h_pin_t = nullptr; double* h_pin_s = nullptr; size_t pin_n = 0, pin_s = 0;
  int sweep_grid = 0;
  SkylineFactor* sky = nullptr;
  SkylineFactor* sky_border = nullptr;
  int64_t factor_nnz = 0;
  // multi-GPU state
  std::unique_ptr<Comm> comm_owned;      // outer
  std::unique_ptr<Solver> inner;         // outer: this rank's local problem
  std::vector<int> loc2glob, loop2glob;  // outer: local node -> global node (-1 = padding), local loop -> global loop
  pgs_dist_stats dstats{};
  DBuf<double> d_xbuf, d_gfull, d_sb, d_diagb, d_zb, d_dampb;   // inner: border exchange buffer, summed gradient, border scale / LM diagonal / solution / damping
  struct ChainState { SkylineFactor* f = nullptr; cudaStream_t st = nullptr; cudaEvent_t done = nullptr; DBuf<double> y; DBuf<int> bmap; int nb = 0, nf = 0; };
  std::vector<ChainState> cstate;
  cudaEvent_t ev_fork = nullptr;
  DBuf<int> d_border_gpos; DBuf<char> d_node_counted;
  DBuf<int> d_fixed_o, d_fixed_r; int n_fixed_o = 0, n_fixed_r = 0;   // residual blocks whose parameter blocks are all constant (Summary::fixed_cost)
  DBuf<const int*> d_flag_ptrs; bool flag_ptrs_ready = false;
  std::vector<double> backward_error;    // per linear solve of the last LM run: ||b - A y|| / ||b||
  void release_chains();
  double ms_comm = 0, ms_eliminate = 0, ms_exchange = 0, ms_border = 0;   // sharded: collectives of the LM loop | this rank's chains | border all-reduce incl. waiting | border factor + solve
  cudaEvent_t ev_ph[4] = {nullptr, nullptr, nullptr, nullptr}; bool ph_pending = false;
  void collect_chain_times(float ms_from_start_to_eliminated);
  double cur_radius = 0.0; bool cur_reuse_diag = false, border_scale_ready = false;

  // phase timers (ms, accumulated per solve)
  double ms_sweep = 0, ms_asm = 0, ms_lin = 0;
  void tic(); double toc();
};

}  // namespace pgs
