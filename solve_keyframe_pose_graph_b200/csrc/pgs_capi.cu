// extern "C" surface declared in include/pgs.h.  Thin: argument checks + dispatch into pgs::Solver.
#include <cstdlib>
#include <exception>
#include <new>
#include <string>
#include "pgs_solver.h"
#include "host/partition.h"

using pgs::Solver;
static thread_local std::string g_create_error;

// A factorisation needs its streams (panel chain, next-panel tiles, trailing update; per elimination chain) to overlap.  With
// the driver's default of 8 hardware work queues per context, streams of a process that runs other streams as well (NCCL, a
// framework) come to share a queue and serialise each other — measured 56 us per panel instead of 38 on every rank of a 2-
// and a 4-GPU solve.  The variable is read when the CUDA context is created: it is set here, when the library is loaded,
// unless the host has chosen a value itself; a host that creates its context before loading libpgs.so has to set it itself.
__attribute__((constructor)) static void pgs_more_hardware_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
struct pgs_solver_s { Solver* s; };

#define H(h) do { if (!(h) || !(h)->s) return PGS_ERR_INVALID_ARGUMENT; } while (0)
// No C++ exception may cross the C boundary (include/pgs.h): host-side containers can throw std::bad_alloc / length_error.
#define GUARDED(h, expr)                                                                                              \
  do {                                                                                                                \
    try { return (expr); }                                                                                            \
    catch (const std::bad_alloc&) { (h)->s->err = "out of host memory"; return PGS_ERR_OUT_OF_MEMORY; }               \
    catch (const std::exception& e) { (h)->s->err = std::string("unexpected C++ exception: ") + e.what(); return PGS_ERR_STATE; } \
    catch (...) { (h)->s->err = "unexpected C++ exception"; return PGS_ERR_STATE; }                                   \
  } while (0)

extern "C" {

int pgs_default_options(pgs_options* o) {
  if (!o) return PGS_ERR_INVALID_ARGUMENT;
  o->max_num_iterations = 10;                 // reference PoseGraphSLAM.cpp:1272
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5; o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->jacobi_scaling = 1; o->switch_init = 0.99;  // reference PoseGraphSLAM.cpp:353
  o->device = 0; o->linear_solver = PGS_SKYLINE_CHOLESKY; o->pcg_max_iterations = 20000; o->pcg_tolerance = 1e-10;
  o->chains = 0; o->check_linear_solves = 1; o->max_factor_bytes = 0.0; o->max_factor_flops = 0.0;
  return PGS_OK;
}

int pgs_create(const pgs_options* opt, pgs_handle* out) {
  if (!out) return PGS_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  pgs_options o;
  if (opt) o = *opt; else pgs_default_options(&o);
  Solver* s = nullptr;
  try {
    s = new Solver(o);
    const int rc = s->init();
    if (rc != PGS_OK) { g_create_error = s->err; delete s; return rc; }
    *out = new pgs_solver_s{s};
    return PGS_OK;
  } catch (const std::bad_alloc&) { delete s; g_create_error = "out of host memory"; return PGS_ERR_OUT_OF_MEMORY; }
  catch (const std::exception& e) { delete s; g_create_error = std::string("unexpected C++ exception: ") + e.what(); return PGS_ERR_STATE; }
}
int pgs_destroy(pgs_handle h) { if (!h) return PGS_OK; delete h->s; delete h; return PGS_OK; }
const char* pgs_last_error(pgs_handle h) { return (h && h->s) ? h->s->err.c_str() : g_create_error.c_str(); }
int pgs_get_sizes(pgs_handle h, pgs_sizes* out) { H(h); if (!out) return PGS_ERR_INVALID_ARGUMENT; h->s->sizes(out); return PGS_OK; }

int pgs_set_nodes(pgs_handle h, int32_t n, const double* q, const double* t) { H(h); GUARDED(h, h->s->set_nodes(n, q, t, false)); }
int pgs_append_nodes(pgs_handle h, int32_t n, const double* q, const double* t) { H(h); GUARDED(h, h->s->set_nodes(n, q, t, true)); }
int pgs_update_nodes(pgs_handle h, int32_t first, int32_t n, const double* q, const double* t) { H(h); GUARDED(h, h->s->update_nodes(first, n, q, t)); }
int pgs_get_poses(pgs_handle h, int32_t first, int32_t n, double* q, double* t) { H(h); GUARDED(h, h->s->get_poses(first, n, q, t)); }
int pgs_set_constant_nodes(pgs_handle h, int32_t first, int32_t n, int32_t constant) { H(h); GUARDED(h, h->s->set_constant(first, n, constant)); }
int pgs_set_switches(pgs_handle h, int32_t first, int32_t n, const double* s) { H(h); GUARDED(h, h->s->set_switches(first, n, s)); }
int pgs_get_switches(pgs_handle h, int32_t first, int32_t n, double* s) { H(h); GUARDED(h, h->s->get_switches(first, n, s)); }
int pgs_add_odom_edges(pgs_handle h, int32_t m, const int32_t* c1, const int32_t* c2, const double* q, const double* t, const double* w) {
  H(h); GUARDED(h, h->s->add_odom(m, c1, c2, q, t, w)); }
int pgs_add_loop_edges(pgs_handle h, int32_t m, const int32_t* a, const int32_t* b, const double* q, const double* t, const double* w) {
  H(h); GUARDED(h, h->s->add_loop(m, a, b, q, t, w)); }
int pgs_set_regularizers(pgs_handle h, int32_t k, const int32_t* node, const double* q, const double* t, const double* w) {
  H(h); GUARDED(h, h->s->set_regs(k, node, q, t, w)); }
int pgs_evaluate(pgs_handle h, double* cost, double* r_o, double* J_o, double* r_l, double* J_l, double* r_r, double* J_r) {
  H(h); GUARDED(h, h->s->evaluate(cost, r_o, J_o, r_l, J_l, r_r, J_r)); }
int pgs_gradient(pgs_handle h, double* gp, double* gs) { H(h); GUARDED(h, h->s->gradient(gp, gs)); }
int pgs_assemble(pgs_handle h, double* diag, int32_t* phi, int32_t* plo, double* off, double* lv, double* lh) {
  H(h); GUARDED(h, h->s->assemble(diag, phi, plo, off, lv, lh)); }
int pgs_linear_step(pgs_handle h, double radius, double* dp, double* ds, double* mcc, int32_t* it) { H(h); GUARDED(h, h->s->linear_step(radius, dp, ds, mcc, it)); }
int pgs_solve(pgs_handle h, pgs_summary* sum, pgs_iteration* iters, int32_t cap) { H(h); GUARDED(h, h->s->solve(sum, iters, cap)); }
int pgs_time_stream_write(pgs_handle h, int64_t bytes, int32_t reps, int32_t flush, double* ms) { H(h); GUARDED(h, h->s->time_stream_write(bytes, reps, flush, ms)); }
int pgs_time_sweep(pgs_handle h, int32_t mode, int32_t reps, int32_t flush, double* ms, double* ms_kernel, int64_t* launches) { H(h); GUARDED(h, h->s->time_sweep(mode, reps, flush, ms, ms_kernel, launches)); }
int pgs_evaluate_from_host(pgs_handle h, const double* q, const double* t, const double* s, double* cost) { H(h); GUARDED(h, h->s->evaluate_from_host(q, t, s, cost)); }
int64_t pgs_sweep_algorithmic_bytes(pgs_handle h) { if (!h || !h->s) return 0; return h->s->sweep_bytes(); }


int pgs_dist_unique_id(void* id128) {
  if (!id128) return PGS_ERR_INVALID_ARGUMENT;
  return pgs::Comm::unique_id(id128, &g_create_error);
}
int pgs_dist_init(pgs_handle h, int32_t rank, int32_t world, const void* id128) { H(h); GUARDED(h, h->s->dist_init(rank, world, id128)); }
int pgs_dist_init_local(pgs_handle h, int32_t rank, int32_t world, const char* group) { H(h); GUARDED(h, h->s->dist_init_local(rank, world, group)); }
int pgs_dist_get_stats(pgs_handle h, pgs_dist_stats* out) { H(h); if (!out) return PGS_ERR_INVALID_ARGUMENT; return h->s->dist_stats(out); }
int pgs_get_linear_backward_errors(pgs_handle h, double* out, int32_t cap, int32_t* n) { H(h); return h->s->get_backward_errors(out, cap, n); }
int pgs_partition(int32_t n_nodes, int32_t world, int32_t n_odom, const int32_t* c1, const int32_t* c2, int32_t n_loop, const int32_t* a,
                  const int32_t* b, int32_t n_reg, const int32_t* reg_node, int32_t* node_owner, int32_t* odom_owner, int32_t* loop_owner,
                  int32_t* reg_owner, int32_t* n_border, int32_t chains_per_rank, int32_t* cut, int32_t* node_chain, int32_t* chain_down,
                  int32_t* n_chains) {
  if (n_nodes < 0 || world < 1 || n_odom < 0 || n_loop < 0 || n_reg < 0) return PGS_ERR_INVALID_ARGUMENT;
  if ((n_odom && (!c1 || !c2)) || (n_loop && (!a || !b)) || (n_reg && !reg_node)) return PGS_ERR_INVALID_ARGUMENT;
  for (int e = 0; e < n_odom; ++e) if (c1[e] < 0 || c1[e] >= n_nodes || c2[e] < 0 || c2[e] >= n_nodes) return PGS_ERR_INVALID_ARGUMENT;
  for (int e = 0; e < n_loop; ++e) if (a[e] < 0 || a[e] >= n_nodes || b[e] < 0 || b[e] >= n_nodes) return PGS_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < n_reg; ++k) if (reg_node[k] < 0 || reg_node[k] >= n_nodes) return PGS_ERR_INVALID_ARGUMENT;
  pgs::Partition P;
  pgs::make_partition(n_nodes, world, n_odom, c1, c2, n_loop, a, b, n_reg, reg_node, &P, chains_per_rank);
  if (cut) for (int k = 0; k <= world; ++k) cut[k] = P.cut[k];
  if (node_chain) for (int i = 0; i < n_nodes; ++i) node_chain[i] = P.node_chain[i];
  if (chain_down) for (size_t c = 0; c < P.ranges.size(); ++c) chain_down[c] = P.ranges[c].down ? 1 : 0;
  if (n_chains) *n_chains = (int)P.ranges.size();
  if (node_owner) for (int i = 0; i < n_nodes; ++i) node_owner[i] = P.node_owner[i];
  if (odom_owner) for (int e = 0; e < n_odom; ++e) odom_owner[e] = P.odom_owner[e];
  if (loop_owner) for (int e = 0; e < n_loop; ++e) loop_owner[e] = P.loop_owner[e];
  if (reg_owner) for (int k = 0; k < n_reg; ++k) reg_owner[k] = P.reg_owner[k];
  if (n_border) *n_border = (int)P.border.size();
  return PGS_OK;
}

}  // extern "C"
