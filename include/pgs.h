/* pgs.h — C-ABI of the B200-native pose-graph hot path (libpgs.so).
 *
 * Drop-in boundary for the Ceres path of mpkuse/solve_keyframe_pose_graph: everything the
 * reference does between "parameter blocks + residual blocks exist" and "optimised poses are
 * readable" (reference src/PoseGraphSLAM.cpp:1340-1367, 1550-1556, 1629-1633, 1847-1849, 1903 and
 * the getters at :178-224) is reachable through these entry points.  Plain pointers and sizes,
 * host memory, fp64, SoA; the caller allocates every output.  No torch / Eigen / ROS types.
 *
 * Conventions (reference file:line):
 *   - quaternions are stored x,y,z,w (PoseGraphSLAM.h:153); a node pose is w_T_c = (q[4], t[3]).
 *   - tangent ordering per node is [dtheta(3), dt(3)], dtheta being the half-angle left-multiplied
 *     increment of ceres::EigenQuaternionParameterization (PoseGraphSLAM.cpp:1351-1353).
 *   - an odometry edge binds parameters (c1, c2) with observation c1_T_c2 and weight w:
 *     SixDOFError (CeresResidues.h:19-90), added at PoseGraphSLAM.cpp:1629-1633 with (c1,c2)=(u,u-f).
 *   - a loop edge (a, b) carries b_T_a and owns one switch variable initialised to 0.99
 *     (PoseGraphSLAM.cpp:353); it binds (c1,c2,s) = (b, a, s): SixDOFErrorWithSwitchingConstraints
 *     (CeresResidues.h:145-222) added at PoseGraphSLAM.cpp:1550-1556.  Its weight is stored but,
 *     as in the reference (CeresResidues.h:198), not applied.
 *   - a regulariser anchors one node to a fixed pose: NodePoseRegularization
 *     (CeresResidues.h:96-141) added at PoseGraphSLAM.cpp:1847-1849.
 *   - cost = 1/2 sum ||r||^2, no loss function (all AddResidualBlock calls pass NULL).
 *
 * Every function returns PGS_OK (0) or a negative pgs_status; pgs_last_error() gives the text.
 * Nothing here calls exit() or throws across the boundary (the reference exit()s on bad state,
 * PoseGraphSLAM.cpp:1680,1711).  All compute runs on the CUDA device chosen in pgs_options.device
 * and fails with PGS_ERR_CUDA when no device is usable — there is no CPU fallback.
 */
#ifndef PGS_H_
#define PGS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgs_solver_s* pgs_handle;

typedef enum pgs_status {
  PGS_OK = 0,
  PGS_ERR_INVALID_ARGUMENT = -1,
  PGS_ERR_CUDA = -2,
  PGS_ERR_OUT_OF_MEMORY = -3,
  PGS_ERR_LINEAR_SOLVER = -4,
  PGS_ERR_STATE = -5
} pgs_status;

typedef enum pgs_termination { PGS_CONVERGENCE = 0, PGS_NO_CONVERGENCE = 1, PGS_FAILURE = 2 } pgs_termination;

typedef enum pgs_linear_solver {
  PGS_SKYLINE_CHOLESKY = 0, /* direct: block skyline LL^T on device (stands in for SPARSE_NORMAL_CHOLESKY) */
  PGS_BLOCK_PCG = 1         /* iterative: block-Jacobi preconditioned CG on device */
} pgs_linear_solver;

/* ceres::Solver::Options in effect in the reference (PoseGraphSLAM.cpp:1268-1272 + Ceres 1.12-1.14
 * defaults) followed by device knobs.  pgs_default_options() fills the
 * reference values. */
typedef struct pgs_options {
  int32_t max_num_iterations;               /* 10, PoseGraphSLAM.cpp:1272 */
  double initial_trust_region_radius;       /* 1e4 */
  double max_trust_region_radius;           /* 1e16 */
  double min_trust_region_radius;           /* 1e-32 */
  double min_relative_decrease;             /* 1e-3 */
  double min_lm_diagonal;                   /* 1e-6 */
  double max_lm_diagonal;                   /* 1e32 */
  int32_t max_num_consecutive_invalid_steps;/* 5 */
  double function_tolerance;                /* 1e-6 */
  double gradient_tolerance;                /* 1e-10 */
  double parameter_tolerance;               /* 1e-8 */
  int32_t jacobi_scaling;                   /* 1 */
  double switch_init;                       /* 0.99, PoseGraphSLAM.cpp:353 */
  int32_t device;                           /* CUDA ordinal */
  int32_t linear_solver;                    /* pgs_linear_solver */
  int32_t pcg_max_iterations;               /* per linear solve */
  double pcg_tolerance;                     /* relative residual ||b-Ax|| / ||b|| */
  int32_t chains;                           /* elimination chains of the skyline solver on ONE GPU: 0 = automatic (two chains
                                               burning from both ends of the keyframe chain from 4096 nodes on, unless the
                                               separator in the middle would be a quarter of the graph), 1 = one natural-order
                                               chain, 2.. = that many; with pgs_dist_init: chains per rank (0 = 1) */
  int32_t check_linear_solves;              /* != 0: measure the componentwise backward error max_i |b - A y|_i / (|A||y| + |b|)_i of
                                               every linear solve (one block SpMV each; pgs_get_linear_backward_errors) */
  double max_factor_bytes;                  /* skyline factor larger than this (0 = 80 % of the free device memory) or ... */
  double max_factor_flops;                  /* ... costlier than this per factorisation (0 = no limit): fall back to PGS_BLOCK_PCG */
} pgs_options;

/* One row of Ceres' minimizer_progress_to_stdout table. */
typedef struct pgs_iteration {
  int32_t iteration;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease, trust_region_radius;
  int32_t step_is_valid, step_is_successful, linear_solver_iterations;
} pgs_iteration;

typedef struct pgs_summary {
  double initial_cost, final_cost;
  int32_t termination;                      /* pgs_termination */
  int32_t num_successful_steps, num_unsuccessful_steps, num_iterations;
  int32_t linear_solver_iterations;         /* total PCG iterations (0 for the direct solver) */
  double ms_sweep, ms_assemble, ms_linear_solve, ms_total; /* device time by phase (CUDA events) */
  int64_t factor_nnz;                       /* scalars stored by the skyline factor(s) (0 for PCG) */
  int32_t linear_solver_used;               /* pgs_linear_solver actually used (PGS_BLOCK_PCG when the skyline estimate exceeded the budget) */
  int32_t n_chains;                         /* elimination chains on this GPU (1 = plain natural order) */
  double factor_flops;                      /* estimated flops of one natural-order factorisation (sum over columns of rows-below^2) */
  double max_linear_backward_error;         /* max over the linear solves of max_i |b - A y|_i / (|A||y| + |b|)_i; -1 when not measured */
  double fixed_cost;                        /* Ceres Summary::fixed_cost: cost of the residual blocks whose parameter blocks are all constant */
  double ms_comm;                           /* device time spent in collectives incl. waiting for other ranks (multi-GPU) */
} pgs_summary;

/* Sizes of the residual/Jacobian outputs of pgs_evaluate. */
typedef struct pgs_sizes {
  int32_t n_nodes, n_odom, n_loop, n_reg;
  int32_t n_pairs;                          /* distinct off-diagonal 6x6 blocks of J^T J */
} pgs_sizes;

int pgs_default_options(pgs_options* opt);
int pgs_create(const pgs_options* opt, pgs_handle* out);
int pgs_destroy(pgs_handle h);
const char* pgs_last_error(pgs_handle h);   /* h may be NULL: error of the last failed pgs_create */
int pgs_get_sizes(pgs_handle h, pgs_sizes* out);

/* ---- variable store: replaces allocate_and_append_new_opt_variable_withpose / update_opt_variable_with
 *      (PoseGraphSLAM.cpp:226-335) and allocate_and_append_new_edge_switch_var (:351-361) ---- */
int pgs_set_nodes(pgs_handle h, int32_t n, const double* q_xyzw, const double* t);      /* replace all */
int pgs_append_nodes(pgs_handle h, int32_t n, const double* q_xyzw, const double* t);   /* grow */
int pgs_update_nodes(pgs_handle h, int32_t first, int32_t n, const double* q_xyzw, const double* t); /* new initial guesses */
int pgs_get_poses(pgs_handle h, int32_t first, int32_t n, double* q_xyzw, double* t);   /* getNodePose, :197-214 */
/* ceres::Problem::SetParameterBlockConstant (constant != 0) / SetParameterBlockVariable on the q and t blocks of nodes
 * [first, first+n): what PoseGraphSLAM::load_state does to every keyframe of a session restored from disk
 * (PoseGraphSLAM.cpp:40-170).  Constant blocks keep their value, get zero Jacobian columns and leave the step and
 * gradient norms, as in Ceres' reduced program; residual blocks on them still count in the cost. */
int pgs_set_constant_nodes(pgs_handle h, int32_t first, int32_t n, int32_t constant);
int pgs_set_switches(pgs_handle h, int32_t first, int32_t n, const double* s);
int pgs_get_switches(pgs_handle h, int32_t first, int32_t n, double* s);                /* get_loopedge_switching_variable_val, PoseGraphSLAM.h:219 */

/* ---- residual blocks: replace ceres::Problem::AddResidualBlock at PoseGraphSLAM.cpp:1629-1633,
 *      :1550-1556, :1847-1849 and RemoveResidualBlock at :1803-1807 ---- */
int pgs_add_odom_edges(pgs_handle h, int32_t m, const int32_t* c1, const int32_t* c2,
                       const double* q_c1Tc2, const double* t_c1Tc2, const double* w);
int pgs_add_loop_edges(pgs_handle h, int32_t m, const int32_t* a, const int32_t* b,
                       const double* q_bTa, const double* t_bTa, const double* w);
int pgs_set_regularizers(pgs_handle h, int32_t k, const int32_t* node, const double* q_f, const double* t_f,
                         const double* w);                                              /* replaces the previous set */

/* ---- evaluation: ceres Evaluate() semantics at the current parameters.  cost always; any other
 *      pointer may be NULL.  Blocks come back in the caller's insertion order, row-major:
 *      r_odom[E][6], J_odom[E][6][12] (cols th_c1,t_c1,th_c2,t_c2); r_loop[E][7], J_loop[E][7][13]
 *      (last col = switch); r_reg[K][6], J_reg[K][6][6].  On the device the sweep writes the same
 *      numbers in a warp-tiled SoA layout; these are re-laid-out copies. ---- */
int pgs_evaluate(pgs_handle h, double* cost, double* r_odom, double* J_odom, double* r_loop, double* J_loop,
                 double* r_reg, double* J_reg);
/* J^T r in tangent space: g_pose[N][6], g_switch[n_loop]. */
int pgs_gradient(pgs_handle h, double* g_pose, double* g_switch);
/* Block-sparse J^T J (pose part, before scaling/damping/switch elimination): diag[N][6][6],
 * pair_hi[n_pairs], pair_lo[n_pairs], offdiag[n_pairs][6][6] = block (row hi, col lo), and the
 * per-loop-edge switch coupling v[n_loop][12] = J_p^T j_s, hss[n_loop] = j_s^T j_s. Any may be NULL. */
int pgs_assemble(pgs_handle h, double* diag, int32_t* pair_hi, int32_t* pair_lo, double* offdiag,
                 double* loop_v, double* loop_hss);
/* One LM linear step at the current parameters with first-iteration Jacobi scaling and the given
 * radius: unscaled tangent step delta_pose[N][6], delta_switch[n_loop], and the model cost change. */
int pgs_linear_step(pgs_handle h, double radius, double* delta_pose, double* delta_switch, double* model_cost_change,
                    int32_t* linear_iterations);

/* ---- the solve: replaces ceres::Solve(reint_options, &reint_problem, &reint_summary)
 *      (PoseGraphSLAM.cpp:1903).  iters may be NULL; at most iters_cap rows are written. ---- */
int pgs_solve(pgs_handle h, pgs_summary* summary, pgs_iteration* iters, int32_t iters_cap);

/* ---- multi-GPU: one process per GPU, node-range sharding.  Every process loads the
 *      SAME full graph through the calls above, then attaches a communicator; pgs_solve() then sweeps/assembles
 *      only this rank's residual blocks, eliminates this rank's interior nodes and all-reduces the border Schur
 *      system once per LM iteration over NCCL.  After the solve every rank holds all optimised poses. ---- */
typedef struct pgs_dist_stats {
  int32_t rank, world;
  int32_t n_interior_nodes, n_border_nodes;         /* this rank's interior; border is global */
  int32_t n_odom_owned, n_loop_owned, n_reg_owned;
  int64_t border_buffer_bytes;                      /* size of the per-iteration all-reduce */
  int64_t n_collectives, bytes_reduced;             /* over the last solve */
  int32_t n_chains, n_local_border_nodes;           /* chains this rank eliminates; border nodes its factors hold */
  int64_t factor_nnz;                               /* this rank's chain factors + its copy of the border factor */
  double ms_comm;                                   /* this rank's time in collectives over the last solve (waiting included) */
  double ms_eliminate;                              /* ... in eliminating its own chains (what the partition balances) */
  double ms_exchange;                               /* ... in the border all-reduce, i.e. mostly waiting for the slowest rank (part of ms_comm) */
  double ms_border;                                 /* ... in factoring and solving the border system */
} pgs_dist_stats;
int pgs_dist_unique_id(void* id128);                /* rank 0: ncclGetUniqueId; distribute the 128 bytes yourself */
int pgs_dist_init(pgs_handle h, int32_t rank, int32_t world, const void* id128);
/* The same sharded solve over an in-process transport: `world` handles of ONE process (same or different devices), each
 * driven by its own host thread, that name the same `group`.  Every collective is a host barrier plus device copies —
 * for hosts that drive several GPUs from one process, and for running any number of ranks on a single GPU. */
int pgs_dist_init_local(pgs_handle h, int32_t rank, int32_t world, const char* group);
int pgs_dist_get_stats(pgs_handle h, pgs_dist_stats* out);   /* also valid after a single-GPU solve that used several chains */
/* Componentwise backward errors (Oettli-Prager) of the linear solves of the last pgs_solve (pgs_options.check_linear_solves),
 * in the order they were done; at most cap values are written, *n gets the number available. */
int pgs_get_linear_backward_errors(pgs_handle h, double* out, int32_t cap, int32_t* n);
/* The partition rule on its own (host only, no device needed): node_owner[N] = owning rank or -1 for a border
 * node; *_owner = rank that evaluates each residual block; cut[world+1] = the node ranges of the ranks (ranges that
 * carry a separator through their elimination get fewer nodes); node_chain[N] = chain that eliminates the node (-1 =
 * border), chain_down[world*chains_per_rank] = 1 where a chain runs in descending keyframe order.  Any output may be
 * NULL.  chains_per_rank = 0: the default of pgs_options.chains.  Loop edge e couples (a[e], b[e]). */
int pgs_partition(int32_t n_nodes, int32_t world, int32_t n_odom, const int32_t* c1, const int32_t* c2, int32_t n_loop,
                  const int32_t* a, const int32_t* b, int32_t n_reg, const int32_t* reg_node, int32_t* node_owner,
                  int32_t* odom_owner, int32_t* loop_owner, int32_t* reg_owner, int32_t* n_border,
                  int32_t chains_per_rank, int32_t* cut, int32_t* node_chain, int32_t* chain_down, int32_t* n_chains);

/* ---- measurement hooks used by bench.py ---- */
/* Runs the residual+Jacobian sweep `reps` times with every input already resident in HBM and
 * returns the mean device time in milliseconds, measured per repetition with CUDA events on the
 * solver's stream (ms_per_sweep and ms_sweep_kernel coincide since the cost reduction moved into the sweep kernel).
 * flush_l2 != 0 writes a 384 MiB scratch buffer (> the 126 MB L2) and then re-reads 256 MiB of it between
 * repetitions, outside the timed spans: the write evicts the solver's data, the read pass drains the dirty
 * lines the write left behind so their write-back is not billed to the sweep.
 * mode: 0 = residuals + Jacobians (mode J), 1 = cost only. */
int pgs_time_sweep(pgs_handle h, int32_t mode, int32_t reps, int32_t flush_l2, double* ms_per_sweep,
                   double* ms_sweep_kernel, int64_t* kernel_launches);
/* Practical ceiling for a store-dominated kernel of this size: a pure streaming write of `bytes` bytes (up to 16 GiB),
 * timed exactly like pgs_time_sweep (per repetition, CUDA events, same L2 flush).  Diagnostic only. */
int pgs_time_stream_write(pgs_handle h, int64_t bytes, int32_t reps, int32_t flush_l2, double* ms_per_write);
/* End-to-end step: host poses/switches (pinned or pageable) -> device, one mode-J sweep, cost back
 * to the host.  q,t,s may be NULL to reuse the current values of that array. */
int pgs_evaluate_from_host(pgs_handle h, const double* q_xyzw, const double* t, const double* s, double* cost);
/* Algorithmic bytes of one mode-J sweep per SURVEY §8(d): 56N + 72 Eo + 72 El + 68 K in,
 * 624 Eo + 784 El + 336 K out. */
int64_t pgs_sweep_algorithmic_bytes(pgs_handle h);

#ifdef __cplusplus
}
#endif
#endif /* PGS_H_ */
