/* pgs_facade.h — C view of the ROS-free C++ facade (NodeDataManager + PoseGraphSLAM,
 * solve_keyframe_pose_graph_b200/csrc/host/) so that non-C++ hosts (the Python tests, bench.py) can drive
 * the same trigger logic the reference node runs: ingest keyframes / loop edges / kidnap signals, call
 * one wake-up of reinit_ceres_problem_onnewloopedge_optimize6DOF(), read the optimised poses.
 * Reference: src/NodeDataManager.cpp:23-215 (ingest), src/PoseGraphSLAM.cpp:1251-1950 (trigger),
 * src/PoseGraphSLAM.cpp:178-224 (getters).  C++ users include the facade headers directly. */
#ifndef PGS_FACADE_H_
#define PGS_FACADE_H_
#include <stdint.h>
#include "pgs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgs_facade_s* pgs_facade_handle;

typedef struct pgs_facade_options {
  int32_t odom_fanout;        /* 5, PoseGraphSLAM.cpp:1577 */
  int32_t derive_odometry;    /* 1: odometry edges derived from the manager's poses (reference behaviour) */
  int32_t dry_run;            /* 1: build problem + initial guesses on the host only (no CUDA) */
  pgs_options solver;
} pgs_facade_options;

int pgs_facade_default_options(pgs_facade_options* o);
int pgs_facade_create(const pgs_facade_options* o, pgs_facade_handle* out);
void pgs_facade_destroy(pgs_facade_handle h);
const char* pgs_facade_last_error(pgs_facade_handle h);

/* ingest (NodeDataManager) */
int pgs_facade_add_nodes(pgs_facade_handle h, int32_t n, const int64_t* stamps_ns, const double* q_xyzw, const double* t);
int pgs_facade_add_loop_edges(pgs_facade_handle h, int32_t m, const int32_t* a, const int32_t* b, const double* q_bTa,
                              const double* t_bTa, const double* w);          /* by node index */
int pgs_facade_add_loop_edge_stamped(pgs_facade_handle h, int64_t stamp_a_ns, int64_t stamp_b_ns, const double* q_bTa,
                                     const double* t_bTa, double w);          /* 1 added, 0 dropped (no such keyframe) */
int pgs_facade_kidnap_indicator(pgs_facade_handle h, int64_t stamp_ns, int32_t kidnapped);
int pgs_facade_add_odometry_edge(pgs_facade_handle h, int32_t a, int32_t b, const double* q_aTb, const double* t_aTb, double w);

/* ROS-message entry points (csrc/host/RosShim.h): the fields of nav_msgs/Odometry, LoopEdge.msg and std_msgs/Header
 * as plain arguments, routed through the reference's callback names.  frame_id: "kidnapped" / "unkidnapped". */
int pgs_facade_camera_pose_callback(pgs_facade_handle h, uint32_t sec, uint32_t nsec, const double* position_xyz, const double* orientation_xyzw,
                                    const double* covariance36 /* pose.covariance, may be NULL */);
int pgs_facade_loopclosure_pose_callback(pgs_facade_handle h, uint32_t sec0, uint32_t nsec0, uint32_t sec1, uint32_t nsec1, const double* position_xyz,
                                         const double* orientation_xyzw, float weight, const char* description);   /* 1 added, 0 dropped */
int pgs_facade_rcvd_kidnap_indicator_callback(pgs_facade_handle h, uint32_t sec, uint32_t nsec, const char* frame_id);

/* PoseGraphSLAM::load_state (src/PoseGraphSLAM.cpp:40-170): after the manager was restored from disk, turn every loaded
 * keyframe into a CONSTANT optimisation variable (pose in its set root's frame) and move solvedUntil to the last one. */
int pgs_facade_load_state(pgs_facade_handle h);

/* one wake-up of the solver thread: 1 solved, 0 not triggered, <0 error */
int pgs_facade_solve_once(pgs_facade_handle h, int32_t force);
int pgs_facade_status(pgs_facade_handle h);
/* The way keyframe_pose_graph_slam_node.cpp runs the solver (:353,475-477,494,505): enable, run
 * reinit_ceres_problem_onnewloopedge_optimize6DOF() on its own std::thread polling at rate_hz (the reference's 0.5 Hz
 * when rate_hz <= 0), disable + join.  Ingest calls and getters may be used concurrently from other threads. */
int pgs_facade_thread_start(pgs_facade_handle h, double rate_hz);
int pgs_facade_thread_stop(pgs_facade_handle h);      /* returns the number of solves the thread triggered */

/* results (PoseGraphSLAM getters) */
int32_t pgs_facade_n_nodes(pgs_facade_handle h);
int32_t pgs_facade_solved_until(pgs_facade_handle h);
/* getAllNodePose: the first min(nNodes(), cap) optimised poses; returns how many were written (the solver thread may
 * append variables between a pgs_facade_n_nodes() call and this one, hence the capacity) */
int pgs_facade_get_poses(pgs_facade_handle h, int32_t cap, double* q_xyzw, double* t);
int pgs_facade_get_switches(pgs_facade_handle h, int32_t n, double* s);       /* per manager loop edge */
int pgs_facade_get_summary(pgs_facade_handle h, pgs_summary* s, pgs_iteration* iters, int32_t cap);

/* Composer (reference src/Composer.cpp:10-292): one pass of pose_assember_thread's loop body on the device
 * (include/pgs_compose.h).  out_T [cap][16] row-major assembled pose per keyframe (global_lmb), out_world [cap] its
 * world id (the key of global_jmb); either may be NULL.  Every keyframe the manager holds at the time of the call is
 * assembled; the first min(that, cap) are written and their count is returned (< 0 on error) — the callbacks may
 * append keyframes between a pgs_facade_n_keyframes() call and this one, hence the capacity. */
int pgs_facade_compose(pgs_facade_handle h, int32_t cap, double* out_T, int32_t* out_world);
int32_t pgs_facade_n_keyframes(pgs_facade_handle h);                          /* manager->getNodeLen() */
/* get_last_known_camerapose (Composer.cpp:264-276): index of the last keyframe or -1; T16 / stamp may be NULL */
int pgs_facade_last_known_camerapose(pgs_facade_handle h, double* T16, int64_t* stamp_ns);
int pgs_facade_compose_timing(pgs_facade_handle h, double* ms_kernel, double* ms_total);

/* On-disk formats of the reference (csrc/host/GraphIO.h): writes dir/log_posegraph.json
 * (NodeDataManager::saveAsJSON, src/NodeDataManager.cpp:503-628), dir/log_optimized_poses.json
 * (PoseGraphSLAM::saveAsJSON, src/PoseGraphSLAM.cpp:1111-1207) and dir/solved_posegraph.json (Composer::saveStateToDisk,
 * src/Composer.cpp:990-1031: SolvedPoseGraph — empty until pgs_facade_compose has run — KidnapTimestamps, WorldsData).
 * Returns 1 | 2, plus 4 when SolvedPoseGraph holds the assembled poses, or < 0. */
int pgs_facade_save_json(pgs_facade_handle h, const char* dir);
/* NodeDataManager::loadFromJSON (src/NodeDataManager.cpp:631-754) into an empty facade, kidnap signals replayed. */
int pgs_facade_load_posegraph_json(pgs_facade_handle h, const char* dir);
/* format helpers (golden-vector tests): prettyprintMatrix4d (src/utils/PoseManipUtils.cpp:206-215) and the matrix
 * string parsers (PoseManipUtils.cpp:272-295, RawFileIO.cpp:372-409).  Return the string length / 1 on success. */
int pgs_io_prettyprint(const double* T16, char* out, int32_t cap);
int pgs_io_mat_to_string(const double* T16, int32_t solved_posegraph_layout, char* out, int32_t cap);
int pgs_io_string_to_mat(const char* s, double* T16);
/* Worlds::loadStateFromDisk (src/Worlds.cpp:499-640) from the "WorldsData" object of a solved_posegraph.json into an
 * EMPTY facade (before any keyframe): relative poses, world stamps, union-find op-log replayed. */
int pgs_facade_load_worlds_state(pgs_facade_handle h, const char* solved_posegraph_json);
/* loads a solved_posegraph.json: n = number of keyframes (call with NULL outputs to size), poses [n][16], stamps, ids */
/* Composer::saveStateToDisk (src/Composer.cpp:955-1105), what the reference node does at shutdown for its `saveStateToDisk` parameter:
 * if the session is not kidnapped the current world is ended at the stamp of the last keyframe (mark_as_kidnapped_and_signal_end_of_world,
 * :967-974 — the session is left in the kidnapped state, as the reference's is), then <dir>/solved_posegraph.json is written.
 * pgs_facade_save_json above writes the same file without touching the session. */
int pgs_facade_save_state_to_disk(pgs_facade_handle h, const char* dir);
/* Composer::loadStateFromDisk (src/Composer.cpp:1109-1177), what the reference node does for its `loadStateFromDisk` parameter:
 * from <dir>/solved_posegraph.json restore the Worlds object ("WorldsData"), the kidnap stamps ("KidnapTimestamps"), every keyframe of
 * "SolvedPoseGraph" (moved from its set root's frame back into its own world's frame) and then PoseGraphSLAM::load_state (the
 * restored keyframes become constant optimisation variables, solvedUntil moves to the last one).  Into an empty handle. */
int pgs_facade_load_state_from_disk(pgs_facade_handle h, const char* dir);
int pgs_io_load_solved_posegraph(const char* json_file, double* T, int64_t* stamp_ns, int32_t* world_id, int32_t* set_id, int32_t cap);

/* introspection of the graph-construction rules (parity tests against the oracle front-end).  The lists behind these
 * calls belong to the solver thread: while it runs (pgs_facade_thread_start) they return PGS_ERR_STATE.  The get_
 * functions write at most cap entries (cap_nodes / cap_edges for the two-sized one) and return how many. */
/* The blocks the reference's switched-off builds would add for this session (kind = pgs_fourdof_kind of pgs_fourdof.h:
 * FourDOFError on the odometry edges, PoseGraphSLAM.cpp:1630; FourDOFErrorWithSwitchingConstraints on the loop edges,
 * :1551; the __USE_YPR_REP build, QinFourDOFWeightError on odometry then loop edges, :1608-1626,1534-1548), and their
 * evaluation on the device.  _size gives the array lengths; every output pointer may be NULL.  get_ returns the number of
 * blocks written; evaluate_ fails with PGS_ERR_INVALID_ARGUMENT when there are more blocks than cap_edges. */
int pgs_facade_alternative_terms_size(pgs_facade_handle h, int32_t kind, int32_t* n_nodes, int32_t* n_edges);
int pgs_facade_get_alternative_terms(pgs_facade_handle h, int32_t kind, int32_t cap_nodes, int32_t cap_edges, double* rot, double* t, int32_t* c1, int32_t* c2,
                                     double* obs_rot, double* obs_t, double* weight, double* sw);
int pgs_facade_evaluate_alternative(pgs_facade_handle h, int32_t kind, int32_t cap_edges, double* r, double* J, double* cost);   /* r, J sized as in pgs_fourdof.h */
int32_t pgs_facade_n_odom_terms(pgs_facade_handle h);      /* count, or PGS_ERR_STATE while the solver thread runs */
int pgs_facade_get_odom_terms(pgs_facade_handle h, int32_t cap, int32_t* u, int32_t* umf, double* q, double* t, double* w);
int32_t pgs_facade_n_reg_terms(pgs_facade_handle h);
int pgs_facade_get_reg_terms(pgs_facade_handle h, int32_t cap, int32_t* node, double* q, double* t, double* w);
int32_t pgs_facade_which_world(pgs_facade_handle h, int64_t stamp_ns);
int32_t pgs_facade_n_worlds(pgs_facade_handle h);
int32_t pgs_facade_world_setid(pgs_facade_handle h, int32_t world);
int32_t pgs_facade_world_start(pgs_facade_handle h, int32_t world);
int32_t pgs_facade_world_end(pgs_facade_handle h, int32_t world);
int pgs_facade_pose_between_worlds(pgs_facade_handle h, int32_t m, int32_t n, double* m_T_n_rowmajor16); /* 1 ok, 0 unknown */

#ifdef __cplusplus
}
#endif
#endif
