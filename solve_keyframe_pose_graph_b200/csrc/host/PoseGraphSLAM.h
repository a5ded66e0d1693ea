// ROS-free drop-in for the reference's PoseGraphSLAM class (reference src/PoseGraphSLAM.h:57-223): same
// constructor, solver-thread entry point, status codes and thread-safe getters, but the body of
// reinit_ceres_problem_onnewloopedge_optimize6DOF() builds the problem through the C-ABI of
// include/pgs.h and solves it on the B200 instead of calling ceres::Solve.
//
// Additions named by BASELINE.json's north_star (absent in the reference): addOdometryEdge / addLoopEdge
// for explicit graphs, and solve_once() = one trigger body without the 0.5 Hz sleep loop.
#pragma once
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../../include/pgs.h"
#include "NodeDataManager.h"
#include "pose_math.h"

namespace pgs {

struct PoseGraphSLAMOptions {
  int odom_fanout = 5;              // f = 1..5, reference PoseGraphSLAM.cpp:1577
  double odom_decay = 0.9;          // pow(0.9, f), :1604
  double odom_yaw_divisor = 6.0;    // exp(-yaw_deg^2 / 6), :1606
  double min_reg_weight = 1.1;      // max(1.1, log(1 + end - start)/2), :1839
  bool derive_odometry = true;      // false: only edges given through addOdometryEdge
  double loop_rate_hz = 0.5;        // ros::Rate loop_rate(0.5), :1257
  bool dry_run = false;             // build the problem and the initial guesses on the host, skip the device
  pgs_options solver;               // Ceres-equivalent + device options (pgs_default_options)
  PoseGraphSLAMOptions() { pgs_default_options(&solver); }
};

class PoseGraphSLAM {
 public:
  explicit PoseGraphSLAM(NodeDataManager* _manager, const PoseGraphSLAMOptions& options = PoseGraphSLAMOptions());
  ~PoseGraphSLAM();

  // Intended to run on its own thread: polls the manager at loop_rate_hz and solves when new loop edges
  // arrived (PoseGraphSLAM.cpp:1251-1950).
  void reinit_ceres_problem_onnewloopedge_optimize6DOF();
  void reinit_ceres_problem_onnewloopedge_optimize6DOF_enable() { isEnabled = true; }
  void reinit_ceres_problem_onnewloopedge_optimize6DOF_disable() { isEnabled = false; }
  // -1 nothing happening, 0 sleeping, 1 setting up, 2 solve in progress, 3 solve finished (PoseGraphSLAM.h:100-105)
  int get_reinit_ceres_problem_onnewloopedge_optimize6DOF_status() { return status; }
  void set_loop_rate_hz(double hz) { if (hz > 0) opt_.loop_rate_hz = hz; }   // the reference hard-codes 0.5 Hz (:1257)
  int n_solves() const { return n_solves_; }

  // One wake-up of the loop above.  Returns true if a solve was triggered.  force = solve even without
  // new loop edges.  On failure returns false and last_error() is non-empty.
  bool solve_once(bool force = false);
  // After a session was restored into the manager (loadFromJSON + Worlds state): one optimisation variable per loaded
  // keyframe, initialised to ws_T_w * w_T_c (its pose in the frame of its world's set root) and marked CONSTANT, and
  // solvedUntil moved to the last of them (reference PoseGraphSLAM.cpp:40-170).  Later keyframes are optimised against
  // this fixed backbone.  false: no keyframes, or a world whose set transform is unknown (the reference exits).
  bool load_state();
  bool saveAsJSON(const std::string base_path) const;      // log_optimized_poses.json (PoseGraphSLAM.cpp:1111-1207); defined in GraphIO.cpp
  std::string last_error() const { std::lock_guard<std::mutex> lk(mutex_error_); return error_; }   // a copy: the solver thread may rewrite it

  // Explicit-graph API (north_star).  Poses are 4x4: a_T_b for odometry (binds SixDOFError(a, b)),
  // b_T_a for loop edges (stored in the manager, bound as (b, a, switch)).
  bool addOdometryEdge(int a, int b, const Matrix4d& a_T_b, double weight);
  bool addLoopEdge(int a, int b, const Matrix4d& b_T_a, double weight);

  // ---- thread-safe getters (PoseGraphSLAM.cpp:178-224)
  const Matrix4d getNodePose(int i) const;
  bool nodePoseExists(int i) const;
  int nNodes() const;
  void getAllNodePose(std::vector<Matrix4d>& vec_w_T_ci) const;
  void getAllNodeRaw(std::vector<double>& quat_xyzw, std::vector<double>& t) const;   // the optimiser's own storage, one lock (used by Composer)
  int solvedUntil() const { std::lock_guard<std::mutex> lk(mutex_opt_vars); return solved_until; }
  double get_loopedge_switching_variable_val(int i) const;      // i = loop-edge index; NaN for a bad index
  const std::tuple<int, int, float, std::string>& get_odomedge_residue_info(int i) const;
  int get_odomedge_residue_info_size() const;
  const std::tuple<int, int, float, std::string, std::string>& get_loopedge_residue_info(int i) const;
  int get_loopedge_residue_info_size() const;

  // ---- introspection for tests / bench
  pgs_summary last_summary() const { std::lock_guard<std::mutex> lk(mutex_summary_); return summary_; }
  std::vector<pgs_iteration> last_iterations() const { std::lock_guard<std::mutex> lk(mutex_summary_); return iterations_; }
  struct RegTerm { int node; Matrix4d anchor; double weight; };
  const std::vector<RegTerm>& regularization_terms() const { return reg_terms_; }
  struct OdomTerm { int u, umf; double q[4], t[3], weight; };
  const std::vector<OdomTerm>& odometry_terms() const { return odom_terms_; }
  pgs_handle device_handle() { return handle_; }
  // The residual blocks the reference's SWITCHED-OFF builds would add for the same session, in the
  // array form include/pgs_fourdof.h takes.  kind 0: FourDOFError::Create(u_M_umf, odom_edge_weight) on (u, u-f), the
  // commented call at PoseGraphSLAM.cpp:1630; kind 1: FourDOFErrorWithSwitchingConstraints::Create(bTa, weight) on
  // (second, first, switch e), :1551; kind 2: the __USE_YPR_REP build — QinFourDOFWeightError on (ypr, t) variables for
  // odometry (:1608-1626) followed by loop edges (:1389-1396,1534-1548; pitch / roll are read from paur.first there).
  struct AlternativeTerms {
    int kind = 0, n_nodes = 0;
    std::vector<double> rot, t;                      // [4|3 n_nodes], [3 n_nodes]: the optimisation variables
    std::vector<int32_t> c1, c2;
    std::vector<double> obs_rot, obs_t, weight, sw;  // per block
  };
  bool alternative_terms(int kind, AlternativeTerms& out) const;

 private:
  bool fail(const std::string& msg) { { std::lock_guard<std::mutex> lk(mutex_error_); error_ = msg; } status = 0; return false; }
  void clear_error() { std::lock_guard<std::mutex> lk(mutex_error_); error_.clear(); }
  void allocate_and_append_new_opt_variable_withpose(const Matrix4d& pose);
  bool update_opt_variable_with(int i, const Matrix4d& pose);
  void allocate_and_append_new_edge_switch_var();
  int n_opt_variables() const;
  int n_opt_switch() const;

  NodeDataManager* manager;
  PoseGraphSLAMOptions opt_;
  std::atomic<bool> isEnabled;
  std::atomic<int> status;
  std::atomic<int> n_solves_{0};
  mutable std::mutex mutex_error_;
  std::string error_;

  mutable std::mutex mutex_opt_vars;
  std::vector<double> _opt_quat_;    // x,y,z,w per node (PoseGraphSLAM.h:153)
  std::vector<double> _opt_t_;
  std::vector<double> _opt_switch_;  // one per loop edge of the manager
  int solved_until = 0;

  mutable std::mutex mutex_residue_info;
  std::vector<std::tuple<int, int, float, std::string>> odometry_edges_terms;
  std::vector<std::tuple<int, int, float, std::string, std::string>> loop_edges_terms;

  // trigger state that persists across wake-ups (locals of the reference's thread function)
  int prev_loopedge_len = 0, prev_node_len = 0;
  std::map<int, std::tuple<int, int>> changes_to_setid_on_set_union;
  std::vector<RegTerm> reg_terms_;
  std::vector<OdomTerm> odom_terms_;
  mutable std::mutex mutex_pending_;          // addOdometryEdge may run on the ingest thread while the solver thread drains the list
  std::vector<OdomTerm> pending_explicit_odom_;
  std::vector<int> loop_slot_;       // manager loop-edge index -> device loop-edge index (-1: skipped, dead zone)
  int n_device_nodes_ = 0, n_device_loops_ = 0, n_device_odom_ = 0;   // how much of the lists below the device already holds
  // Loop-closure blocks in device (slot) order.  The block lists only ever grow, and the device is brought up to them at the
  // start of every solve, so a trigger whose device update or solve fails adds nothing twice when it is tried again.
  std::vector<int> loop_a_, loop_b_; std::vector<double> loop_q_, loop_t_, loop_w_;
  int loops_taken_until_ = 0;        // manager loop edges already turned into blocks
  int odom_added_until_ = 0;         // keyframes whose odometry blocks exist (== solvedUntil() after a successful solve)
  bool retry_pending_ = false;       // the last trigger built its blocks but the device failed: the next wake-up solves again
  int odom_scanned_until_ = 0;       // odometry edges exist for u <= this
  int n_constant_ = 0;               // variables [0, n_constant_) are constant blocks (load_state)
  int n_constant_on_device_ = 0;

  pgs_handle handle_ = nullptr;
  mutable std::mutex mutex_summary_;   // summary_ / iterations_ are published by the solver thread at the end of a solve
  pgs_summary summary_{};
  std::vector<pgs_iteration> iterations_;
};

}  // namespace pgs
