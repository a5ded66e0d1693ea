// Relative poses between odometry worlds + set merging (ROS-free restatement of the reference's
// Worlds class: src/Worlds.h:50-79, src/Worlds.cpp:6-275).  Timestamps are int64 nanoseconds.
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "DisjointSet.h"
#include "Json.h"
#include "pose_math.h"

namespace pgs {

class Worlds {
 public:
  // m_T_n between two worlds of the same set.  Direct entry, inverse of (n,m), else the product along
  // the BFS path over known pairs (memoised).  The reference's BFS branch falls off the end without a
  // return (Worlds.cpp:137-149, undefined behaviour); the intended value `ans` is returned here.
  // ok=false replaces the reference's exit(5)/exit(2).
  Matrix4d getPoseBetweenWorlds(int m, int n, bool* ok = nullptr) const;
  bool setPoseBetweenWorlds(int m, int n, const Matrix4d& m_T_n, const std::string& info);
  bool is_exist(int m, int n) const;
  void getAllKeys(std::vector<std::pair<int, int>>& keys) const;
  void getWorld2SetIDMap(std::map<int, int>& out) const;
  void world_starts(int64_t stamp_ns);
  void world_ends(int64_t stamp_ns);
  int find_setID_of_world_i(int i) const;
  int n_worlds() const;
  int n_sets() const;
  std::string disjoint_set_status() const;   // "element_count=2   set_count=1;world#0 is in setID=0;..." (Worlds.cpp:333-363)
  std::string disjoint_set_log() const;   // "add_element:0;union_sets:1,0;" op-log (Worlds.cpp:164-170,230-240)
  // The "WorldsData" object of solved_posegraph.json (Worlds.cpp:442-497): relative poses with their info strings,
  // world start / end stamps and the union-find op-log, which loadStateFromDisk replays (Worlds.cpp:499-640).
  Json saveStateToDisk() const;
  bool loadStateFromDisk(const Json& obj, std::string* err = nullptr);   // into an empty Worlds

 private:
  mutable std::mutex mutex_world;
  mutable std::map<std::pair<int, int>, Matrix4d> rel_pose;   // (m,n) -> m_T_n
  mutable std::map<std::pair<int, int>, std::string> rel_pose_info;
  std::vector<int64_t> vec_world_starts, vec_world_ends;
  mutable DisjointSetForest disjoint_set;
  mutable std::string log_;
  std::string debug_;                      // disjoint_set_debug of the reference (Worlds.cpp:169,238), saved as "debug_string"
};

}  // namespace pgs
