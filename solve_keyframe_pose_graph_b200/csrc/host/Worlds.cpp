#include "Worlds.h"

#include <algorithm>
#include <deque>

namespace pgs {

Matrix4d Worlds::getPoseBetweenWorlds(int m, int n, bool* ok) const {
  if (ok) *ok = true;
  if (m == n) return Matrix4d::Identity();
  if (!is_exist(m, n)) { if (ok) *ok = false; return Matrix4d::Identity(); }
  std::vector<int> path;
  Matrix4d ans = Matrix4d::Identity();
  {
    std::lock_guard<std::mutex> lk(mutex_world);
    auto direct = rel_pose.find({m, n});
    if (direct != rel_pose.end()) return direct->second;
    auto rev = rel_pose.find({n, m});
    if (rev != rel_pose.end()) return rev->second.inverse();

    // Neither orientation is stored: chain the known pairs of this set.  Adjacency lists are filled in
    // the map's key order, both directions per key, and the BFS is rooted at n, exactly as the reference
    // does (Worlds.cpp:62-92), so the same path is found when several exist.
    const int setID = disjoint_set.find_set(m);
    const int W = disjoint_set.element_count();
    std::vector<std::vector<int>> adj(W);
    for (const auto& kv : rel_pose) {
      const int a = kv.first.first, b = kv.first.second;
      if (a < 0 || b < 0 || a >= W || b >= W) continue;
      if (disjoint_set.find_set(a) != setID || disjoint_set.find_set(b) != setID) continue;
      adj[a].push_back(b); adj[b].push_back(a);
    }
    std::vector<int> parent(W, -1); std::vector<char> seen(W, 0);
    std::deque<int> queue; queue.push_back(n); seen[n] = 1; parent[n] = -2;
    while (!queue.empty()) {
      const int s = queue.front(); queue.pop_front();
      for (int v : adj[s]) if (!seen[v]) { seen[v] = 1; parent[v] = s; queue.push_back(v); }
    }
    if (!seen[m]) { if (ok) *ok = false; return Matrix4d::Identity(); }
    for (int v = m, guard = 0; guard < 100; ++guard) {   // the reference caps the walk at 100 hops (MyDirectionalGraph.h:80)
      path.push_back(v);
      if (parent[v] == -2) break;
      v = parent[v];
    }
    for (size_t h = 0; h + 1 < path.size(); ++h) {
      auto f = rel_pose.find({path[h], path[h + 1]});
      if (f != rel_pose.end()) ans = ans * f->second;
      else {
        auto r = rel_pose.find({path[h + 1], path[h]});
        if (r == rel_pose.end()) { if (ok) *ok = false; return Matrix4d::Identity(); }
        ans = ans * r->second.inverse();
      }
    }
  }
  // memoise under (path.front(), path.back()) == (m, n)   (Worlds.cpp:137)
  const_cast<Worlds*>(this)->setPoseBetweenWorlds(path.front(), path.back(), ans, "pose set by inference with BFS");
  return ans;
}

bool Worlds::setPoseBetweenWorlds(int m, int n, const Matrix4d& m_T_n, const std::string& info) {
  std::lock_guard<std::mutex> lk(mutex_world);
  if (!disjoint_set.exists(m) || !disjoint_set.exists(n)) return false;
  rel_pose[{m, n}] = m_T_n;
  rel_pose_info[{m, n}] += ";" + info;
  // larger id first: on rank ties the smaller id stays the root (Worlds.cpp:168)
  disjoint_set.union_sets(std::max(m, n), std::min(m, n));
  log_ += "union_sets:" + std::to_string(std::max(m, n)) + "," + std::to_string(std::min(m, n)) + ";";
  return true;
}

bool Worlds::is_exist(int m, int n) const {
  if (m < 0 || n < 0) return false;
  if (m == n) return true;
  if (m >= n_worlds() || n >= n_worlds()) return false;
  std::lock_guard<std::mutex> lk(mutex_world);
  const int sm = disjoint_set.find_set(m), sn = disjoint_set.find_set(n);
  return sm >= 0 && sn >= 0 && sm == sn;
}

void Worlds::getAllKeys(std::vector<std::pair<int, int>>& keys) const {
  std::lock_guard<std::mutex> lk(mutex_world);
  keys.clear();
  for (const auto& kv : rel_pose) keys.push_back(kv.first);
}

void Worlds::getWorld2SetIDMap(std::map<int, int>& out) const {
  out.clear();
  const int W = n_worlds();
  for (int w = 0; w < W; ++w) out[w] = find_setID_of_world_i(w);
}

void Worlds::world_starts(int64_t stamp_ns) {
  std::lock_guard<std::mutex> lk(mutex_world);
  vec_world_starts.push_back(stamp_ns);
  const int id = (int)vec_world_starts.size() - 1;
  disjoint_set.add_element(id);
  log_ += "add_element:" + std::to_string(id) + ";";
}

void Worlds::world_ends(int64_t stamp_ns) {
  std::lock_guard<std::mutex> lk(mutex_world);
  vec_world_ends.push_back(stamp_ns);
}

int Worlds::find_setID_of_world_i(int i) const {
  std::lock_guard<std::mutex> lk(mutex_world);
  return disjoint_set.exists(i) ? disjoint_set.find_set(i) : -1;
}

int Worlds::n_worlds() const { std::lock_guard<std::mutex> lk(mutex_world); return disjoint_set.element_count(); }
int Worlds::n_sets() const { std::lock_guard<std::mutex> lk(mutex_world); return disjoint_set.set_count(); }
std::string Worlds::disjoint_set_log() const { std::lock_guard<std::mutex> lk(mutex_world); return log_; }

}  // namespace pgs
