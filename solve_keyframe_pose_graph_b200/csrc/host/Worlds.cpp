#include "Worlds.h"

#include <algorithm>
#include <cstdio>
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
  debug_ += "\t\t\tunion_sets( " + std::to_string(std::max(m, n)) + "," + std::to_string(std::min(m, n)) + ")\n";   // Worlds.cpp:169
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
  debug_ += "\t\t\tadd_element( " + std::to_string(id) + ")\n";                                                     // Worlds.cpp:238
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
// Worlds::disjoint_set_status (Worlds.cpp:333-363): what log_posegraph.json stores under "disjoint_set_status"
std::string Worlds::disjoint_set_status() const {
  std::lock_guard<std::mutex> lk(mutex_world);
  std::map<int, std::string> ff;
  std::string out = "element_count=" + std::to_string(disjoint_set.element_count()) + "   set_count=" + std::to_string(disjoint_set.set_count()) + ";";
  for (int i = 0; i < disjoint_set.element_count(); ++i) {
    const int setID = disjoint_set.find_set(i);
    out += "world#" + std::to_string(i) + " is in setID=" + std::to_string(setID) + ";";
    if (ff.count(setID)) ff[setID] += "," + std::to_string(i); else ff[setID] = std::to_string(i);
  }
  out += ";";
  for (const auto& kv : ff) out += "set#" + std::to_string(kv.first) + " contains worlds: " + kv.second + ";";
  return out;
}

// ---------------------------------------------------------------- state file (Worlds.cpp:442-640)
static std::string mat_rows_string(const Matrix4d& M) {   // RawFileIO::eigen_matrix_to_json: ", " between coefficients, "\n" between rows
  std::string out; char b[40];
  for (int r = 0; r < 4; ++r) { for (int c = 0; c < 4; ++c) { snprintf(b, sizeof(b), "%.16g", M(r, c)); out += b; if (c < 3) out += ", "; } if (r < 3) out += "\n"; }
  return out;
}
static bool parse_mat_rows(const std::string& s, Matrix4d& M) {
  std::vector<double> v; std::string tok;
  auto flush = [&]() -> bool { if (tok.empty()) return true; try { v.push_back(std::stod(tok)); } catch (...) { return false; } tok.clear(); return true; };
  for (char c : s) { if (c == ',' || c == ';' || c == '\n') { if (!flush()) return false; } else if (c != ' ' && c != '\t' && c != '\r') tok.push_back(c); }
  if (!flush() || v.size() != 16) return false;
  for (int i = 0; i < 16; ++i) M.m[i] = v[i];
  return true;
}

Json Worlds::saveStateToDisk() const {
  std::lock_guard<std::mutex> lk(mutex_world);
  Json obb;
  // nlohmann leaves a key that was only ever indexed, never pushed to, as null: empty lists are written as null (Worlds.cpp:451-497)
  obb["rel_pose_between_worlds__wb_T_wa"] = rel_pose.empty() ? Json() : Json::array();
  for (const auto& kv : rel_pose) {
    Json item;
    item["node_b"] = Json(kv.first.first); item["node_a"] = Json(kv.first.second);
    item["wb_T_wa"]["rows"] = Json(4); item["wb_T_wa"]["cols"] = Json(4);
    item["wb_T_wa"]["data"] = Json(mat_rows_string(kv.second));
    double ypr[3]; R2ypr(kv.second, ypr); char b[200];
    snprintf(b, sizeof(b), ":YPR(deg)=(%4.3f,%4.3f,%4.3f)  :TxTyTz=(%4.3f,%4.3f,%4.3f)", ypr[0], ypr[1], ypr[2], kv.second(0, 3), kv.second(1, 3), kv.second(2, 3));
    item["wb_T_wa"]["data_pretty"] = Json(std::string(b));
    auto info = rel_pose_info.find(kv.first);
    item["info_wb_T_wa"] = Json(info == rel_pose_info.end() ? std::string() : info->second);
    obb["rel_pose_between_worlds__wb_T_wa"].push_back(item);
  }
  Json A = Json::array(), B = Json::array();
  for (int64_t t : vec_world_starts) { Json a; a["stampNSec"] = Json(t); A.push_back(a); }
  for (int64_t t : vec_world_ends) { Json b; b["stampNSec"] = Json(t); B.push_back(b); }
  obb["vec_world_starts"] = vec_world_starts.empty() ? Json() : A; obb["vec_world_ends"] = vec_world_ends.empty() ? Json() : B;
  obb["disjoint_set"]["debug_string"] = Json(debug_);
  obb["disjoint_set"]["log_string"] = Json(log_);
  return obb;
}

bool Worlds::loadStateFromDisk(const Json& o, std::string* err) {
  auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
  std::lock_guard<std::mutex> lk(mutex_world);
  if (disjoint_set.element_count() != 0 || !rel_pose.empty()) return fail("Worlds::loadStateFromDisk: not empty");
  const Json& rp = o.at("rel_pose_between_worlds__wb_T_wa");
  for (size_t i = 0; i < rp.size(); ++i) {
    Matrix4d T;
    if (!parse_mat_rows(rp[i].at("wb_T_wa").at("data").as_string(), T)) return fail("Worlds::loadStateFromDisk: bad wb_T_wa");
    const std::pair<int, int> p((int)rp[i].at("node_b").as_int(), (int)rp[i].at("node_a").as_int());
    rel_pose[p] = T; rel_pose_info[p] = rp[i].at("info_wb_T_wa").as_string();
  }
  // replay the union-find op-log: "add_element:0;add_element:1;union_sets:1,0;" (Worlds.cpp:549-620)
  const std::string log = o.at("disjoint_set").at("log_string").as_string();
  size_t p0 = 0;
  while (p0 < log.size()) {
    size_t p1 = log.find(';', p0); if (p1 == std::string::npos) p1 = log.size();
    const std::string cmd = log.substr(p0, p1 - p0); p0 = p1 + 1;
    if (cmd.size() < 4) break;
    const size_t colon = cmd.find(':');
    if (colon == std::string::npos) return fail("Worlds::loadStateFromDisk: bad op-log entry '" + cmd + "'");
    const std::string op = cmd.substr(0, colon), arg = cmd.substr(colon + 1);
    try {
      if (op == "add_element") disjoint_set.add_element(std::stoi(arg));
      else if (op == "union_sets") {
        const size_t comma = arg.find(',');
        if (comma == std::string::npos) return fail("Worlds::loadStateFromDisk: bad union_sets operands");
        const int x = std::stoi(arg.substr(0, comma)), y = std::stoi(arg.substr(comma + 1));
        disjoint_set.union_sets(std::max(x, y), std::min(x, y));
      } else return fail("Worlds::loadStateFromDisk: unknown op '" + op + "'");
    } catch (...) { return fail("Worlds::loadStateFromDisk: bad number in op-log"); }
  }
  log_ = log;
  if (o.at("disjoint_set").contains("debug_string")) debug_ = o.at("disjoint_set").at("debug_string").as_string();   // Worlds.cpp:557
  const Json& ws = o.at("vec_world_starts"); const Json& we = o.at("vec_world_ends");
  for (size_t i = 0; i < ws.size(); ++i) vec_world_starts.push_back(ws[i].at("stampNSec").as_int());
  for (size_t i = 0; i < we.size(); ++i) vec_world_ends.push_back(we[i].at("stampNSec").as_int());
  return true;
}

}  // namespace pgs
