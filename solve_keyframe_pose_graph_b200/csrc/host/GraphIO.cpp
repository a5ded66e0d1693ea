#include "GraphIO.h"

#include <cstdio>
#include <fstream>
#include <sstream>

namespace pgs {

static bool set_err(std::string* err, const std::string& m) { if (err) *err = m; return false; }

std::string mat_to_string(const Matrix4d& M, const char* coeff_sep, const char* row_sep) {
  std::string s; char b[40];
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) { snprintf(b, sizeof(b), "%.16g", M(r, c)); s += b; if (c < 3) s += coeff_sep; }
    if (r < 3) s += row_sep;
  }
  return s;
}

bool string_to_mat(const std::string& s, Matrix4d& M) {
  std::vector<double> v; std::string tok;
  auto flush = [&]() -> bool { if (tok.empty()) return true; try { v.push_back(std::stod(tok)); } catch (...) { return false; } tok.clear(); return true; };
  for (char c : s) {
    if (c == ',' || c == ';' || c == '\n') { if (!flush()) return false; }
    else if (c != ' ' && c != '\t' && c != '\r') tok.push_back(c);
  }
  if (!flush() || v.size() != 16) return false;
  for (int i = 0; i < 16; ++i) M.m[i] = v[i];
  return true;
}

std::string prettyprintMatrix4d(const Matrix4d& M) {
  double ypr[3]; R2ypr(M, ypr);
  char b[200];
  snprintf(b, sizeof(b), ":YPR(deg)=(%4.3f,%4.3f,%4.3f)  :TxTyTz=(%4.3f,%4.3f,%4.3f)", ypr[0], ypr[1], ypr[2], M(0, 3), M(1, 3), M(2, 3));
  return b;
}

static bool write_file(const std::string& path, const std::string& text, std::string* err) {
  std::ofstream f(path);
  if (!f.is_open()) return set_err(err, "cannot open " + path + " for writing");
  f << text << std::endl;
  return (bool)f;
}
static bool read_file(const std::string& path, std::string* text, std::string* err) {
  std::ifstream f(path);
  if (!f.is_open()) return set_err(err, "cannot open " + path);
  std::stringstream ss; ss << f.rdbuf(); *text = ss.str();
  return true;
}
static double to_sec(int64_t ns) { return (double)(ns / 1000000000LL) + 1e-9 * (double)(ns % 1000000000LL); }   // ros::Time::toSec
static int64_t from_sec(double s) { const int64_t sec = (int64_t)std::floor(s); return sec * 1000000000LL + (int64_t)std::llround((s - (double)sec) * 1e9); }   // ros::Time(double)

// ------------------------------------------------------------------ log_posegraph.json (NodeDataManager.cpp:503-628)
bool saveAsJSON(const NodeDataManager& m, const std::string& base_path, std::string* err) {
  Json all;
  all["meta_data"]["getNodeLen"] = Json(m.getNodeLen());
  all["meta_data"]["getEdgeLen"] = Json(m.getEdgeLen());
  auto cov_string = [](const double* c36) { std::string s; char b[40];
    for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) { snprintf(b, sizeof(b), "%.16g", c36[6 * r + c]); s += b; if (c < 5) s += ","; } if (r < 5) s += ";"; } return s; };
  all["nodes"] = Json::array();
  for (int i = 0; i < m.getNodeLen(); ++i) {
    Json node;
    node["timestamp"] = Json(to_sec(m.getNodeTimestamp(i)));
    node["idx"] = Json(i);
    node["world_id"] = Json(m.which_world_is_this(m.getNodeTimestamp(i)));
    const Matrix4d& wTc = m.getNodePose(i);
    node["wTc"] = Json(mat_to_string(wTc));
    node["wTc_pretty"] = Json(prettyprintMatrix4d(wTc));
    double cov[36] = {}; m.getNodeCov(i, cov);
    node["cov"] = Json(cov_string(cov));                // the reference serialises cov BEFORE reading it (:531-535): its files hold garbage here
    all["nodes"].push_back(node);
  }
  all["loopedges"] = Json::array();
  for (int i = 0; i < m.getEdgeLen(); ++i) {
    Json e;
    const std::pair<int, int> p = m.getEdgeIdxInfo(i);
    e["idx0"] = Json(p.first); e["idx1"] = Json(p.second);
    e["timestamp0"] = Json(to_sec(m.getNodeTimestamp(p.first))); e["timestamp1"] = Json(to_sec(m.getNodeTimestamp(p.second)));
    const int w0 = m.which_world_is_this(m.getNodeTimestamp(p.first)), w1 = m.which_world_is_this(m.getNodeTimestamp(p.second));
    e["world0_id"] = Json(w0); e["world1_id"] = Json(w1);
    e["code"] = Json((w0 < 0 || w1 < 0) ? -1 : (w0 == w1 ? 1 : 2));
    e["b_T_a"] = Json(mat_to_string(m.getEdgePose(i)));
    e["b_T_a_pretty"] = Json(prettyprintMatrix4d(m.getEdgePose(i)));
    e["weight"] = Json(m.getEdgeWeight(i));
    e["description"] = Json(m.getEdgeDescriptionString(i));
    all["loopedges"].push_back(e);
  }
  all["world_info"] = Json::array();
  for (int i = 0; i < m.n_worlds(); ++i) {
    Json w; w["id"] = Json(i);
    w["nodeidx_of_world_i_started"] = Json(m.nodeidx_of_world_i_started(i));
    w["nodeidx_of_world_i_ended"] = Json(m.nodeidx_of_world_i_ended(i));
    all["world_info"].push_back(w);
  }
  all["meta_data"]["n_worlds"] = Json(m.n_kidnaps());   // sic: the reference overwrites n_worlds with n_kidnaps (:599)
  all["kidnap_info"] = m.n_kidnaps() ? Json::array() : Json();   // never pushed to -> null in the reference's file
  for (int i = 0; i < m.n_kidnaps(); ++i) {
    Json k; k["idx"] = Json(i);
    k["stamp_of_kidnap_i_started"] = Json(to_sec(m.stamp_of_kidnap_i_started(i)));
    k["stamp_of_kidnap_i_ended"] = Json(to_sec(m.stamp_of_kidnap_i_ended(i)));
    // extension (ignored by the reference's loader): exact nanosecond stamps, doubles lose them above 2^53 ns
    k["stampNSec_started"] = Json(m.stamp_of_kidnap_i_started(i)); k["stampNSec_ended"] = Json(m.stamp_of_kidnap_i_ended(i));
    all["kidnap_info"].push_back(k);
  }
  all["disjoint_set_status"] = Json(m.getWorldsConstPtr()->disjoint_set_status());   // NodeDataManager.cpp:611
  return write_file(base_path + "/log_posegraph.json", all.dump(4), err);
}

bool loadFromJSON(NodeDataManager& m, const std::string& base_path, const std::vector<bool>& edge_mask, bool restore_kidnaps, std::string* err) {
  if (m.getNodeLen() != 0 || m.getEdgeLen() != 0) return set_err(err, "loadFromJSON: the manager must be empty (the reference resets it, :642-643)");
  std::string text;
  if (!read_file(base_path + "/log_posegraph.json", &text, err)) return false;
  Json all; std::string perr;
  if (!Json::parse(text, &all, &perr)) return set_err(err, "log_posegraph.json: " + perr);
  const Json& nodes = all.at("nodes"); const Json& edges = all.at("loopedges");
  if (all.at("meta_data").at("getEdgeLen").as_int() != (int64_t)edges.size() || all.at("meta_data").at("getNodeLen").as_int() != (int64_t)nodes.size())
    return set_err(err, "The meta data and the json file is not consistant");          // :660-667
  // kidnap signals in time order, interleaved with the nodes exactly as they arrived
  std::vector<std::pair<int64_t, int>> ev;
  if (restore_kidnaps) {
    const Json& ki = all.at("kidnap_info");
    for (size_t i = 0; i < ki.size(); ++i) {
      const int64_t s = ki[i].contains("stampNSec_started") ? ki[i].at("stampNSec_started").as_int() : from_sec(ki[i].at("stamp_of_kidnap_i_started").as_double());
      const int64_t e = ki[i].contains("stampNSec_ended") ? ki[i].at("stampNSec_ended").as_int() : from_sec(ki[i].at("stamp_of_kidnap_i_ended").as_double());
      ev.push_back({s, 1}); ev.push_back({e, 0});
    }
  }
  size_t k = 0;
  std::vector<int64_t> stamps(nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) {
    const int64_t st = nodes[i].contains("stampNSec") ? nodes[i].at("stampNSec").as_int() : from_sec(nodes[i].at("timestamp").as_double());
    stamps[i] = st;
    while (k < ev.size() && ev[k].first < st) { m.rcvd_kidnap_indicator(ev[k].first, ev[k].second != 0); ++k; }
    Matrix4d wTc;
    if (!string_to_mat(nodes[i].at("wTc").as_string(), wTc)) return set_err(err, "node " + std::to_string(i) + ": wTc is not a 4x4 matrix string");
    double cov[36] = {}; bool have_cov = false;          // cov: 6x6 as "a,..;..." (PoseManipUtils.cpp:298-320); tolerate the reference's garbage
    if (nodes[i].contains("cov")) {
      std::vector<double> v; std::string tok;
      for (char c : nodes[i].at("cov").as_string() + ";") { if (c == ',' || c == ';') { if (!tok.empty()) { try { v.push_back(std::stod(tok)); } catch (...) { v.clear(); break; } tok.clear(); } } else if (c != ' ') tok.push_back(c); }
      if (v.size() == 36) { for (int k = 0; k < 36; ++k) cov[k] = v[k]; have_cov = true; }
    }
    m.add_node(st, wTc, have_cov ? cov : nullptr);
  }
  while (k < ev.size()) { m.rcvd_kidnap_indicator(ev[k].first, ev[k].second != 0); ++k; }
  for (size_t i = 0; i < edges.size(); ++i) {
    if (!edge_mask.empty() && i < edge_mask.size() && !edge_mask[i]) continue;          // :700-701
    const int idx0 = (int)edges[i].at("idx0").as_int(), idx1 = (int)edges[i].at("idx1").as_int();
    Matrix4d bTa;
    if (!string_to_mat(edges[i].at("b_T_a").as_string(), bTa)) return set_err(err, "edge " + std::to_string(i) + ": b_T_a is not a 4x4 matrix string");
    if (idx0 < 0 || idx1 < 0 || idx0 >= (int)nodes.size() || idx1 >= (int)nodes.size()) return set_err(err, "[Insonsistent json] edge index out of range");
    // the reference exits when the claimed timestamps differ from the nodes' (:736-747)
    if (std::llabs(from_sec(edges[i].at("timestamp0").as_double()) - stamps[idx0]) > 1000 || std::llabs(from_sec(edges[i].at("timestamp1").as_double()) - stamps[idx1]) > 1000)
      return set_err(err, "[Insonsistent json] node_timestamps[idx] != stamp of edge " + std::to_string(i));
    m.add_loop_edge_by_index(idx0, idx1, bTa, edges[i].at("weight").as_double(), edges[i].at("description").as_string());
  }
  return true;
}

// ------------------------------------------------------------------ log_optimized_poses.json (PoseGraphSLAM.cpp:1111-1207)
bool saveAsJSON(const PoseGraphSLAM& slam, const NodeDataManager& m, const std::string& base_path, std::string* err) {
  Json all;
  const int n = slam.nNodes();
  all["meta_data"]["nNodes"] = Json(n);
  all["PoseGraphSLAM_nodes"] = Json::array();
  for (int i = 0; i < n; ++i) {
    Json v;
    const Matrix4d opt = slam.getNodePose(i);
    v["wTc_opt"] = Json(mat_to_string(opt)); v["wTc_opt_prettyprint"] = Json(prettyprintMatrix4d(opt));
    const Matrix4d odom = i < m.getNodeLen() ? m.getNodePose(i) : Matrix4d::Identity();
    v["w_T_c_odom"] = Json(mat_to_string(odom)); v["w_T_c_odom_prettyprint"] = Json(prettyprintMatrix4d(odom));
    v["node_i"] = Json(i);
    all["PoseGraphSLAM_nodes"].push_back(v);
  }
  all["PoseGraphSLAM_loopedgeinfo"] = Json::array();
  for (int i = 0; i < m.getEdgeLen(); ++i) {
    Json e;
    const int a = m.getEdgeIdxInfo(i).first, b = m.getEdgeIdxInfo(i).second;
    e["getEdge_i"] = Json(i); e["a"] = Json(a); e["b"] = Json(b);
    e["world_of_a"] = Json(m.which_world_is_this(m.getNodeTimestamp(a))); e["world_of_b"] = Json(m.which_world_is_this(m.getNodeTimestamp(b)));
    e["weight"] = Json(m.getEdgeWeight(i)); e["description_string"] = Json(m.getEdgeDescriptionString(i));
    e["getEdgePose"] = Json(prettyprintMatrix4d(m.getEdgePose(i)));
    if (slam.nodePoseExists(a) && slam.nodePoseExists(b)) e["getEdgePose_after_opt"] = Json(prettyprintMatrix4d(slam.getNodePose(b).inverse() * slam.getNodePose(a)));
    const double s = slam.get_loopedge_switching_variable_val(i);
    if (s == s) e["switching_var_after_opt"] = Json(s);     // only for edges that own a switch (:1171-1172)
    all["PoseGraphSLAM_loopedgeinfo"].push_back(e);
  }
  return write_file(base_path + "/log_optimized_poses.json", all.dump(4), err);
}

// ------------------------------------------------------------------ solved_posegraph.json (Composer.cpp:990-1031)
bool saveSolvedPoseGraph(const Composer* composer, const NodeDataManager& m, const std::string& dir, std::string* err) {
  const std::vector<Matrix4d> lmb = composer ? composer->get_global_lmb() : std::vector<Matrix4d>();   // no pass yet: empty list, as global_lmb is
  Json obj;
  obj["SolvedPoseGraph"] = Json::array();
  for (size_t i = 0; i < lmb.size(); ++i) {
    Json node;
    node["w_T_c"]["rows"] = Json(4); node["w_T_c"]["cols"] = Json(4);
    node["w_T_c"]["data"] = Json(mat_to_string(lmb[i], ", ", "\n"));
    node["w_T_c"]["data_pretty"] = Json(prettyprintMatrix4d(lmb[i]));
    const int w = m.which_world_is_this(m.getNodeTimestamp((int)i));
    node["worldID"] = Json(w);
    node["setID_of_worldID"] = Json(m.getWorldsConstPtr()->find_setID_of_world_i(w));
    node["stampNSec"] = Json((int64_t)m.getNodeTimestamp((int)i));
    node["seq"] = Json((int)i);
    obj["SolvedPoseGraph"].push_back(node);
  }
  Json ks = Json::array(), ke = Json::array();                                      // NodeDataManager::kidnap_data_to_json (:854-888)
  const int nk_started = m.curr_kidnap_status() ? m.n_kidnaps() + 1 : m.n_kidnaps();
  for (int i = 0; i < nk_started; ++i) { Json a; a["stampNSec"] = Json((int64_t)m.stamp_of_kidnap_i_started(i)); ks.push_back(a); }
  for (int i = 0; i < m.n_kidnaps(); ++i) { Json b; b["stampNSec"] = Json((int64_t)m.stamp_of_kidnap_i_ended(i)); ke.push_back(b); }
  obj["KidnapTimestamps"]["kidnap_starts"] = ks; obj["KidnapTimestamps"]["kidnap_ends"] = ke;
  obj["WorldsData"] = m.getWorldsConstPtr()->saveStateToDisk();                   // Composer.cpp:1031
  return write_file(dir + "/solved_posegraph.json", obj.dump(4), err);
}

bool loadSolvedPoseGraph(const std::string& file, SolvedPoseGraph* out, std::string* err) {
  std::string text;
  if (!read_file(file, &text, err)) return false;
  Json obj; std::string perr;
  if (!Json::parse(text, &obj, &perr)) return set_err(err, file + ": " + perr);
  *out = SolvedPoseGraph();
  const Json& pg = obj.at("SolvedPoseGraph");
  for (size_t i = 0; i < pg.size(); ++i) {
    Matrix4d T;
    if (!string_to_mat(pg[i].at("w_T_c").at("data").as_string(), T)) return set_err(err, "SolvedPoseGraph[" + std::to_string(i) + "]: bad matrix");
    out->w_T_c.push_back(T); out->stamp_ns.push_back(pg[i].at("stampNSec").as_int());
    out->world_id.push_back((int)pg[i].at("worldID").as_int()); out->set_id.push_back((int)pg[i].at("setID_of_worldID").as_int());
  }
  const Json& kt = obj.at("KidnapTimestamps");
  for (size_t i = 0; i < kt.at("kidnap_starts").size(); ++i) out->kidnap_starts.push_back(kt.at("kidnap_starts")[i].at("stampNSec").as_int());
  for (size_t i = 0; i < kt.at("kidnap_ends").size(); ++i) out->kidnap_ends.push_back(kt.at("kidnap_ends")[i].at("stampNSec").as_int());
  return true;
}

// member spellings of the reference
bool NodeDataManager::saveAsJSON(const std::string& base_path) const { return pgs::saveAsJSON(*this, base_path, nullptr); }
bool NodeDataManager::loadFromJSON(const std::string& base_path, const std::vector<bool>& edge_mask) { return pgs::loadFromJSON(*this, base_path, edge_mask, true, nullptr); }
bool PoseGraphSLAM::saveAsJSON(const std::string base_path) const { return pgs::saveAsJSON(*this, *manager, base_path, nullptr); }

}  // namespace pgs
