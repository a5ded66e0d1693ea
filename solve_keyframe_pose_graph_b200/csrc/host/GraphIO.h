// On-disk formats of the reference, so that graphs recorded by the reference node can be
// replayed through this solver and results written here can be read by the reference's tooling:
//   log_posegraph.json        NodeDataManager::saveAsJSON / loadFromJSON   (src/NodeDataManager.cpp:503-754)
//   log_optimized_poses.json  PoseGraphSLAM::saveAsJSON                     (src/PoseGraphSLAM.cpp:1111-1207)
//   solved_posegraph.json     Composer::saveStateToDisk: SolvedPoseGraph, KidnapTimestamps and WorldsData
//                             (src/Composer.cpp:952-1106, src/NodeDataManager.cpp:854-888, src/Worlds.cpp:442-497)
// Matrices are strings: "a,b,c,d;e,f,g,h;..." (Eigen IOFormat(FullPrecision, DontAlignCols, ",", ";")) in the two
// log files, "a, b, c, d\n..." inside {"rows","cols","data"} in solved_posegraph.json (src/utils/RawFileIO.h:95-106).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "Composer.h"
#include "Json.h"
#include "NodeDataManager.h"
#include "PoseGraphSLAM.h"

namespace pgs {

// Eigen's FullPrecision for double = 16 significant digits; default float formatting (%.16g)
std::string mat_to_string(const Matrix4d& M, const char* coeff_sep = ",", const char* row_sep = ";");
// accepts both layouts above (src/utils/PoseManipUtils.cpp:272-295, src/utils/RawFileIO.cpp:372-409)
bool string_to_mat(const std::string& s, Matrix4d& M);
// ":YPR(deg)=(y,p,r)  :TxTyTz=(x,y,z)" with %4.3f (src/utils/PoseManipUtils.cpp:206-215)
std::string prettyprintMatrix4d(const Matrix4d& M);

bool saveAsJSON(const NodeDataManager& manager, const std::string& base_path, std::string* err = nullptr);            // -> base_path/log_posegraph.json
// Loads nodes and loop edges like the reference (edge_mask: empty = all).  restore_kidnaps additionally replays
// "kidnap_info" so that which_world_is_this() answers as it did when the file was written (the reference restores
// those from solved_posegraph.json instead, Composer.cpp:1137-1151).
bool loadFromJSON(NodeDataManager& manager, const std::string& base_path, const std::vector<bool>& edge_mask = {}, bool restore_kidnaps = true,
                  std::string* err = nullptr);
bool saveAsJSON(const PoseGraphSLAM& slam, const NodeDataManager& manager, const std::string& base_path, std::string* err = nullptr);   // -> log_optimized_poses.json
bool saveSolvedPoseGraph(const Composer* composer, const NodeDataManager& manager, const std::string& save_dir_path, std::string* err = nullptr);   // -> solved_posegraph.json (composer may be null)
struct SolvedPoseGraph { std::vector<Matrix4d> w_T_c; std::vector<int64_t> stamp_ns; std::vector<int> world_id, set_id; std::vector<int64_t> kidnap_starts, kidnap_ends; };
bool loadSolvedPoseGraph(const std::string& json_file, SolvedPoseGraph* out, std::string* err = nullptr);

}  // namespace pgs
