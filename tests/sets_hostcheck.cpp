// TEST INFRASTRUCTURE ONLY — never linked into libpgs.so.
// Compiles the product's host-side union-find and Worlds classes (solve_keyframe_pose_graph_b200/csrc/host) into a
// small library with a C surface, so that tests/test_reference_sets.py can drive them op by op against the reference's
// own classes (oracle/_ref/libref_sets.so).
#include "../solve_keyframe_pose_graph_b200/csrc/host/Worlds.cpp"

using pgs::DisjointSetForest;
using pgs::Matrix4d;
using pgs::Worlds;

extern "C" {

void* ours_dsf_create() { return new DisjointSetForest(); }
void ours_dsf_destroy(void* p) { delete (DisjointSetForest*)p; }
void ours_dsf_add_element(void* p, int x) { ((DisjointSetForest*)p)->add_element(x); }
int ours_dsf_exists(void* p, int x) { return ((DisjointSetForest*)p)->exists(x) ? 1 : 0; }
int ours_dsf_element_count(void* p) { return ((DisjointSetForest*)p)->element_count(); }
int ours_dsf_set_count(void* p) { return ((DisjointSetForest*)p)->set_count(); }
int ours_dsf_find_set(void* p, int x) { return ((DisjointSetForest*)p)->find_set(x); }
void ours_dsf_union_sets(void* p, int x, int y) { ((DisjointSetForest*)p)->union_sets(x, y); }

void* ours_worlds_create() { return new Worlds(); }
void ours_worlds_destroy(void* p) { delete (Worlds*)p; }
void ours_worlds_world_starts(void* p, long long stamp) { ((Worlds*)p)->world_starts(stamp); }
int ours_worlds_set_pose(void* p, int m, int n, const double* T16) { Matrix4d T; for (int i = 0; i < 16; ++i) T.m[i] = T16[i]; return ((Worlds*)p)->setPoseBetweenWorlds(m, n, T, "test") ? 1 : 0; }
int ours_worlds_is_exist(void* p, int m, int n) { return ((Worlds*)p)->is_exist(m, n) ? 1 : 0; }
int ours_worlds_get_pose(void* p, int m, int n, double* T16) { bool ok = false; const Matrix4d T = ((Worlds*)p)->getPoseBetweenWorlds(m, n, &ok); for (int i = 0; i < 16; ++i) T16[i] = T.m[i]; return ok ? 1 : 0; }
int ours_worlds_find_setid(void* p, int i) { return ((Worlds*)p)->find_setID_of_world_i(i); }
int ours_worlds_n_keys(void* p) { std::vector<std::pair<int, int>> k; ((Worlds*)p)->getAllKeys(k); return (int)k.size(); }

}  // extern "C"
