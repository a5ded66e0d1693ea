// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C wrapper around the REFERENCE'S OWN set-bookkeeping classes, compiled from where they lie under /root/reference
// (never copied into this repository): mp::DisjointSetForest<int> (src/utils/DisjointSet.h) and MyDirectionalGraph
// (src/utils/MyDirectionalGraph.h).  They are the only part of the reference's hot path that builds without Ceres,
// Eigen, ROS or OpenCV; `make -C oracle` turns them into oracle/_ref/libref_sets.so when /root/reference is present.
// tests/test_reference_sets.py checks the product's DisjointSet.h / Worlds.cpp and the Python front-end restatement
// against this library, which pins the rule that decides every set root and the BFS path the
// reference's Worlds::getPoseBetweenWorlds follows (src/Worlds.cpp:62-100) to the real code.
#include <vector>

#include "utils/DisjointSet.h"
#include "utils/MyDirectionalGraph.h"

extern "C" {

typedef mp::DisjointSetForest<int> Dsf;

void* ref_dsf_create() { return new Dsf(); }
void ref_dsf_destroy(void* p) { delete (Dsf*)p; }
void ref_dsf_add_element(void* p, int x) { ((Dsf*)p)->add_element(x, 0); }
int ref_dsf_exists(void* p, int x) { return ((Dsf*)p)->exists(x) ? 1 : 0; }
int ref_dsf_element_count(void* p) { return ((Dsf*)p)->element_count(); }
int ref_dsf_set_count(void* p) { return ((Dsf*)p)->set_count(); }
// the reference's get_element does a bare `throw;` for an unknown element (terminate): guard like its callers do
int ref_dsf_find_set(void* p, int x) { return ((Dsf*)p)->exists(x) ? ((Dsf*)p)->find_set(x) : -1; }
void ref_dsf_union_sets(void* p, int x, int y) { if (((Dsf*)p)->exists(x) && ((Dsf*)p)->exists(y)) ((Dsf*)p)->union_sets(x, y); }

void* ref_graph_create(int V) { return new MyDirectionalGraph(V); }
void ref_graph_destroy(void* p) { delete (MyDirectionalGraph*)p; }
void ref_graph_add_edge(void* p, int v, int w) { ((MyDirectionalGraph*)p)->add_edge(v, w); }
void ref_graph_bfs(void* p, int s) { ((MyDirectionalGraph*)p)->BFS(s); }
// path from v back to the BFS root; returns its length (0: v was not reached)
int ref_graph_get_path_from(void* p, int v, int* out, int cap) {
  std::vector<int> path;
  ((MyDirectionalGraph*)p)->get_path_from(v, path);
  for (int i = 0; i < (int)path.size() && i < cap; ++i) out[i] = path[i];
  return (int)path.size();
}

}  // extern "C"
