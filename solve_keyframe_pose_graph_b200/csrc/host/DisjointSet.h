// Union-find over world ids with the exact rank rule of the reference's mp::DisjointSetForest
// (reference src/utils/DisjointSet.h:152-161 find_set with path compression, :185-190 union_sets,
// :241-257 link).  The rule decides which world becomes a set's root, which in turn decides the
// frame every node is initialised in and where regularisers go, so it is
// restated precisely: link(X, Y): rank[X] > rank[Y] ? parent[Y] = X : (parent[X] = Y, tie -> ++rank[Y]).
#pragma once
#include <map>

namespace pgs {

class DisjointSetForest {
 public:
  void add_element(int x) {
    if (nodes_.count(x)) return;
    nodes_[x] = Node{x, 0};
    ++sets_;
  }
  bool exists(int x) const { return nodes_.count(x) != 0; }
  int element_count() const { return (int)nodes_.size(); }
  int set_count() const { return sets_; }
  // -1 if x is unknown (the reference throws; its callers guard with exists()).
  int find_set(int x) const {
    auto it = nodes_.find(x);
    if (it == nodes_.end()) return -1;
    int root = x;
    while (nodes_.at(root).parent != root) root = nodes_.at(root).parent;
    int cur = x;   // path compression (cached mutable state, as in the reference)
    while (cur != root) { Node& n = nodes_.at(cur); const int next = n.parent; n.parent = root; cur = next; }
    return root;
  }
  void union_sets(int x, int y) {
    const int sx = find_set(x), sy = find_set(y);
    if (sx < 0 || sy < 0 || sx == sy) return;
    Node& X = nodes_.at(sx); Node& Y = nodes_.at(sy);
    if (X.rank > Y.rank) Y.parent = sx;
    else { X.parent = sy; if (X.rank == Y.rank) ++Y.rank; }
    --sets_;
  }

 private:
  struct Node { int parent; int rank; };
  mutable std::map<int, Node> nodes_;
  int sets_ = 0;
};

}  // namespace pgs
