// Node-range partition of a pose graph for the multi-GPU solve.
//
// Nodes are split into `world` contiguous ranges.  A node is a BORDER node when it has a neighbour in a lower
// range (it lies on the upper side of a cut and an edge crosses that cut); removing the border nodes
// disconnects the ranges, so every other node is INTERIOR to exactly one rank.  A residual block belongs to
// the rank whose interior holds one of its endpoints (unique), blocks between two border nodes and
// regularisers on border nodes to the rank whose range holds their lowest node.  Pure host code, no CUDA.
#pragma once
#include <vector>

namespace pgs {

struct Partition {
  int N = 0, world = 1;
  std::vector<int> cut;          // world + 1 entries; rank k's range is [cut[k], cut[k+1])
  std::vector<int> node_owner;   // per node: owning rank, -1 for a border node
  std::vector<int> border;       // border nodes, ascending
  std::vector<int> odom_owner, loop_owner, reg_owner;
  int range_of(int node) const;
};

// Odometry edge e couples (oc1[e], oc2[e]); loop edge e couples (la[e], lb[e]); regulariser k sits on rnode[k].
void make_partition(int N, int world, int n_odom, const int* oc1, const int* oc2, int n_loop, const int* la, const int* lb,
                    int n_reg, const int* rnode, Partition* out);

}  // namespace pgs
