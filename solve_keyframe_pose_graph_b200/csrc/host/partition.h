// Node-range plan of a pose graph for the sharded linear solve.
//
// Nodes are split into contiguous RANGES (keyframe order).  A node is a BORDER node when it has a neighbour in a
// lower range (it lies on the upper side of a cut and an edge crosses that cut); removing the border nodes
// disconnects the ranges, so every other node is INTERIOR to exactly one range.  The interior of a range is a
// CHAIN: it is eliminated on its own, in ascending ("up") or descending ("down") keyframe order, by the rank
// that owns the range; what is left is the Schur complement on the border nodes the chain touches.  A chain
// CARRIES the border nodes next to where its elimination starts (they stay in its front all the way) and MEETS
// the ones next to where it ends — so the first range goes up and the last one down (nothing to carry: "burn at
// both ends"), and ranges in between, which carry, are given fewer nodes.
//
// A residual block belongs to the chain whose interior holds one of its endpoints (unique); blocks between two
// border nodes and regularisers on border nodes to the chain of the range that holds their lowest node.
// Pure host code, no CUDA.
#pragma once
#include <vector>

namespace pgs {

struct PlanRange { int lo = 0, hi = 0, rank = 0; bool down = false; };

struct Partition {
  int N = 0, world = 1;
  std::vector<PlanRange> ranges;          // chain c eliminates the interior of ranges[c]
  std::vector<int> cut;                   // world + 1 entries; rank k's nodes are [cut[k], cut[k+1]) (a union of its ranges)
  std::vector<int> node_chain;            // per node: chain, -1 for a border node
  std::vector<int> node_owner;            // per node: owning rank, -1 for a border node
  std::vector<int> border;                // border nodes, ascending
  std::vector<int> border_index;          // per node: position in `border`, -1 for interior nodes
  std::vector<int> odom_chain, loop_chain, reg_chain;
  std::vector<int> odom_owner, loop_owner, reg_owner;   // rank of the owning chain
  std::vector<std::vector<int>> chain_border;   // per chain: positions in `border` of the border nodes its factor holds, ascending
  std::vector<int> border_env;            // per border position: first border position of its row envelope in the border system
};

// Time per eliminated node of a range that carries a separator relative to one that does not.  The carried separator
// roughly triples the flops of a panel's trailing update; a free range is bound by its panel chain (diagonal block +
// panel solve).  Measured with the round's kernels: 122 us per panel carried (BASELINE config 5 over 8 GPUs) against
// 38.8 us free (chain mode 1, profiles/r02_timeline_c3_one_chain_mode1.txt), plus 2.2 us per panel of backward sweep on
// either since it became a chain of programmatic dependent launches: (122 + 2.2) / (38.8 + 2.2) = 3.0 (2.75 with the
// 6.6 us backward launches and the 40.6 us chain; 2 with round 1's 32 us diagonal kernel).  PGS_CARRY_COST overrides it.
constexpr double kCarryCost = 3.0;

// chains_per_rank: 0 = default (world == 1: two chains burning from both ends when the graph is large enough,
// else one; world > 1: one chain per rank).  Odometry edge e couples (oc1[e], oc2[e]); loop edge e couples
// (la[e], lb[e]); regulariser k sits on rnode[k].
void make_partition(int N, int world, int n_odom, const int* oc1, const int* oc2, int n_loop, const int* la, const int* lb,
                    int n_reg, const int* rnode, Partition* out, int chains_per_rank = 0);

}  // namespace pgs
