#include "partition.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pgs {

namespace {
// below this many nodes a single GPU keeps one natural-order chain: the middle separator would not pay for itself
constexpr int kMinNodesForTwoChains = 4096;
}

void make_partition(int N, int world, int n_odom, const int* oc1, const int* oc2, int n_loop, const int* la, const int* lb,
                    int n_reg, const int* rnode, Partition* P, int chains_per_rank) {
  P->N = N; P->world = world;
  int C = chains_per_rank;
  if (C <= 0) C = (world == 1 && N >= kMinNodesForTwoChains) ? 2 : 1;
  const int R = world * C;
  // ---- ranges of equal predicted cost: the first goes up and the last down (free ends, weight 1 per node), the
  // ones in between go up and carry the separator at their low end (weight kCarryCost per node)
  std::vector<double> units(R);
  double total = 0.0;
  double carry = kCarryCost;
  if (const char* e = std::getenv("PGS_CARRY_COST")) { const double v = std::atof(e); if (v >= 1.0 && v <= 16.0) carry = v; }
  for (int c = 0; c < R; ++c) { units[c] = (c == 0 || c == R - 1) ? carry : 1.0; total += units[c]; }
  P->ranges.assign(R, PlanRange());
  double acc = 0.0;
  for (int c = 0; c < R; ++c) {
    P->ranges[c].lo = (int)std::llround(acc / total * N);
    acc += units[c];
    P->ranges[c].hi = c == R - 1 ? N : (int)std::llround(acc / total * N);
    P->ranges[c].rank = c / C;
    P->ranges[c].down = (R > 1 && c == R - 1);
  }
  P->cut.resize(world + 1);
  for (int k = 0; k < world; ++k) P->cut[k] = P->ranges[k * C].lo;
  P->cut[world] = N;
  std::vector<int> range(N);
  for (int c = 0; c < R; ++c) for (int i = P->ranges[c].lo; i < P->ranges[c].hi; ++i) range[i] = c;
  // ---- border nodes: the upper endpoint of every edge that crosses a cut
  std::vector<char> is_border(N, 0);
  auto mark = [&](int i, int j) { if (range[i] > range[j]) is_border[i] = 1; else if (range[j] > range[i]) is_border[j] = 1; };
  for (int e = 0; e < n_odom; ++e) mark(oc1[e], oc2[e]);
  for (int e = 0; e < n_loop; ++e) mark(la[e], lb[e]);
  P->node_chain.resize(N); P->node_owner.resize(N); P->border_index.assign(N, -1); P->border.clear();
  for (int i = 0; i < N; ++i) {
    P->node_chain[i] = is_border[i] ? -1 : range[i];
    P->node_owner[i] = is_border[i] ? -1 : P->ranges[range[i]].rank;
    if (is_border[i]) { P->border_index[i] = (int)P->border.size(); P->border.push_back(i); }
  }
  // ---- residual blocks -> chains; the border nodes every chain's factor has to hold
  P->chain_border.assign(R, std::vector<int>());
  auto edge_chain = [&](int i, int j) {
    int c;
    if (!is_border[i]) { c = range[i]; if (is_border[j]) P->chain_border[c].push_back(P->border_index[j]); }
    else if (!is_border[j]) { c = range[j]; P->chain_border[c].push_back(P->border_index[i]); }
    else { c = range[std::min(i, j)]; P->chain_border[c].push_back(P->border_index[i]); P->chain_border[c].push_back(P->border_index[j]); }
    return c;
  };
  P->odom_chain.resize(n_odom); P->loop_chain.resize(n_loop); P->reg_chain.resize(n_reg);
  P->odom_owner.resize(n_odom); P->loop_owner.resize(n_loop); P->reg_owner.resize(n_reg);
  for (int e = 0; e < n_odom; ++e) { P->odom_chain[e] = edge_chain(oc1[e], oc2[e]); P->odom_owner[e] = P->ranges[P->odom_chain[e]].rank; }
  for (int e = 0; e < n_loop; ++e) { P->loop_chain[e] = edge_chain(la[e], lb[e]); P->loop_owner[e] = P->ranges[P->loop_chain[e]].rank; }
  for (int k = 0; k < n_reg; ++k) {
    const int c = range[rnode[k]];
    P->reg_chain[k] = c; P->reg_owner[k] = P->ranges[c].rank;
    if (is_border[rnode[k]]) P->chain_border[c].push_back(P->border_index[rnode[k]]);
  }
  const int nb = (int)P->border.size();
  P->border_env.resize(nb);
  for (int b = 0; b < nb; ++b) P->border_env[b] = b;
  for (int c = 0; c < R; ++c) {
    std::vector<int>& L = P->chain_border[c];
    std::sort(L.begin(), L.end());
    L.erase(std::unique(L.begin(), L.end()), L.end());
    // eliminating the chain couples every pair of border nodes it touches: one clique of the border system
    for (int b : L) P->border_env[b] = std::min(P->border_env[b], L[0]);
  }
}

}  // namespace pgs
