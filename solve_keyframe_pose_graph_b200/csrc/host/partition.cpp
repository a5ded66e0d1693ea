#include "partition.h"

#include <algorithm>

namespace pgs {

int Partition::range_of(int node) const {
  return (int)(std::upper_bound(cut.begin(), cut.end(), node) - cut.begin()) - 1;
}

void make_partition(int N, int world, int n_odom, const int* oc1, const int* oc2, int n_loop, const int* la, const int* lb,
                    int n_reg, const int* rnode, Partition* P) {
  P->N = N; P->world = world;
  P->cut.resize(world + 1);
  for (int k = 0; k <= world; ++k) P->cut[k] = (int)((long long)k * N / world);
  std::vector<int> range(N);
  for (int k = 0; k < world; ++k) for (int i = P->cut[k]; i < P->cut[k + 1]; ++i) range[i] = k;
  std::vector<char> is_border(N, 0);
  auto mark = [&](int i, int j) { if (range[i] > range[j]) is_border[i] = 1; else if (range[j] > range[i]) is_border[j] = 1; };
  for (int e = 0; e < n_odom; ++e) mark(oc1[e], oc2[e]);
  for (int e = 0; e < n_loop; ++e) mark(la[e], lb[e]);
  P->node_owner.resize(N); P->border.clear();
  for (int i = 0; i < N; ++i) { P->node_owner[i] = is_border[i] ? -1 : range[i]; if (is_border[i]) P->border.push_back(i); }
  auto edge_owner = [&](int i, int j) {
    if (!is_border[i]) return range[i];
    if (!is_border[j]) return range[j];
    return range[std::min(i, j)];
  };
  P->odom_owner.resize(n_odom); P->loop_owner.resize(n_loop); P->reg_owner.resize(n_reg);
  for (int e = 0; e < n_odom; ++e) P->odom_owner[e] = edge_owner(oc1[e], oc2[e]);
  for (int e = 0; e < n_loop; ++e) P->loop_owner[e] = edge_owner(la[e], lb[e]);
  for (int k = 0; k < n_reg; ++k) P->reg_owner[k] = range[rnode[k]];
}

}  // namespace pgs
