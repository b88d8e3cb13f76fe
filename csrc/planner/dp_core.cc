// Galvatron layer-wise strategy search: knapsack dynamic programme over a memory
// budget with intra-layer execution cost and inter-layer (re-sharding) transition cost.
// (capability parity: tools/Galvatron/csrc/dp_core.cpp:22-90, re-implemented on plain
//  vectors with a rolling cost table and an explicit back-pointer tensor)
#include "dp_core.h"

#include <algorithm>
#include <limits>

namespace hb {

DpResult galvatron_dp(int layer_num, int max_mem, int strategy_num, const std::vector<int>& mem_cost,
                      const std::vector<double>& intra_cost, const std::vector<double>& inter_cost) {
  const double INF = std::numeric_limits<double>::infinity();
  HB_CHECK((int)mem_cost.size() == layer_num * strategy_num) << "mem_cost must be [layers, strategies]";
  HB_CHECK((int)intra_cost.size() == layer_num * strategy_num) << "intra_cost must be [layers, strategies]";
  HB_CHECK((int)inter_cost.size() == layer_num * strategy_num * strategy_num) << "inter_cost must be [layers, S, S]";
  DpResult res;
  res.cost = INF;
  res.mem_remaining = -1;
  if (layer_num == 0 || max_mem <= 0) return res;
  // f[v][s]: best cost of the layers processed so far using at most v memory units, last layer on strategy s
  std::vector<double> f((size_t)max_mem * strategy_num, 0.0);
  std::vector<int> mark((size_t)layer_num * max_mem * strategy_num, -1);
  std::vector<double> cand(strategy_num);
  for (int i = 0; i < layer_num; ++i) {
    for (int v = max_mem - 1; v >= 0; --v) {
      for (int s = 0; s < strategy_num; ++s) {
        const int need = mem_cost[i * strategy_num + s];
        double& cell = f[(size_t)v * strategy_num + s];
        int& bp = mark[((size_t)i * max_mem + v) * strategy_num + s];
        if (v < need) { cell = INF; bp = -1; continue; }
        int best = 0;
        double best_c = INF;
        for (int p = 0; p < strategy_num; ++p) {
          const double c = f[(size_t)(v - need) * strategy_num + p] + inter_cost[((size_t)i * strategy_num + p) * strategy_num + s] +
                           intra_cost[i * strategy_num + s];
          if (c < best_c) { best_c = c; best = p; }
        }
        cell = best_c;
        bp = best;
      }
    }
  }
  const double* last = &f[(size_t)(max_mem - 1) * strategy_num];
  int s = int(std::min_element(last, last + strategy_num) - last);
  if (!(last[s] < INF)) return res;
  res.cost = last[s];
  res.strategies.assign(layer_num, 0);
  int v = max_mem - 1;
  res.strategies[layer_num - 1] = s;
  for (int i = layer_num - 1; i > 0; --i) {
    const int prev = mark[((size_t)i * max_mem + v) * strategy_num + s];
    v -= mem_cost[i * strategy_num + s];
    s = prev;
    res.strategies[i - 1] = s;
  }
  res.mem_remaining = v - mem_cost[s];
  return res;
}

}  // namespace hb
