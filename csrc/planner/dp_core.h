#pragma once
#include <vector>

#include "../core/base.h"

namespace hb {

struct DpResult {
  double cost;
  std::vector<int> strategies;  // chosen strategy index per layer (empty when infeasible)
  int mem_remaining;
};
// mem_cost / intra_cost: [layer, strategy] row-major; inter_cost: [layer, prev_strategy, strategy].
DpResult galvatron_dp(int layer_num, int max_mem, int strategy_num, const std::vector<int>& mem_cost,
                      const std::vector<double>& intra_cost, const std::vector<double>& inter_cost);

}  // namespace hb
