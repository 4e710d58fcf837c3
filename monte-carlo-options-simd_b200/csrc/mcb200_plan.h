// mcb200_plan.h -- launch planner interface (host only).
#pragma once
#include "../../include/mcb200.h"

namespace mcb200 {
constexpr int kPlanWarps = 8;                  // warps per CTA; must equal kWarps in mcb200_pair.cuh
int make_plan(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out);
int make_shard(int n_options, float steps, float num_trials, int n_shards, int shard, int grid_slots, mcb200_shard* out);
void plan_warp_range(const mcb200_plan* p, int warp, long long* first, long long* last);
}
