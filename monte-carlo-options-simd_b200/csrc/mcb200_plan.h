// mcb200_plan.h -- launch planner interface (host only).
#pragma once
#include "../../include/mcb200.h"

namespace mcb200 {
constexpr int kPlanThreads = 256;              // must equal kThreads in mcb200_kernels.cuh
constexpr int kPlanMaxPathsPerThread = 4096;   // keeps the per-thread fp32 payoff sums well conditioned
int make_plan(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out);
}
