// mcb200_plan.h -- launch planner interface (host only).
#pragma once
#include "../../include/mcb200.h"

namespace mcb200 {
constexpr int kPlanWarps = 8;                  // warps per CTA; must equal kWarps in mcb200_pair.cuh
constexpr int kPlanWaves = 6;                  // most waves of CTAs a batch is launched as (tuned: profiles/r2_waves.txt);
                                               // the generator pool holds resident threads x this many substreams
constexpr int kWaveMinRounds = 64;             // every logical warp of a multi-wave launch owns at least this many rounds
long long plan_waves(long long total_rounds, long long resident_warps);
int make_plan(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out);
int make_shard(int n_options, float steps, float num_trials, int n_shards, int shard, int grid_slots, mcb200_shard* out);
void plan_warp_range(const mcb200_plan* p, int warp, long long* first, long long* last);
}
