// mcb200_plan.cpp -- launch planner: maps (options x trials) onto the logical warps.  Pure host logic, no CUDA.
//
// The reference fans 8-path chunks of ONE option out over a thread pool (rayon, /root/reference/src/mc_simd.rs:49-51)
// and prices options one call at a time (benches/benchmark.rs:40-42).  Here a whole batch is one launch.  The unit of
// work is a "round": 32 consecutive paths of one option (one per lane of a warp).  The n_options * ceil(n_paths/32)
// rounds are dealt to the W logical warps in contiguous ranges  [w*R/W, (w+1)*R/W)  -- equal to within one round --
// so the load of every SM sub-partition is the same whatever the shape of the batch: 1e5 options x 1e4 trials and
// 1 option x 1e8 trials both fill all 148 SMs.  The kernel (mcb200_kernels.cuh) evaluates the same formulas.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include "../../include/mcb200.h"
#include "mcb200_plan.h"

namespace mcb200 {

static inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// Waves of a launch: up to kPlanWaves, as many as leave every logical warp at least kWaveMinRounds rounds (small
// batches: one wave, W = resident warps, as before).  MCB200_WAVES caps the count for tuning runs (tools/waves.sh).
long long plan_waves(long long total_rounds, long long resident_warps) {
    long long cap = kPlanWaves;
    if (const char* env = std::getenv("MCB200_WAVES")) {
        const long v = std::strtol(env, nullptr, 10);
        if (v >= 1 && v <= kPlanWaves) cap = v;
    }
    long long min_rounds = kWaveMinRounds;
    if (const char* env = std::getenv("MCB200_WAVE_MIN_ROUNDS")) {
        const long v = std::strtol(env, nullptr, 10);
        if (v >= 1) min_rounds = v;
    }
    long long waves = total_rounds / (resident_warps * min_rounds);
    if (waves > cap) waves = cap;
    return waves < 1 ? 1 : waves;
}

int make_plan(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out) {
    if (!out || n_options < 0 || grid_slots <= 0) return MCB200_ERR_INVALID;
    // steps in (0,1) and negative num_trials are legal in the reference: no normals are drawn / the chunk range is
    // empty (mc_simd.rs:47,49) and a finite value comes back.  steps <= 0 gives dt = T/steps = inf or NaN there; here
    // it is refused (the drop-ins return NaN either way).
    if (!(steps > 0.0f) || !std::isfinite(steps) || !std::isfinite(num_trials)) return MCB200_ERR_INVALID;
    if (steps > 2.0e9f || num_trials > 2.0e9f || num_trials < -2.0e9f) return MCB200_ERR_INVALID;
    const long long chunks = (long long)num_trials / 8;                 // (num_trials as i32) / 8 chunks of 8
    const long long n_paths = chunks > 0 ? 8ll * chunks : 0;            // an empty or negative range simulates nothing
    const int half_steps = (int)steps / 2;                              // (steps as i32) / 2
    mcb200_plan p{};
    p.n_paths = (int)n_paths;
    p.half_steps = half_steps;
    p.path_steps = (double)n_options * (double)n_paths * 2.0 * (double)half_steps;
    if (n_options == 0 || n_paths == 0) {
        *out = p;                                                        // nothing to launch
        return MCB200_OK;
    }
    p.rounds_per_option = (int)ceil_div(n_paths, 32);
    p.total_rounds = (long long)n_options * p.rounds_per_option;
    // W <= R, so that every warp owns at least one round (the arrival counts of the kernel rely on it).
    // Large batches are launched as several WAVES of CTAs (kPlanWaves times the resident count, each logical warp with
    // its own generator substream): the SM's warp schedulers favour the CTA that became resident first, so with one
    // wave the favoured CTAs finish at a third of the launch and the last one runs alone at the end; with several
    // waves a fresh CTA takes every freed slot and only the last, short, wave tails off.
    const long long resident = (long long)grid_slots * kPlanWarps;
    const long long waves = plan_waves(p.total_rounds, resident);
    const long long logical = resident * waves;
    p.warps = (int)(p.total_rounds < logical ? p.total_rounds : logical);
    p.grid = (int)ceil_div(p.warps, kPlanWarps);
    p.waves = (int)waves;
    p.max_rounds_per_warp = (int)ceil_div(p.total_rounds, p.warps);
    if (p.total_rounds > (1ll << 62) / p.warps) return MCB200_ERR_INVALID;   // (r+1)*W must fit in 64 bits
    *out = p;
    return MCB200_OK;
}

// Sharding policy of a batch over `n_shards` GPUs (one process with several devices, or one process per GPU): replaces
// the reference's single fan-out over a thread pool (mc_simd.rs:49-51) and its reduce (mc_simd.rs:71-74) at node scale.
//   single         the batch does not fill one GPU (rounds <= its resident warps), or one shard: shard 0 runs it alone;
//   options-major  n >= n_shards and the largest block is within 10 % of the mean (always true for n >= 10 * n_shards):
//                  contiguous blocks of options, no exchange step;
//   trials-minor   otherwise: every shard simulates a contiguous range of the 8-path chunks of ALL options; the fp64
//                  moment sums are added in shard order and finalised once.
int make_shard(int n_options, float steps, float num_trials, int n_shards, int shard, int grid_slots, mcb200_shard* out) {
    if (!out || n_shards <= 0 || shard < 0 || shard >= n_shards) return MCB200_ERR_INVALID;
    mcb200_plan whole;
    int rc = make_plan(n_options, steps, num_trials, grid_slots, &whole);
    if (rc != MCB200_OK) return rc;
    mcb200_shard s{};
    s.n_shards = n_shards; s.shard = shard;
    s.total_paths = (double)whole.n_paths;
    const long long chunks = whole.n_paths / 8;
    const long long resident = (long long)grid_slots * kPlanWarps;
    if (n_shards == 1 || n_options == 0 || chunks == 0 || whole.total_rounds <= resident) {
        s.mode = MCB200_SHARD_SINGLE;
        if (shard == 0) { s.option_begin = 0; s.option_end = n_options; s.num_trials_local = num_trials; s.active = 1; }
    } else if (n_options >= n_shards &&
               10ll * (((long long)n_options + n_shards - 1) / n_shards) * n_shards <= 11ll * n_options) {
        s.mode = MCB200_SHARD_OPTIONS;
        s.option_begin = (int)((long long)shard * n_options / n_shards);
        s.option_end = (int)(((long long)shard + 1) * n_options / n_shards);
        s.num_trials_local = num_trials;
        s.active = s.option_end > s.option_begin;
    } else {
        s.mode = MCB200_SHARD_TRIALS;
        const long long c0 = shard * chunks / n_shards, c1 = ((long long)shard + 1) * chunks / n_shards;
        s.option_begin = 0; s.option_end = n_options;
        s.num_trials_local = (float)(8 * (c1 - c0));
        s.active = c1 > c0;
    }
    *out = s;
    return MCB200_OK;
}

void plan_warp_range(const mcb200_plan* p, int warp, long long* first, long long* last) {
    *first = (long long)warp * p->total_rounds / p->warps;
    *last = ((long long)warp + 1) * p->total_rounds / p->warps;
}

}  // namespace mcb200

extern "C" int mcb200_plan_query(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out) {
    return mcb200::make_plan(n_options, steps, num_trials, grid_slots, out);
}

extern "C" int mcb200_shard_query(int n_options, float steps, float num_trials, int n_shards, int shard, int grid_slots,
                                  mcb200_shard* out) {
    return mcb200::make_shard(n_options, steps, num_trials, n_shards, shard, grid_slots, out);
}

extern "C" int mcb200_plan_warp_range(const mcb200_plan* plan, int warp, int64_t* first_round, int64_t* end_round) {
    if (!plan || !first_round || !end_round || plan->warps <= 0 || warp < 0 || warp >= plan->warps) return MCB200_ERR_INVALID;
    long long a, b;
    mcb200::plan_warp_range(plan, warp, &a, &b);
    *first_round = a; *end_round = b;
    return MCB200_OK;
}
