// mcb200_plan.cpp -- launch planner: maps (options x trials) onto the persistent grid.  Pure host logic, no CUDA.
//
// The reference fans 8-path chunks of ONE option out over a thread pool (rayon, /root/reference/src/mc_simd.rs:49-51)
// and prices options one call at a time (benches/benchmark.rs:40-42).  Here a whole batch is one launch: each option
// is cut into `slices` equal trial ranges, each (option, slice) item is simulated by one 256-thread CTA, and the
// persistent grid walks the items round-robin.  The slice count is chosen to minimise the makespan
//     waves(items / grid_slots) * paths_per_thread
// so that both many-option batches (slices = 1) and few-option / huge-trial requests fill all 148 SMs.
#include <cmath>
#include <cstdint>
#include "../../include/mcb200.h"
#include "mcb200_plan.h"

namespace mcb200 {

static inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

int make_plan(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out) {
    if (!out || n_options < 0 || grid_slots <= 0) return MCB200_ERR_INVALID;
    if (!(steps >= 1.0f) || !(num_trials >= 0.0f) || !std::isfinite(steps) || !std::isfinite(num_trials))
        return MCB200_ERR_INVALID;
    if (steps > 2.0e9f || num_trials > 2.0e9f) return MCB200_ERR_INVALID;
    const long long n_paths = 8ll * ((long long)num_trials / 8);        // (num_trials as i32) / 8 chunks of 8
    const int half_steps = (int)steps / 2;                              // (steps as i32) / 2
    mcb200_plan p{};
    p.n_paths = (int)n_paths;
    p.half_steps = half_steps;
    p.path_steps = (double)n_options * (double)n_paths * 2.0 * (double)half_steps;
    if (n_options == 0 || n_paths == 0) {
        p.slices = 1; p.paths_per_slice = 0; p.paths_per_thread = 0; p.grid = 0; p.items = n_options;
        *out = p;
        return MCB200_OK;
    }
    const long long T = kPlanThreads;
    const long long s_min = ceil_div(n_paths, T * kPlanMaxPathsPerThread);
    long long s_max = ceil_div(n_paths, T);                              // one path per thread
    if (s_max > s_min + 4ll * grid_slots) s_max = s_min + 4ll * grid_slots;
    // per-item fixed cost (constants, reduction) expressed in issue slots, per-pair cost ~28 slots
    const long long pair_cost = 28, path_cost = 48, item_cost = 900;
    long long best_s = s_min, best_cost = -1;
    for (long long s = s_min; s <= s_max; ++s) {
        const long long pps = ceil_div(n_paths, s);
        const long long ppt = ceil_div(pps, T);
        const long long waves = ceil_div((long long)n_options * s, grid_slots);
        const long long cost = waves * (ppt * (half_steps * pair_cost + path_cost) + item_cost);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; }
    }
    p.slices = (int)best_s;
    p.paths_per_slice = (int)ceil_div(n_paths, best_s);
    p.paths_per_thread = (int)ceil_div(p.paths_per_slice, T);
    p.items = (long long)n_options * best_s;
    if (p.items > 0x7fffffffll) return MCB200_ERR_INVALID;
    p.grid = (int)(p.items < grid_slots ? p.items : grid_slots);
    *out = p;
    return MCB200_OK;
}

}  // namespace mcb200

extern "C" int mcb200_plan_query(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan* out) {
    return mcb200::make_plan(n_options, steps, num_trials, grid_slots, out);
}
