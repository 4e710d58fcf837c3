// mcb200_internal.h -- what the host runtime's translation units share (not part of the C ABI).
#pragma once
#include <cstdint>
#include "../../include/mcb200.h"

namespace mcb200 {

struct MultiRuntime;                       // mcb200_multi.cu: one worker thread + one per-device context per entry

// mcb200_api.cu
int set_error(int code, const char* fmt, ...);                 // records the thread-local message, returns code
MultiRuntime* ctx_multi(mcb200_ctx* c);                        // null for a plain (single-device) context
mcb200_ctx* ctx_new_parent(MultiRuntime* m, int device0, uint64_t seed, uint32_t first_block);   // owns no device memory
// Host-buffer building blocks of the trials-minor split (each GPU simulates a disjoint trial range of the SAME options):
//   host_moments : one fused launch on c's device, per-option fp64 moment sums [n][2 * n_est] copied to the host;
//   host_finalize: summed moments -> values and standard errors with the device's own finalisation arithmetic.
int host_moments(mcb200_ctx* c, int n, const float* const in[6], float steps, float num_trials_local, int greeks,
                 unsigned flags, const mcb200_bumps* bumps, double* moments_host);
int host_finalize(mcb200_ctx* c, int n, int greeks, const float* rate, const float* years, const double* moments_host,
                  double total_paths, double total_num_trials, const mcb200_bumps* bumps, float* const* out,
                  float* const* out_se);

// mcb200_multi.cu
void multi_destroy(MultiRuntime* m);
int multi_run_host(MultiRuntime* m, int n, const float* const in[6], float steps, float num_trials, int greeks,
                   unsigned flags, const mcb200_bumps* bumps, float* const* out, float* const* out_se);
int multi_set_seed(MultiRuntime* m, uint64_t seed, uint32_t first_block);
int multi_device_count(MultiRuntime* m);
mcb200_ctx* multi_child(MultiRuntime* m, int i);
uint64_t multi_launches(MultiRuntime* m);

}  // namespace mcb200
