/* A plain C99 client of include/mcb200.h: proves the header is valid C, that the library links without any C++ or
 * torch dependency on the caller's side, and that the boundary behaves without a GPU (planner works; the drop-ins
 * return NaN with a reason -- there is no CPU fallback).  With a GPU (argv[1] == "gpu") it prices through the drop-in
 * and the batched entry point, like a caller of the reference's mc_simd::call_price would. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "mcb200.h"

int main(int argc, char **argv) {
    mcb200_plan p;
    if (mcb200_plan_query(112, 100.0f, 20000.0f, 592, &p) != MCB200_OK) return 10;
    if (p.n_paths != 20000 || p.half_steps != 50 || p.rounds_per_option != 625 || p.total_rounds != 70000) return 11;
    int64_t a, b;
    if (mcb200_plan_warp_range(&p, p.warps - 1, &a, &b) != MCB200_OK || b != p.total_rounds) return 12;
    if (mcb200_plan_query(1, 0.0f, 1000.0f, 592, &p) != MCB200_ERR_INVALID) return 13;

    mcb200_shard sh;                                                /* 8 GPUs: few options -> trials-minor; a sweep -> options-major */
    if (mcb200_shard_query(5, 100.0f, 8e6f, 8, 3, 592, &sh) != MCB200_OK || sh.mode != MCB200_SHARD_TRIALS ||
        sh.num_trials_local != 1e6f || sh.total_paths != 8e6) return 14;
    if (mcb200_shard_query(100000, 252.0f, 1e4f, 8, 7, 592, &sh) != MCB200_OK || sh.mode != MCB200_SHARD_OPTIONS ||
        sh.option_begin != 87500 || sh.option_end != 100000) return 15;
    if (mcb200_shard_query(1, 100.0f, 1000.0f, 8, 1, 592, &sh) != MCB200_OK || sh.mode != MCB200_SHARD_SINGLE || sh.active) return 16;

    const int want_gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    float v = mc_simd_call_price(100.0f, 110.0f, 0.25f, 0.05f, 0.5f, 0.02f, 100.0f, 20000.0f);
    if (!want_gpu) {
        if (!isnan(v)) return 20;                                   /* no device: NaN, never a CPU result */
        if (strstr(mcb200_last_error(), "no CPU path") == NULL) return 21;
        mcb200_ctx *ctx = NULL;
        if (mcb200_ctx_create(&ctx, 0, 1u, 0u) != MCB200_ERR_NO_DEVICE || ctx != NULL) return 22;
        int devs0[2] = {0, 1};
        if (mcb200_ctx_create_multi(&ctx, devs0, 2, 1u, 0u) != MCB200_ERR_NO_DEVICE || ctx != NULL) return 23;
        printf("c client ok (no device): %s\n", mcb200_last_error());
        return 0;
    }
    if (isnan(v) || fabsf(v - 3.8598f) > 0.5f) { printf("drop-in price %f (%s)\n", v, mcb200_last_error()); return 30; }
    mcb200_ctx *ctx = NULL;
    if (mcb200_ctx_create(&ctx, 0, 1234u, 0u) != MCB200_OK) return 31;
    float spot[3] = {90.0f, 100.0f, 110.0f}, strike[3] = {110.0f, 110.0f, 110.0f}, vol[3] = {0.25f, 0.25f, 0.25f};
    float rate[3] = {0.05f, 0.05f, 0.05f}, years[3] = {0.5f, 0.5f, 0.5f}, div[3] = {0.02f, 0.02f, 0.02f};
    float call[3], put[3], cse[3], pse[3];
    if (mcb200_price_batch(ctx, 3, spot, strike, vol, rate, years, div, 100.0f, 40000.0f, 0u, call, put, cse, pse) != MCB200_OK) return 32;
    float bs_call[3];
    float *bs_out[MCB200_N_GREEK_OUTPUTS] = {bs_call, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (mcb200_bs_batch(ctx, 3, spot, strike, vol, rate, years, div, bs_out) != MCB200_OK) return 33;
    for (int i = 0; i < 3; ++i)
        if (fabsf(call[i] - bs_call[i]) > 4.0f * cse[i] + 1e-3f) { printf("%d: %f vs %f +- %f\n", i, call[i], bs_call[i], cse[i]); return 34; }
    /* the multi-device context behind the same host-buffer call: two per-device contexts (here both on device 0, on
     * substream blocks 0 and 1); 3 options x 400000 trials -> trials-minor split, host sum in device order */
    int devs[2] = {0, 0};
    mcb200_ctx *multi = NULL;
    if (mcb200_ctx_create_multi(&multi, devs, 2, 1234u, 0u) != MCB200_OK) { printf("%s\n", mcb200_last_error()); return 40; }
    if (mcb200_ctx_device_count(multi) != 2 || mcb200_ctx_device_count(ctx) != 1) return 41;
    float call2[3], put2[3], cse2[3], pse2[3];
    if (mcb200_price_batch(multi, 3, spot, strike, vol, rate, years, div, 100.0f, 400000.0f, 0u, call2, put2, cse2, pse2) != MCB200_OK) {
        printf("%s\n", mcb200_last_error()); return 42;
    }
    for (int i = 0; i < 3; ++i) {
        if (fabsf(call2[i] - bs_call[i]) > 4.0f * cse2[i] + 1e-3f) { printf("multi %d: %f vs %f +- %f\n", i, call2[i], bs_call[i], cse2[i]); return 43; }
        if (!(cse2[i] < 0.45f * cse[i])) return 44;                 /* ten times the trials: SE shrinks by sqrt(10) */
    }
    if (mcb200_price_batch_device(multi, 3, spot, strike, vol, rate, years, div, 100.0f, 1000.0f, 0u, call2, put2, cse2, pse2, 0)
        != MCB200_ERR_INVALID) return 45;                           /* device pointers belong to one device */
    mcb200_ctx_destroy(multi);
    mcb200_ctx_destroy(ctx);
    printf("c client ok (gpu): call %.4f %.4f %.4f  drop-in %.4f\n", call[0], call[1], call[2], v);
    return 0;
}
