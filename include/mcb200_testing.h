/* mcb200_testing.h -- parity, inspection and measurement hooks of libmcb200.so.  NOT part of the drop-in surface:
 * nothing here replaces a reference function.  Used by tests/, bench.py (kernel timing for the roofline) and tools/.
 */
#ifndef MCB200_TESTING_H
#define MCB200_TESTING_H
#include "mcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ parity / inspection ------------------------ */
/* Reference-stream mode: consume the generator exactly like the reference (a 256-byte simd_rand seed per 8-path
 * chunk, two 64-bit outputs per Box-Muller pair, 53-bit -> f64 -> f32 uniforms; rand32x8.rs:5-21, mc_simd.rs:9-21).
 * chunk_seeds: n * floor(num_trials/8) * 256 bytes (host).  growth_out (host, nullable): n * 8*floor(num_trials/8)
 * per-path terminal growth factors S_T / spot.  Test-only: this is the one mode that writes path data to HBM. */
int mcb200_price_batch_refstream(mcb200_ctx *ctx, int n, const float *spot, const float *strike,
                                 const float *volatility, const float *risk_free_rate, const float *years_to_expiry,
                                 const float *dividend_yield, float steps, float num_trials, unsigned flags,
                                 const uint8_t *chunk_seeds, float *call_out, float *put_out, float *call_se,
                                 float *put_se, float *growth_out);
int mcb200_greeks_batch_refstream(mcb200_ctx *ctx, int n, const float *spot, const float *strike,
                                  const float *volatility, const float *risk_free_rate, const float *years_to_expiry,
                                  const float *dividend_yield, float steps, float num_trials,
                                  const mcb200_bumps *bumps, const uint8_t *chunk_seeds,
                                  float *const out[MCB200_N_GREEK_OUTPUTS], float *const out_se[MCB200_N_GREEK_OUTPUTS]);
/* Production stream with the per-path growth factors dumped (test-only).  growth_out: n * n_paths floats (host). */
int mcb200_price_batch_dump(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                            const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                            float steps, float num_trials, unsigned flags,
                            float *call_out, float *put_out, float *call_se, float *put_se, float *growth_out);
/* Copy generator states [first, first+count) of the pool to the host (4 x u64 per thread slot). */
int mcb200_ctx_export_states(mcb200_ctx *ctx, int64_t first, int64_t count, uint64_t *states_out);

/* Kernel timing for roofline accounting: when enabled, CUDA events are recorded on the launch stream immediately
 * before and after every simulate-kernel launch; mcb200_ctx_last_sim_ms() waits for the latest one and returns its
 * device duration in milliseconds (the fused path + reduction + finalisation kernel, the only kernel of a batch). */
int mcb200_ctx_set_timing(mcb200_ctx *ctx, int enabled);
int mcb200_ctx_last_sim_ms(mcb200_ctx *ctx, float *ms_out);

/* Pipe-throughput microbenchmarks (MUFU / FMA / ALU / issue) that fix the roofline denominator with measured numbers.
 * Writes up to `cap` results; returns the number written (negative = error). */
typedef struct { char name[32]; double ops_per_clk_per_sm; double elapsed_ms; double sm_clock_mhz; } mcb200_pipe_rate;
int mcb200_microbench(mcb200_ctx *ctx, mcb200_pipe_rate *out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_TESTING_H */
