/* mcb200.h -- C ABI of libmcb200.so, the B200-native (sm_100a) Monte Carlo European-option hot path.
 *
 * Drop-in boundary.  The reference (ashayp22/monte-carlo-options-simd) has no FFI layer: its boundary is the Rust
 * `pub fn` surface of module mc_simd (src/lib.rs:11).  Every mc_simd_* function below replaces the reference function
 * of the same name with the same f32 parameter list in the same ORDER (the order differs per function); the
 * reference-side binding a maintainer would add is shown in INTEGRATION.md.  Plain pointers and sizes only; the
 * caller owns every buffer; no CPU fallback exists -- without a CUDA device every entry point fails loudly.
 *
 * Errors.  The reference's functions return a bare f32 (no Result, src/mc_simd.rs:477-779).  The drop-ins keep that
 * shape: on failure they return NaN and record a thread-local message readable with mcb200_last_error(); every other
 * entry point returns an mcb200_status.
 *
 * Reference quirks kept on purpose: steps and num_trials are f32; 8*floor(num_trials/8) paths
 * are simulated but the mean divides by num_trials (mc_simd.rs:49,77); 2*floor(steps/2) normals are consumed while
 * the drift uses steps (mc_simd.rs:44,47), so steps in (0,1) prices the deterministic forward; a negative num_trials is
 * an empty chunk range and returns zeros (mc_simd.rs:49); vega and rho are per 1 point (/200h), theta per year; gamma
 * and vega are call-only.
 * Deviations: steps <= 0 (dt = T/steps is inf or NaN in the reference) is refused -- the drop-ins return NaN either
 * way; the production stream draws the Box-Muller radius from a 23-bit uniform (largest single normal 5.65 sigma
 * against ~8.6 sigma for the reference's 53-bit -> f32 uniforms) and the angle from an equispaced grid of 2^15
 * directions; the reference-stream entry points of mcb200_testing.h consume the generator exactly like the reference.
 */
#ifndef MCB200_H
#define MCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcb200_ctx mcb200_ctx;

typedef enum {
    MCB200_OK = 0,
    MCB200_ERR_INVALID = 1,     /* bad argument: null pointer, n < 0, steps <= 0, non-finite or beyond +-2e9; a device-pointer
                                   entry point on a multi-device context; xgpu launch / results out of turn */
    MCB200_ERR_CUDA = 2,        /* a CUDA runtime call failed; see mcb200_last_error() */
    MCB200_ERR_NO_DEVICE = 3,   /* no usable CUDA device (the library has no CPU path) */
    MCB200_ERR_NOMEM = 4
} mcb200_status;

/* flags for the price entry points */
#define MCB200_ANTITHETIC 1u    /* call_price_av / put_price_av estimator (mc_simd.rs:82-149) */

/* ------------------------------------------------------------------ context ------------------------------------ */
/* One context = one device, one stream, one pool of per-thread xoshiro256++ substreams (one per LOGICAL thread of the
 * largest launch plan: sm_count * ctas_per_sm * 256 resident threads x 6 waves; smaller plans use a prefix of it).
 * Thread slot i of the device uses substream jump^i(long_jump^substream_block(splitmix64(seed))): 2^128 draws per
 * thread, up to 2^40 thread slots per block, one block per GPU/rank, reached in O(log block) polynomial applications
 * (replaces per-chunk OS-entropy seeding, mc_simd.rs:52-54).  Launches continue their substreams.
 *
 * Reproducibility.  A fixed (seed, block, call sequence) reproduces every result bit for bit ON THE SAME GPU MODEL: the
 * launch plan -- which thread simulates which path, and the order of the fixed-order sums -- is a function of the
 * batch shape and of sm_count * ctas_per_sm (mcb200_plan_query), nothing else.  A GPU with a different SM count draws
 * the same law from differently assigned substreams.
 *
 * Threads and streams.  Every entry point may be called from any thread (a mutex per context).  The pool, the partial-sum
 * slots and the arrival counters are ONE set per context, so launches of one context never overlap on the device: a
 * launch enqueued on a different stream than its predecessor first waits (cudaStreamWaitEvent, device side) for an
 * event recorded behind the predecessor.  Two caller streams therefore interleave safely but do not run concurrently;
 * use one context per stream (different substream blocks) for concurrency.  Stream handles passed to the *_device
 * entry points must stay valid until the work enqueued on them has finished. */
int mcb200_ctx_create(mcb200_ctx **ctx, int device, uint64_t seed, uint32_t substream_block);
void mcb200_ctx_destroy(mcb200_ctx *ctx);
int mcb200_ctx_set_seed(mcb200_ctx *ctx, uint64_t seed, uint32_t substream_block);
const char *mcb200_last_error(void);

typedef struct {
    int device, sm_count, ctas_per_sm, threads_per_cta;
    int grid_slots;              /* resident CTAs the planner fills for the price kernels: sm_count * ctas_per_sm (4) */
    int grid_slots_greeks;       /* ... and for the Greeks kernel, which runs 3 fatter CTAs per SM */
    int64_t pool_threads;        /* generator states held in HBM (32 bytes each) */
    uint64_t kernel_launches;    /* kernels launched by this context so far */
    uint64_t seed;
    uint32_t substream_block;
    int sm_clock_khz;            /* cudaDevAttrClockRate */
} mcb200_info;
int mcb200_ctx_info(mcb200_ctx *ctx, mcb200_info *out);

/* Launch plan for one batch: how the work maps onto the logical warps.  Pure host logic.
 * A "round" is 32 consecutive paths of one option (one per lane).  Round r belongs to option r / rounds_per_option;
 * logical warp w simulates rounds [w*R/W, (w+1)*R/W), R = total_rounds, W = warps, in order, and its lane l draws from
 * generator substream 32*w + l of the pool.  (Hardware warp h of CTA c is logical warp 8c + 2*(h%4) + h/4, so the two
 * warps of an SM sub-partition own neighbouring ranges.)
 * Small batches run as ONE wave: W = min(8 * grid_slots, R) resident warps.  Large batches are launched as up to 6 waves
 * of CTAs (W = 8 * grid_slots * waves, every warp still owning >= 64 rounds): the SM's warp schedulers favour the CTA
 * that became resident first, so a single wave tails off with one CTA per SM running alone; with several waves a fresh
 * CTA takes every freed slot and only the last, short, wave tails off (1.5-2.7 % faster).  The environment variable
 * MCB200_WAVES=1..6 caps the wave count (tuning and tests; it changes the plan, hence which substream draws which path). */
typedef struct {
    int n_paths;                 /* 8 * floor(num_trials / 8) */
    int half_steps;              /* floor(steps) / 2 */
    int rounds_per_option;       /* ceil(n_paths / 32) */
    int warps;                   /* W = min(8 * grid_slots * waves, total_rounds): every warp owns at least one round */
    int grid;                    /* CTAs launched = ceil(W / 8) */
    int max_rounds_per_warp;     /* ceil(R / W): rounds differ by at most one between warps */
    int64_t total_rounds;        /* R = n_options * rounds_per_option */
    double path_steps;           /* n_options * n_paths * 2 * half_steps: the throughput numerator */
    int waves;                   /* min(6, R / (8 * grid_slots * 64)), at least 1 */
} mcb200_plan;
int mcb200_plan_query(int n_options, float steps, float num_trials, int grid_slots, mcb200_plan *out);
int mcb200_plan_warp_range(const mcb200_plan *plan, int warp, int64_t *first_round, int64_t *end_round);

/* ------------------------------------------------------------------ several GPUs -------------------------------- */
/* The reference parallelises INSIDE a call: chunks of 8 paths fan out over a thread pool (src/mc_simd.rs:49-51) and
 * are reduced (src/mc_simd.rs:71-74).  The node-scale equivalent is a multi-device context: one process, one per-device
 * context + host worker thread + stream per listed device (a device may be listed more than once), device i on
 * substream block first_substream_block + i.  The HOST-BUFFER entry points (mcb200_price_batch, mcb200_greeks_batch,
 * mcb200_bs_batch and the twelve drop-ins through MCB200_DEVICES=0,1,..) accept it and split every batch by the
 * policy of mcb200_shard_query; the *_device, xgpu and inspection entry points are per device (mcb200_ctx_child).
 * Results are a function of (seed, first block, device COUNT, call sequence): reproducible bit for bit, and the
 * trials-minor sum runs in device order on the host.  Destroy with mcb200_ctx_destroy. */
int mcb200_ctx_create_multi(mcb200_ctx **ctx, const int *devices, int n_devices, uint64_t seed,
                            uint32_t first_substream_block);
int mcb200_ctx_device_count(mcb200_ctx *ctx);                       /* 1 for a plain context */
int mcb200_ctx_child(mcb200_ctx *ctx, int i, mcb200_ctx **child);  /* borrowed per-device context of a multi-device one */

/* How one batch is split over n_shards GPUs -- the same policy for the multi-device context (shard = device index) and
 * for one-process-per-GPU operation (shard = rank).  Pure host logic.
 *   MCB200_SHARD_SINGLE   the batch does not fill one GPU (or n_shards == 1): shard 0 runs all of it;
 *   MCB200_SHARD_OPTIONS  n_options >= n_shards and ceil(n/S) within 10 % of n/S (always so for n >= 10 S): shard s owns
 *                         options [s*n/S, (s+1)*n/S), no exchange step;
 *   MCB200_SHARD_TRIALS   otherwise: shard s simulates chunks [s*C/S, (s+1)*C/S) of the C = floor(num_trials/8) 8-path
 *                         chunks of EVERY option; per-option fp64 moment sums are added in shard order and finalised
 *                         once with total_paths and the caller's num_trials (the divisor of mc_simd.rs:77). */
enum { MCB200_SHARD_SINGLE = 0, MCB200_SHARD_OPTIONS = 1, MCB200_SHARD_TRIALS = 2 };
typedef struct {
    int mode, n_shards, shard;
    int active;                  /* 0: this shard has nothing to do */
    int option_begin, option_end;
    float num_trials_local;      /* num_trials to pass to this shard's launch */
    double total_paths;          /* 8 * floor(num_trials / 8): paths of the whole sample per option */
} mcb200_shard;
int mcb200_shard_query(int n_options, float steps, float num_trials, int n_shards, int shard, int grid_slots,
                       mcb200_shard *out);

/* ------------------------------------------------------------------ batched entry points (new) ----------------- */
/* Host buffers: the call returns when the outputs are valid.  Inputs reach the device through a pinned staging copy
 * (batches of <= 128 options ride in the launch parameters instead); outputs come back through a D2H copy (batches of
 * <= 4096 options are written by the kernel straight into mapped pinned memory).  One kernel launch per call.
 * Every option gets its own fresh paths (no sharing across options).  Output pointers may be NULL.
 * num_trials < 8 simulates nothing and returns zeros, like the reference's empty chunk range (mc_simd.rs:49). */
int mcb200_price_batch(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                       const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                       float steps, float num_trials, unsigned flags,
                       float *call_out, float *put_out, float *call_se, float *put_se);

typedef struct { float delta_spot, delta_volatility, delta_risk_free_rate, delta_years_to_expiry; } mcb200_bumps;

/* All ten estimators from ONE set of paths per option (common random numbers across every bump).
 * Order of out[] / out_se[]: call, put, call_delta, put_delta, gamma, vega, call_rho, put_rho, call_theta, put_theta. */
#define MCB200_N_GREEK_OUTPUTS 10
int mcb200_greeks_batch(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                        const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                        float steps, float num_trials, const mcb200_bumps *bumps,
                        float *const out[MCB200_N_GREEK_OUTPUTS], float *const out_se[MCB200_N_GREEK_OUTPUTS]);

/* Device-resident variants: every pointer is a device pointer on the context's device; work is enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = the context's own stream) and the call does not synchronize. */
int mcb200_price_batch_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                              const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                              float steps, float num_trials, unsigned flags,
                              float *call_out, float *put_out, float *call_se, float *put_se, void *cuda_stream);
int mcb200_greeks_batch_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike,
                               const float *volatility, const float *risk_free_rate, const float *years_to_expiry,
                               const float *dividend_yield, float steps, float num_trials, const mcb200_bumps *bumps,
                               float *const out[MCB200_N_GREEK_OUTPUTS], float *const out_se[MCB200_N_GREEK_OUTPUTS],
                               void *cuda_stream);

/* Trials-sharded operation (several GPUs simulate disjoint trial ranges of the same options):
 *   1. each GPU: mcb200_price_moments_device() -> moments[n][4] fp64 = {sum call, sum call^2, sum put, sum put^2}
 *   2. caller sums the moment arrays across GPUs (ncclAllReduce / torch.distributed.all_reduce, fp64, sum)
 *   3. mcb200_price_finalize_device() with the TOTAL number of simulated paths and the total num_trials. */
int mcb200_price_moments_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike,
                                const float *volatility, const float *risk_free_rate, const float *years_to_expiry,
                                const float *dividend_yield, float steps, float num_trials_local, unsigned flags,
                                double *moments, void *cuda_stream);
int mcb200_price_finalize_device(mcb200_ctx *ctx, int n, const float *risk_free_rate, const float *years_to_expiry,
                                 const double *moments, double total_paths, double total_num_trials,
                                 float *call_out, float *put_out, float *call_se, float *put_se, void *cuda_stream);

/* The same three steps for the ten Greek estimators: moments[n][20] fp64 = {sum, sum^2} per estimator in the order of
 * mcb200_greeks_batch (the per-path central differences are what is summed, so the moments add across GPUs). */
int mcb200_greeks_moments_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike,
                                 const float *volatility, const float *risk_free_rate, const float *years_to_expiry,
                                 const float *dividend_yield, float steps, float num_trials_local,
                                 const mcb200_bumps *bumps, double *moments, void *cuda_stream);
int mcb200_greeks_finalize_device(mcb200_ctx *ctx, int n, const float *risk_free_rate, const float *years_to_expiry,
                                  const double *moments, double total_paths, double total_num_trials,
                                  const mcb200_bumps *bumps, float *const out[MCB200_N_GREEK_OUTPUTS],
                                  float *const out_se[MCB200_N_GREEK_OUTPUTS], void *cuda_stream);

/* Fused variant of the trials-sharded operation: NO collective and no second kernel.  One GPU (the owner) holds an
 * exchange buffer that the other processes map through CUDA IPC (peer memory over NVLink / NVSwitch).  Each rank's
 * path kernel stores its finished per-option moment sums into its own row there and bumps a cross-GPU arrival counter;
 * the warp that completes an option's count -- on whichever GPU -- sums the rows in rank order (deterministic) and
 * writes price and standard error of the WHOLE sample into the buffer.
 *   owner : mcb200_xgpu_create(ctx, n_max, world, 0, handle)      -> ship the 64 handle bytes to the other ranks
 *   others: mcb200_xgpu_open(ctx, handle, n_max, world, 0)
 *   all   : mcb200_price_xgpu_device(..., num_trials_local, flags, rank, total_paths, total_num_trials, stream)
 *   all   : mcb200_xgpu_results(ctx, n, call, put, call_se, put_se)  -- waits ON THE DEVICE (a one-thread kernel spinning
 *           on the buffer's epoch word) until the round is complete on every rank, then copies the results out.
 * Every rank must issue the same sequence of rounds (the round number is counted locally, after a successful launch).
 * Launch and results ALTERNATE on every rank: a second mcb200_*_xgpu_device before the pending round's results were
 * fetched returns MCB200_ERR_INVALID, so round k+1 can never mix with round k in the buffer.  Ranks of ONE process
 * (one context per GPU, or several contexts on one GPU) share the owner's allocation with mcb200_xgpu_attach -- plain
 * peer access, no IPC handle.  The device-side wait gives up after mcb200_xgpu_set_timeout seconds (default 120).
 * The kernel of mcb200_price_xgpu_device is enqueued on `cuda_stream`; the wait and the copy run on the context's
 * own stream behind an event recorded after that kernel. */
#define MCB200_IPC_HANDLE_BYTES 64
int mcb200_xgpu_create(mcb200_ctx *ctx, int n_max, int world, int greeks, uint8_t handle_out[MCB200_IPC_HANDLE_BYTES]);
int mcb200_xgpu_open(mcb200_ctx *ctx, const uint8_t handle[MCB200_IPC_HANDLE_BYTES], int n_max, int world, int greeks);
int mcb200_xgpu_attach(mcb200_ctx *ctx, mcb200_ctx *owner, int n_max, int world, int greeks);
int mcb200_xgpu_set_timeout(mcb200_ctx *ctx, double seconds);
int mcb200_xgpu_close(mcb200_ctx *ctx);
int mcb200_price_xgpu_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                             const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                             float steps, float num_trials_local, unsigned flags, int rank, double total_paths,
                             double total_num_trials, void *cuda_stream);
int mcb200_xgpu_results(mcb200_ctx *ctx, int n, float *call_out, float *put_out, float *call_se, float *put_se);
/* the same for the ten Greek estimators (exchange buffer created / opened with greeks = 1) */
int mcb200_greeks_xgpu_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                              const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                              float steps, float num_trials_local, const mcb200_bumps *bumps, int rank,
                              double total_paths, double total_num_trials, void *cuda_stream);
int mcb200_xgpu_results_greeks(mcb200_ctx *ctx, int n, float *const out[MCB200_N_GREEK_OUTPUTS],
                               float *const out_se[MCB200_N_GREEK_OUTPUTS]);

/* ------------------------------------------------------------------ closed form (analytic companion) ----------- */
/* Black-Scholes closed form with dividend yield for a batch -- the analytic value the reference's tests compare every
 * Monte Carlo estimate with (src/bs.rs:42-184, pub(crate) there).  Same units as bs.rs and as the estimators above:
 * vega and rho per 1 point, theta per year.  out[] in the order of mcb200_greeks_batch; entries may be NULL. */
int mcb200_bs_batch(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                    const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                    float *const out[MCB200_N_GREEK_OUTPUTS]);
int mcb200_bs_batch_device(mcb200_ctx *ctx, int n, const float *spot, const float *strike, const float *volatility,
                           const float *risk_free_rate, const float *years_to_expiry, const float *dividend_yield,
                           float *const out[MCB200_N_GREEK_OUTPUTS], void *cuda_stream);

/* ------------------------------------------------------------------ the 12 drop-ins ---------------------------- */
/* Same names, parameter lists and semantics as /root/reference/src/mc_simd.rs:477-779.  They run on a lazily
 * created process-wide default context (device MCB200_DEVICE or 0, seed MCB200_SEED or a fixed constant), are
 * blocking and callable from any thread. */
float mc_simd_call_price(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                         float dividend_yield, float steps, float num_trials);               /* mc_simd.rs:477-498 */
float mc_simd_put_price(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                        float dividend_yield, float steps, float num_trials);                /* mc_simd.rs:500-521 */
float mc_simd_call_price_av(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                            float dividend_yield, float steps, float num_trials);            /* mc_simd.rs:523-544 */
float mc_simd_put_price_av(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                           float dividend_yield, float steps, float num_trials);             /* mc_simd.rs:546-567 */
float mc_simd_call_delta(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                         float years_to_expiry, float dividend_yield, float steps, float num_trials); /* :569-593 */
float mc_simd_put_delta(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                        float years_to_expiry, float dividend_yield, float steps, float num_trials);  /* :595-619 */
float mc_simd_gamma(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                    float years_to_expiry, float dividend_yield, float steps, float num_trials);      /* :621-645 */
float mc_simd_vega(float spot, float strike, float volatility, float delta_volatility, float risk_free_rate,
                   float years_to_expiry, float dividend_yield, float steps, float num_trials);       /* :647-671 */
float mc_simd_call_rho(float spot, float strike, float volatility, float risk_free_rate, float delta_risk_free_rate,
                       float years_to_expiry, float dividend_yield, float steps, float num_trials);   /* :673-699 */
float mc_simd_put_rho(float spot, float strike, float volatility, float risk_free_rate, float delta_risk_free_rate,
                      float years_to_expiry, float dividend_yield, float steps, float num_trials);    /* :701-726 */
float mc_simd_call_theta(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                         float delta_years_to_expiry, float dividend_yield, float steps, float num_trials); /* :728-753 */
float mc_simd_put_theta(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                        float delta_years_to_expiry, float dividend_yield, float steps, float num_trials);  /* :755-779 */

/* Select / inspect the default context used by the drop-ins.  MCB200_DEVICES=0,1,.. makes it a multi-device context.
 * The pointer returned by mcb200_default_ctx is borrowed: valid until the next mcb200_default_ctx_reset (a drop-in
 * running on another thread keeps the old context alive until it returns). */
int mcb200_default_ctx(mcb200_ctx **out);
int mcb200_default_ctx_reset(int device, uint64_t seed, uint32_t substream_block);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
