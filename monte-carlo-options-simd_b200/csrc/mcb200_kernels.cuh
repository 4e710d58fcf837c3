// mcb200_kernels.cuh -- the fused GBM path + payoff kernels (sm_100a) and their fixed-order second pass.
//
// One kernel family replaces the reference's inner step and its six simulation cores:
//   speed_update                    /root/reference/src/mc_simd.rs:9-21      -> tri_fast() / pair_fast() / pair_ref()
//   get_rand_uniform_f32x8          /root/reference/src/rand32x8.rs:5-21     -> uniforms built in pair_*()
//   monte_carlo_pricing             src/mc_simd.rs:25-79                      -> simulate_kernel<*, PAYOFF_PRICE>
//   monte_carlo_av_pricing          src/mc_simd.rs:82-149                     -> simulate_kernel<*, PAYOFF_AV>
//   monte_carlo_{spot,volatility,interest,time}_pricing  src/mc_simd.rs:152-473 -> simulate_kernel<*, PAYOFF_GREEKS>
//   rayon map/reduce + reduce_add + discount (mc_simd.rs:49-51,71-78)        -> warp butterfly, per-(warp, option) fp64
//                                                                               slots, fixed-order sum and finalisation by
//                                                                               the last-arriving warp, in the same kernel
// Path data never leaves registers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "mcb200_pair.cuh"

namespace mcb200 {

enum Stream { STREAM_FAST = 0, STREAM_REF = 1 };
enum Payoff { PAYOFF_PRICE = 0, PAYOFF_AV = 1, PAYOFF_GREEKS = 2 };

// estimator slots (each has a sum and a sum of squares in the moment arrays)
enum Estimator {
    EST_CALL = 0, EST_PUT = 1,
    EST_CALL_DELTA = 2, EST_PUT_DELTA = 3, EST_GAMMA = 4, EST_VEGA = 5,
    EST_CALL_RHO = 6, EST_PUT_RHO = 7, EST_CALL_THETA = 8, EST_PUT_THETA = 9,
    EST_COUNT_PRICE = 2, EST_COUNT_GREEKS = 10
};

template <int PAYOFF> struct PayoffTraits {
    static constexpr int kEst = 2; static constexpr int kConst = 4;
    static constexpr int kCtasPerSm = kPriceCtasPerSm; static constexpr int kQuadUnroll = kPriceQuadUnroll;
};
template <> struct PayoffTraits<PAYOFF_GREEKS> {
    static constexpr int kEst = 10; static constexpr int kConst = 20;
    static constexpr int kCtasPerSm = kGreeksCtasPerSm; static constexpr int kQuadUnroll = kGreeksQuadUnroll;
};

struct OptionArrays {          // struct-of-arrays option batch in HBM (f32, n entries each)
    const float* spot; const float* strike; const float* vol; const float* rate; const float* years; const float* div;
};

struct Bumps { float dspot, dvol, drate, dyears; };

struct FinalizeSpec {       // how per-option moment sums become prices / Greeks (shared by the fused and the stand-alone pass)
    double n_paths;          // paths simulated in total (all shards)
    double num_trials;       // divisor, as passed by the caller (mc_simd.rs:77 divides by num_trials)
    const float* rate; const float* years;
    Bumps bumps;
    float* out[EST_COUNT_GREEKS];      // estimator values (nullable)
    float* out_se[EST_COUNT_GREEKS];   // their standard errors (nullable)
};

struct SimArgs {
    OptionArrays opt;
    int n_options;
    int n_paths;             // 8 * floor(num_trials / 8): paths actually simulated per option (mc_simd.rs:49)
    int half_steps;          // floor(steps) / 2 Box-Muller pairs per path (mc_simd.rs:47)
    float steps;             // as passed (drift uses steps * nudt, mc_simd.rs:44)
    int rounds_per_option;   // ceil(n_paths / 32): a round = 32 consecutive paths of one option = one pass of a warp
    long long total_rounds;  // n_options * rounds_per_option
    int n_warps;             // W <= min(gridDim.x * kWarps, R); warp w owns rounds [w*R/W, (w+1)*R/W)
    Bumps bumps;
    uint64_t* pool;          // STREAM_FAST: per-thread generator states, 4 x u64 each, indexed by global thread id
    const uint64_t* chunk_seeds;   // STREAM_REF: n_options * n_paths/8 seeds of 32 x u64 (simd_rand layout)
    double* slots;           // [(n_warps + n_options)][2 * kEst]: partial sums of (warp w, option o) live in slot w + o
    unsigned* arrivals;      // [n_options], zero on entry and on exit: segments of the option that have been flushed
    double* moments_out;     // nullable: per-option moment sums [n_options][2 * kEst] (trials-sharded multi-GPU operation)
    int finalize;            // non-zero: the last warp to finish an option also writes fin.out / fin.out_se
    FinalizeSpec fin;
    float* growth_out;       // DUMP only: [n_options][n_paths] terminal growth factors S_T / spot
    // Small host batches (n_options <= kInlineMaxOptions) carry their six input columns in the launch parameters, so
    // the call needs no H2D copy at all.
    int use_inline;
    unsigned one;            // the value 1, unknown to ptxas (tri_fast: carry adds stay on the FMA-heavy pipe)
#ifdef MCB_TRACE
    unsigned long long* trace;   // tools/kernel_timeline.py builds only: [n_warps][8] %globaltimer stamps (lane 0)
#endif
#ifdef MCB_CHECKED
    unsigned* check;             // tools/checked_build.py builds only: OR of failed index checks (mcb200_debug_check)
    unsigned long long slots_cap, xg_doubles;
    long long pool_threads;
#endif
    // Cross-GPU fused combine (trials-sharded operation without a collective): every GPU stores the finished moment
    // sums of an option into ITS row of an exchange buffer that lives on one GPU (peer memory over NVLink, mapped
    // through CUDA IPC) and bumps the option's cross-GPU arrival counter; the warp -- on whichever GPU -- that completes
    // the count sums the rows in rank order and finalises into the exchange buffer.  Null xg_slots = single-GPU launch.
    double* xg_slots;        // [xg_world][xg_n_max][2 * kEst] on the owner GPU
    unsigned* xg_arrivals;   // [xg_n_max] on the owner GPU, zero between launches
    unsigned* xg_ctl;        // owner GPU: [0] = options finalised in this launch round (self-resetting), [1] = epoch of
                             // the last COMPLETE round, published by the warp that finalises the round's last option
    unsigned xg_epoch;       // this round's epoch (same on every rank)
    int xg_rank, xg_world, xg_n_max;
    float inline_opt[6][kInlineMaxOptions];     // spot, strike, vol, rate, years, div
};

// f32 constants exactly as the reference computes them (mc_simd.rs:36-45), then rescaled to base-2 exponent units.
struct LegConst { float c2, d2; };
template <int STREAM>
__device__ __forceinline__ LegConst leg_constants(float vol, float rate, float div, float dt, float steps) {
    const float nudt = (rate - div - 0.5f * (vol * vol)) * dt;
    const float sidt = vol * sqrtf(dt);
    const float drift = steps * nudt;
    const float diff = kSqrt2 * sidt;
    LegConst c;
    c.c2 = diff * (STREAM == STREAM_FAST ? kSqrt2Ln2 * kLog2e : kLog2e);
    c.d2 = drift * kLog2e;
    return c;
}

// ------------------------------------------------------------------------------------------------ per-option constants
// PRICE / AV : [0]=spot [1]=strike [2]=c2 [3]=d2
// GREEKS     : [0]=spot [1]=strike [2]=c2 [3]=d2 [4]=dspot
//              [5..8]  = c2,d2 for vol+h ; c2,d2 for vol-h
//              [9,10]  = d2 for r+h, r-h          [11,12] = exp(-(r+h)T), exp(-(r-h)T)
//              [13..16]= c2,d2 for T+h ; c2,d2 for T-h    [17,18] = exp(-r(T+h)), exp(-r(T-h))
template <int STREAM, int PAYOFF>
__device__ __forceinline__ void load_option_constants(const SimArgs& a, int o, float* sc) {
    float spot, strike, vol, rate, years, div;
    if (a.use_inline) {
        spot = a.inline_opt[0][o]; strike = a.inline_opt[1][o]; vol = a.inline_opt[2][o];
        rate = a.inline_opt[3][o]; years = a.inline_opt[4][o]; div = a.inline_opt[5][o];
    } else {
        spot = a.opt.spot[o]; strike = a.opt.strike[o]; vol = a.opt.vol[o];
        rate = a.opt.rate[o]; years = a.opt.years[o]; div = a.opt.div[o];
    }
    const LegConst base = leg_constants<STREAM>(vol, rate, div, years / a.steps, a.steps);
    sc[0] = spot; sc[1] = strike; sc[2] = base.c2; sc[3] = base.d2;
    if (PAYOFF == PAYOFF_GREEKS) {
        const Bumps b = a.bumps;
        sc[4] = b.dspot;
        const LegConst vp = leg_constants<STREAM>(vol + b.dvol, rate, div, years / a.steps, a.steps);
        const LegConst vm = leg_constants<STREAM>(vol - b.dvol, rate, div, years / a.steps, a.steps);
        sc[5] = vp.c2; sc[6] = vp.d2; sc[7] = vm.c2; sc[8] = vm.d2;
        const float rp = rate + b.drate, rm = rate - b.drate;
        sc[9] = leg_constants<STREAM>(vol, rp, div, years / a.steps, a.steps).d2;
        sc[10] = leg_constants<STREAM>(vol, rm, div, years / a.steps, a.steps).d2;
        sc[11] = expf(-rp * years); sc[12] = expf(-rm * years);
        const float tp = years + b.dyears, tm = years - b.dyears;
        const LegConst yp = leg_constants<STREAM>(vol, rate, div, tp / a.steps, a.steps);
        const LegConst ym = leg_constants<STREAM>(vol, rate, div, tm / a.steps, a.steps);
        sc[13] = yp.c2; sc[14] = yp.d2; sc[15] = ym.c2; sc[16] = ym.d2;
        sc[17] = expf(-rate * tp); sc[18] = expf(-rate * tm);
        sc[19] = 0.0f;
    }
}

// acc[2e] += x ; acc[2e+1] += x*x  -- per-thread fp32 moment registers of estimator e
#define MCB_ACC(e, x) do { const float x_ = (x); acc[2 * (e)] += x_; acc[2 * (e) + 1] = fmaf(x_, x_, acc[2 * (e) + 1]); } while (0)

// Payoff evaluation for one finished path; m = accumulated Box-Muller sum.  Returns the base growth factor.
// sc points at registers (PRICE / AV) or at the warp's shared-memory constants (GREEKS).
template <int PAYOFF>
__device__ __forceinline__ float evaluate_path(float m, const float* sc, float (&acc)[2 * PayoffTraits<PAYOFF>::kEst]) {
    const float spot = sc[0], strike = sc[1];
    const float g0 = mufu_ex2(fmaf(m, sc[2], sc[3]));
    const float a0 = fmaf(spot, g0, -strike);                // S_T - K
    if (PAYOFF == PAYOFF_PRICE) {
        MCB_ACC(EST_CALL, fmaxf(a0, 0.0f));                  // max(S_T - K, 0)            (mc_simd.rs:62-69)
        MCB_ACC(EST_PUT, fmaxf(-a0, 0.0f));                  // max(-S_T + K, 0): call_mult = -1
    } else if (PAYOFF == PAYOFF_AV) {
        const float g1 = mufu_ex2(fmaf(-m, sc[2], sc[3]));   // antithetic leg: -sqrt(2)*sidt (mc_simd.rs:103)
        const float a1 = fmaf(spot, g1, -strike);
        MCB_ACC(EST_CALL, 0.5f * (fmaxf(a0, 0.0f) + fmaxf(a1, 0.0f)));
        MCB_ACC(EST_PUT, 0.5f * (fmaxf(-a0, 0.0f) + fmaxf(-a1, 0.0f)));
    } else {
        MCB_ACC(EST_CALL, fmaxf(a0, 0.0f));
        MCB_ACC(EST_PUT, fmaxf(-a0, 0.0f));
        // spot +-h share g0 (mc_simd.rs:196-213).  With b = h*g0: pay(S+h) - pay(S-h) = min(max(a0+b,0), 2b) and
        // pay(S+h) - 2 pay(S) + pay(S-h) = max(b - |a0|, 0): the same central differences, without cancellation.
        const float b = sc[4] * g0;
        const float dcall = fminf(fmaxf(a0 + b, 0.0f), 2.0f * b);
        MCB_ACC(EST_CALL_DELTA, dcall);
        MCB_ACC(EST_PUT_DELTA, dcall - 2.0f * b);            // put(S+h) - put(S-h) = call diff - 2 h g0
        MCB_ACC(EST_GAMMA, fmaxf(b - fabsf(a0), 0.0f));
        // vol +-h: both drift and diffusion change (mc_simd.rs:242-250), call only
        const float gvp = mufu_ex2(fmaf(m, sc[5], sc[6])), gvm = mufu_ex2(fmaf(m, sc[7], sc[8]));
        MCB_ACC(EST_VEGA, fmaxf(fmaf(spot, gvp, -strike), 0.0f) - fmaxf(fmaf(spot, gvm, -strike), 0.0f));
        // r +-h: drift changes, each leg discounted with its own rate (mc_simd.rs:328-329,384-385)
        const float arp = fmaf(spot, mufu_ex2(fmaf(m, sc[2], sc[9])), -strike);
        const float arm = fmaf(spot, mufu_ex2(fmaf(m, sc[2], sc[10])), -strike);
        MCB_ACC(EST_CALL_RHO, sc[11] * fmaxf(arp, 0.0f) - sc[12] * fmaxf(arm, 0.0f));
        MCB_ACC(EST_PUT_RHO, sc[11] * fmaxf(-arp, 0.0f) - sc[12] * fmaxf(-arm, 0.0f));
        // T +-h: dt, drift, diffusion and discount change (mc_simd.rs:402-413,470-471); theta = (P(T-h) - P(T+h)) / 2h
        const float atp = fmaf(spot, mufu_ex2(fmaf(m, sc[13], sc[14])), -strike);
        const float atm = fmaf(spot, mufu_ex2(fmaf(m, sc[15], sc[16])), -strike);
        MCB_ACC(EST_CALL_THETA, sc[18] * fmaxf(atm, 0.0f) - sc[17] * fmaxf(atp, 0.0f));
        MCB_ACC(EST_PUT_THETA, sc[18] * fmaxf(-atm, 0.0f) - sc[17] * fmaxf(-atp, 0.0f));
    }
    return g0;
}

// ------------------------------------------------------------------------------------------------ finalisation
// mean = sum / num_trials * scale ; se = sample std of the per-path estimator / sqrt(n) * (n / num_trials) * scale
__device__ __forceinline__ void finalize_estimator(const FinalizeSpec& f, int o, int e, double s, double ss, float rate,
                                                   float years) {
    const float disc = expf(-rate * years);                       // mc_simd.rs:77
    double scale;
    switch (e) {
        case EST_CALL: case EST_PUT: scale = disc; break;
        case EST_CALL_DELTA: case EST_PUT_DELTA: scale = (double)disc / (2.0 * f.bumps.dspot); break;      // :592,618
        case EST_GAMMA: scale = (double)disc / ((double)f.bumps.dspot * f.bumps.dspot); break;              // :644
        case EST_VEGA: scale = (double)disc / (200.0 * f.bumps.dvol); break;                                // :670
        case EST_CALL_RHO: case EST_PUT_RHO: scale = 1.0 / (200.0 * f.bumps.drate); break;                  // :698,725
        default: scale = 1.0 / (2.0 * f.bumps.dyears); break;                                               // :752,778
    }
    const double n = f.n_paths;
    const double mu = n > 0.0 ? s / n : 0.0;
    double var = n > 1.0 ? (ss - n * mu * mu) / (n - 1.0) : 0.0;
    var = var > 0.0 ? var : 0.0;
    if (f.out[e]) f.out[e][o] = (float)(s / f.num_trials * scale);
    if (f.out_se[e]) f.out_se[e][o] = n > 0.0 ? (float)(sqrt(var / n) * (n / f.num_trials) * fabs(scale)) : 0.0f;
}

// Stand-alone pass over moment sums that are already complete (after a cross-GPU all-reduce, or when no path was run).
__global__ void finalize_kernel(const double* __restrict__ moments, int n_options, int n_est, const FinalizeSpec f) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_options * n_est) return;
    const int o = idx / n_est, e = idx - o * n_est;
    finalize_estimator(f, o, e, moments[(size_t)idx * 2], moments[(size_t)idx * 2 + 1], f.rate[o], f.years[o]);
}

// Rounds [w*R/W, (w+1)*R/W) belong to warp w; the warp that owns round r is therefore ceil((r+1)*W/R) - 1.
__device__ __forceinline__ long long round_begin(long long w, long long R, long long W) { return w * R / W; }
__device__ __forceinline__ int owner_warp(long long r, long long R, long long W) { return (int)(((r + 1) * W - 1) / R); }

// ------------------------------------------------------------------------------------------------ the fused kernel
// One launch does the whole batch: simulate, reduce, finalise.
//   * work unit = a "round": 32 consecutive paths of one option, one path per lane.  The n_options * ceil(n_paths/32)
//     rounds are dealt to the resident warps in contiguous, equal (+-1) ranges, so every SM sub-partition gets the same
//     load whatever the batch shape (no CTA-level quantisation, no block barrier anywhere in the path loop);
//   * the inner step is tri_fast(): two xoshiro256++ outputs feed three Box-Muller pairs (mcb200_pair.cuh);
//   * every thread owns one generator stream for the whole launch (state loaded from / stored back to the pool, so
//     consecutive launches continue the same substreams) and keeps its fp32 moment sums in registers;
//   * when a warp leaves an option it reduces its sums (fp64, butterfly) into slot (warp + option) and bumps the
//     option's arrival counter; the warp that completes the count sums the option's slots in warp order (fixed order:
//     bit-reproducible, independent of which warp arrives last) and writes moments and/or final values and SEs.
// Path data never leaves registers.
#ifndef MCB_LAZY_SLOT
#define MCB_LAZY_SLOT(payoff) ((payoff) == PAYOFF_GREEKS)
#endif

#ifdef MCB_CHECKED
#define MCB_CHECK(cond, bit) do { if (!(cond)) atomicOr(a.check, (unsigned)(bit)); } while (0)
#else
#define MCB_CHECK(cond, bit) do { } while (0)
#endif

#ifdef MCB_TRACE
#define MCB_STAMP(k) do { if (a.trace && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); \
                                                     a.trace[(size_t)w * 8 + (k)] = t_; } } while (0)
#else
#define MCB_STAMP(k) do { } while (0)
#endif

template <int STREAM, int PAYOFF, bool DUMP>
__global__ void __launch_bounds__(kThreads, PayoffTraits<PAYOFF>::kCtasPerSm) simulate_kernel(const __grid_constant__ SimArgs a) {
    constexpr int kEst = PayoffTraits<PAYOFF>::kEst;
    constexpr int kAcc = 2 * kEst;
    constexpr int kConst = PayoffTraits<PAYOFF>::kConst;
    constexpr bool kConstInSmem = PAYOFF == PAYOFF_GREEKS;
    constexpr int kQuadUnroll_ = PayoffTraits<PAYOFF>::kQuadUnroll;
    __shared__ float s_const[kConstInSmem ? kWarps * kConst : 1];

    const int lane = threadIdx.x & 31;
    // Logical warp index.  Hardware warp h of a CTA runs on SM sub-partition h % 4; giving the two warps of a
    // sub-partition CONSECUTIVE logical indices keeps the +-1-round differences of the range partition from lining up
    // with the sub-partitions (with R/W = 1.5 every odd logical warp has the extra round).
    const int hw = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarps + ((hw & 3) * 2 + (hw >> 2));
    const long long R = a.total_rounds, W = a.n_warps;
    if (w >= a.n_warps) return;                              // tail of the last CTA when W is not a multiple of 8
    long long r = round_begin(w, R, W);
    const long long r_end = round_begin(w + 1, R, W);        // W <= R: never empty
    MCB_STAMP(0);
#ifdef MCB_TRACE
    if (a.trace && lane == 0) {                               // where this warp runs: SM, hardware warp slot, CTA, warp in CTA
        unsigned smid_, warpid_;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid_));
        a.trace[(size_t)w * 8 + 5] = ((unsigned long long)smid_ << 32) | warpid_;
        a.trace[(size_t)w * 8 + 6] = ((unsigned long long)blockIdx.x << 32) | (unsigned)hw;
    }
#endif

    Xoshiro256 g;
    const size_t gtid = (size_t)w * 32 + lane;              // the generator stream belongs to the LOGICAL thread
    MCB_CHECK(STREAM != STREAM_FAST || (long long)gtid < a.pool_threads, 1);
    if (STREAM == STREAM_FAST) xoshiro_load(g, a.pool + 4ull * gtid);

    float creg[kConstInSmem ? 1 : kConst];
    float* sc = kConstInSmem ? s_const + (threadIdx.x >> 5) * kConst : creg;
    float acc[kAcc];
    int o = (int)(r / a.rounds_per_option);
    int k = (int)(r - (long long)o * a.rounds_per_option);   // round index inside option o
    bool enter = true;                                       // constants of option o not loaded yet
    bool fresh = true;                                       // next flush is the first of this (warp, option) segment
    MCB_STAMP(1);

    while (r < r_end) {
        MCB_CHECK(o >= 0 && o < a.n_options && k >= 0 && k < a.rounds_per_option, 2);
        if (enter) {
            if (kConstInSmem) {
                __syncwarp();
                if (lane == 0) load_option_constants<STREAM, PAYOFF>(a, o, sc);
                __syncwarp();
            } else {
                load_option_constants<STREAM, PAYOFF>(a, o, sc);
            }
        }
#pragma unroll
        for (int i = 0; i < kAcc; ++i) acc[i] = 0.0f;

        // ---- a run of this warp's rounds inside option o (bounded so that the fp32 sums stay well conditioned)
        long long left = r_end - r;
        int n_run = a.rounds_per_option - k;
        if ((long long)n_run > left) n_run = (int)left;
        if (n_run > kMaxRoundsPerFlush) n_run = kMaxRoundsPerFlush;
        for (int j = 0; j < n_run; ++j, ++k) {
            const int p = k * 32 + lane;
            if (p < a.n_paths) {
                if (STREAM == STREAM_REF) {
                    // simd_rand seed layout: word j of lane l at u64 index 8*j + l of the chunk's 32-word seed
                    const uint64_t* seed = a.chunk_seeds + ((size_t)o * (a.n_paths >> 3) + (p >> 3)) * 32 + (p & 7);
                    const uint64_t w0 = seed[0], w1 = seed[8], w2 = seed[16], w3 = seed[24];
                    g.s0l = (uint32_t)w0; g.s0h = (uint32_t)(w0 >> 32); g.s1l = (uint32_t)w1; g.s1h = (uint32_t)(w1 >> 32);
                    g.s2l = (uint32_t)w2; g.s2h = (uint32_t)(w2 >> 32); g.s3l = (uint32_t)w3; g.s3h = (uint32_t)(w3 >> 32);
                }
                float m = 0.0f;
                if (STREAM == STREAM_FAST) {
                    // three pairs per two generator outputs, then the 0..2 left-over pairs one output each
                    const int tris = a.half_steps / 3;
#pragma unroll kQuadUnroll_
                    for (int i = 0; i < tris; ++i) m = tri_fast(g, m, a.one);
                    for (int i = tris * 3; i < a.half_steps; ++i) m = pair_fast(g, m);
                } else {
#pragma unroll 2
                    for (int i = 0; i < a.half_steps; ++i) m = pair_ref(g, m);
                }
                const float g0 = evaluate_path<PAYOFF>(m, sc, acc);
                if (DUMP) a.growth_out[(size_t)o * a.n_paths + p] = g0;
            }
        }
        r += n_run;
        MCB_STAMP(2);
        const bool option_done = (k == a.rounds_per_option);
        const bool segment_done = option_done || r == r_end;

        // ---- flush: fp32 thread sums -> fp64, butterfly within the warp; lane i keeps accumulator i
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < kAcc; ++i) {
            double v = (double)acc[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == i) mine = v;
        }
        // A (warp, option) segment longer than kMaxRoundsPerFlush rounds is flushed in pieces that accumulate in the
        // slot.  An option owned by one warp alone ("solo", most options of a large sweep) is finalised straight from
        // registers: no fence, no arrival.  The Greeks kernel also skips the solo option's slot store (160 bytes); the
        // price kernels keep that 32-byte store, because hoisting the ownership test above it makes ptxas schedule
        // their path loop 3 % slower (profiles/r2_ab_kernel_variants.txt) -- the loop is ALU-pipe bound and its
        // instruction order is what the time depends on.
        constexpr bool kLazySlot = MCB_LAZY_SLOT(PAYOFF);
        long long first_round = 0;
        int w_first = 0, w_last = 0;
        bool solo = false;                                   // the whole option lives in this warp's range
        if (kLazySlot) {
            first_round = (long long)o * a.rounds_per_option;
            w_first = owner_warp(first_round, R, W);
            w_last = owner_warp(first_round + a.rounds_per_option - 1, R, W);
            solo = (w_first == w_last);
        }
        double* slot = a.slots + ((size_t)w + (size_t)o) * kAcc;
        MCB_CHECK(((size_t)w + (size_t)o + 1) * kAcc <= a.slots_cap, 4);
        if (lane < kAcc) {
            if (!fresh) mine += slot[lane];                  // only this warp touches the slot before its arrival
            if (!kLazySlot || !(solo && segment_done)) slot[lane] = mine;
        }
        if (segment_done) {
            if (!kLazySlot) {
                first_round = (long long)o * a.rounds_per_option;
                w_first = owner_warp(first_round, R, W);
                w_last = owner_warp(first_round + a.rounds_per_option - 1, R, W);
                solo = (w_first == w_last);
            }
            bool complete = solo;
            double s = mine;                                 // lane i < kAcc holds accumulator i
            if (!solo) {
                // publish the slot and count the arrival; the warp that completes the option finishes it
                __threadfence();
                __syncwarp();
                unsigned prev = 0;
                if (lane == 0) prev = atomicAdd(a.arrivals + o, 1u);
                prev = __shfl_sync(0xffffffffu, prev, 0);
                MCB_CHECK((int)prev <= w_last - w_first, 8);
                complete = ((int)prev == w_last - w_first);
            }
            if (complete && !solo) {
                __threadfence();
                // Fixed-order sum of the option's segments.  Lane l = (group, accumulator): kGroups groups stride over
                // the segments with independent loads in flight (a serial walk would expose one L2 round trip per
                // segment at the very end of the kernel), then the groups are combined by a fixed butterfly.
                constexpr int kAccPad = kAcc <= 4 ? 4 : (kAcc <= 8 ? 8 : (kAcc <= 16 ? 16 : 32));
                constexpr int kGroups = 32 / kAccPad;
                const int i = lane % kAccPad, grp = lane / kAccPad;
                const int n_seg = w_last - w_first + 1;
                MCB_CHECK(((size_t)w_last + (size_t)o + 1) * kAcc <= a.slots_cap && w_first >= 0 && w_last < a.n_warps, 16);
                s = 0.0;
                if (i < kAcc) {
                    const double* base = a.slots + ((size_t)w_first + (size_t)o) * kAcc + i;
                    int j = grp;
                    for (; j + 3 * kGroups < n_seg; j += 4 * kGroups) {
                        const double v0 = __ldcg(base + (size_t)j * kAcc);
                        const double v1 = __ldcg(base + (size_t)(j + kGroups) * kAcc);
                        const double v2 = __ldcg(base + (size_t)(j + 2 * kGroups) * kAcc);
                        const double v3 = __ldcg(base + (size_t)(j + 3 * kGroups) * kAcc);
                        s += v0; s += v1; s += v2; s += v3;
                    }
                    for (; j < n_seg; j += kGroups) s += __ldcg(base + (size_t)j * kAcc);
                }
#pragma unroll
                for (int off = kAccPad; off < 32; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) a.arrivals[o] = 0u;           // leave the counter clean for the next launch
            }
            if (complete) {
                if (a.moments_out && lane < kAcc) a.moments_out[(size_t)o * kAcc + lane] = s;
                bool finish = a.finalize != 0;
                if (a.xg_slots) {
                    // this GPU's share of option o is complete: publish it to the owner GPU, count the arrival there
                    MCB_CHECK(((size_t)a.xg_rank * a.xg_n_max + (size_t)o + 1) * kAcc <= a.xg_doubles && o < a.xg_n_max, 32);
                    volatile double* row = a.xg_slots + ((size_t)a.xg_rank * a.xg_n_max + (size_t)o) * kAcc;
                    if (lane < kAcc) row[lane] = s;
                    __threadfence_system();
                    __syncwarp();
                    unsigned seen = 0;
                    if (lane == 0) seen = atomicAdd_system(a.xg_arrivals + o, 1u);
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                    MCB_CHECK((int)seen < a.xg_world, 32);
                    finish = finish && (int)seen == a.xg_world - 1;
                    if (finish) {
                        __threadfence_system();
                        s = 0.0;
                        if (lane < kAcc)
                            for (int rk = 0; rk < a.xg_world; ++rk) {       // rank order: deterministic
                                const volatile double* src = a.xg_slots + ((size_t)rk * a.xg_n_max + (size_t)o) * kAcc;
                                s += src[lane];
                            }
                        if (lane == 0) *(volatile unsigned*)(a.xg_arrivals + o) = 0u;
                    }
                }
                if (finish) {
                    const double sum = __shfl_sync(0xffffffffu, s, (2 * lane) & 31);
                    const double sq = __shfl_sync(0xffffffffu, s, (2 * lane + 1) & 31);
                    if (lane < kEst)
                        finalize_estimator(a.fin, o, lane, sum, sq, a.use_inline ? a.inline_opt[3][o] : a.opt.rate[o],
                                           a.use_inline ? a.inline_opt[4][o] : a.opt.years[o]);
                }
                if (a.xg_slots && finish) {
                    // results of option o are in the exchange buffer: count it; the last one publishes the epoch that
                    // every rank's xgpu_wait_kernel is spinning on
                    __threadfence_system();
                    __syncwarp();
                    if (lane == 0 && atomicAdd_system(a.xg_ctl, 1u) == (unsigned)a.n_options - 1u) {
                        *(volatile unsigned*)a.xg_ctl = 0u;
                        __threadfence_system();
                        *(volatile unsigned*)(a.xg_ctl + 1) = a.xg_epoch;
                    }
                }
            }
        }
        MCB_STAMP(3);
        if (option_done) { ++o; k = 0; enter = true; fresh = true; }
        else { enter = false; fresh = false; }
    }
    if (STREAM == STREAM_FAST) xoshiro_store(g, a.pool + 4ull * gtid);
    MCB_STAMP(4);
}

// Device-side completion wait of the cross-GPU fused combine: one thread spins on the epoch word of the exchange
// buffer (peer memory) until the round `epoch` is complete on ALL ranks; work queued behind it on the stream (the copy
// of the results) then sees the whole sample.  Gives up after ~timeout_ns and reports through *status.
__global__ void xgpu_wait_kernel(const unsigned* ctl, unsigned epoch, unsigned long long timeout_ns, int* status) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    int ok = 1;
    while (*(volatile const unsigned*)(ctl + 1) != epoch) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { ok = 0; break; }
        __nanosleep(200);
    }
    __threadfence_system();
    *status = ok;
}

// ------------------------------------------------------------------------------------------------ closed form
// Black-Scholes with continuous dividend yield for a whole batch: the analytic value every reference test compares a
// Monte Carlo estimate with (/root/reference/src/bs.rs:42-184; units as there: vega and rho per 1 point, theta per
// year).  Evaluated in fp64 with erfc (bs.rs uses an f32 Abramowitz-Stegun erf; the difference is below 1e-6 relative),
// so batched results can be turned into z-scores on the device.  out[e] (nullable) in Estimator order.
struct BsArgs {
    OptionArrays opt;
    int n;
    float* out[EST_COUNT_GREEKS];
};

__global__ void bs_closed_form_kernel(const BsArgs a) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= a.n) return;
    const double s = a.opt.spot[o], k = a.opt.strike[o], v = a.opt.vol[o];
    const double r = a.opt.rate[o], t = a.opt.years[o], q = a.opt.div[o];
    const double sq = sqrt(t), vs = v * sq;
    const double d1 = (log(s / k) + (r - q + 0.5 * v * v) * t) / vs, d2 = d1 - vs;
    const double nd1 = 0.5 * erfc(-d1 * 0.70710678118654752440), nd2 = 0.5 * erfc(-d2 * 0.70710678118654752440);
    const double md1 = 0.5 * erfc(d1 * 0.70710678118654752440), md2 = 0.5 * erfc(d2 * 0.70710678118654752440);   // N(-d)
    const double pdf = exp(-0.5 * d1 * d1) * 0.39894228040143267794;
    const double eq = exp(-q * t), er = exp(-r * t);
    const double decay = -eq * s * pdf * v / (2.0 * sq);
    double val[EST_COUNT_GREEKS];
    val[EST_CALL] = s * eq * nd1 - k * er * nd2;
    val[EST_PUT] = k * er * md2 - s * eq * md1;
    val[EST_CALL_DELTA] = eq * nd1;
    val[EST_PUT_DELTA] = -eq * md1;
    val[EST_GAMMA] = eq * pdf / (s * vs);
    val[EST_VEGA] = eq * pdf * s * sq / 100.0;
    val[EST_CALL_RHO] = k * t * er * nd2 / 100.0;
    val[EST_PUT_RHO] = -k * t * er * md2 / 100.0;
    val[EST_CALL_THETA] = decay - r * k * er * nd2 + q * s * eq * nd1;
    val[EST_PUT_THETA] = decay + r * k * er * md2 - q * s * eq * md1;
#pragma unroll
    for (int e = 0; e < EST_COUNT_GREEKS; ++e)
        if (a.out[e]) a.out[e][o] = (float)val[e];
}

// ------------------------------------------------------------------------------------------------ RNG substreams
// pool[i] = jump^(first + i)(base): thread i applies the precomputed polynomials x^(2^(128+k)) for the set bits of
// (first + i).  Replaces per-chunk OS-entropy seeding (mc_simd.rs:52-54) with non-overlapping 2^128-long substreams.
__global__ void seed_pool_kernel(uint64_t* __restrict__ pool, int count, uint64_t first, ulonglong4 base,
                                 const uint64_t* __restrict__ jump_pow2, int n_pow2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Xoshiro256 g;
    g.s0l = (uint32_t)base.x; g.s0h = (uint32_t)(base.x >> 32); g.s1l = (uint32_t)base.y; g.s1h = (uint32_t)(base.y >> 32);
    g.s2l = (uint32_t)base.z; g.s2h = (uint32_t)(base.z >> 32); g.s3l = (uint32_t)base.w; g.s3h = (uint32_t)(base.w >> 32);
    const uint64_t idx = first + (uint64_t)i;
    for (int k = 0; k < n_pow2; ++k) {
        if ((idx >> k) & 1ull) {
            uint64_t poly[4] = {jump_pow2[4 * k], jump_pow2[4 * k + 1], jump_pow2[4 * k + 2], jump_pow2[4 * k + 3]};
            xoshiro_apply_poly(g, poly);
        }
    }
    xoshiro_store(g, pool + 4ull * i);
}

}  // namespace mcb200
