// mcb200_kernels.cuh -- the fused GBM path + payoff kernels (sm_100a) and their fixed-order second pass.
//
// One kernel family replaces the reference's inner step and its six simulation cores:
//   speed_update                    /root/reference/src/mc_simd.rs:9-21      -> pair_fast() / pair_ref()
//   get_rand_uniform_f32x8          /root/reference/src/rand32x8.rs:5-21     -> uniforms built in pair_*()
//   monte_carlo_pricing             src/mc_simd.rs:25-79                      -> simulate_kernel<*, PAYOFF_PRICE>
//   monte_carlo_av_pricing          src/mc_simd.rs:82-149                     -> simulate_kernel<*, PAYOFF_AV>
//   monte_carlo_{spot,volatility,interest,time}_pricing  src/mc_simd.rs:152-473 -> simulate_kernel<*, PAYOFF_GREEKS>
//   rayon map/reduce + reduce_add + discount (mc_simd.rs:49-51,71-78)        -> block_reduce() + finalize kernels
// Path data never leaves registers; per option-slice a CTA writes NACC fp64 partial sums.
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

template <int PAYOFF> struct PayoffTraits { static constexpr int kEst = 2; static constexpr int kConst = 4; };
template <> struct PayoffTraits<PAYOFF_GREEKS> { static constexpr int kEst = 10; static constexpr int kConst = 20; };

struct OptionArrays {          // struct-of-arrays option batch in HBM (f32, n entries each)
    const float* spot; const float* strike; const float* vol; const float* rate; const float* years; const float* div;
};

struct Bumps { float dspot, dvol, drate, dyears; };

struct SimArgs {
    OptionArrays opt;
    int n_options;
    int n_paths;           // 8 * floor(num_trials / 8): paths actually simulated per option (mc_simd.rs:49)
    int half_steps;        // floor(steps) / 2 Box-Muller pairs per path (mc_simd.rs:47)
    float steps;           // as passed (drift uses steps * nudt, mc_simd.rs:44)
    int slices;            // CTAs' worth of work per option
    int paths_per_slice;
    Bumps bumps;
    uint64_t* pool;        // STREAM_FAST: per-thread generator states, 4 x u64 each, indexed by resident thread slot
    const uint64_t* chunk_seeds;   // STREAM_REF: n_options * n_paths/8 seeds of 32 x u64 (simd_rand layout)
    double* partials;      // [n_options * slices][2 * kEst]
    float* growth_out;     // DUMP only: [n_options][n_paths] terminal growth factors S_T / spot
    uint32_t one;          // the value 1, delivered at run time so ptxas keeps IMAD.WIDE adds (mcb200_rng.cuh)
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

// ------------------------------------------------------------------------------------------------ constants in smem
// PRICE / AV : [0]=spot [1]=strike [2]=c2 [3]=d2
// GREEKS     : [0]=spot [1]=strike [2]=c2 [3]=d2 [4]=dspot
//              [5..8]  = c2,d2 for vol+h ; c2,d2 for vol-h
//              [9,10]  = d2 for r+h, r-h          [11,12] = exp(-(r+h)T), exp(-(r-h)T)
//              [13..16]= c2,d2 for T+h ; c2,d2 for T-h    [17,18] = exp(-r(T+h)), exp(-r(T-h))
template <int STREAM, int PAYOFF>
__device__ __forceinline__ void load_item_constants(const SimArgs& a, int o, float* sc) {
    const float spot = a.opt.spot[o], strike = a.opt.strike[o], vol = a.opt.vol[o];
    const float rate = a.opt.rate[o], years = a.opt.years[o], div = a.opt.div[o];
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

__device__ __forceinline__ void accumulate(float* acc, int est, float x) {
    // acc is this thread's column of the [2*kEst][kThreads] shared accumulator tile
    acc[(2 * est) * kThreads] += x;
    acc[(2 * est + 1) * kThreads] = fmaf(x, x, acc[(2 * est + 1) * kThreads]);
}

// Payoff evaluation for one finished path; m = accumulated Box-Muller sum.  Returns the base growth factor.
template <int PAYOFF>
__device__ __forceinline__ float evaluate_path(float m, const float* sc, float* acc) {
    const float spot = sc[0], strike = sc[1];
    const float g0 = mufu_ex2(fmaf(m, sc[2], sc[3]));
    const float a0 = fmaf(spot, g0, -strike);                // S_T - K
    if (PAYOFF == PAYOFF_PRICE) {
        accumulate(acc, EST_CALL, fmaxf(a0, 0.0f));          // max(S_T - K, 0)            (mc_simd.rs:62-69)
        accumulate(acc, EST_PUT, fmaxf(-a0, 0.0f));          // max(-S_T + K, 0): call_mult = -1
    } else if (PAYOFF == PAYOFF_AV) {
        const float g1 = mufu_ex2(fmaf(-m, sc[2], sc[3]));   // antithetic leg: -sqrt(2)*sidt (mc_simd.rs:103)
        const float a1 = fmaf(spot, g1, -strike);
        accumulate(acc, EST_CALL, 0.5f * (fmaxf(a0, 0.0f) + fmaxf(a1, 0.0f)));
        accumulate(acc, EST_PUT, 0.5f * (fmaxf(-a0, 0.0f) + fmaxf(-a1, 0.0f)));
    } else {
        accumulate(acc, EST_CALL, fmaxf(a0, 0.0f));
        accumulate(acc, EST_PUT, fmaxf(-a0, 0.0f));
        // spot +-h share g0 (mc_simd.rs:196-213).  With b = h*g0: pay(S+h) - pay(S-h) = min(max(a0+b,0), 2b) and
        // pay(S+h) - 2 pay(S) + pay(S-h) = max(b - |a0|, 0): the same central differences, without cancellation.
        const float b = sc[4] * g0;
        const float dcall = fminf(fmaxf(a0 + b, 0.0f), 2.0f * b);
        accumulate(acc, EST_CALL_DELTA, dcall);
        accumulate(acc, EST_PUT_DELTA, dcall - 2.0f * b);    // put(S+h) - put(S-h) = call diff - 2 h g0
        accumulate(acc, EST_GAMMA, fmaxf(b - fabsf(a0), 0.0f));
        // vol +-h: both drift and diffusion change (mc_simd.rs:242-250), call only
        const float gvp = mufu_ex2(fmaf(m, sc[5], sc[6])), gvm = mufu_ex2(fmaf(m, sc[7], sc[8]));
        accumulate(acc, EST_VEGA, fmaxf(fmaf(spot, gvp, -strike), 0.0f) - fmaxf(fmaf(spot, gvm, -strike), 0.0f));
        // r +-h: drift changes, each leg discounted with its own rate (mc_simd.rs:328-329,384-385)
        const float arp = fmaf(spot, mufu_ex2(fmaf(m, sc[2], sc[9])), -strike);
        const float arm = fmaf(spot, mufu_ex2(fmaf(m, sc[2], sc[10])), -strike);
        accumulate(acc, EST_CALL_RHO, sc[11] * fmaxf(arp, 0.0f) - sc[12] * fmaxf(arm, 0.0f));
        accumulate(acc, EST_PUT_RHO, sc[11] * fmaxf(-arp, 0.0f) - sc[12] * fmaxf(-arm, 0.0f));
        // T +-h: dt, drift, diffusion and discount change (mc_simd.rs:402-413,470-471); theta = (P(T-h) - P(T+h)) / 2h
        const float atp = fmaf(spot, mufu_ex2(fmaf(m, sc[13], sc[14])), -strike);
        const float atm = fmaf(spot, mufu_ex2(fmaf(m, sc[15], sc[16])), -strike);
        accumulate(acc, EST_CALL_THETA, sc[18] * fmaxf(atm, 0.0f) - sc[17] * fmaxf(atp, 0.0f));
        accumulate(acc, EST_PUT_THETA, sc[18] * fmaxf(-atm, 0.0f) - sc[17] * fmaxf(-atp, 0.0f));
    }
    return g0;
}

// ------------------------------------------------------------------------------------------------ the fused kernel
// Persistent CTAs: CTA c works on items c, c + gridDim.x, ... where item = (option, slice).  Thread t of the CTA
// simulates paths slice_base + t + 256*j.  In STREAM_FAST every thread owns one generator stream for the whole
// launch (state loaded from / stored back to the pool, so consecutive launches continue the same substreams).
template <int STREAM, int PAYOFF, bool DUMP, int MIN_CTAS>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) simulate_kernel(const SimArgs a) {
    constexpr int kEst = PayoffTraits<PAYOFF>::kEst;
    constexpr int kAcc = 2 * kEst;
    __shared__ float s_acc[kAcc * kThreads];
    __shared__ float s_const[PayoffTraits<PAYOFF>::kConst];
    __shared__ double s_warp[kWarps][kAcc];

    const int t = threadIdx.x;
    const uint32_t one = a.one;
    Xoshiro256 g;
    if (STREAM == STREAM_FAST) xoshiro_load(g, a.pool + 4ull * ((size_t)blockIdx.x * kThreads + t));

    const int n_items = a.n_options * a.slices;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int o = item / a.slices, slice = item - o * a.slices;
        __syncthreads();                                   // previous item's s_const / s_warp readers are done
        if (t == 0) load_item_constants<STREAM, PAYOFF>(a, o, s_const);
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s_acc[k * kThreads + t] = 0.0f;
        __syncthreads();

        const int p_begin = slice * a.paths_per_slice;
        const int p_end = min(p_begin + a.paths_per_slice, a.n_paths);
        for (int p = p_begin + t; p < p_end; p += kThreads) {
            if (STREAM == STREAM_REF) {
                // simd_rand seed layout: word k of lane l at u64 index 8*k + l of the chunk's 32-word seed
                const uint64_t* seed = a.chunk_seeds + ((size_t)o * (a.n_paths >> 3) + (p >> 3)) * 32 + (p & 7);
                const uint64_t w0 = seed[0], w1 = seed[8], w2 = seed[16], w3 = seed[24];
                g.s0l = (uint32_t)w0; g.s0h = (uint32_t)(w0 >> 32); g.s1l = (uint32_t)w1; g.s1h = (uint32_t)(w1 >> 32);
                g.s2l = (uint32_t)w2; g.s2h = (uint32_t)(w2 >> 32); g.s3l = (uint32_t)w3; g.s3h = (uint32_t)(w3 >> 32);
            }
            float m = 0.0f;
            if (STREAM == STREAM_FAST) {
#pragma unroll 4
                for (int i = 0; i < a.half_steps; ++i) m = pair_fast(g, m, one);
            } else {
#pragma unroll 2
                for (int i = 0; i < a.half_steps; ++i) m = pair_ref(g, m);
            }
            const float g0 = evaluate_path<PAYOFF>(m, s_const, s_acc + t);
            if (DUMP) a.growth_out[(size_t)o * a.n_paths + p] = g0;
        }

        // deterministic reduction: fp32 thread sums -> fp64, butterfly within the warp (4 accumulators at a time to
        // bound register use), then a fixed-order sum across the 8 warps
#pragma unroll 1
        for (int k0 = 0; k0 < kAcc; k0 += 4) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (double)s_acc[(k0 + k) * kThreads + t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
            }
            if ((t & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s_warp[t >> 5][k0 + k] = v[k];
            }
        }
        __syncthreads();
        if (t < kAcc) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += s_warp[w][t];
            a.partials[(size_t)item * kAcc + t] = s;
        }
    }
    if (STREAM == STREAM_FAST) xoshiro_store(g, a.pool + 4ull * ((size_t)blockIdx.x * kThreads + t));
}

// ------------------------------------------------------------------------------------------------ second pass
// Fixed-order sum over an option's slices (index order, fp64) -> per-option moment sums [n_options][2*kEst].
__global__ void sum_slices_kernel(const double* __restrict__ partials, int n_options, int slices, int n_acc,
                                  double* __restrict__ moments) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_options * n_acc) return;
    const int o = idx / n_acc, k = idx - o * n_acc;
    double s = 0.0;
    for (int sl = 0; sl < slices; ++sl) s += partials[((size_t)o * slices + sl) * n_acc + k];
    moments[idx] = s;
}

struct FinalizeArgs {
    const double* moments;   // [n_options][2*n_est], possibly summed over several devices
    int n_options;
    int n_est;
    double n_paths;          // paths simulated in total (all shards)
    double num_trials;       // divisor, as passed by the caller (mc_simd.rs:77 divides by num_trials)
    const float* rate; const float* years;
    Bumps bumps;
    float* out[EST_COUNT_GREEKS];      // estimator values (nullable)
    float* out_se[EST_COUNT_GREEKS];   // their standard errors (nullable)
};

// mean = sum / num_trials * scale ; se = sample std of the per-path estimator / sqrt(n) * (n / num_trials) * scale
__global__ void finalize_kernel(const FinalizeArgs f) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= f.n_options) return;
    const float disc = expf(-f.rate[o] * f.years[o]);             // mc_simd.rs:77
    const double n = f.n_paths;
    for (int e = 0; e < f.n_est; ++e) {
        double scale;
        switch (e) {
            case EST_CALL: case EST_PUT: scale = disc; break;
            case EST_CALL_DELTA: case EST_PUT_DELTA: scale = (double)disc / (2.0 * f.bumps.dspot); break;      // :592,618
            case EST_GAMMA: scale = (double)disc / ((double)f.bumps.dspot * f.bumps.dspot); break;              // :644
            case EST_VEGA: scale = (double)disc / (200.0 * f.bumps.dvol); break;                                // :670
            case EST_CALL_RHO: case EST_PUT_RHO: scale = 1.0 / (200.0 * f.bumps.drate); break;                  // :698,725
            default: scale = 1.0 / (2.0 * f.bumps.dyears); break;                                               // :752,778
        }
        const double s = f.moments[((size_t)o * f.n_est + e) * 2], ss = f.moments[((size_t)o * f.n_est + e) * 2 + 1];
        const double mu = n > 0.0 ? s / n : 0.0;
        double var = n > 1.0 ? (ss - n * mu * mu) / (n - 1.0) : 0.0;
        var = var > 0.0 ? var : 0.0;
        if (f.out[e]) f.out[e][o] = (float)(s / f.num_trials * scale);
        if (f.out_se[e]) f.out_se[e][o] = n > 0.0 ? (float)(sqrt(var / n) * (n / f.num_trials) * fabs(scale)) : 0.0f;
    }
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
