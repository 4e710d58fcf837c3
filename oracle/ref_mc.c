/* oracle/ref_mc.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * Seed-injectable CPU restatement of the reference's Monte Carlo hot path, module mc_simd
 * (/root/reference/src/mc_simd.rs).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call into this file; the product (libmcb200.so) never links or loads it.
 *
 * What is restated, and where it lives in the reference:
 *   - the 8-lane uniform adapter                 src/rand32x8.rs:5-21    -> ref_uniform_f32()
 *   - the pair update ("speed_update")           src/mc_simd.rs:9-21     -> ref_pair_increment()
 *   - the six simulation cores                   src/mc_simd.rs:25-79 (plain), :82-149 (antithetic),
 *     :152-227 (spot bumps), :230-309 (vol bumps), :312-387 (rate bumps), :390-473 (expiry bumps)
 *                                                                         -> ref_mc_run() with a bump table
 *   - the 12 public wrappers / finite differences src/mc_simd.rs:477-779 -> ref_mc_call_price() ... ref_mc_put_theta()
 *
 * The reference draws a fresh OS-entropy seed per 8-path chunk (mc_simd.rs:52-54) and is therefore not
 * reproducible; here the 256 seed bytes of every chunk are an INPUT (chunk_seeds), which is the only change.
 *
 * PARITY UNPINNED at the third-party boundary (crate sources absent from /root/reference, no exact vectors in the
 * reference's tests -- only the moment tests rand32x8.rs:68-81):
 *   - simd_rand @ b8c0597 (Cargo.lock:716-718): seed layout assumed struct-of-arrays (state word k of lane l =
 *     little-endian u64 at byte 64*k + 8*l); next_f64x8 assumed (x >> 11) * 2^-53.
 *   - wide 0.7.13 (Cargo.lock:867-869): f32x8 ln / sin_cos / exp are ~1-2 ulp polynomials; libm's logf/sinf/cosf/
 *     expf (<= 1 ulp) stand in for them; mul_add / mul_sub are taken as fused (README.md:87-92 builds with
 *     target-cpu=native).
 *   - rayon's reduction order is unspecified (mc_simd.rs:71-74); the oracle adds chunk vectors in chunk order.
 * The reference's own tests compare against the bs.rs closed form with absolute tolerances (mc_simd.rs:781-907);
 * tests/ replays those through this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "xoshiro256pp.h"

#define REF_LANES 8
#define REF_SEED_BYTES 256

/* The two conventions of simd_rand that cannot be read offline (see PARITY UNPINNED above) are switchable, so that
 * vectors dumped from the real crate (tools/ref_vectors.rs -> tests/golden/ref_vectors.json) can falsify them in one
 * run: tests/test_oracle.py::test_reference_vectors_when_present tries every combination and names the one that fits.
 *   seed layout   0 (assumed): struct-of-arrays, state word k of lane l at byte 64*k + 8*l
 *                 1          : array-of-structs, state word k of lane l at byte 32*l + 8*k
 *   uniform mode  0 (assumed): (x >> 11) * 2^-53
 *                 1          : the exponent trick, bits(0x3ff << 52 | x >> 12) - 1.0  (52 bits)               */
static int g_seed_layout = 0, g_uniform_mode = 0;
void ref_mc_set_conventions(int seed_layout, int uniform_mode) { g_seed_layout = seed_layout; g_uniform_mode = uniform_mode; }

/* ---- rand32x8.rs:5-21 : one uniform per lane, u64 -> f64 in [0,1) -> f32 (round to nearest; may give 1.0f) ---- */
static inline float ref_uniform_f32(xo256_t *lane) {
    const uint64_t x = xo256pp_next(lane);
    double d;
    if (g_uniform_mode == 0) {
        d = (double)(x >> 11) * 0x1.0p-53;
    } else {
        const uint64_t b = (0x3ffULL << 52) | (x >> 12);
        memcpy(&d, &b, 8);
        d -= 1.0;
    }
    return (float)d;
}

/* ---- mc_simd.rs:9-21 : two time steps for one lane ---- */
static inline float ref_pair_increment(float two_pi, float acc, float u1, float u2) {
    const float ang = two_pi * u2;
    const float sn = sinf(ang), cs = cosf(ang);
    return fmaf(sqrtf(-logf(u1)), sn + cs, acc);
}

static void ref_seed_chunk(const uint8_t *seed256, xo256_t lanes[REF_LANES]) {
    for (int k = 0; k < 4; k++)
        for (int l = 0; l < REF_LANES; l++) {
            uint64_t w;
            memcpy(&w, seed256 + (g_seed_layout == 0 ? 64 * k + 8 * l : 32 * l + 8 * k), 8);   /* little-endian host */
            lanes[l].s[k] = w;
        }
}

/* Accumulated Box-Muller sum M for the 8 paths of one chunk (the reference's stock_price_mult before scaling). */
void ref_mc_chunk_sums(const uint8_t *seed256, int half_steps, float out_m[REF_LANES]) {
    xo256_t lanes[REF_LANES];
    ref_seed_chunk(seed256, lanes);
    const float two_pi = 2.0f * 3.14159265358979323846f;
    float acc[REF_LANES] = {0};
    float u1[REF_LANES], u2[REF_LANES];
    for (int i = 0; i < half_steps; i++) {
        for (int l = 0; l < REF_LANES; l++) u1[l] = ref_uniform_f32(&lanes[l]);   /* first_rand  (mc_simd.rs:12) */
        for (int l = 0; l < REF_LANES; l++) u2[l] = ref_uniform_f32(&lanes[l]);   /* second_rand (mc_simd.rs:13) */
        for (int l = 0; l < REF_LANES; l++) acc[l] = ref_pair_increment(two_pi, acc[l], u1[l], u2[l]);
    }
    memcpy(out_m, acc, sizeof(acc));
}

/* One evaluated scenario of a simulation core: payoff = max(sspot * exp(M*diff + drift) - sstrike, 0). */
typedef struct {
    float sspot, sstrike;   /* call_mult * spot, call_mult * strike */
    float diff;             /* sqrt(2) * sidt (sign-flipped for the antithetic leg) */
    float drift;            /* steps * nudt */
} ref_leg_t;

#define REF_MAX_LEGS 3

/* Generic simulation core.  For each chunk (in order) and each leg, accumulates the per-lane f32 payoff vector
 * exactly like the reference's map/reduce, plus fp64 first/second moments for standard errors.
 *   totals_f32[leg]   : horizontal sum of the f32 lane accumulators (what reduce_add() returns)
 *   sum64/sumsq64[leg]: fp64 sum and sum of squares of the per-path payoffs
 *   terminal_out      : optional, n_chunks*8 values of exp(M*diff0 + drift0) (leg 0's S_T / spot)
 *   m_out             : optional, n_chunks*8 raw sums M                                                      */
static void ref_mc_run(const uint8_t *chunk_seeds, int n_chunks, int half_steps, int n_legs, const ref_leg_t *legs,
                       float *totals_f32, double *sum64, double *sumsq64, float *terminal_out, float *m_out) {
    float lane_tot[REF_MAX_LEGS][REF_LANES];
    memset(lane_tot, 0, sizeof(lane_tot));
    for (int g = 0; g < n_legs; g++) { sum64[g] = 0.0; sumsq64[g] = 0.0; }
    for (int c = 0; c < n_chunks; c++) {
        float m[REF_LANES];
        ref_mc_chunk_sums(chunk_seeds + (size_t)c * REF_SEED_BYTES, half_steps, m);
        for (int g = 0; g < n_legs; g++) {
            for (int l = 0; l < REF_LANES; l++) {
                const float growth = expf(fmaf(m[l], legs[g].diff, legs[g].drift));
                float pay = fmaf(legs[g].sspot, growth, -legs[g].sstrike);      /* mul_sub */
                pay = pay > 0.0f ? pay : 0.0f;                                   /* fast_max(., 0) */
                lane_tot[g][l] += pay;
                sum64[g] += (double)pay;
                sumsq64[g] += (double)pay * (double)pay;
                if (g == 0 && terminal_out) terminal_out[(size_t)c * REF_LANES + l] = growth;
            }
        }
        if (m_out) memcpy(m_out + (size_t)c * REF_LANES, m, sizeof(m));
    }
    for (int g = 0; g < n_legs; g++) {
        float t = 0.0f;
        for (int l = 0; l < REF_LANES; l++) t += lane_tot[g][l];
        totals_f32[g] = t;
    }
}

/* scalar precompute shared by every core (mc_simd.rs:36-38 and the bumped variants) */
static inline void ref_constants(float vol, float r, float q, float dt, float steps, float *drift, float *diff) {
    const float nudt = (r - q - 0.5f * (vol * vol)) * dt;
    const float sidt = vol * sqrtf(dt);
    *drift = steps * nudt;
    *diff = 1.41421356237309504880f * sidt;
}

static inline int ref_half_steps(float steps) { return ((int)steps) / 2; }
static inline int ref_chunks(float num_trials) { return ((int)num_trials) / 8; }

/* ---------------- simulation cores ---------------- */

/* mc_simd.rs:25-79.  Returns the f32 price; optional fp64 mean/SE of the discounted payoff estimator
 * (same divide-by-num_trials convention), optional per-path growth factors S_T/spot and raw sums M. */
float ref_mc_pricing(float spot, float strike, float vol, float r, float years, float q, float steps,
                     float num_trials, float call_mult, const uint8_t *chunk_seeds,
                     double *mean64, double *se64, float *terminal_out, float *m_out) {
    ref_leg_t leg;
    ref_constants(vol, r, q, years / steps, steps, &leg.drift, &leg.diff);
    leg.sspot = call_mult * spot;
    leg.sstrike = call_mult * strike;
    float tot; double s, ss;
    const int n_chunks = ref_chunks(num_trials);
    ref_mc_run(chunk_seeds, n_chunks, ref_half_steps(steps), 1, &leg, &tot, &s, &ss, terminal_out, m_out);
    const float disc = expf(-r * years);
    if (mean64 || se64) {
        const double n = 8.0 * n_chunks, scale = (double)disc * n / (double)num_trials;
        const double mu = s / n, var = n > 1 ? (ss - n * mu * mu) / (n - 1.0) : 0.0;
        if (mean64) *mean64 = mu * scale;
        if (se64) *se64 = sqrt((var > 0 ? var : 0) / n) * scale;
    }
    return (tot / num_trials) * disc;
}

/* mc_simd.rs:82-149 (antithetic).  The SE is that of the per-path estimator 0.5*(pay+ + pay-). */
float ref_mc_av_pricing(float spot, float strike, float vol, float r, float years, float q, float steps,
                        float num_trials, float call_mult, const uint8_t *chunk_seeds) {
    ref_leg_t legs[2];
    ref_constants(vol, r, q, years / steps, steps, &legs[0].drift, &legs[0].diff);
    legs[0].sspot = call_mult * spot;
    legs[0].sstrike = call_mult * strike;
    legs[1] = legs[0];
    legs[1].diff = -legs[0].diff;                      /* sqrt(2) * -sidt (mc_simd.rs:103) */
    float tot[2]; double s[2], ss[2];
    ref_mc_run(chunk_seeds, ref_chunks(num_trials), ref_half_steps(steps), 2, legs, tot, s, ss, NULL, NULL);
    return ((0.5f * (tot[0] + tot[1])) / num_trials) * expf(-r * years);
}

/* mc_simd.rs:152-227: (P(S-h), P(S), P(S+h)) from one growth factor per path */
void ref_mc_spot_pricing(float spot, float dspot, float strike, float vol, float r, float years, float q,
                         float steps, float num_trials, float call_mult, const uint8_t *chunk_seeds,
                         float out[3]) {
    ref_leg_t legs[3];
    ref_constants(vol, r, q, years / steps, steps, &legs[0].drift, &legs[0].diff);
    legs[0].sstrike = strike * call_mult;
    legs[1] = legs[0]; legs[2] = legs[0];
    legs[0].sspot = call_mult * spot;
    legs[1].sspot = call_mult * (spot + dspot);
    legs[2].sspot = call_mult * (spot - dspot);
    float tot[3]; double s[3], ss[3];
    ref_mc_run(chunk_seeds, ref_chunks(num_trials), ref_half_steps(steps), 3, legs, tot, s, ss, NULL, NULL);
    const float final_mult = expf(-r * years) / num_trials;
    out[0] = tot[2] * final_mult;   /* minus */
    out[1] = tot[0] * final_mult;   /* base  */
    out[2] = tot[1] * final_mult;   /* plus  */
}

/* mc_simd.rs:230-309: (P(vol-h), P(vol+h)), call only */
void ref_mc_volatility_pricing(float spot, float strike, float vol, float dvol, float r, float years, float q,
                               float steps, float num_trials, const uint8_t *chunk_seeds, float out[2]) {
    const float dt = years / steps;
    ref_leg_t legs[2];
    ref_constants(vol + dvol, r, q, dt, steps, &legs[0].drift, &legs[0].diff);
    ref_constants(vol - dvol, r, q, dt, steps, &legs[1].drift, &legs[1].diff);
    legs[0].sspot = legs[1].sspot = spot;
    legs[0].sstrike = legs[1].sstrike = strike;
    float tot[2]; double s[2], ss[2];
    ref_mc_run(chunk_seeds, ref_chunks(num_trials), ref_half_steps(steps), 2, legs, tot, s, ss, NULL, NULL);
    const float final_mult = expf(-r * years) / num_trials;
    out[0] = tot[1] * final_mult;
    out[1] = tot[0] * final_mult;
}

/* mc_simd.rs:312-387: (P(r-h), P(r+h)); drift and discount use the bumped rate */
void ref_mc_interest_pricing(float spot, float strike, float vol, float r, float dr, float years, float q,
                             float steps, float num_trials, float call_mult, const uint8_t *chunk_seeds,
                             float out[2]) {
    const float dt = years / steps;
    const float r_plus = r + dr, r_minus = r - dr;
    ref_leg_t legs[2];
    ref_constants(vol, r_plus, q, dt, steps, &legs[0].drift, &legs[0].diff);
    ref_constants(vol, r_minus, q, dt, steps, &legs[1].drift, &legs[1].diff);
    legs[0].sspot = legs[1].sspot = spot * call_mult;
    legs[0].sstrike = legs[1].sstrike = strike * call_mult;
    float tot[2]; double s[2], ss[2];
    ref_mc_run(chunk_seeds, ref_chunks(num_trials), ref_half_steps(steps), 2, legs, tot, s, ss, NULL, NULL);
    out[0] = (tot[1] * expf(-r_minus * years)) / num_trials;
    out[1] = (tot[0] * expf(-r_plus * years)) / num_trials;
}

/* mc_simd.rs:390-473: (P(T-h), P(T+h)); dt, drift, diffusion and discount all use the bumped expiry */
void ref_mc_time_pricing(float spot, float strike, float vol, float r, float years, float dyears, float q,
                         float steps, float num_trials, float call_mult, const uint8_t *chunk_seeds,
                         float out[2]) {
    const float t_plus = years + dyears, t_minus = years - dyears;
    ref_leg_t legs[2];
    ref_constants(vol, r, q, t_plus / steps, steps, &legs[0].drift, &legs[0].diff);
    ref_constants(vol, r, q, t_minus / steps, steps, &legs[1].drift, &legs[1].diff);
    legs[0].sspot = legs[1].sspot = spot * call_mult;
    legs[0].sstrike = legs[1].sstrike = strike * call_mult;
    float tot[2]; double s[2], ss[2];
    ref_mc_run(chunk_seeds, ref_chunks(num_trials), ref_half_steps(steps), 2, legs, tot, s, ss, NULL, NULL);
    out[0] = (tot[1] * expf(-r * t_minus)) / num_trials;
    out[1] = (tot[0] * expf(-r * t_plus)) / num_trials;
}

/* ---------------- the 12 public entry points (mc_simd.rs:477-779), argument order as in the reference ------------- */

float ref_mc_call_price(float spot, float strike, float vol, float r, float years, float q, float steps,
                        float num_trials, const uint8_t *seeds) {
    return ref_mc_pricing(spot, strike, vol, r, years, q, steps, num_trials, 1.0f, seeds, NULL, NULL, NULL, NULL);
}
float ref_mc_put_price(float spot, float strike, float vol, float r, float years, float q, float steps,
                       float num_trials, const uint8_t *seeds) {
    return ref_mc_pricing(spot, strike, vol, r, years, q, steps, num_trials, -1.0f, seeds, NULL, NULL, NULL, NULL);
}
float ref_mc_call_price_av(float spot, float strike, float vol, float r, float years, float q, float steps,
                           float num_trials, const uint8_t *seeds) {
    return ref_mc_av_pricing(spot, strike, vol, r, years, q, steps, num_trials, 1.0f, seeds);
}
float ref_mc_put_price_av(float spot, float strike, float vol, float r, float years, float q, float steps,
                          float num_trials, const uint8_t *seeds) {
    return ref_mc_av_pricing(spot, strike, vol, r, years, q, steps, num_trials, -1.0f, seeds);
}
float ref_mc_call_delta(float spot, float dspot, float strike, float vol, float r, float years, float q,
                        float steps, float num_trials, const uint8_t *seeds) {
    float p[3];
    ref_mc_spot_pricing(spot, dspot, strike, vol, r, years, q, steps, num_trials, 1.0f, seeds, p);
    return (p[2] - p[0]) / (2.0f * dspot);
}
float ref_mc_put_delta(float spot, float dspot, float strike, float vol, float r, float years, float q,
                       float steps, float num_trials, const uint8_t *seeds) {
    float p[3];
    ref_mc_spot_pricing(spot, dspot, strike, vol, r, years, q, steps, num_trials, -1.0f, seeds, p);
    return (p[2] - p[0]) / (2.0f * dspot);
}
float ref_mc_gamma(float spot, float dspot, float strike, float vol, float r, float years, float q, float steps,
                   float num_trials, const uint8_t *seeds) {
    float p[3];
    ref_mc_spot_pricing(spot, dspot, strike, vol, r, years, q, steps, num_trials, 1.0f, seeds, p);
    return (p[2] - 2.0f * p[1] + p[0]) / (dspot * dspot);
}
float ref_mc_vega(float spot, float strike, float vol, float dvol, float r, float years, float q, float steps,
                  float num_trials, const uint8_t *seeds) {
    float p[2];
    ref_mc_volatility_pricing(spot, strike, vol, dvol, r, years, q, steps, num_trials, seeds, p);
    return (p[1] - p[0]) / (200.0f * dvol);           /* per 1 vol point */
}
float ref_mc_call_rho(float spot, float strike, float vol, float r, float dr, float years, float q, float steps,
                      float num_trials, const uint8_t *seeds) {
    float p[2];
    ref_mc_interest_pricing(spot, strike, vol, r, dr, years, q, steps, num_trials, 1.0f, seeds, p);
    return (p[1] - p[0]) / (200.0f * dr);             /* per 1 rate point */
}
float ref_mc_put_rho(float spot, float strike, float vol, float r, float dr, float years, float q, float steps,
                     float num_trials, const uint8_t *seeds) {
    float p[2];
    ref_mc_interest_pricing(spot, strike, vol, r, dr, years, q, steps, num_trials, -1.0f, seeds, p);
    return (p[1] - p[0]) / (200.0f * dr);
}
float ref_mc_call_theta(float spot, float strike, float vol, float r, float years, float dyears, float q,
                        float steps, float num_trials, const uint8_t *seeds) {
    float p[2];
    ref_mc_time_pricing(spot, strike, vol, r, years, dyears, q, steps, num_trials, 1.0f, seeds, p);
    return (p[0] - p[1]) / (2.0f * dyears);           /* per year; shorter expiry first */
}
float ref_mc_put_theta(float spot, float strike, float vol, float r, float years, float dyears, float q,
                       float steps, float num_trials, const uint8_t *seeds) {
    float p[2];
    ref_mc_time_pricing(spot, strike, vol, r, years, dyears, q, steps, num_trials, -1.0f, seeds, p);
    return (p[0] - p[1]) / (2.0f * dyears);
}

/* ---------------- helpers for tests ---------------- */

/* Draw n uniforms from lane `lane` of a chunk seed (rand32x8.rs:23-38 uses lane 0). */
void ref_uniform_lane_stream(const uint8_t *seed256, int lane, int n, float *out) {
    xo256_t lanes[REF_LANES];
    ref_seed_chunk(seed256, lanes);
    for (int i = 0; i < n; i++)
        for (int l = 0; l < REF_LANES; l++) {
            const float u = ref_uniform_f32(&lanes[l]);
            if (l == lane) out[i] = u;
        }
}

/* raw generator access for the known-answer and jump tests */
void ref_xoshiro_outputs(uint64_t state[4], int n, uint64_t *out) {
    xo256_t g; memcpy(g.s, state, 32);
    for (int i = 0; i < n; i++) out[i] = xo256pp_next(&g);
    memcpy(state, g.s, 32);
}
void ref_xoshiro_apply_poly(uint64_t state[4], const uint64_t poly[4]) {
    xo256_t g; memcpy(g.s, state, 32);
    xo256_apply_poly(&g, poly);
    memcpy(state, g.s, 32);
}
uint64_t ref_splitmix64(uint64_t *x) { return splitmix64_next(x); }
