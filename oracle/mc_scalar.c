/* oracle/mc_scalar.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * Restatement of the reference's scalar comparison pricer /root/reference/src/mc.rs:5-73:
 * one standard normal and one exp per time step, multiplicative path, single thread, f32 arithmetic:
 *     dt = T/steps; nudt = (r - q - 0.5 v^2) dt; sidt = v sqrt(dt)                       (mc.rs:15-17)
 *     for each of (num_trials as i32) trials: mult = prod over (steps as i32) steps of exp(nudt + sidt z)   (mc.rs:20-26)
 *     payoff = max(+-(spot * mult - strike), 0); price = sum / num_trials * exp(-r T)    (mc.rs:28-37)
 * The reference draws z from rand_distr::StandardNormal (ziggurat) on an OS-seeded thread_rng; here z comes from a
 * seedable xoshiro256++ through the polar Box-Muller method in fp64, rounded to f32 -- a different generator AND a
 * different path construction from the mc_simd path, which is what makes it an independent cross-check.
 * Unlike mc_simd it uses every one of the `steps` normals (odd step counts are not under-dispersed) and ALL
 * (num_trials as i32) trials (no rounding down to a multiple of 8).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "xoshiro256pp.h"

static inline double u01(xo256_t *g) { return (double)(xo256pp_next(g) >> 11) * 1.1102230246251565e-16; }

static inline float std_normal(xo256_t *g, double *spare, int *has_spare) {
    if (*has_spare) { *has_spare = 0; return (float)*spare; }
    double a, b, s;
    do { a = 2.0 * u01(g) - 1.0; b = 2.0 * u01(g) - 1.0; s = a * a + b * b; } while (s >= 1.0 || s == 0.0);
    const double f = sqrt(-2.0 * log(s) / s);
    *spare = b * f; *has_spare = 1;
    return (float)(a * f);
}

/* call_mult = +1 for mc::call_price, -1 for mc::put_price.  mean_out / se_out (nullable): fp64 mean and standard error
 * of the discounted payoff over the simulated trials. */
float mc_scalar_price(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                      float dividend_yield, float steps, float num_trials, float call_mult, uint64_t seed,
                      double *mean_out, double *se_out) {
    const float dt = years_to_expiry / steps;
    const float nudt = (risk_free_rate - dividend_yield - 0.5f * (volatility * volatility)) * dt;
    const float sidt = volatility * sqrtf(dt);
    xo256_t g; uint64_t sm = seed;
    for (int k = 0; k < 4; k++) g.s[k] = splitmix64_next(&sm);
    double spare = 0.0; int has_spare = 0;
    float total = 0.0f;
    double sum = 0.0, sumsq = 0.0;
    const int n = (int)num_trials, m = (int)steps;
    for (int t = 0; t < n; t++) {
        float mult = 1.0f;
        for (int i = 0; i < m; i++) mult *= expf(nudt + sidt * std_normal(&g, &spare, &has_spare));
        const float price = call_mult * (spot * mult - strike);
        if (price > 0.0f) { total += price; sum += price; sumsq += (double)price * price; }
    }
    const float disc = expf(-risk_free_rate * years_to_expiry);
    if (mean_out) *mean_out = n > 0 ? sum / n * disc : 0.0;
    if (se_out) {
        const double mu = n > 0 ? sum / n : 0.0;
        const double var = n > 1 ? (sumsq - n * mu * mu) / (n - 1) : 0.0;
        *se_out = n > 0 ? sqrt((var > 0 ? var : 0) / n) * disc : 0.0;
    }
    return (total / num_trials) * disc;
}
