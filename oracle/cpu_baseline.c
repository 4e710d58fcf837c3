/* oracle/cpu_baseline.c -- TEST/BENCH INFRASTRUCTURE (timed CPU baseline, kind "port"). Never on the product path.
 *
 * A SIMD + threads CPU restatement of the reference's mc_simd pricers, written for speed so the host-core baseline
 * bench.py reports is a fair stand-in for the reference binary (which cannot be built here: no Rust toolchain, nightly
 * feature gate src/lib.rs:1, unvendored git dependency Cargo.toml:13).  Same algorithm and work per path as
 * /root/reference/src/mc_simd.rs:
 *   - 8-lane chunks, a freshly seeded 8-lane xoshiro256++ per chunk (mc_simd.rs:52-54; seeding here is splitmix64
 *     of (seed, call, chunk) instead of 256 bytes of ChaCha12 -- cheaper, which favours the baseline),
 *   - two 53-bit uniforms -> f32 per lane per pair (rand32x8.rs:5-21), sqrt(-ln u1)*(sin+cos) FMA-accumulated
 *     (mc_simd.rs:9-21), one exp + payoff per path (mc_simd.rs:62-69), f32 lane sums, divide by num_trials, discount
 *     (mc_simd.rs:76-78),
 *   - chunks fanned out over all host threads (OpenMP standing in for rayon, mc_simd.rs:49-51,71-74),
 *   - options priced by a serial loop of calls, like benches/benchmark.rs:40-42.
 * The 8-wide inner loops are written so gcc vectorises them to AVX2 (the width of wide::f32x8) and maps logf/sinf/
 * cosf/expf to libmvec's 8-lane routines (-O3 -ffast-math -march=x86-64-v3), the analogue of wide's polynomial
 * transcendentals.  Results are checked against the oracle statistically in tests/ (not bit-exact: fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <omp.h>

#define L 8

static inline uint64_t sm64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

typedef struct { uint64_t s0[L], s1[L], s2[L], s3[L]; } xo8_t;

static inline void xo8_seed(xo8_t *g, uint64_t seed, uint64_t call, uint64_t chunk) {
    uint64_t x = seed ^ (call * 0xd1342543de82ef95ULL) ^ (chunk * 0x2545f4914f6cdd1dULL);
    for (int l = 0; l < L; l++) g->s0[l] = sm64(&x);
    for (int l = 0; l < L; l++) g->s1[l] = sm64(&x);
    for (int l = 0; l < L; l++) g->s2[l] = sm64(&x);
    for (int l = 0; l < L; l++) g->s3[l] = sm64(&x);
}

/* 8 uniforms in [0,1): the top 53 bits as f64 * 2^-53, narrowed to f32 */
static inline void xo8_uniform(xo8_t *g, float u[L]) {
#pragma omp simd
    for (int l = 0; l < L; l++) {
        const uint64_t a = g->s0[l] + g->s3[l];
        const uint64_t r = ((a << 23) | (a >> 41)) + g->s0[l];
        const uint64_t t = g->s1[l] << 17;
        g->s2[l] ^= g->s0[l];
        g->s3[l] ^= g->s1[l];
        g->s1[l] ^= g->s2[l];
        g->s0[l] ^= g->s3[l];
        g->s2[l] ^= t;
        g->s3[l] = (g->s3[l] << 45) | (g->s3[l] >> 19);
        /* exact u64(53 bit) -> f64 without AVX-512DQ: split into two exactly representable halves */
        const uint64_t m = r >> 11;
        const double hi = (double)(int64_t)(m >> 26), lo = (double)(int64_t)(m & 0x3ffffffULL);
        u[l] = (float)((hi * 67108864.0 + lo) * 0x1.0p-53);
    }
}

/* Accumulated payoff sums of one call: leg 0 uses +diff, leg 1 (antithetic) uses -diff when av != 0. */
static void price_call(float spot, float strike, float vol, float r, float years, float q, float steps,
                       float num_trials, float call_mult, int av, uint64_t seed, uint64_t call_id,
                       float *price_out) {
    const float dt = years / steps;
    const float nudt = (r - q - 0.5f * (vol * vol)) * dt;
    const float sidt = vol * sqrtf(dt);
    const float drift = steps * nudt, diff = 1.41421356237309504880f * sidt;
    const float two_pi = 6.28318530717958647692f;
    const float sspot = call_mult * spot, sstrike = call_mult * strike;
    const int half_steps = ((int)steps) / 2, n_chunks = ((int)num_trials) / 8;
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int c = 0; c < n_chunks; c++) {
        xo8_t g;
        xo8_seed(&g, seed, call_id, (uint64_t)c);
        float acc[L] = {0}, u1[L], u2[L];
        for (int i = 0; i < half_steps; i++) {
            xo8_uniform(&g, u1);
            xo8_uniform(&g, u2);
#pragma omp simd
            for (int l = 0; l < L; l++) {
                const float ang = two_pi * u2[l];
                acc[l] = sqrtf(-logf(u1[l])) * (sinf(ang) + cosf(ang)) + acc[l];
            }
        }
        float pay[L];
#pragma omp simd
        for (int l = 0; l < L; l++) {
            float p = sspot * expf(acc[l] * diff + drift) - sstrike;
            p = p > 0.0f ? p : 0.0f;
            if (av) {
                float p2 = sspot * expf(drift - acc[l] * diff) - sstrike;
                p = 0.5f * (p + (p2 > 0.0f ? p2 : 0.0f));
            }
            pay[l] = p;
        }
        float s = 0.0f;
        for (int l = 0; l < L; l++) s += pay[l];
        total += (double)s;
    }
    *price_out = ((float)total / num_trials) * expf(-r * years);
}

/* Price n options with a serial loop of calls (benches/benchmark.rs:40-42).  kind: 0 call, 1 put, 2 call_av, 3 put_av.
 * Returns the number of path-steps simulated: n * 8*floor(trials/8) * 2*floor(steps/2). */
double cpu_baseline_price_loop(int n, const float *spot, const float *strike, const float *vol, const float *r,
                               const float *years, const float *q, float steps, float num_trials, int kind,
                               uint64_t seed, float *out) {
    for (int i = 0; i < n; i++)
        price_call(spot[i], strike[i], vol[i], r[i], years[i], q[i], steps, num_trials,
                   (kind & 1) ? -1.0f : 1.0f, kind >= 2, seed, (uint64_t)i, &out[i]);
    return (double)n * (8.0 * (double)(((int)num_trials) / 8)) * (2.0 * (double)(((int)steps) / 2));
}

int cpu_baseline_threads(void) { return omp_get_max_threads(); }

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the bench sets the team size explicitly to all host cores. */
void cpu_baseline_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
