/* oracle/fast_model.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * CPU model of the B200 production ("fast") stream convention, used to pin the CUDA kernel per path.
 * The reference's pair update (/root/reference/src/mc_simd.rs:9-21) adds sqrt(-ln u1) * (sin 2*pi*u2 + cos 2*pi*u2)
 * per two time steps, with u1,u2 taken from two 64-bit generator outputs per lane (src/rand32x8.rs:5-21).
 * The production kernel keeps that estimator but restates it for the GPU's MUFU/ALU pipes:
 *     sqrt(-ln u1) (sin t + cos t)  ==  sqrt(2 ln 2) * sqrt(-log2 u1) * sin(t + pi/4)         (exact identity)
 * and draws u1,u2 from ONE 64-bit output (high word -> u1 in (0,1], low word -> the angle), 23 bits each.
 * Pairs are taken four at a time from THREE outputs (quad_fast in mcb200_pair.cuh): pairs A, B, C use the top 23 bits
 * of the words of outputs 0, 1, 2; pair D uses the six low bytes those leave over (u1 from the low bytes of the three
 * high words, the angle from those of the three low words; the byte of the LAST output is the most significant).
 * The 0..3 pairs left when half_steps is not a multiple of 4 use one output each.
 * The constant sqrt(2 ln 2) is folded into the terminal scale.  This file states that convention on the CPU with
 * libm (log2f / sqrtf / sinf / exp2f); the GPU differs only by the MUFU approximations (tests give the tolerance).
 *
 * One generator stream serves many paths back to back (the per-thread stream of the kernel), instead of a fresh
 * 256-byte seed per 8 paths (mc_simd.rs:52-54); substreams come from jump()/long_jump() (xoshiro256pp.h).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "xoshiro256pp.h"

static inline float bits_to_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

/* one Box-Muller term from the bit patterns of two floats in [1,2): sqrt(|log2 u1|) * sin(2*pi*f2 + pi/4) + acc */
static inline float bm_term(uint32_t b1, uint32_t b2, float acc) {
    const float f1 = bits_to_float(b1), f2 = bits_to_float(b2);
    const float u1 = 2.0f - f1;                                  /* (0,1], exact */
    const float rad = sqrtf(fabsf(log2f(u1)));
    const float ang = fmaf(f2, 6.28318530717958647692f, 0.78539816339744830962f);
    return fmaf(rad, sinf(ang), acc);
}
static inline uint32_t top23(uint32_t w) { return (w >> 9) | 0x3f800000u; }
/* 23 bits from the low bytes of three words: wc's byte is the most significant, then wb's, then wa's top 7 bits */
static inline uint32_t low_bytes23(uint32_t wa, uint32_t wb, uint32_t wc) {
    const uint32_t x = ((wc & 0xffu) << 24) | ((wb & 0xffu) << 16) | ((wa & 0xffu) << 8);
    return top23(x);
}

/* one pair from one output (the left-over pairs of a path) */
static inline float fast_pair_increment(xo256_t *g, float acc) {
    const uint64_t x = xo256pp_next(g);
    return bm_term(top23((uint32_t)(x >> 32)), top23((uint32_t)x), acc);
}

/* four pairs from three outputs */
static inline float fast_quad_increment(xo256_t *g, float acc) {
    const uint64_t x0 = xo256pp_next(g), x1 = xo256pp_next(g), x2 = xo256pp_next(g);
    const uint32_t w0 = (uint32_t)(x0 >> 32), w1 = (uint32_t)x0, w2 = (uint32_t)(x1 >> 32), w3 = (uint32_t)x1;
    const uint32_t w4 = (uint32_t)(x2 >> 32), w5 = (uint32_t)x2;
    acc = bm_term(top23(w0), top23(w1), acc);
    acc = bm_term(top23(w2), top23(w3), acc);
    acc = bm_term(top23(w4), top23(w5), acc);
    return bm_term(low_bytes23(w0, w2, w4), low_bytes23(w1, w3, w5), acc);
}

/* The four Box-Muller terms of `n` consecutive quads, separately (out[4*i + j], j = 0..3 for pairs A, B, C, D):
 * lets the tests check that pair D -- built from the low bytes pairs A-C leave over -- is a standard Box-Muller
 * variate of the same law and uncorrelated with the others. */
void fast_model_quad_terms(uint64_t state[4], int n, float *out) {
    xo256_t g; memcpy(g.s, state, 32);
    for (int i = 0; i < n; i++) {
        const uint64_t x0 = xo256pp_next(&g), x1 = xo256pp_next(&g), x2 = xo256pp_next(&g);
        const uint32_t w0 = (uint32_t)(x0 >> 32), w1 = (uint32_t)x0, w2 = (uint32_t)(x1 >> 32), w3 = (uint32_t)x1;
        const uint32_t w4 = (uint32_t)(x2 >> 32), w5 = (uint32_t)x2;
        out[4 * i + 0] = bm_term(top23(w0), top23(w1), 0.0f);
        out[4 * i + 1] = bm_term(top23(w2), top23(w3), 0.0f);
        out[4 * i + 2] = bm_term(top23(w4), top23(w5), 0.0f);
        out[4 * i + 3] = bm_term(low_bytes23(w0, w2, w4), low_bytes23(w1, w3, w5), 0.0f);
    }
    memcpy(state, g.s, 32);
}

/* Run `n_paths` consecutive paths of `half_steps` pairs each from one stream.  state is advanced in place.
 * out_m[p] = accumulated sum M' of path p (units: sqrt(2 ln 2) smaller than the reference's M). */
void fast_model_stream_paths(uint64_t state[4], int half_steps, int n_paths, float *out_m) {
    xo256_t g; memcpy(g.s, state, 32);
    for (int p = 0; p < n_paths; p++) {
        float acc = 0.0f;
        const int quads = half_steps / 4;
        for (int i = 0; i < quads; i++) acc = fast_quad_increment(&g, acc);
        for (int i = 4 * quads; i < half_steps; i++) acc = fast_pair_increment(&g, acc);
        out_m[p] = acc;
    }
    memcpy(state, g.s, 32);
}

/* Terminal-scale constants of the fast convention from the reference's f32 constants (mc_simd.rs:36-45):
 *   growth = exp(M * sqrt(2)*sidt + steps*nudt) = exp2(M' * c2 + d2),
 *   c2 = sqrt(2)*sidt * (sqrt(2 ln 2) * log2 e),  d2 = steps*nudt * log2 e, all in f32 like the kernel prologue. */
void fast_model_constants(float vol, float r, float q, float years, float steps, float *c2, float *d2) {
    const float dt = years / steps;
    const float nudt = (r - q - 0.5f * (vol * vol)) * dt;
    const float sidt = vol * sqrtf(dt);
    const float drift = steps * nudt;
    const float diff = 1.41421356237309504880f * sidt;
    const float k = 1.17741002251547469101f /* sqrt(2 ln 2) */ * 1.44269504088896340736f /* log2 e */;
    *c2 = diff * k;
    *d2 = drift * 1.44269504088896340736f;
}

float fast_model_growth(float m, float c2, float d2) { return exp2f(fmaf(m, c2, d2)); }
