/* oracle/fast_model.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * CPU model of the B200 production ("fast") stream convention, used to pin the CUDA kernel per path.
 * The reference's pair update (/root/reference/src/mc_simd.rs:9-21) adds sqrt(-ln u1) * (sin 2*pi*u2 + cos 2*pi*u2)
 * per two time steps, with u1,u2 taken from two 64-bit generator outputs per lane (src/rand32x8.rs:5-21).
 * The production kernel keeps that estimator but restates it for the GPU's MUFU/ALU pipes:
 *     sqrt(-ln u1) (sin t + cos t)  ==  sqrt(2 ln 2) * sqrt(-log2 u1) * sin(t + pi/4)         (exact identity)
 * Pairs are taken three at a time from TWO outputs (tri_fast in mcb200_pair.cuh), words (w0, w1), (w2, w3) = (hi, lo):
 * u1 of pairs A, B, C = top 23 bits of w0, w1, w2 as a float in [1,2), u1 = 2 - f in (0,1]; the angles are 16-bit
 * fields -- A: top half of w3, B: bottom half of w3, C: low byte of w1 (high) and of w0 (low) -- read as the float with
 * bit pattern 0x3f000000 | field << 8, i.e. 1 + m in [1,2) or (1 + m)/2 in [0.5,1) by the field's top bit, times 4 pi:
 * whole turns either way, so the angle is uniform on an equispaced grid of 2^15 or 2^14 directions.
 * The 0..2 pairs left when half_steps is not a multiple of 3 use one output each (high word -> u1, low word -> the
 * angle, 23 bits each, angle scale 2 pi).
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

/* one Box-Muller term of a triple: the angle float lies in [0.5,2) and is scaled by 4 pi */
static inline float bm_term16(uint32_t b1, uint32_t ang_bits, float acc) {
    const float u1 = 2.0f - bits_to_float(b1);
    const float rad = sqrtf(fabsf(log2f(u1)));
    const float ang = fmaf(bits_to_float(ang_bits), 2.0f * 6.28318530717958647692f, 0.78539816339744830962f);
    return fmaf(rad, sinf(ang), acc);
}
static inline void tri_fields(const uint32_t w[4], uint32_t b1[3], uint32_t ang[3]) {
    b1[0] = top23(w[0]); b1[1] = top23(w[1]); b1[2] = top23(w[2]);
    ang[0] = 0x3f000000u | ((w[3] >> 16) << 8);
    ang[1] = 0x3f000000u | ((w[3] & 0xffffu) << 8);
    ang[2] = 0x3f000000u | ((w[1] & 0xffu) << 16) | ((w[0] & 0xffu) << 8);
}
/* three pairs from two outputs */
static inline float fast_tri_increment(xo256_t *g, float acc) {
    const uint64_t x0 = xo256pp_next(g), x1 = xo256pp_next(g);
    const uint32_t w[4] = {(uint32_t)(x0 >> 32), (uint32_t)x0, (uint32_t)(x1 >> 32), (uint32_t)x1};
    uint32_t b1[3], ang[3];
    tri_fields(w, b1, ang);
    for (int j = 0; j < 3; j++) acc = bm_term16(b1[j], ang[j], acc);
    return acc;
}

/* The three Box-Muller terms of `n` consecutive triples, separately (out[3*i + j], j = 0..2 for pairs A, B, C), and
 * (nullable) their angles in turns modulo 1: lets the tests check that every term is a standard Box-Muller variate
 * of the same law, uncorrelated with the others, on an equispaced angle grid. */
void fast_model_tri_terms(uint64_t state[4], int n, float *out, float *turns_out) {
    xo256_t g; memcpy(g.s, state, 32);
    for (int i = 0; i < n; i++) {
        const uint64_t x0 = xo256pp_next(&g), x1 = xo256pp_next(&g);
        const uint32_t w[4] = {(uint32_t)(x0 >> 32), (uint32_t)x0, (uint32_t)(x1 >> 32), (uint32_t)x1};
        uint32_t b1[3], ang[3];
        tri_fields(w, b1, ang);
        for (int j = 0; j < 3; j++) {
            out[3 * i + j] = bm_term16(b1[j], ang[j], 0.0f);
            if (turns_out) { const double t = 2.0 * (double)bits_to_float(ang[j]); turns_out[3 * i + j] = (float)(t - floor(t)); }
        }
    }
    memcpy(state, g.s, 32);
}

/* one pair from one output (the left-over pairs of a path) */
static inline float fast_pair_increment(xo256_t *g, float acc) {
    const uint64_t x = xo256pp_next(g);
    return bm_term(top23((uint32_t)(x >> 32)), top23((uint32_t)x), acc);
}

/* Run `n_paths` consecutive paths of `half_steps` pairs each from one stream.  state is advanced in place.
 * out_m[p] = accumulated sum M' of path p (units: sqrt(2 ln 2) smaller than the reference's M). */
void fast_model_stream_paths(uint64_t state[4], int half_steps, int n_paths, float *out_m) {
    xo256_t g; memcpy(g.s, state, 32);
    for (int p = 0; p < n_paths; p++) {
        float acc = 0.0f;
        const int tris = half_steps / 3;
        for (int i = 0; i < tris; i++) acc = fast_tri_increment(&g, acc);
        for (int i = 3 * tris; i < half_steps; i++) acc = fast_pair_increment(&g, acc);
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
