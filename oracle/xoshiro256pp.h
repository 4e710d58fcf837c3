/* oracle/xoshiro256pp.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * Scalar restatement of the xoshiro256++ generator the reference drives through the third-party crate
 * simd_rand (git ashayp22/simd-rand @ b8c059782bca4fb0ca35aa7ecc565d32fad58e40, pinned in the reference's
 * Cargo.lock:716-718; call sites src/mc_simd.rs:52-54 and src/rand32x8.rs:5-6).  The crate source is not
 * under /root/reference, so the recurrence below is the published public-domain algorithm (Blackman & Vigna,
 * "xoshiro256++ 1.0") of which Xoshiro256PlusPlusX8 is eight interleaved lanes; it is pinned here by the
 * known-answer vector for state {1,2,3,4} (tests/golden/xoshiro256pp_kat.json) and by the jump()/long_jump()
 * consistency tests.  PARITY UNPINNED at the simd_rand boundary for (a) the byte layout of the 256-byte seed
 * and (b) the u64 -> f64 conversion of next_f64x8: both are stated conventions (see ref_mc.c).
 */
#ifndef ORACLE_XOSHIRO256PP_H
#define ORACLE_XOSHIRO256PP_H
#include <stdint.h>

typedef struct { uint64_t s[4]; } xo256_t;

static inline uint64_t xo_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* one generator step: returns the ++-scrambled output, advances the linear engine */
static inline uint64_t xo256pp_next(xo256_t *g) {
    uint64_t *s = g->s;
    const uint64_t result = xo_rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = xo_rotl(s[3], 45);
    return result;
}

/* splitmix64: the conventional seeder for xoshiro state words */
static inline uint64_t splitmix64_next(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* Apply a jump polynomial (4 x 64 coefficient bits, LSB first): state := sum_{b: poly_b = 1} T^b state.
 * With poly = JUMP this is Vigna's jump() (2^128 steps); with LONG_JUMP it is long_jump() (2^192 steps). */
static inline void xo256_apply_poly(xo256_t *g, const uint64_t poly[4]) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int w = 0; w < 4; w++) {
        for (int b = 0; b < 64; b++) {
            if (poly[w] & (1ULL << b)) {
                a0 ^= g->s[0]; a1 ^= g->s[1]; a2 ^= g->s[2]; a3 ^= g->s[3];
            }
            (void)xo256pp_next(g);
        }
    }
    g->s[0] = a0; g->s[1] = a1; g->s[2] = a2; g->s[3] = a3;
}

#endif
