// mcb200_pair.cuh -- the inner step: one Box-Muller pair = two GBM time steps (sm_100a MUFU + FP32 + integer pipes).
//
// Replaces speed_update (/root/reference/src/mc_simd.rs:9-21) and the uniform adapter it calls
// (/root/reference/src/rand32x8.rs:5-21).  Shared by the path kernels and the pipe microbenchmarks.
//   tri_fast()   production: two generator outputs -> three pairs (the path loop of the kernels)
//   pair_fast()  one generator output -> one pair (the 0..2 pairs left over when floor(steps/2) % 3 != 0)
//   quad_fast()  round-1 production step (three outputs -> four pairs), kept for the microbenchmark comparison
//   pair_ref()   the reference's own consumption of the stream (parity mode)
#pragma once
#include <cstdint>
#include "mcb200_rng.cuh"

namespace mcb200 {

constexpr int kThreads = 256;          // threads per CTA (8 warps, 2 per SM sub-partition)
constexpr int kWarps = kThreads / 32;
// Resident CTAs per SM and tri_fast calls (3 pairs each) per loop iteration the path kernels are compiled for.  The
// inner step is bound by the ALU pipe, not by latency (4.3 pairs/clk/SM at 8 or 4 CTAs/SM alike, 4.1 at 2:
// profiles/r2_microbench_tri_vs_quad.json), so low occupancy costs nothing and buys register-resident accumulators.
// Measured with tools/ab_flags.sh (profiles/r2_ab_kernel_variants.txt), roofline.frac on C2 / C3 / C5R / C4:
//   price / antithetic kernels: 4 CTAs/SM (64 registers); 2 -> 0.856 / 0.865 / 1.022, 4 -> 0.877 / 0.890 / 1.057,
//                               8 -> 0.895 / 0.912 / 1.080, 16 tris per iteration -> 0.913 / 0.930 / 1.082
//                               (5 CTAs/SM: 0.835 / 0.690 / 1.009)
//   Greeks kernel (20 moment registers): 3 CTAs/SM; 4 tris per iteration -> C4 1.024, 6 -> 1.043, 8 -> 1.035
// The MCB_* macros exist for that sweep only (MCB200_NVCC_FLAGS in build.py).
#ifndef MCB_PRICE_CTAS
#define MCB_PRICE_CTAS 4
#endif
#ifndef MCB_PRICE_UNROLL
#define MCB_PRICE_UNROLL 16
#endif
#ifndef MCB_GREEKS_CTAS
#define MCB_GREEKS_CTAS 3
#endif
#ifndef MCB_GREEKS_UNROLL
#define MCB_GREEKS_UNROLL 6
#endif
constexpr int kPriceCtasPerSm = MCB_PRICE_CTAS, kPriceQuadUnroll = MCB_PRICE_UNROLL;
constexpr int kGreeksCtasPerSm = MCB_GREEKS_CTAS, kGreeksQuadUnroll = MCB_GREEKS_UNROLL;
constexpr int kMaxCtasPerSm = kPriceCtasPerSm > kGreeksCtasPerSm ? kPriceCtasPerSm : kGreeksCtasPerSm;
constexpr int kInlineMaxOptions = 128;      // host batches up to this size travel in the launch parameters (3 KB)
constexpr int kMaxRoundsPerFlush = 4096;   // paths a thread sums in fp32 before its moments go to fp64

// ------------------------------------------------------------------------------------------------ math helpers
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kQuarterPi = 0.78539816339744830962f;
constexpr float kLn2 = 0.69314718055994530942f;
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kSqrt2 = 1.41421356237309504880f;
constexpr float kSqrt2Ln2 = 1.17741002251547469101f;   // sqrt(2 ln 2)

// One-output pair update: ONE generator output -> one Box-Muller pair -> two time steps.
//   sqrt(-ln u1) * (sin t + cos t) == sqrt(2 ln 2) * sqrt(-log2 u1) * sin(t + pi/4);  the constant is folded into the
//   terminal scale, so the loop body is 3 MUFU (LG2, SQRT, SIN) + 4 FP32 + 2 bit-casts + the generator.
// FMODE selects how the two [1,2) floats are built from the generator output (bit-identical results):
//   0: LEA.HI  ((x >> 9) | 0x3f800000) on the ALU pipe
//   1: IMAD.HI ((x * 2^23) >> 32 + 0x3f800000) on the FMA-heavy pipe at half rate
//   2: IMAD.WIDE (x * 2^23 + (0x3f800000 << 32), upper word) on the FMA-heavy pipe at full rate
//   3: IMAD.WIDE upper word k = x >> 9 read as an fp32 DENORMAL (value k * 2^-149) and rescaled by the FMA pipe, which
//      handles denormal inputs at full rate: u1 = fma(d, -2^126, 1) = 1 - k 2^-23 (the same value as 2 - f1, exactly)
//      and angle = fma(d, 2 pi 2^126, pi/4) -- the same angle minus one full turn.  No ALU-pipe instruction at all.
//   4: I2FP.F32.U32.RZ of the whole 32-bit word (24 significant bits, truncated): u1 = fma(f, -2^-32, 1) in
//      [2^-24, 1], angle = fma(f, 2 pi 2^-32, pi/4).  A different (finer) uniform grid than modes 0-3.
// For 1 and 2 the multiplier 2^23 must arrive at run time (two23), or ptxas folds it back into LEA.HI.
template <int V, int FMODE>
__device__ __forceinline__ float pair_fast_v(Xoshiro256& g, float acc, uint32_t two23 = 8388608u, uint32_t one = 1u) {
    uint32_t hi, lo;
    xoshiro_next_v<V>(g, hi, lo, one);
    uint32_t b1, b2;
    if (FMODE == 1) {
        asm("mad.hi.u32 %0, %1, %2, 1065353216;" : "=r"(b1) : "r"(hi), "r"(two23));
        asm("mad.hi.u32 %0, %1, %2, 1065353216;" : "=r"(b2) : "r"(lo), "r"(two23));
    } else if (FMODE == 2) {
        uint32_t dump;
        unpack64(mad_wide(hi, two23, 0x3f80000000000000ull), dump, b1);
        unpack64(mad_wide(lo, two23, 0x3f80000000000000ull), dump, b2);
    } else if (FMODE == 3) {
        uint32_t dump;
        unpack64(mad_wide(hi, two23, 0ull), dump, b1);
        unpack64(mad_wide(lo, two23, 0ull), dump, b2);
        const float u1 = fmaf(__uint_as_float(b1), -8.507059173023462e37f, 1.0f);                    // 2^126
        const float ang = fmaf(__uint_as_float(b2), 6.28318530717958647692f * 8.507059173023462e37f, kQuarterPi);
        return fmaf(mufu_sqrt(fabsf(mufu_lg2(u1))), mufu_sin(ang), acc);
    } else if (FMODE == 4) {
        const float u1 = fmaf(__uint2float_rz(hi), -2.3283064365386963e-10f, 1.0f);                  // 2^-32
        const float ang = fmaf(__uint2float_rz(lo), 6.28318530717958647692f * 2.3283064365386963e-10f, kQuarterPi);
        return fmaf(mufu_sqrt(fabsf(mufu_lg2(u1))), mufu_sin(ang), acc);
    } else {
        b1 = (hi >> 9) | 0x3f800000u;
        b2 = (lo >> 9) | 0x3f800000u;
    }
    const float f1 = __uint_as_float(b1);                           // [1,2)
    const float f2 = __uint_as_float(b2);                           // [1,2) == one full turn
    const float u1 = 2.0f - f1;                                     // (0,1], exact
    const float rad = mufu_sqrt(fabsf(mufu_lg2(u1)));
    const float ang = fmaf(f2, kTwoPi, kQuarterPi);
    return fmaf(rad, mufu_sin(ang), acc);
}

constexpr int kFloatMode = 0;  // production choice

__device__ __forceinline__ float pair_fast(Xoshiro256& g, float acc) {
    return pair_fast_v<kXoVariant, kFloatMode>(g, acc);
}

// Dense use of the generator: THREE 64-bit outputs -> FOUR Box-Muller pairs (eight time steps).
//   words w0..w5 = (hi, lo) of the three outputs.  Pairs A, B, C take the top 23 bits of (w0,w1), (w2,w3), (w4,w5) like
//   pair_fast.  Pair D is built from the six low BYTES, which pair_fast throws away: u1 from the low bytes of
//   w0, w2, w4 and the angle from those of w1, w3, w5 (two PRMT byte gathers + one LEA.HI each; bit 8 of every word
//   and one bit of each gathered field stay unused).  No bit is used twice, every field is 23 uniform bits, and the
//   generator -- 15 of the 17 ALU-pipe instructions of a pair, the binding resource -- runs 3/4 as often.
__device__ __forceinline__ float bm_term(uint32_t b1, uint32_t b2, float acc) {
    const float u1 = 2.0f - __uint_as_float(b1);                    // b1, b2: bit patterns of floats in [1,2)
    const float rad = mufu_sqrt(fabsf(mufu_lg2(u1)));
    const float ang = fmaf(__uint_as_float(b2), kTwoPi, kQuarterPi);
    return fmaf(rad, mufu_sin(ang), acc);
}
__device__ __forceinline__ uint32_t top23(uint32_t w) { return (w >> 9) | 0x3f800000u; }
__device__ __forceinline__ uint32_t low_bytes23(uint32_t wa, uint32_t wb, uint32_t wc) {
    uint32_t t, x;
    asm("prmt.b32 %0, %1, %2, 0x4400;" : "=r"(t) : "r"(wa), "r"(wb));       // t = [ . , wb.b0, wa.b0, . ]
    asm("prmt.b32 %0, %1, %2, 0x4210;" : "=r"(x) : "r"(t), "r"(wc));        // x = [wc.b0, wb.b0, wa.b0, . ]
    return top23(x);
}
__device__ __forceinline__ float quad_fast(Xoshiro256& g, float acc) {
    uint32_t w0, w1, w2, w3, w4, w5;
    xoshiro_next(g, w0, w1);
    xoshiro_next(g, w2, w3);
    xoshiro_next(g, w4, w5);
    acc = bm_term(top23(w0), top23(w1), acc);
    acc = bm_term(top23(w2), top23(w3), acc);
    acc = bm_term(top23(w4), top23(w5), acc);
    return bm_term(low_bytes23(w0, w2, w4), low_bytes23(w1, w3, w5), acc);
}

// Denser still: TWO 64-bit outputs -> THREE Box-Muller pairs (six time steps): 12.33 ALU-pipe instructions per pair
// instead of 14.25, and the ALU pipe is what binds the kernel.
//   words (w0, w1), (w2, w3) = (hi, lo) of the two outputs.  u1 of pairs A, B, C = top 23 bits of w0, w1, w2 (LEA.HI,
//   the same (0,1] grid as before).  The ANGLES are 16-bit fields placed by one byte permute each: A = top half of w3,
//   B = bottom half of w3, C = low bytes of w0 and w1 (one extra PRMT to gather them).  The permute takes the top byte
//   0x3f from the bit pattern of pair A's [1,2) float, so the result is the float  0 01111111b mmmmmmm mmmmmmmm 0..0 :
//   b = 1 -> 1 + m in [1,2), b = 0 -> (1 + m)/2 in [0.5,1), m on a 2^-15 grid.  With the angle scale 4 pi both
//   halves are uniform over whole turns (2 turns + 2 m turns, or 1 + m turns): an equispaced angle grid of 2^15
//   (b = 0) or 2^14 (b = 1) directions.  An equispaced grid of N directions reproduces every moment of the Gaussian
//   up to order N - 1 exactly (the discrete average of exp(i j theta) vanishes unless N | j), so the angle needs far
//   fewer bits than the radius; the radius keeps all 23.  No generator bit is used twice; 11 of 128 stay unused.
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(SEL));
    return r;
}
__device__ __forceinline__ float bm_term16(uint32_t b1, uint32_t ang_bits, float acc) {
    const float u1 = 2.0f - __uint_as_float(b1);                    // b1: bit pattern of a float in [1,2)
    const float rad = mufu_sqrt(fabsf(mufu_lg2(u1)));
    const float ang = fmaf(__uint_as_float(ang_bits), 2.0f * kTwoPi, kQuarterPi);   // ang_bits: float in [0.5,2)
    return fmaf(rad, mufu_sin(ang), acc);
}
// `one` is the value 1 delivered at run time (a kernel parameter): it keeps the carry halves of the generator's two
// 64-bit adds on IMAD.X (FMA-heavy pipe); left alone, ptxas moves a quarter of them to IADD3.X on the ALU pipe.
__device__ __forceinline__ float tri_fast(Xoshiro256& g, float acc, uint32_t one) {
    uint32_t w0, w1, w2, w3;
    xoshiro_next_v<XO_CARRY_MAD>(g, w0, w1, one);
    xoshiro_next_v<XO_CARRY_MAD>(g, w2, w3, one);
    const uint32_t fa = top23(w0), fb = top23(w1), fc = top23(w2);
    // prmt<SEL>(x, fa): result bytes, most significant first = [fa.b3 (0x3f), x.bJ, x.bK, 0x00 (sign fill of 0x3f)]
    const uint32_t aa = prmt<0x732F>(w3, fa);                        // w3.b3, w3.b2
    const uint32_t ab = prmt<0x710F>(w3, fa);                        // w3.b1, w3.b0
    const uint32_t t = prmt<0x0040>(w0, w1);                         // t.b1 = w1.b0, t.b0 = w0.b0
    const uint32_t ac = prmt<0x710F>(t, fa);                         // w1.b0, w0.b0
    acc = bm_term16(fa, aa, acc);
    acc = bm_term16(fb, ab, acc);
    return bm_term16(fc, ac, acc);
}

// Reference-stream pair update: consumes the generator exactly like rand32x8.rs:5-21 + mc_simd.rs:9-21
// (two 64-bit outputs, top 53 bits -> f64 -> f32) and evaluates the reference's formula with MUFU math.
__device__ __forceinline__ float u53_to_f32(uint32_t hi, uint32_t lo) {
    const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return __ull2float_rn(x) * 1.1102230246251565e-16f;             // 2^-53 (exact scaling)
}
__device__ __forceinline__ float pair_ref(Xoshiro256& g, float acc) {
    uint32_t h1, l1, h2, l2;
    xoshiro_next(g, h1, l1);
    xoshiro_next(g, h2, l2);
    const float u1 = u53_to_f32(h1, l1), u2 = u53_to_f32(h2, l2);
    const float ang = kTwoPi * u2;
    const float rad = mufu_sqrt(-(mufu_lg2(u1) * kLn2));
    return fmaf(rad, mufu_sin(ang) + mufu_cos(ang), acc);
}

}  // namespace mcb200
