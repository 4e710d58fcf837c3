// mcb200_pair.cuh -- the inner step: one Box-Muller pair = two GBM time steps (sm_100a MUFU + FP32 + integer pipes).
//
// Replaces speed_update (/root/reference/src/mc_simd.rs:9-21) and the uniform adapter it calls
// (/root/reference/src/rand32x8.rs:5-21).  Shared by the path kernels and the pipe microbenchmarks.
#pragma once
#include <cstdint>
#include "mcb200_rng.cuh"

namespace mcb200 {

constexpr int kThreads = 256;          // threads per CTA (8 warps, 2 per SM sub-partition)
constexpr int kWarps = kThreads / 32;

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

// Production pair update: ONE generator output -> one Box-Muller pair -> two time steps.
//   sqrt(-ln u1) * (sin t + cos t) == sqrt(2 ln 2) * sqrt(-log2 u1) * sin(t + pi/4);  the constant is folded into the
//   terminal scale, so the loop body is 3 MUFU (LG2, SQRT, SIN) + 4 FP32 + 2 bit-casts + the generator.
// FLOAT_MADHI builds the [1,2) floats with IMAD.HI ((x * 2^23) >> 32 + 0x3f800000) on the FMA-heavy pipe instead of
// LEA.HI on the ALU pipe; the multiplier 2^23 must arrive at run time (two23), or ptxas folds it back into LEA.HI.
template <int V, bool FLOAT_MADHI>
__device__ __forceinline__ float pair_fast_v(Xoshiro256& g, float acc, uint32_t two23 = 8388608u) {
    uint32_t hi, lo;
    xoshiro_next_v<V>(g, hi, lo);
    uint32_t b1, b2;
    if (FLOAT_MADHI) {
        asm("mad.hi.u32 %0, %1, %2, 1065353216;" : "=r"(b1) : "r"(hi), "r"(two23));
        asm("mad.hi.u32 %0, %1, %2, 1065353216;" : "=r"(b2) : "r"(lo), "r"(two23));
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

constexpr bool kFloatMadHi = false;  // production choice

__device__ __forceinline__ float pair_fast(Xoshiro256& g, float acc, uint32_t two23 = 8388608u) {
    return pair_fast_v<kXoVariant, kFloatMadHi>(g, acc, two23);
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
