// mcb200_rng.cuh -- per-thread xoshiro256++ for sm_100a, state held in eight 32-bit registers.
//
// Replaces the reference's 8-lane generator adapter (/root/reference/src/rand32x8.rs:5-21, simd_rand's
// Xoshiro256PlusPlusX8 seeded per 8-path chunk at src/mc_simd.rs:52-54) with one scalar generator per GPU thread.
// The 64-bit recurrence is written on 32-bit halves so that ptxas emits exactly: 2 carry adds (IADD3 + IMAD.X) x2,
// 5 funnel shifts (SHF), 1 IMAD.SHL and 8 three-input LOP3 per 64-bit output -- the integer pipes are the binding
// resource of the whole kernel ("issue budget").
#pragma once
#include <cstdint>

namespace mcb200 {

struct Xoshiro256 {
    uint32_t s0l, s0h, s1l, s1h, s2l, s2h, s3l, s3h;
};

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Pipe-balance variants of the generator step.  The SM has three relevant issue targets per sub-partition: the ALU pipe
// (LOP3 / SHF / IADD3 / LEA, 16 lanes), the FMA-heavy pipe (IMAD*, 16 lanes) and the XU pipe (MUFU, 4 lanes).  The
// straightforward step is ALU-bound (measured, profiles/microbench_*.json), so XO_T_WIDE moves the s1 << 17 shift onto
// IMAD.  (Moving the 64-bit adds to IMAD.WIDE was tried: ptxas splits them back into IADD3 + IMAD.X.)  All variants
// compute bit-identical results.
enum XoVariant : int {
    XO_T_WIDE = 1,     // t = s1 << 17 as IMAD.WIDE.U32 (s1.lo * 2^17) + IMAD (s1.hi * 2^17 + hi) instead of SHL + SHF
};

__device__ __forceinline__ void add64(uint32_t al, uint32_t ah, uint32_t bl, uint32_t bh, uint32_t& rl, uint32_t& rh) {
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(rl), "=r"(rh) : "r"(al), "r"(ah), "r"(bl), "r"(bh));
}

// Advance the linear engine one step and return the ++-scrambled 64-bit output as (hi, lo).
template <int V>
__device__ __forceinline__ void xoshiro_next_v(Xoshiro256& g, uint32_t& out_hi, uint32_t& out_lo) {
    // a = s0 + s3
    uint32_t al, ah;
    add64(g.s0l, g.s0h, g.s3l, g.s3h, al, ah);
    // r = rotl(a, 23) + s0
    const uint32_t rh = __funnelshift_l(al, ah, 23);
    const uint32_t rl = __funnelshift_l(ah, al, 23);
    add64(g.s0l, g.s0h, rl, rh, out_lo, out_hi);
    // t = s1 << 17
    uint32_t tl, th;
    if (V & XO_T_WIDE) {
        asm("{\n\t.reg .b64 w;\n\tmul.wide.u32 w, %2, 131072;\n\tmov.b64 {%0, %1}, w;\n\t}" : "=r"(tl), "=r"(th) : "r"(g.s1l));
        asm("mad.lo.u32 %0, %1, 131072, %0;" : "+r"(th) : "r"(g.s1h));
    } else {
        tl = g.s1l << 17;
        th = __funnelshift_l(g.s1l, g.s1h, 17);
    }
    // s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 45)   (fused into 3-input XORs)
    const uint32_t x3l = g.s3l ^ g.s1l, x3h = g.s3h ^ g.s1h;        // s3 ^ s1
    const uint32_t n1l = xor3(g.s1l, g.s2l, g.s0l), n1h = xor3(g.s1h, g.s2h, g.s0h);
    const uint32_t n2l = xor3(g.s2l, g.s0l, tl), n2h = xor3(g.s2h, g.s0h, th);
    g.s0l ^= x3l; g.s0h ^= x3h;
    g.s1l = n1l; g.s1h = n1h;
    g.s2l = n2l; g.s2h = n2h;
    // rotl(x3, 45) = rotl(swap halves, 13)
    g.s3h = __funnelshift_l(x3h, x3l, 13);
    g.s3l = __funnelshift_l(x3l, x3h, 13);
}

constexpr int kXoVariant = 0;      // production choice (see profiles/ for the measurements behind it)

__device__ __forceinline__ void xoshiro_next(Xoshiro256& g, uint32_t& out_hi, uint32_t& out_lo) {
    xoshiro_next_v<kXoVariant>(g, out_hi, out_lo);
}

__device__ __forceinline__ void xoshiro_load(Xoshiro256& g, const uint64_t* __restrict__ p) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(p + 2);
    g.s0l = (uint32_t)a.x; g.s0h = (uint32_t)(a.x >> 32);
    g.s1l = (uint32_t)a.y; g.s1h = (uint32_t)(a.y >> 32);
    g.s2l = (uint32_t)b.x; g.s2h = (uint32_t)(b.x >> 32);
    g.s3l = (uint32_t)b.y; g.s3h = (uint32_t)(b.y >> 32);
}

__device__ __forceinline__ void xoshiro_store(const Xoshiro256& g, uint64_t* __restrict__ p) {
    ulonglong2 a, b;
    a.x = ((uint64_t)g.s0h << 32) | g.s0l; a.y = ((uint64_t)g.s1h << 32) | g.s1l;
    b.x = ((uint64_t)g.s2h << 32) | g.s2l; b.y = ((uint64_t)g.s3h << 32) | g.s3l;
    *reinterpret_cast<ulonglong2*>(p) = a;
    *reinterpret_cast<ulonglong2*>(p + 2) = b;
}

// state := sum over set coefficient bits b of T^b(state)  -- Vigna's jump() procedure for an arbitrary polynomial.
__device__ inline void xoshiro_apply_poly(Xoshiro256& g, const uint64_t poly[4]) {
    uint32_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < 4; ++w) {
        const uint64_t word = poly[w];
        for (int b = 0; b < 64; ++b) {
            const uint32_t m = (uint32_t)0 - (uint32_t)((word >> b) & 1ull);
            a[0] ^= g.s0l & m; a[1] ^= g.s0h & m; a[2] ^= g.s1l & m; a[3] ^= g.s1h & m;
            a[4] ^= g.s2l & m; a[5] ^= g.s2h & m; a[6] ^= g.s3l & m; a[7] ^= g.s3h & m;
            uint32_t hi, lo;
            xoshiro_next(g, hi, lo);
        }
    }
    g.s0l = a[0]; g.s0h = a[1]; g.s1l = a[2]; g.s1h = a[3];
    g.s2l = a[4]; g.s2h = a[5]; g.s3l = a[6]; g.s3h = a[7];
}

}  // namespace mcb200
