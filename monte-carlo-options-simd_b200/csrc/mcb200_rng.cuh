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

// Pipe-balance variants of the generator step.  Per SM sub-partition the relevant issue targets are the ALU pipe
// (LOP3 / SHF / IADD3 / LEA: 16 lanes, measured 64 thread-ops/clk/SM), the FMA-heavy pipe (IMAD*: also 64/clk/SM),
// the FMA-lite pipe (FFMA/FADD/FMUL share 128/clk/SM with the heavy one) and the XU pipe (MUFU, 16/clk/SM).  The
// straightforward step puts 15 of its 18 instructions on the ALU pipe, which then binds the whole path kernel
// (profiles/r1_microbench.json), so the variants below move 64-bit adds, shifts and rotates onto IMAD.WIDE / IMAD
// (FMA-heavy) without changing the instruction count.  All variants compute bit-identical results; `one` must be the
// value 1 delivered at run time (a kernel parameter), otherwise ptxas folds the multiply back into IADD3.
enum XoVariant : int {
    XO_T_WIDE = 1,      // t = s1 << 17          : IMAD.WIDE(s1.lo, 2^17) + IMAD(s1.hi, 2^17, hi)        (was IMAD.SHL + SHF)
    XO_ADD_WIDE = 2,    // a = s0 + s3           : IMAD.WIDE(s3.lo, one, s0) + IMAD(s3.hi, one, hi)       (was IADD3 + IMAD.X)
    XO_ROT_WIDE = 4,    // r = rotl(a, 23) + s0  : IMAD.WIDE(a.lo, 2^23, s0) + IMAD(a.hi, 2^23, hi) + SHF.R + IMAD.WIDE(a.hi >> 9, one, .)
                        //                                                                        (was 2 SHF + IADD3 + IMAD.X)
    XO_ROT45_WIDE = 8,  // s3 = rotl(s3 ^ s1, 45): IMAD.WIDE + IMAD + IMAD.HI                              (was 2 SHF)
    XO_CARRY_MAD = 16,  // high words of both 64-bit adds: madc.lo(x, one, y) = IMAD.X on the FMA-heavy pipe, always
                        //   (ptxas otherwise turns some of them into IADD3.X on the binding ALU pipe); likewise the low
                        //   word of t = s1 << 17 as IMAD(s1.lo, one << 17) (ptxas otherwise turns some into SHF)
};

__device__ __forceinline__ void add64(uint32_t al, uint32_t ah, uint32_t bl, uint32_t bh, uint32_t& rl, uint32_t& rh) {
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(rl), "=r"(rh) : "r"(al), "r"(ah), "r"(bl), "r"(bh));
}
__device__ __forceinline__ void add64_mad(uint32_t al, uint32_t ah, uint32_t bl, uint32_t bh, uint32_t one, uint32_t& rl,
                                          uint32_t& rh) {
    asm("add.cc.u32 %0, %2, %4;\n\tmadc.lo.u32 %1, %3, %6, %5;"
        : "=r"(rl), "=r"(rh) : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(one));
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack64(uint64_t v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Advance the linear engine one step and return the ++-scrambled 64-bit output as (hi, lo).
template <int V>
__device__ __forceinline__ void xoshiro_next_v(Xoshiro256& g, uint32_t& out_hi, uint32_t& out_lo, uint32_t one = 1u) {
    // a = s0 + s3
    uint32_t al, ah;
    if (V & XO_ADD_WIDE) {
        unpack64(mad_wide(g.s3l, one, pack64(g.s0l, g.s0h)), al, ah);
        ah = mad_lo(g.s3h, one, ah);
    } else if (V & XO_CARRY_MAD) {
        add64_mad(g.s0l, g.s0h, g.s3l, g.s3h, one, al, ah);
    } else {
        add64(g.s0l, g.s0h, g.s3l, g.s3h, al, ah);
    }
    // r = rotl(a, 23) + s0
    if (V & XO_ROT_WIDE) {
        // rotl(a,23) = (a.lo * 2^23 as 64 bits) + ((a.hi << 23) << 32) + (a.hi >> 9): three disjoint bit fields
        uint32_t tl, th;
        unpack64(mad_wide(al, 1u << 23, pack64(g.s0l, g.s0h)), tl, th);
        th = mad_lo(ah, 1u << 23, th);
        unpack64(mad_wide(ah >> 9, one, pack64(tl, th)), out_lo, out_hi);
    } else {
        const uint32_t rh = __funnelshift_l(al, ah, 23);
        const uint32_t rl = __funnelshift_l(ah, al, 23);
        if (V & XO_CARRY_MAD) add64_mad(rl, rh, g.s0l, g.s0h, one, out_lo, out_hi);
        else add64(g.s0l, g.s0h, rl, rh, out_lo, out_hi);
    }
    // t = s1 << 17
    uint32_t tl, th;
    if (V & XO_T_WIDE) {
        unpack64(mad_wide(g.s1l, 1u << 17, 0ull), tl, th);
        th = mad_lo(g.s1h, 1u << 17, th);
    } else if (V & XO_CARRY_MAD) {
        tl = mad_lo(g.s1l, one << 17, 0u);                          // IMAD with a run-time 2^17: never an ALU-pipe SHF
        th = __funnelshift_l(g.s1l, g.s1h, 17);
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
    // rotl(x3, 45) = rotl(y, 13) with y = x3 with its halves swapped (y.lo = x3h, y.hi = x3l)
    if (V & XO_ROT45_WIDE) {
        uint32_t rl, rh;
        unpack64(mad_wide(x3h, 1u << 13, 0ull), rl, rh);            // (y.lo >> 19 : y.lo << 13)
        g.s3h = mad_lo(x3l, 1u << 13, rh);                          // + y.hi << 13
        g.s3l = mad_hi(x3l, 1u << 13, rl);                          // + y.hi >> 19 (disjoint from y.lo << 13)
    } else {
        g.s3h = __funnelshift_l(x3h, x3l, 13);
        g.s3l = __funnelshift_l(x3l, x3h, 13);
    }
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
