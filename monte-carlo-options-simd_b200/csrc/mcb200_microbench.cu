// mcb200_microbench.cu -- measured per-SM pipe rates (MUFU / FMA / ALU / issue) on the device at hand.
//
// The GBM path kernel is bound by instruction issue and the XU (MUFU) pipe, not by HBM or tensor cores, and
// MEASURED_PEAKS.json only holds HBM and tensor peaks; these microbenchmarks supply measured denominators for the
// roofline (build-plan step 3).  Each kernel runs kChains independent dependency chains of one
// instruction per thread at full occupancy; rate = thread-ops / (SM cycles from clock64) / SMs.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/mcb200_testing.h"
#include "mcb200_pair.cuh"

namespace mcb200 {

constexpr int kChains = 8;
constexpr int kIters = 8192;

// single-instruction kinds (one hex digit each in a mix code)
enum Kind { K_LOP3 = 1, K_SHF = 2, K_IADD3 = 3, K_IMAD = 4, K_IMAD_HI = 5, K_FFMA = 6, K_FADD_IMM = 7, K_FMUL_IMM = 8,
            K_EX2 = 9, K_I2FP = 10, K_LG2 = 11, K_SIN = 12, K_SQRT = 13, K_FFMA_IMM = 14, K_LEA_HI = 15 };

template <int KIND>
__device__ __forceinline__ void op_step(uint32_t& x, uint32_t y, uint32_t z) {
    float fx = __uint_as_float(x);
    const float fy = __uint_as_float(y), fz = __uint_as_float(z);
    if (KIND == K_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (KIND == K_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (KIND == K_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (KIND == K_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fx) : "f"(fy), "f"(fz));
    if (KIND == K_FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F490FDB;" : "+f"(fx) : "f"(fy));
    if (KIND == K_FADD_IMM) asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+f"(fx));
    if (KIND == K_FMUL_IMM) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(fx));
    if (KIND == K_LG2 || KIND == K_SIN || KIND == K_SQRT || KIND == K_EX2 || KIND == K_FFMA || KIND == K_FFMA_IMM ||
        KIND == K_FADD_IMM || KIND == K_FMUL_IMM) x = __float_as_uint(fx);
    if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(x) : "r"(y));
    if (KIND == K_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(y), "r"(z));
    if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (KIND == K_IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (KIND == K_I2FP) asm volatile("cvt.rz.f32.u32 %0, %0;" : "+r"(x));
    if (KIND == K_LEA_HI) asm volatile("{.reg .u32 t; shr.u32 t, %0, 9; or.b32 %0, t, 0x3f800000;}" : "+r"(x));
}

// CODE = eight hex digits, one Kind per slot (most significant digit = slot 0); every slot is its own dependency chain.
template <unsigned CODE>
__global__ void __launch_bounds__(kThreads) pipe_kernel(uint32_t* sink, long long* cycles, uint32_t seed) {
    uint32_t v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = 0x3f800000u + (threadIdx.x + c * 977u + seed) % 4096u;
    const uint32_t y = 0x3f800001u + (seed & 7u), z = 0x3a000000u + (seed & 3u);
    __syncthreads();
    const long long t0 = clock64();
    constexpr int k0 = (int)((CODE >> 28) & 15u), k1 = (int)((CODE >> 24) & 15u), k2 = (int)((CODE >> 20) & 15u),
                  k3 = (int)((CODE >> 16) & 15u), k4 = (int)((CODE >> 12) & 15u), k5 = (int)((CODE >> 8) & 15u),
                  k6 = (int)((CODE >> 4) & 15u), k7 = (int)(CODE & 15u);
#pragma unroll 2
    for (int i = 0; i < kIters; ++i) {
        op_step<k0>(v[0], y, z); op_step<k1>(v[1], y, z); op_step<k2>(v[2], y, z); op_step<k3>(v[3], y, z);
        op_step<k4>(v[4], y, z); op_step<k5>(v[5], y, z); op_step<k6>(v[6], y, z); op_step<k7>(v[7], y, z);
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) r ^= v[c];
    if (r == 0x12345678u) sink[0] = r;          // keeps the chains alive without a store per thread
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Bench {
    cudaStream_t st; int sm_count; uint32_t* d_sink; long long* d_cycles; std::vector<long long> h;
    mcb200_pipe_rate* out; int cap; int n = 0; int rc = 0;
};

// Times the second of two launches; rate = ops_per_thread * threads resident per SM / cycles of the slowest CTA.
template <typename Launch>
static void timed(Bench& b, const char* name, int ctas_per_sm, double ops_per_thread, Launch launch, int threads = kThreads) {
    if (b.rc != 0 || b.n >= b.cap) return;
    const int grid = b.sm_count * ctas_per_sm;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { b.rc = -1; return; }
    launch(grid, 1u);
    cudaEventRecord(e0, b.st);
    launch(grid, 2u);
    cudaEventRecord(e1, b.st);
    if (cudaStreamSynchronize(b.st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { b.rc = -1; return; }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cudaMemcpy(b.h.data(), b.d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost) != cudaSuccess) { b.rc = -1; return; }
    long long max_cycles = 0;
    for (int i = 0; i < grid; ++i) if (b.h[i] > max_cycles) max_cycles = b.h[i];
    mcb200_pipe_rate* o = &b.out[b.n++];
    std::memset(o, 0, sizeof(*o));
    std::snprintf(o->name, sizeof(o->name), "%s", name);
    // warps finish at different times under the greedy scheduler: the slowest CTA bounds the sustained rate
    o->ops_per_clk_per_sm = ops_per_thread * threads * ctas_per_sm / (double)max_cycles;
    o->elapsed_ms = ms;
    o->sm_clock_mhz = (double)max_cycles / (ms * 1e3);
}

template <unsigned CODE>
static void run_mix(Bench& b, const char* name) {
    timed(b, name, 8, (double)kIters * kChains, [&](int grid, uint32_t seed) {
        pipe_kernel<CODE><<<grid, kThreads, 0, b.st>>>(b.d_sink, b.d_cycles, seed);
    });
}

// ------------------------------------------------------------------------------------------------ pair / generator loops
// The production inner step (and the bare generator) for the pipe-balance variants of mcb200_rng.cuh / mcb200_pair.cuh.
// Rate = calls (one call = one Box-Muller pair = 2 path-steps, or one 64-bit output).  EXTRA adds instructions of one
// kind per call on a side chain, to read the marginal cost of that kind at the kernel's operating point:
//   1: +4 FFMA   2: +2 LOP3   3: +2 IMAD   4: +2 IMAD.WIDE
constexpr int kPairIters = 8192;
extern __shared__ uint32_t dyn_smem[];

template <int V, int FMODE, bool GEN_ONLY, int UNROLL, int EXTRA>
__global__ void __launch_bounds__(kThreads, 8) pair_kernel(uint32_t* sink, long long* cycles, uint32_t seed) {
    const uint32_t t = threadIdx.x + seed;
    Xoshiro256 g{0x3f800000u + t % 4096u, t * 977u, t * 31u + 7u, t ^ 0x9e3779b9u, (t * 5u) | 1u, t + 11u, t * 3u, ~t};
    const uint32_t one = seed >> 1 | 1u;                    // 1, known only at run time (seed = 1, 2)
    const uint32_t two23 = one << 23;
    float acc = 0.0f;
    uint32_t mix = 0, e0 = t, e1 = t * 3u;
    const uint32_t y = 0x3f800001u + (seed & 7u), z = 0x3a000000u + (seed & 3u);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll UNROLL
    for (int i = 0; i < kPairIters; ++i) {
        if (GEN_ONLY) { uint32_t hi, lo; xoshiro_next_v<V>(g, hi, lo, one); mix ^= hi + lo; }
        else if (FMODE == 9) acc = quad_fast(g, acc);      // 4 pairs per call: the host scales the rate
        else if (FMODE == 10) acc = tri_fast(g, acc, one);      // 3 pairs per call
        else acc = pair_fast_v<V, FMODE>(g, acc, two23, one);
        if (EXTRA == 1) { op_step<K_FFMA>(e0, y, z); op_step<K_FFMA>(e1, y, z); op_step<K_FFMA>(e0, y, z); op_step<K_FFMA>(e1, y, z); }
        if (EXTRA == 2) { op_step<K_LOP3>(e0, y, z); op_step<K_LOP3>(e1, y, z); }
        if (EXTRA == 3) { op_step<K_IMAD>(e0, y, z); op_step<K_IMAD>(e1, y, z); }
        if (EXTRA == 4) {
            asm volatile("{.reg .b64 w; mul.wide.u32 w, %0, %2; mov.b64 {%0, %1}, w;}" : "+r"(e0), "+r"(e1) : "r"(y));
            asm volatile("{.reg .b64 w; mul.wide.u32 w, %1, %2; mov.b64 {%0, %1}, w;}" : "+r"(e0), "+r"(e1) : "r"(z));
        }
    }
    const long long t1 = clock64();
    const uint32_t r = __float_as_uint(acc) ^ mix ^ g.s0l ^ g.s1h ^ g.s2l ^ g.s3h ^ e0 ^ e1;
    if (r == 0x12345678u) sink[0] = r ^ dyn_smem[0];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ctas_per_sm < 8 is forced with dynamic shared memory (the kernel itself always fits 8 CTAs of 256 threads per SM).
template <int V, int FMODE, bool GEN_ONLY, int UNROLL, int EXTRA>
static void run_pair(Bench& b, int ctas_per_sm = 8, int threads = kThreads) {
    char name[32], occ[16] = "";
    if (ctas_per_sm != 8 || threads != kThreads) std::snprintf(occ, sizeof(occ), "_w%d", ctas_per_sm * threads / 32);
    std::snprintf(name, sizeof(name), "%s_xo%d_f%d_u%d%s%s", GEN_ONLY ? "gen" : "pair", V, FMODE, UNROLL,
                  EXTRA == 1 ? "+4ffma" : EXTRA == 2 ? "+2lop3" : EXTRA == 3 ? "+2imad" : EXTRA == 4 ? "+2wide" : "", occ);
    auto kern = pair_kernel<V, FMODE, GEN_ONLY, UNROLL, EXTRA>;
    size_t smem = 0;
    if (ctas_per_sm < 8) {
        smem = (size_t)(227 * 1024) / ctas_per_sm - 2048;          // k CTAs fit, k+1 do not
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { b.rc = -1; return; }
    }
    timed(b, name, ctas_per_sm, (double)kPairIters * (FMODE == 9 ? 4 : FMODE == 10 ? 3 : 1), [&](int grid, uint32_t seed) {
        kern<<<grid, threads, smem, b.st>>>(b.d_sink, b.d_cycles, seed);
    }, threads);
}

#define MIX8(a, b, c, d, e, f, g, h) ((unsigned)(a) << 28 | (b) << 24 | (c) << 20 | (d) << 16 | (e) << 12 | (f) << 8 | (g) << 4 | (h))
#define ALL8(k) MIX8(k, k, k, k, k, k, k, k)
#define ALT8(j, k) MIX8(j, k, j, k, j, k, j, k)

int run_microbench(cudaStream_t st, int sm_count, mcb200_pipe_rate* out, int cap) {
    Bench b{st, sm_count, nullptr, nullptr, std::vector<long long>((size_t)sm_count * 8), out, cap};
    if (cudaMalloc(&b.d_sink, 64) != cudaSuccess) return -1;
    if (cudaMalloc(&b.d_cycles, sizeof(long long) * sm_count * 8) != cudaSuccess) { cudaFree(b.d_sink); return -1; }
    // single pipes
    run_mix<ALL8(K_LG2)>(b, "mufu_lg2"); run_mix<ALL8(K_SIN)>(b, "mufu_sin(+fmul)"); run_mix<ALL8(K_SQRT)>(b, "mufu_sqrt");
    run_mix<ALL8(K_EX2)>(b, "mufu_ex2");
    run_mix<ALL8(K_FFMA)>(b, "ffma"); run_mix<ALL8(K_FFMA_IMM)>(b, "ffma_imm"); run_mix<ALL8(K_FADD_IMM)>(b, "fadd_imm");
    run_mix<ALL8(K_FMUL_IMM)>(b, "fmul_imm");
    run_mix<ALL8(K_LOP3)>(b, "lop3"); run_mix<ALL8(K_SHF)>(b, "shf_funnel"); run_mix<ALL8(K_IADD3)>(b, "iadd3_3in");
    run_mix<ALL8(K_LEA_HI)>(b, "lea_hi"); run_mix<ALL8(K_IMAD)>(b, "imad"); run_mix<ALL8(K_IMAD_HI)>(b, "imad_hi");
    run_mix<ALL8(K_I2FP)>(b, "i2fp_u32_rz");
    // two- and three-pipe mixes (rate counts every instruction)
    run_mix<ALT8(K_LOP3, K_IMAD)>(b, "mix_lop3+imad"); run_mix<ALT8(K_LOP3, K_FFMA)>(b, "mix_lop3+ffma");
    run_mix<ALT8(K_LOP3, K_FFMA_IMM)>(b, "mix_lop3+ffma_imm"); run_mix<ALT8(K_LOP3, K_FADD_IMM)>(b, "mix_lop3+fadd_imm");
    run_mix<ALT8(K_IMAD, K_FFMA)>(b, "mix_imad+ffma"); run_mix<ALT8(K_LOP3, K_I2FP)>(b, "mix_lop3+i2fp");
    run_mix<ALT8(K_IMAD, K_I2FP)>(b, "mix_imad+i2fp"); run_mix<ALT8(K_LOP3, K_IMAD_HI)>(b, "mix_lop3+imad_hi");
    run_mix<MIX8(K_LOP3, K_LOP3, K_LOP3, K_FFMA, K_LOP3, K_LOP3, K_LOP3, K_FFMA)>(b, "mix_3lop3+ffma");
    run_mix<MIX8(K_LOP3, K_LOP3, K_LOP3, K_EX2, K_LOP3, K_LOP3, K_LOP3, K_LOP3)>(b, "mix_7lop3+mufu");
    run_mix<MIX8(K_LOP3, K_IMAD, K_FFMA, K_LOP3, K_IMAD, K_FFMA, K_LOP3, K_IMAD)>(b, "mix_lop3+imad+ffma");
    run_mix<MIX8(K_IMAD, K_IMAD, K_IMAD, K_EX2, K_IMAD, K_IMAD, K_IMAD, K_IMAD)>(b, "mix_7imad+mufu");
    // the inner step: generator alone, production pair, variants, marginal costs, occupancy
    run_pair<0, 0, true, 4, 0>(b);
    run_pair<0, 0, false, 4, 0>(b); run_pair<0, 0, false, 8, 0>(b);
    run_pair<0, 4, false, 4, 0>(b); run_pair<0, 4, false, 8, 0>(b); run_pair<1, 4, false, 4, 0>(b);
    run_pair<0, 0, false, 4, 1>(b); run_pair<0, 0, false, 4, 2>(b); run_pair<0, 0, false, 4, 3>(b); run_pair<0, 0, false, 4, 4>(b);
    run_pair<0, 0, false, 4, 0>(b, 7); run_pair<0, 0, false, 4, 0>(b, 6); run_pair<0, 0, false, 4, 0>(b, 5);
    run_pair<0, 0, false, 4, 0>(b, 4); run_pair<0, 0, false, 4, 0>(b, 3); run_pair<0, 0, false, 4, 0>(b, 2);
    run_pair<0, 0, false, 8, 0>(b, 4); run_pair<0, 0, false, 8, 0>(b, 2);
    run_pair<0, 9, false, 2, 0>(b, 4); run_pair<0, 9, false, 1, 0>(b, 4); run_pair<0, 9, false, 2, 0>(b, 8);   // quad_fast
    run_pair<0, 9, false, 4, 0>(b, 4); run_pair<0, 9, false, 2, 0>(b, 2); run_pair<0, 9, false, 2, 0>(b, 1, 256);
    run_pair<0, 9, false, 4, 0>(b, 1, 256); run_pair<0, 9, false, 2, 0>(b, 1, 128); run_pair<0, 9, false, 4, 0>(b, 1, 128);
    // tri_fast (production): two outputs -> three pairs, 16-bit angle fields placed by PRMT
    run_pair<0, 10, false, 1, 0>(b, 4); run_pair<0, 10, false, 2, 0>(b, 4); run_pair<0, 10, false, 3, 0>(b, 4);
    run_pair<0, 10, false, 4, 0>(b, 4); run_pair<0, 10, false, 2, 0>(b, 8); run_pair<0, 10, false, 2, 0>(b, 3);
    run_pair<0, 10, false, 2, 0>(b, 2); run_pair<0, 10, false, 2, 0>(b, 1, 256); run_pair<0, 10, false, 2, 0>(b, 1, 128);
    run_pair<0, 10, false, 2, 1>(b, 4); run_pair<0, 10, false, 2, 2>(b, 4); run_pair<0, 10, false, 2, 3>(b, 4);
    // few warps per SM (named by resident warps per SM): how much of the pipe can 1 or 2 warps per sub-partition use?
    run_pair<0, 0, false, 8, 0>(b, 1, 256); run_pair<0, 0, false, 8, 0>(b, 1, 128); run_pair<0, 0, false, 4, 0>(b, 1, 128);
    run_pair<0, 0, false, 8, 0>(b, 1, 32);
    cudaFree(b.d_sink); cudaFree(b.d_cycles);
    return b.rc == 0 ? b.n : -1;
}

}  // namespace mcb200
