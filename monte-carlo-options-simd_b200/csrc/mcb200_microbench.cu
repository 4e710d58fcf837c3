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
#include "../../include/mcb200.h"
#include "mcb200_pair.cuh"

namespace mcb200 {

constexpr int kChains = 8;
constexpr int kIters = 8192;

enum PipeOp { OP_MUFU_LG2, OP_MUFU_SIN, OP_MUFU_SQRT, OP_MUFU_EX2, OP_FFMA, OP_FADD, OP_FMUL, OP_LOP3, OP_SHF,
              OP_IADD3, OP_IADD_CARRY, OP_IMAD, OP_IMAD_HI, OP_IMAD_WIDE,
              OP_PAIR_V0, OP_PAIR_V1, OP_PAIR_V0H, OP_PAIR_V1H, OP_XOSHIRO_V0, OP_XOSHIRO_V1, OP_COUNT };

static const char* kOpNames[OP_COUNT] = {"mufu_lg2", "mufu_sin(+fmul)", "mufu_sqrt", "mufu_ex2", "ffma", "fadd", "fmul",
                                         "lop3", "shf_funnel", "iadd3_3in", "iadd64_carry_pair", "imad", "imad_hi",
                                         "imad_wide",
                                         "pair_v0", "pair_v1_twide", "pair_v0_madhi", "pair_v1_twide_madhi",
                                         "xoshiro_v0", "xoshiro_v1_twide"};

template <int OP>
__device__ __forceinline__ void op_step(uint32_t& x, uint32_t y, uint32_t z) {
    float fx = __uint_as_float(x);
    const float fy = __uint_as_float(y), fz = __uint_as_float(z);
    if (OP == OP_MUFU_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (OP == OP_MUFU_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (OP == OP_MUFU_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (OP == OP_MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fx));
    if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fx) : "f"(fy), "f"(fz));
    if (OP == OP_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fx) : "f"(fy));
    if (OP == OP_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(fx) : "f"(fy));
    if (OP <= OP_FMUL) x = __float_as_uint(fx);
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(x) : "r"(y));
    if (OP == OP_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    if (OP == OP_IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
}

template <int OP>
__global__ void __launch_bounds__(kThreads) pipe_kernel(uint32_t* sink, long long* cycles, uint32_t seed) {
    uint32_t v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = 0x3f800000u + (threadIdx.x + c * 977u + seed) % 4096u;
    const uint32_t y = 0x3f800001u + (seed & 7u), z = 0x3a000000u + (seed & 3u);
    __syncthreads();
    const long long t0 = clock64();
    if (OP >= OP_PAIR_V0) {
        Xoshiro256 g{v[0], v[1], v[2], v[3], v[4] | 1u, v[5], v[6], v[7]};
        float acc = 0.0f;
        uint32_t mix = 0;
        const uint32_t two23 = (seed >> 1 | 1u) << 23;      // 2^23, known only at run time (seed = 1, 2)
#pragma unroll 4
        for (int i = 0; i < kIters * kChains / 8; ++i) {
            if (OP == OP_PAIR_V0) acc = pair_fast_v<0, false>(g, acc, two23);
            if (OP == OP_PAIR_V1) acc = pair_fast_v<1, false>(g, acc, two23);
            if (OP == OP_PAIR_V0H) acc = pair_fast_v<0, true>(g, acc, two23);
            if (OP == OP_PAIR_V1H) acc = pair_fast_v<1, true>(g, acc, two23);
            if (OP == OP_XOSHIRO_V0) { uint32_t hi, lo; xoshiro_next_v<0>(g, hi, lo); mix ^= hi + lo; }
            if (OP == OP_XOSHIRO_V1) { uint32_t hi, lo; xoshiro_next_v<1>(g, hi, lo); mix ^= hi + lo; }
        }
        v[0] = __float_as_uint(acc) ^ mix ^ g.s0l;
    } else if (OP == OP_IMAD_WIDE) {
        unsigned long long w[kChains];
#pragma unroll
        for (int c = 0; c < kChains; ++c) w[c] = ((unsigned long long)v[c] << 32) | y;
        for (int i = 0; i < kIters; ++i) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(y), "r"(seed));
        }
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    } else if (OP == OP_IADD_CARRY) {
        uint32_t lo[kChains / 2], hi[kChains / 2];
#pragma unroll
        for (int c = 0; c < kChains / 2; ++c) { lo[c] = v[2 * c]; hi[c] = v[2 * c + 1]; }
        for (int i = 0; i < kIters; ++i) {
#pragma unroll
            for (int c = 0; c < kChains / 2; ++c)
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(y), "r"(z));
        }
#pragma unroll
        for (int c = 0; c < kChains / 2; ++c) { v[2 * c] = lo[c]; v[2 * c + 1] = hi[c]; }
    } else {
        for (int i = 0; i < kIters; ++i) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) op_step<OP>(v[c], y, z);
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) r ^= v[c];
    if (r == 0x12345678u) sink[0] = r;          // keeps the chains alive without a store per thread
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static int run_one(cudaStream_t st, int sm_count, int ctas_per_sm, uint32_t* d_sink, long long* d_cycles,
                   std::vector<long long>& h_cycles, mcb200_pipe_rate* out) {
    const int grid = sm_count * ctas_per_sm;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
    pipe_kernel<OP><<<grid, kThreads, 0, st>>>(d_sink, d_cycles, 1u);          // warm-up
    cudaEventRecord(e0, st);
    pipe_kernel<OP><<<grid, kThreads, 0, st>>>(d_sink, d_cycles, 2u);
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cudaMemcpy(h_cycles.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    double mean_cycles = 0.0; long long max_cycles = 0;
    for (int i = 0; i < grid; ++i) { mean_cycles += (double)h_cycles[i]; if (h_cycles[i] > max_cycles) max_cycles = h_cycles[i]; }
    mean_cycles /= grid;
    // ops per thread: kIters*kChains single instructions (iadd64_carry_pair: pairs, i.e. half as many; pair_* and
    // xoshiro_*: calls, kIters*kChains/8 of them).  Rate = thread-ops per SM / cycles of the longest-running CTA (warps
    // finish at different times under the greedy scheduler, so the mean CTA time would overstate the rate).
    double ops_per_thread = (double)kIters * kChains;
    if (OP >= OP_PAIR_V0) ops_per_thread /= 8.0;
    if (OP == OP_IADD_CARRY) ops_per_thread /= 2.0;
    const double ops_per_sm = ops_per_thread * kThreads * ctas_per_sm;
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->name, sizeof(out->name), "%s", kOpNames[OP]);
    out->ops_per_clk_per_sm = ops_per_sm / (double)max_cycles;
    (void)mean_cycles;
    out->elapsed_ms = ms;
    out->sm_clock_mhz = (double)max_cycles / (ms * 1e3);
    return 0;
}

int run_microbench(cudaStream_t st, int sm_count, mcb200_pipe_rate* out, int cap) {
    const int ctas_per_sm = 8;
    uint32_t* d_sink = nullptr; long long* d_cycles = nullptr;
    const int grid = sm_count * ctas_per_sm;
    if (cudaMalloc(&d_sink, 64) != cudaSuccess) return -1;
    if (cudaMalloc(&d_cycles, sizeof(long long) * grid) != cudaSuccess) { cudaFree(d_sink); return -1; }
    std::vector<long long> h(grid);
    int n = 0, rc = 0;
#define MCB_RUN(OP) if (rc == 0 && n < cap) { rc = run_one<OP>(st, sm_count, ctas_per_sm, d_sink, d_cycles, h, &out[n]); if (rc == 0) ++n; }
    MCB_RUN(OP_MUFU_LG2) MCB_RUN(OP_MUFU_SIN) MCB_RUN(OP_MUFU_SQRT) MCB_RUN(OP_MUFU_EX2)
    MCB_RUN(OP_FFMA) MCB_RUN(OP_FADD) MCB_RUN(OP_FMUL)
    MCB_RUN(OP_LOP3) MCB_RUN(OP_SHF) MCB_RUN(OP_IADD3) MCB_RUN(OP_IADD_CARRY) MCB_RUN(OP_IMAD) MCB_RUN(OP_IMAD_HI)
    MCB_RUN(OP_IMAD_WIDE)
    MCB_RUN(OP_PAIR_V0) MCB_RUN(OP_PAIR_V1) MCB_RUN(OP_PAIR_V0H) MCB_RUN(OP_PAIR_V1H)
    MCB_RUN(OP_XOSHIRO_V0) MCB_RUN(OP_XOSHIRO_V1)
#undef MCB_RUN
    cudaFree(d_sink); cudaFree(d_cycles);
    return rc == 0 ? n : -1;
}

}  // namespace mcb200
