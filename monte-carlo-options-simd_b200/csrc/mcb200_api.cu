// mcb200_api.cu -- host runtime and C ABI of libmcb200.so (declared in include/mcb200.h).
//
// Replaces, for the hot path only, the Rust `pub fn` surface of the reference's mc_simd module
// (/root/reference/src/mc_simd.rs:477-779) and its rayon/thread_rng plumbing (mc_simd.rs:49-54,71-78):
//   context  = device + stream + HBM pool of per-thread xoshiro256++ substreams + pinned staging,
//   planner  = mcb200_plan.cpp, kernels = mcb200_kernels.cuh.
// There is no CPU path: without a CUDA device every entry point reports MCB200_ERR_NO_DEVICE.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/mcb200.h"
#include "../../include/mcb200_testing.h"
#include "../../include/xoshiro256_jump_tables.h"
#include "mcb200_kernels.cuh"
#include "mcb200_plan.h"
#include "mcb200_internal.h"

namespace mcb200 {

int run_microbench(cudaStream_t st, int sm_count, mcb200_pipe_rate* out, int cap);   // mcb200_microbench.cu

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? MCB200_ERR_NOMEM : MCB200_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                      \
    } while (0)

// ------------------------------------------------------------------------------------------------ device allocations
// Production: plain cudaMalloc / cudaFree.  -DMCB_CHECKED builds (tools/checked_build.py; compute-sanitizer is closed
// on the GPU pool this was developed on) put a 256-byte guard filled with 0xA5 on either side of every long-lived device
// buffer of a context, and mcb200_debug_check() verifies the guards, the in-kernel index checks and that every
// arrival counter is back to zero.
#ifdef MCB_CHECKED
static std::mutex g_guard_mu;
static std::map<void*, size_t> g_guarded;          // user pointer -> user bytes
constexpr size_t kGuard = 256;
static cudaError_t dev_malloc(void** p, size_t bytes) {
    char* raw = nullptr;
    cudaError_t e = cudaMalloc((void**)&raw, bytes + 2 * kGuard);
    if (e != cudaSuccess) return e;
    cudaMemset(raw, 0xA5, kGuard);
    cudaMemset(raw + kGuard + bytes, 0xA5, kGuard);
    *p = raw + kGuard;
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guarded[*p] = bytes;
    return cudaSuccess;
}
static cudaError_t dev_free(void* p) {
    if (!p) return cudaSuccess;
    { std::lock_guard<std::mutex> lk(g_guard_mu); g_guarded.erase(p); }
    return cudaFree((char*)p - kGuard);
}
static int guards_damaged() {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    int bad = 0;
    unsigned char h[2 * kGuard];
    for (auto& kv : g_guarded) {
        cudaMemcpy(h, (char*)kv.first - kGuard, kGuard, cudaMemcpyDeviceToHost);
        cudaMemcpy(h + kGuard, (char*)kv.first + kv.second, kGuard, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < 2 * kGuard; ++i) if (h[i] != 0xA5) { ++bad; break; }
    }
    return bad;
}
#else
static cudaError_t dev_malloc(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
static cudaError_t dev_free(void* p) { return cudaFree(p); }
#endif

// ------------------------------------------------------------------------------------------------ host xoshiro (seeding only)
struct HostXo { uint64_t s[4]; };
static inline uint64_t h_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline void h_step(HostXo& g) {
    const uint64_t t = g.s[1] << 17;
    g.s[2] ^= g.s[0]; g.s[3] ^= g.s[1]; g.s[1] ^= g.s[2]; g.s[0] ^= g.s[3]; g.s[2] ^= t; g.s[3] = h_rotl(g.s[3], 45);
}
static void h_apply_poly(HostXo& g, const unsigned long long poly[4]) {
    uint64_t a[4] = {0, 0, 0, 0};
    for (int w = 0; w < 4; ++w)
        for (int b = 0; b < 64; ++b) {
            if (poly[w] & (1ull << b)) for (int k = 0; k < 4; ++k) a[k] ^= g.s[k];
            h_step(g);
        }
    for (int k = 0; k < 4; ++k) g.s[k] = a[k];
}
static inline uint64_t h_splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

}  // namespace mcb200

using namespace mcb200;

// ------------------------------------------------------------------------------------------------ context
struct mcb200_ctx {
    int device = 0;
    int sm_count = 0;
    int ctas_per_sm = 0;          // resident 256-thread CTAs per SM of the price / antithetic kernels
    int grid_slots = 0;           // ... and the CTAs the planner fills for them
    int grid_slots_greeks = 0;    // the Greeks kernel runs at fewer, fatter CTAs per SM (mcb200_pair.cuh)
    int clock_khz = 0;
    cudaStream_t stream = nullptr;
    uint64_t seed = 0;
    uint32_t block = 0;
    uint64_t* pool = nullptr;     // grid_slots * kThreads * kPlanWaves states (one per logical thread of the largest plan)
    int64_t pool_threads = 0;
    uint64_t* d_jump = nullptr;   // jump polynomial table
    double* slots = nullptr; size_t slots_cap = 0;            // doubles: per (warp, option) partial moment sums
    unsigned* arrivals = nullptr; size_t arrivals_cap = 0;    // per-option arrival counters, zero between launches
    double* moments = nullptr; size_t moments_cap = 0;        // doubles (only for the zero-path corner)
    float* h_stage = nullptr; float* d_stage = nullptr; size_t stage_cap = 0;   // floats (pinned+mapped host / device)
    float* h_stage_dev = nullptr;  // device-side address of h_stage: small batches write their results straight to it
    // cross-GPU exchange buffer (trials-sharded fused combine): owned (cudaMalloc) or opened through CUDA IPC
    // (same-process ranks attach to the owner's allocation directly: plain peer access, mcb200_xgpu_attach)
    void* xg_base = nullptr; bool xg_owner = false; bool xg_attached = false; int xg_world = 0, xg_n_max = 0, xg_n_acc = 0;
    unsigned xg_epoch = 0;        // rounds launched through the exchange buffer (all ranks call in lockstep)
    bool xg_pending = false;      // a round was launched and its results not fetched yet: launch and results alternate
    unsigned long long xg_timeout_ns = 120ull * 1000 * 1000 * 1000;   // device-side wait for the other ranks
    int* xg_status = nullptr;     // device word written by xgpu_wait_kernel
    cudaEvent_t xg_event = nullptr;   // recorded behind the latest xgpu launch (it may sit on a caller stream)
    uint64_t launches = 0;
    // Launch ordering.  Pool, slots and arrival counters are one set per context, so two launches of one context must
    // never overlap on the device: every launch records order_ev behind itself, and a launch that arrives on a
    // different stream than its predecessor first waits for that event (cudaStreamWaitEvent, device side only).
    cudaEvent_t order_ev = nullptr; cudaStream_t order_stream = nullptr; bool order_valid = false;
#ifdef MCB_TRACE
    unsigned long long* trace = nullptr;    // device buffer of the timeline build (tools/kernel_timeline.py)
#endif
#ifdef MCB_CHECKED
    unsigned* check = nullptr;              // device word: OR of the in-kernel index checks that failed
#endif
    bool timing = false; bool timed_once = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    mcb200::MultiRuntime* multi = nullptr;   // non-null: this is the parent of a multi-device context (mcb200_multi.cu)
    std::mutex mu;
};

namespace mcb200 {
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
MultiRuntime* ctx_multi(mcb200_ctx* c) { return c ? c->multi : nullptr; }
mcb200_ctx* ctx_new_parent(MultiRuntime* m, int device0, uint64_t seed, uint32_t first_block) {
    mcb200_ctx* c = new (std::nothrow) mcb200_ctx();
    if (c) { c->multi = m; c->device = device0; c->seed = seed; c->block = first_block; }
    return c;
}
}  // namespace mcb200

// Device pointers and streams belong to ONE device: the *_device, xgpu and inspection entry points refuse a parent.
#define SINGLE_DEVICE_ONLY(c, what)                                                                        \
    do {                                                                                                   \
        if ((c) && (c)->multi)                                                                             \
            return fail(MCB200_ERR_INVALID, what ": not available on a multi-device context (use the host-buffer " \
                        "entry points, or mcb200_ctx_child for one device)");                             \
    } while (0)

static int ensure(double** p, size_t* cap, size_t need) {
    if (*cap >= need) return MCB200_OK;
    if (*p) dev_free(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 2 + 1024;
    CU(dev_malloc((void**)p, want * sizeof(double)));
    *cap = want;
    return MCB200_OK;
}

static int ensure_arrivals(mcb200_ctx* c, size_t need) {
    if (c->arrivals_cap >= need) return MCB200_OK;
    if (c->arrivals) dev_free(c->arrivals);
    c->arrivals = nullptr; c->arrivals_cap = 0;
    size_t want = need + need / 2 + 1024;
    CU(dev_malloc((void**)&c->arrivals, want * sizeof(unsigned)));
    CU(cudaMemsetAsync(c->arrivals, 0, want * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));        // launches may arrive on caller streams
    c->arrivals_cap = want;
    return MCB200_OK;
}

static int ensure_stage(mcb200_ctx* c, size_t need_floats) {
    if (c->stage_cap >= need_floats) return MCB200_OK;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) dev_free(c->d_stage);
    c->h_stage = nullptr; c->d_stage = nullptr; c->h_stage_dev = nullptr; c->stage_cap = 0;
    size_t want = need_floats + need_floats / 2 + 4096;
    CU(cudaHostAlloc((void**)&c->h_stage, want * sizeof(float), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void**)&c->h_stage_dev, c->h_stage, 0));
    CU(dev_malloc((void**)&c->d_stage, want * sizeof(float)));
    c->stage_cap = want;
    return MCB200_OK;
}

static int seed_pool(mcb200_ctx* c) {
    uint64_t x = c->seed;
    HostXo base;
    for (int k = 0; k < 4; ++k) base.s[k] = h_splitmix(x);
    // long_jump^block in O(log block): one polynomial x^(2^(192+k)) per set bit of the block index
    for (int k = 0; k < XOSHIRO256_LONG_JUMP_POW2_COUNT; ++k)
        if ((c->block >> k) & 1u) h_apply_poly(base, XOSHIRO256_LONG_JUMP_POW2[k]);
    const ulonglong4 bs = make_ulonglong4(base.s[0], base.s[1], base.s[2], base.s[3]);
    const int threads = 128;
    const int blocks = (int)((c->pool_threads + threads - 1) / threads);
    seed_pool_kernel<<<blocks, threads, 0, c->stream>>>(c->pool, (int)c->pool_threads, 0ull, bs, c->d_jump,
                                                         XOSHIRO256_JUMP_POW2_COUNT);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return MCB200_OK;
}

// Called with the context mutex held, before / after enqueuing work that touches the context's shared device state.
// A launch on a CALLER stream records the event right away (the caller may destroy that stream at any time); a launch
// on the context's own stream records nothing -- the event is recorded on that stream lazily, at the moment a launch
// arrives on a different stream -- so the host-buffer entry points and the drop-ins pay no event per call.
static int order_before(mcb200_ctx* c, cudaStream_t st) {
    if (!c->order_valid || c->order_stream == st) return MCB200_OK;
    if (!c->order_ev) CU(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
    if (c->order_stream == c->stream) CU(cudaEventRecord(c->order_ev, c->stream));     // lazy record for the own stream
    CU(cudaStreamWaitEvent(st, c->order_ev, 0));
    return MCB200_OK;
}
static int order_after(mcb200_ctx* c, cudaStream_t st) {
    if (st != c->stream) {
        if (!c->order_ev) CU(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
        CU(cudaEventRecord(c->order_ev, st));
    }
    c->order_stream = st; c->order_valid = true;
    return MCB200_OK;
}

template <typename K>
static int occupancy(K kernel, int* out) {
    int n = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0));
    *out = n;
    return MCB200_OK;
}

extern "C" int mcb200_ctx_create(mcb200_ctx** out, int device, uint64_t seed, uint32_t substream_block) {
    if (!out) return fail(MCB200_ERR_INVALID, "mcb200_ctx_create: null output pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(MCB200_ERR_NO_DEVICE, "no CUDA device (%s); libmcb200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n_dev) return fail(MCB200_ERR_INVALID, "device %d out of range [0,%d)", device, n_dev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MCB200_ERR_NO_DEVICE, "device %d is sm_%d%d; libmcb200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    mcb200_ctx* c = new (std::nothrow) mcb200_ctx();
    if (!c) return fail(MCB200_ERR_NOMEM, "out of host memory");
    c->device = device; c->sm_count = prop.multiProcessorCount; c->seed = seed; c->block = substream_block;
    cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device);
    int rc = MCB200_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(MCB200_ERR_CUDA, "stream create failed"); break; }
        int occ_price = 0, occ_av = 0, occ_greeks = 0;
        if ((rc = occupancy(simulate_kernel<STREAM_FAST, PAYOFF_PRICE, false>, &occ_price))) break;
        if ((rc = occupancy(simulate_kernel<STREAM_FAST, PAYOFF_AV, false>, &occ_av))) break;
        if ((rc = occupancy(simulate_kernel<STREAM_FAST, PAYOFF_GREEKS, false>, &occ_greeks))) break;
        int occ = occ_price < occ_av ? occ_price : occ_av;
        if (occ < 1 || occ_greeks < 1) { rc = fail(MCB200_ERR_CUDA, "kernel occupancy query returned %d / %d", occ, occ_greeks); break; }
        if (occ > kPriceCtasPerSm) occ = kPriceCtasPerSm;         // the design points (mcb200_pair.cuh); more residency buys nothing
        if (occ_greeks > kGreeksCtasPerSm) occ_greeks = kGreeksCtasPerSm;
        c->ctas_per_sm = occ;
        c->grid_slots = c->sm_count * occ;
        c->grid_slots_greeks = c->sm_count * occ_greeks;
        c->pool_threads = (int64_t)(c->grid_slots > c->grid_slots_greeks ? c->grid_slots : c->grid_slots_greeks) * kThreads *
                          kPlanWaves;                  // one substream per LOGICAL thread of a multi-wave launch
        if (dev_malloc((void**)&c->pool, (size_t)c->pool_threads * 32) != cudaSuccess) { rc = fail(MCB200_ERR_NOMEM, "pool alloc failed"); break; }
        if (cudaMalloc((void**)&c->d_jump, sizeof(XOSHIRO256_JUMP_POW2)) != cudaSuccess) { rc = fail(MCB200_ERR_NOMEM, "jump table alloc failed"); break; }
        if (cudaMemcpy(c->d_jump, XOSHIRO256_JUMP_POW2, sizeof(XOSHIRO256_JUMP_POW2), cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(MCB200_ERR_CUDA, "jump table copy failed"); break; }
        if ((rc = seed_pool(c))) break;
    } while (0);
    if (rc != MCB200_OK) { mcb200_ctx_destroy(c); return rc; }
    *out = c;
    return MCB200_OK;
}

extern "C" void mcb200_ctx_destroy(mcb200_ctx* c) {
    if (!c) return;
    if (c->multi) { multi_destroy(c->multi); delete c; return; }
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->order_ev) cudaEventDestroy(c->order_ev);
    if (c->pool) dev_free(c->pool);
    if (c->d_jump) cudaFree(c->d_jump);
    if (c->slots) dev_free(c->slots);
    if (c->arrivals) dev_free(c->arrivals);
    if (c->moments) dev_free(c->moments);
#ifdef MCB_CHECKED
    if (c->check) cudaFree(c->check);
#endif
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) dev_free(c->d_stage);
    if (c->xg_base) { if (c->xg_owner) cudaFree(c->xg_base); else if (!c->xg_attached) cudaIpcCloseMemHandle(c->xg_base); }
    if (c->xg_status) cudaFree(c->xg_status);
    if (c->xg_event) cudaEventDestroy(c->xg_event);
    delete c;
}

extern "C" int mcb200_ctx_set_seed(mcb200_ctx* c, uint64_t seed, uint32_t substream_block) {
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    if (c->multi) { c->seed = seed; c->block = substream_block; return multi_set_seed(c->multi, seed, substream_block); }
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    c->seed = seed; c->block = substream_block;
    return seed_pool(c);
}

extern "C" const char* mcb200_last_error(void) { return g_err; }

extern "C" int mcb200_ctx_info(mcb200_ctx* c, mcb200_info* out) {
    if (!c || !out) return fail(MCB200_ERR_INVALID, "null argument");
    if (c->multi) {                          // the first device's figures; launches summed over all devices
        int rc = mcb200_ctx_info(multi_child(c->multi, 0), out);
        if (rc == MCB200_OK) { out->kernel_launches = multi_launches(c->multi); out->seed = c->seed; out->substream_block = c->block; }
        return rc;
    }
    out->device = c->device; out->sm_count = c->sm_count; out->ctas_per_sm = c->ctas_per_sm;
    out->threads_per_cta = kThreads; out->grid_slots = c->grid_slots; out->grid_slots_greeks = c->grid_slots_greeks;
    out->pool_threads = c->pool_threads;
    out->kernel_launches = c->launches; out->seed = c->seed; out->substream_block = c->block;
    out->sm_clock_khz = c->clock_khz;
    return MCB200_OK;
}

// ------------------------------------------------------------------------------------------------ launches
struct Request {
    int n = 0;
    OptionArrays opt{};            // device pointers
    float steps = 0, num_trials = 0;
    int payoff = PAYOFF_PRICE;
    int stream_mode = STREAM_FAST;
    bool dump = false;
    Bumps bumps{1.0f, 1.0f, 1.0f, 1.0f};
    const uint64_t* chunk_seeds = nullptr;   // device
    float* growth_out = nullptr;             // device
    const float* const* inline_cols = nullptr;   // host columns of a small batch: they travel in the launch parameters
    bool xg = false; int xg_rank = 0;            // cross-GPU fused combine through the context's exchange buffer
    double xg_total_paths = 0, xg_total_trials = 0;
};

// Exchange buffer layout on the owner GPU: [world][n_max][n_acc] fp64 rows | [2 * n_est][n_max] f32 results | [n_max] u32
struct XgLayout { double* slots; float* results; unsigned* arrivals; unsigned* ctl; size_t bytes; };
static XgLayout xg_layout_for(void* base, int world, int n_max, int n_acc) {
    XgLayout L;
    char* p = (char*)base;
    L.slots = (double*)p; p += sizeof(double) * (size_t)world * n_max * n_acc;
    L.results = (float*)p; p += sizeof(float) * (size_t)n_acc * n_max;
    L.arrivals = (unsigned*)p; p += sizeof(unsigned) * (size_t)n_max;
    L.ctl = (unsigned*)p; p += sizeof(unsigned) * 4;
    L.bytes = (size_t)(p - (char*)base);
    return L;
}
static XgLayout xg_layout(mcb200_ctx* c) { return xg_layout_for(c->xg_base, c->xg_world, c->xg_n_max, c->xg_n_acc); }

template <int STREAM, int PAYOFF, bool DUMP>
static void launch_sim(const SimArgs& a, int grid, cudaStream_t st) {
    simulate_kernel<STREAM, PAYOFF, DUMP><<<grid, kThreads, 0, st>>>(a);
}

static FinalizeSpec make_finalize(int n_est, double n_paths, double num_trials, const float* rate, const float* years,
                                  Bumps bumps, float* const* out, float* const* out_se) {
    FinalizeSpec f{};
    f.n_paths = n_paths; f.num_trials = num_trials; f.rate = rate; f.years = years; f.bumps = bumps;
    for (int e = 0; e < EST_COUNT_GREEKS; ++e) {
        f.out[e] = (e < n_est && out) ? out[e] : nullptr;
        f.out_se[e] = (e < n_est && out_se) ? out_se[e] : nullptr;
    }
    return f;
}

static int run_finalize(mcb200_ctx* c, int n, int n_est, const double* moments, double n_paths, double num_trials,
                        const float* rate, const float* years, Bumps bumps, float* const* out, float* const* out_se,
                        cudaStream_t st) {
    if (n == 0) return MCB200_OK;
    const FinalizeSpec f = make_finalize(n_est, n_paths, num_trials, rate, years, bumps, out, out_se);
    const int total = n * n_est;
    finalize_kernel<<<(total + 127) / 128, 128, 0, st>>>(moments, n, n_est, f);
    c->launches++;
    CU(cudaGetLastError());
    return MCB200_OK;
}

// ONE launch per batch: simulate, reduce deterministically, then (finalize) write values and SEs and / or (moments_out)
// the per-option moment sums [n][2*n_est].
static int run_sim(mcb200_ctx* c, const Request& rq, double* moments_out, bool finalize, float* const* out,
                   float* const* out_se, cudaStream_t st, mcb200_plan* plan_out) {
    mcb200_plan pl;
    int rc = make_plan(rq.n, rq.steps, rq.num_trials, rq.payoff == PAYOFF_GREEKS ? c->grid_slots_greeks : c->grid_slots, &pl);
    if (rc != MCB200_OK) return fail(rc, "invalid request: n=%d steps=%g num_trials=%g", rq.n, rq.steps, rq.num_trials);
    if (plan_out) *plan_out = pl;
    const int n_est = rq.payoff == PAYOFF_GREEKS ? EST_COUNT_GREEKS : EST_COUNT_PRICE;
    const int n_acc = 2 * n_est;
    if (rq.n == 0) return MCB200_OK;
    if (pl.n_paths == 0) {
        // nothing to simulate (num_trials < 8): all sums are zero, like the reference's empty map/reduce
        double* mo = moments_out;
        if (!mo) {
            if ((rc = ensure(&c->moments, &c->moments_cap, (size_t)rq.n * n_acc))) return rc;
            mo = c->moments;
        }
        if ((rc = order_before(c, st))) return rc;
        CU(cudaMemsetAsync(mo, 0, sizeof(double) * (size_t)rq.n * n_acc, st));
        if (finalize)
            rc = run_finalize(c, rq.n, n_est, mo, 0.0, (double)rq.num_trials, rq.opt.rate, rq.opt.years, rq.bumps, out,
                              out_se, st);
        if (rc) return rc;
        return order_after(c, st);
    }
    if ((rc = ensure(&c->slots, &c->slots_cap, ((size_t)pl.warps + (size_t)rq.n) * n_acc))) return rc;
    if ((rc = ensure_arrivals(c, (size_t)rq.n))) return rc;
    SimArgs a{};
    a.opt = rq.opt; a.n_options = rq.n; a.n_paths = pl.n_paths; a.half_steps = pl.half_steps; a.steps = rq.steps;
    a.rounds_per_option = pl.rounds_per_option; a.total_rounds = pl.total_rounds; a.n_warps = pl.warps;
    a.bumps = rq.bumps; a.pool = c->pool; a.chunk_seeds = rq.chunk_seeds; a.slots = c->slots; a.arrivals = c->arrivals;
    a.moments_out = moments_out; a.finalize = finalize ? 1 : 0;
    a.fin = make_finalize(n_est, (double)pl.n_paths, (double)rq.num_trials, rq.opt.rate, rq.opt.years, rq.bumps, out, out_se);
    a.growth_out = rq.growth_out; a.one = 1u;
#ifdef MCB_TRACE
    a.trace = c->trace;
#endif
#ifdef MCB_CHECKED
    if (!c->check) { CU(cudaMalloc((void**)&c->check, sizeof(unsigned))); CU(cudaMemset(c->check, 0, sizeof(unsigned))); }
    a.check = c->check; a.slots_cap = (unsigned long long)c->slots_cap; a.pool_threads = c->pool_threads;
    if (std::getenv("MCB200_CHECK_INJECT")) a.slots_cap = 1;   // self-test of the detector: the slot check must fire
    a.xg_doubles = rq.xg ? (unsigned long long)c->xg_world * c->xg_n_max * n_acc : 0ull;
#endif
    if (rq.xg) {
        XgLayout L = xg_layout(c);
        a.xg_slots = L.slots; a.xg_arrivals = L.arrivals; a.xg_ctl = L.ctl; a.xg_epoch = c->xg_epoch + 1;
        a.xg_rank = rq.xg_rank; a.xg_world = c->xg_world;
        a.xg_n_max = c->xg_n_max;
        a.fin.n_paths = rq.xg_total_paths; a.fin.num_trials = rq.xg_total_trials;
    }
    if (rq.inline_cols) {
        a.use_inline = 1;
        for (int k = 0; k < 6; ++k) std::memcpy(a.inline_opt[k], rq.inline_cols[k], sizeof(float) * rq.n);
    }
    const int key = rq.stream_mode * 100 + rq.payoff * 10 + (rq.dump ? 1 : 0);
    if ((rc = order_before(c, st))) return rc;
    if (c->timing) CU(cudaEventRecord(c->ev0, st));
    switch (key) {
        case 0:   launch_sim<STREAM_FAST, PAYOFF_PRICE, false>(a, pl.grid, st); break;
        case 1:   launch_sim<STREAM_FAST, PAYOFF_PRICE, true>(a, pl.grid, st); break;
        case 10:  launch_sim<STREAM_FAST, PAYOFF_AV, false>(a, pl.grid, st); break;
        case 11:  launch_sim<STREAM_FAST, PAYOFF_AV, true>(a, pl.grid, st); break;
        case 20:  launch_sim<STREAM_FAST, PAYOFF_GREEKS, false>(a, pl.grid, st); break;
        case 100: launch_sim<STREAM_REF, PAYOFF_PRICE, false>(a, pl.grid, st); break;
        case 101: launch_sim<STREAM_REF, PAYOFF_PRICE, true>(a, pl.grid, st); break;
        case 110: launch_sim<STREAM_REF, PAYOFF_AV, false>(a, pl.grid, st); break;
        case 111: launch_sim<STREAM_REF, PAYOFF_AV, true>(a, pl.grid, st); break;
        case 120: launch_sim<STREAM_REF, PAYOFF_GREEKS, false>(a, pl.grid, st); break;
        default: return fail(MCB200_ERR_INVALID, "unsupported kernel variant %d", key);
    }
    c->launches++;
    CU(cudaGetLastError());
    if (c->timing) { CU(cudaEventRecord(c->ev1, st)); c->timed_once = true; }
    return order_after(c, st);
}

static int check_common(mcb200_ctx* c, int n, const float* const* arrays, int n_arrays) {
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    SINGLE_DEVICE_ONLY(c, "device-pointer entry point");
    if (n < 0) return fail(MCB200_ERR_INVALID, "n = %d < 0", n);
    if (n > 0)
        for (int i = 0; i < n_arrays; ++i)
            if (!arrays[i]) return fail(MCB200_ERR_INVALID, "null input array %d", i);
    return MCB200_OK;
}

static Bumps to_bumps(const mcb200_bumps* b) {
    return Bumps{b->delta_spot, b->delta_volatility, b->delta_risk_free_rate, b->delta_years_to_expiry};
}

// Device-resident core shared by every entry point.
static int run_device(mcb200_ctx* c, const Request& rq, float* const* out, float* const* out_se, cudaStream_t st) {
    return run_sim(c, rq, nullptr, true, out, out_se, st, nullptr);
}

extern "C" int mcb200_price_batch_device(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                         const float* vol, const float* rate, const float* years, const float* div,
                                         float steps, float num_trials, unsigned flags, float* call_out,
                                         float* put_out, float* call_se, float* put_se, void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps; rq.num_trials = num_trials;
    rq.payoff = (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE;
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return run_device(c, rq, out, se, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

extern "C" int mcb200_greeks_batch_device(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                          const float* vol, const float* rate, const float* years, const float* div,
                                          float steps, float num_trials, const mcb200_bumps* bumps,
                                          float* const out[MCB200_N_GREEK_OUTPUTS],
                                          float* const out_se[MCB200_N_GREEK_OUTPUTS], void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps; rq.num_trials = num_trials;
    rq.payoff = PAYOFF_GREEKS; rq.bumps = to_bumps(bumps);
    return run_device(c, rq, out, out_se, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

extern "C" int mcb200_price_moments_device(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                           const float* vol, const float* rate, const float* years, const float* div,
                                           float steps, float num_trials_local, unsigned flags, double* moments,
                                           void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (n > 0 && !moments) return fail(MCB200_ERR_INVALID, "null moments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps;
    rq.num_trials = num_trials_local;
    rq.payoff = (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE;
    return run_sim(c, rq, moments, false, nullptr, nullptr, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, nullptr);
}

extern "C" int mcb200_price_finalize_device(mcb200_ctx* c, int n, const float* rate, const float* years,
                                            const double* moments, double total_paths, double total_num_trials,
                                            float* call_out, float* put_out, float* call_se, float* put_se,
                                            void* cuda_stream) {
    const float* in[2] = {rate, years};
    int rc = check_common(c, n, in, 2);
    if (rc) return rc;
    if (n > 0 && !moments) return fail(MCB200_ERR_INVALID, "null moments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return run_finalize(c, n, EST_COUNT_PRICE, moments, total_paths, total_num_trials, rate, years,
                        Bumps{1, 1, 1, 1}, out, se, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

extern "C" int mcb200_greeks_moments_device(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                            const float* vol, const float* rate, const float* years, const float* div,
                                            float steps, float num_trials_local, const mcb200_bumps* bumps,
                                            double* moments, void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    if (n > 0 && !moments) return fail(MCB200_ERR_INVALID, "null moments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps;
    rq.num_trials = num_trials_local; rq.payoff = PAYOFF_GREEKS; rq.bumps = to_bumps(bumps);
    return run_sim(c, rq, moments, false, nullptr, nullptr, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, nullptr);
}

extern "C" int mcb200_greeks_finalize_device(mcb200_ctx* c, int n, const float* rate, const float* years,
                                             const double* moments, double total_paths, double total_num_trials,
                                             const mcb200_bumps* bumps, float* const out[MCB200_N_GREEK_OUTPUTS],
                                             float* const out_se[MCB200_N_GREEK_OUTPUTS], void* cuda_stream) {
    const float* in[2] = {rate, years};
    int rc = check_common(c, n, in, 2);
    if (rc) return rc;
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    if (n > 0 && !moments) return fail(MCB200_ERR_INVALID, "null moments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    return run_finalize(c, n, EST_COUNT_GREEKS, moments, total_paths, total_num_trials, rate, years, to_bumps(bumps), out,
                        out_se, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

// ------------------------------------------------------------------------------------------------ cross-GPU fused combine
static_assert(sizeof(cudaIpcMemHandle_t) == MCB200_IPC_HANDLE_BYTES, "IPC handle size");

static int xg_release(mcb200_ctx* c) {
    if (c->xg_base) {
        cudaStreamSynchronize(c->stream);
        if (c->xg_owner) cudaFree(c->xg_base); else if (!c->xg_attached) cudaIpcCloseMemHandle(c->xg_base);
    }
    c->xg_base = nullptr; c->xg_owner = false; c->xg_attached = false; c->xg_world = c->xg_n_max = c->xg_n_acc = 0;
    c->xg_epoch = 0; c->xg_pending = false;
    return MCB200_OK;
}

extern "C" int mcb200_xgpu_create(mcb200_ctx* c, int n_max, int world, int greeks, uint8_t handle_out[MCB200_IPC_HANDLE_BYTES]) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_create");
    if (!c || !handle_out || n_max <= 0 || world <= 0) return fail(MCB200_ERR_INVALID, "bad xgpu_create arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    xg_release(c);
    const int n_acc = 2 * (greeks ? EST_COUNT_GREEKS : EST_COUNT_PRICE);
    const XgLayout L = xg_layout_for(nullptr, world, n_max, n_acc);
    CU(cudaMalloc(&c->xg_base, L.bytes));
    CU(cudaMemset(c->xg_base, 0, L.bytes));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->xg_base));
    std::memcpy(handle_out, &h, sizeof(h));
    c->xg_owner = true; c->xg_world = world; c->xg_n_max = n_max; c->xg_n_acc = n_acc;
    return MCB200_OK;
}

extern "C" int mcb200_xgpu_open(mcb200_ctx* c, const uint8_t handle[MCB200_IPC_HANDLE_BYTES], int n_max, int world, int greeks) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_open");
    if (!c || !handle || n_max <= 0 || world <= 0) return fail(MCB200_ERR_INVALID, "bad xgpu_open arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    xg_release(c);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(&c->xg_base, h, cudaIpcMemLazyEnablePeerAccess));
    c->xg_owner = false; c->xg_world = world; c->xg_n_max = n_max; c->xg_n_acc = 2 * (greeks ? EST_COUNT_GREEKS : EST_COUNT_PRICE);
    return MCB200_OK;
}

// Same-process variant of mcb200_xgpu_open: several contexts of ONE process (one per GPU, or several on one GPU) share
// the owner's allocation through plain peer access -- no IPC handle.  The owner must outlive the attached contexts'
// use of the buffer.
extern "C" int mcb200_xgpu_attach(mcb200_ctx* c, mcb200_ctx* owner, int n_max, int world, int greeks) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_attach");
    if (!c || !owner || c == owner || n_max <= 0 || world <= 0) return fail(MCB200_ERR_INVALID, "bad xgpu_attach arguments");
    const int n_acc = 2 * (greeks ? EST_COUNT_GREEKS : EST_COUNT_PRICE);
    void* base = nullptr;
    unsigned epoch = 0;
    {
        std::lock_guard<std::mutex> lk(owner->mu);
        if (!owner->xg_base || !owner->xg_owner || owner->xg_n_max != n_max || owner->xg_world != world || owner->xg_n_acc != n_acc)
            return fail(MCB200_ERR_INVALID, "xgpu_attach: the owner holds no exchange buffer of this shape (n_max=%d world=%d)", n_max, world);
        if (owner->xg_pending) return fail(MCB200_ERR_INVALID, "xgpu_attach: the owner has a round pending");
        base = owner->xg_base;
        epoch = owner->xg_epoch;           // join the owner's round count (attach between rounds, on every rank alike)
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    xg_release(c);
    if (c->device != owner->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, c->device, owner->device));
        if (!can) return fail(MCB200_ERR_CUDA, "xgpu_attach: device %d cannot access device %d", c->device, owner->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail(MCB200_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    }
    c->xg_base = base; c->xg_owner = false; c->xg_attached = true; c->xg_world = world; c->xg_n_max = n_max; c->xg_n_acc = n_acc;
    c->xg_epoch = epoch;
    return MCB200_OK;
}

extern "C" int mcb200_xgpu_set_timeout(mcb200_ctx* c, double seconds) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_set_timeout");
    if (!c || !(seconds > 0.0) || seconds > 86400.0) return fail(MCB200_ERR_INVALID, "bad xgpu timeout");
    std::lock_guard<std::mutex> lk(c->mu);
    c->xg_timeout_ns = (unsigned long long)(seconds * 1e9);
    return MCB200_OK;
}

extern "C" int mcb200_xgpu_close(mcb200_ctx* c) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_close");
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    return xg_release(c);
}

// One launch per GPU, no collective: this GPU simulates num_trials_local trials of every option; the results of the
// WHOLE sample (all ranks) appear in the exchange buffer once every rank's launch has finished.
extern "C" int mcb200_price_xgpu_device(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                                        const float* rate, const float* years, const float* div, float steps,
                                        float num_trials_local, unsigned flags, int rank, double total_paths,
                                        double total_num_trials, void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (!c->xg_base) return fail(MCB200_ERR_INVALID, "no exchange buffer: call mcb200_xgpu_create / mcb200_xgpu_open first");
    if (n > c->xg_n_max || rank < 0 || rank >= c->xg_world || c->xg_n_acc != 2 * EST_COUNT_PRICE)
        return fail(MCB200_ERR_INVALID, "xgpu: n=%d (max %d), rank=%d (world %d)", n, c->xg_n_max, rank, c->xg_world);
    if (c->xg_pending)
        return fail(MCB200_ERR_INVALID, "xgpu: round %u is still pending -- fetch it with mcb200_xgpu_results before launching the next", c->xg_epoch);
    mcb200_plan pl;
    if (make_plan(n, steps, num_trials_local, c->grid_slots, &pl) != MCB200_OK || pl.n_paths == 0)
        return fail(MCB200_ERR_INVALID, "xgpu: every rank must simulate at least 8 trials (steps=%g trials=%g)", steps, num_trials_local);
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps; rq.num_trials = num_trials_local;
    rq.payoff = (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE;
    rq.xg = true; rq.xg_rank = rank; rq.xg_total_paths = total_paths; rq.xg_total_trials = total_num_trials;
    const XgLayout L = xg_layout(c);
    float* out[2] = {L.results, L.results + c->xg_n_max};
    float* se[2] = {L.results + 2 * (size_t)c->xg_n_max, L.results + 3 * (size_t)c->xg_n_max};
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    rc = run_sim(c, rq, nullptr, true, out, se, st, nullptr);      // launches round xg_epoch + 1
    if (rc) return rc;
    ++c->xg_epoch; c->xg_pending = true;  // counted only once the launch is in the stream: ranks stay in step on errors
    if (!c->xg_event) CU(cudaEventCreateWithFlags(&c->xg_event, cudaEventDisableTiming));
    CU(cudaEventRecord(c->xg_event, st));
    return MCB200_OK;
}

extern "C" int mcb200_greeks_xgpu_device(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                                         const float* rate, const float* years, const float* div, float steps,
                                         float num_trials_local, const mcb200_bumps* bumps, int rank, double total_paths,
                                         double total_num_trials, void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (!c->xg_base) return fail(MCB200_ERR_INVALID, "no exchange buffer: call mcb200_xgpu_create / mcb200_xgpu_open first");
    if (n > c->xg_n_max || rank < 0 || rank >= c->xg_world || c->xg_n_acc != 2 * EST_COUNT_GREEKS)
        return fail(MCB200_ERR_INVALID, "xgpu (greeks): n=%d (max %d), rank=%d (world %d), buffer created for %d sums",
                    n, c->xg_n_max, rank, c->xg_world, c->xg_n_acc);
    if (c->xg_pending)
        return fail(MCB200_ERR_INVALID, "xgpu: round %u is still pending -- fetch it with mcb200_xgpu_results_greeks before launching the next", c->xg_epoch);
    mcb200_plan pl;
    if (make_plan(n, steps, num_trials_local, c->grid_slots_greeks, &pl) != MCB200_OK || pl.n_paths == 0)
        return fail(MCB200_ERR_INVALID, "xgpu: every rank must simulate at least 8 trials (steps=%g trials=%g)", steps, num_trials_local);
    Request rq;
    rq.n = n; rq.opt = OptionArrays{spot, strike, vol, rate, years, div}; rq.steps = steps; rq.num_trials = num_trials_local;
    rq.payoff = PAYOFF_GREEKS; rq.bumps = to_bumps(bumps);
    rq.xg = true; rq.xg_rank = rank; rq.xg_total_paths = total_paths; rq.xg_total_trials = total_num_trials;
    const XgLayout L = xg_layout(c);
    float* out[EST_COUNT_GREEKS]; float* se[EST_COUNT_GREEKS];
    for (int e = 0; e < EST_COUNT_GREEKS; ++e) {
        out[e] = L.results + (size_t)e * c->xg_n_max;
        se[e] = L.results + (size_t)(EST_COUNT_GREEKS + e) * c->xg_n_max;
    }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    rc = run_sim(c, rq, nullptr, true, out, se, st, nullptr);
    if (rc) return rc;
    ++c->xg_epoch; c->xg_pending = true;
    if (!c->xg_event) CU(cudaEventCreateWithFlags(&c->xg_event, cudaEventDisableTiming));
    CU(cudaEventRecord(c->xg_event, st));
    return MCB200_OK;
}

// Wait (on the device, no collective) until the latest round is complete on ALL ranks, then copy the combined results
// out of the exchange buffer (a peer read on non-owner ranks): rows [value e][n_max] then [SE e][n_max].
static int xg_fetch(mcb200_ctx* c, int n, int n_est, float* const* out, float* const* out_se) {
    if (!c->xg_base || n < 0 || n > c->xg_n_max || c->xg_n_acc != 2 * n_est)
        return fail(MCB200_ERR_INVALID, "xgpu_results: no matching exchange buffer or bad n");
    if (!c->xg_pending) return fail(MCB200_ERR_INVALID, "xgpu_results: no round is pending (launch and results alternate)");
    const XgLayout L = xg_layout(c);
    if (!c->xg_status) CU(cudaMalloc((void**)&c->xg_status, sizeof(int)));
    const size_t out_f = (size_t)c->xg_n_acc * c->xg_n_max;
    int rc = ensure_stage(c, out_f + 16);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    // the wait kernel must not share the GPU with this rank's own path kernel (it would take a CTA slot from it)
    if (c->xg_event) CU(cudaStreamWaitEvent(st, c->xg_event, 0));
    xgpu_wait_kernel<<<1, 1, 0, st>>>(L.ctl, c->xg_epoch, c->xg_timeout_ns, c->xg_status);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_stage, L.results, sizeof(float) * out_f, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_stage + out_f, c->xg_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int ok = 0;
    std::memcpy(&ok, c->h_stage + out_f, sizeof(int));
    c->xg_pending = false;
    if (!ok) return fail(MCB200_ERR_CUDA, "xgpu_results: round %u was not completed by all ranks within %.0f s", c->xg_epoch,
                         (double)c->xg_timeout_ns * 1e-9);
    for (int e = 0; e < n_est; ++e) {
        if (out && out[e] && n > 0) std::memcpy(out[e], c->h_stage + (size_t)e * c->xg_n_max, sizeof(float) * n);
        if (out_se && out_se[e] && n > 0) std::memcpy(out_se[e], c->h_stage + (size_t)(n_est + e) * c->xg_n_max, sizeof(float) * n);
    }
    return MCB200_OK;
}

extern "C" int mcb200_xgpu_results(mcb200_ctx* c, int n, float* call_out, float* put_out, float* call_se, float* put_se) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_results");
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return xg_fetch(c, n, EST_COUNT_PRICE, out, se);
}

extern "C" int mcb200_xgpu_results_greeks(mcb200_ctx* c, int n, float* const out[MCB200_N_GREEK_OUTPUTS],
                                          float* const out_se[MCB200_N_GREEK_OUTPUTS]) {
    SINGLE_DEVICE_ONLY(c, "mcb200_xgpu_results_greeks");
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    return xg_fetch(c, n, EST_COUNT_GREEKS, out, out_se);
}

// ------------------------------------------------------------------------------------------------ host-buffer entry points
constexpr int kDirectWriteMaxOptions = 4096;
// Staging layout (floats): [6][n] inputs | [2*n_est][n] outputs (value rows then SE rows).
static int run_host(mcb200_ctx* c, int n, const float* const in[6], float steps, float num_trials, int payoff,
                    int stream_mode, Bumps bumps, const uint8_t* chunk_seeds_host, float* const* out,
                    float* const* out_se, float* growth_out_host, bool dump) {
    if (c && c->multi) {
        if (stream_mode != STREAM_FAST || dump)
            return fail(MCB200_ERR_INVALID, "reference-stream / dump modes run on one device: use mcb200_ctx_child");
        if (n < 0) return fail(MCB200_ERR_INVALID, "n = %d < 0", n);
        if (n > 0) for (int i = 0; i < 6; ++i) if (!in[i]) return fail(MCB200_ERR_INVALID, "null input array %d", i);
        const mcb200_bumps mb{bumps.dspot, bumps.dvol, bumps.drate, bumps.dyears};
        return multi_run_host(c->multi, n, in, steps, num_trials, payoff == PAYOFF_GREEKS ? 1 : 0,
                              payoff == PAYOFF_AV ? MCB200_ANTITHETIC : 0u, &mb, out, out_se);
    }
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    mcb200_plan pl;
    rc = make_plan(n, steps, num_trials, payoff == PAYOFF_GREEKS ? c->grid_slots_greeks : c->grid_slots, &pl);
    if (rc) return fail(rc, "invalid request: n=%d steps=%g num_trials=%g", n, steps, num_trials);
    if (n == 0) return MCB200_OK;
    const int n_est = payoff == PAYOFF_GREEKS ? EST_COUNT_GREEKS : EST_COUNT_PRICE;
    const size_t in_f = 6 * (size_t)n, out_f = 2 * (size_t)n_est * n;
    if ((rc = ensure_stage(c, in_f + out_f))) return rc;
    cudaStream_t st = c->stream;
    Request rq;
    rq.n = n; rq.steps = steps; rq.num_trials = num_trials; rq.payoff = payoff; rq.stream_mode = stream_mode;
    rq.bumps = bumps; rq.dump = dump;
    // Small batches skip the copy engines: up to kInlineMaxOptions options carry their six input columns in the launch
    // parameters, and up to kDirectWriteMaxOptions options have their results written by the kernel straight into the
    // mapped pinned staging buffer (posted PCIe writes).  One launch + one stream synchronize is then the whole call.
    const bool inline_in = (n <= kInlineMaxOptions && pl.n_paths > 0);
    const bool direct_out = (n <= kDirectWriteMaxOptions);
    if (inline_in) {
        rq.inline_cols = in;
    } else {
        for (int k = 0; k < 6; ++k) std::memcpy(c->h_stage + (size_t)k * n, in[k], sizeof(float) * n);
        CU(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof(float) * in_f, cudaMemcpyHostToDevice, st));
    }
    const float* d = c->d_stage;
    rq.opt = OptionArrays{d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, d + 5 * (size_t)n};
    uint64_t* d_seeds = nullptr; float* d_growth = nullptr;
    const size_t seed_bytes = (size_t)n * (pl.n_paths / 8) * 256, growth_bytes = (size_t)n * pl.n_paths * sizeof(float);
    if (stream_mode == STREAM_REF) {
        if (!chunk_seeds_host) return fail(MCB200_ERR_INVALID, "null chunk_seeds");
        if (seed_bytes) {
            CU(cudaMalloc((void**)&d_seeds, seed_bytes));
            cudaError_t e = cudaMemcpyAsync(d_seeds, chunk_seeds_host, seed_bytes, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { cudaFree(d_seeds); return fail(MCB200_ERR_CUDA, "seed copy: %s", cudaGetErrorString(e)); }
        }
        rq.chunk_seeds = d_seeds;
    }
    if (dump && growth_bytes) {
        cudaError_t e = cudaMalloc((void**)&d_growth, growth_bytes);
        if (e != cudaSuccess) { if (d_seeds) cudaFree(d_seeds); return fail(MCB200_ERR_NOMEM, "growth buffer: %s", cudaGetErrorString(e)); }
        rq.growth_out = d_growth;
    }
    float* d_out[EST_COUNT_GREEKS]; float* d_se[EST_COUNT_GREEKS];
    float* out_base = (direct_out ? c->h_stage_dev : c->d_stage) + in_f;
    for (int e = 0; e < n_est; ++e) {
        d_out[e] = out_base + (size_t)e * n;
        d_se[e] = out_base + (size_t)(n_est + e) * n;
    }
    rc = run_device(c, rq, d_out, d_se, st);
    if (rc == MCB200_OK) {
        cudaError_t e = cudaSuccess;
        if (!direct_out)
            e = cudaMemcpyAsync(c->h_stage + in_f, c->d_stage + in_f, sizeof(float) * out_f, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && dump && growth_bytes)
            e = cudaMemcpyAsync(growth_out_host, d_growth, growth_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(MCB200_ERR_CUDA, "result copy / sync: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(st);
    }
    if (d_seeds) cudaFree(d_seeds);
    if (d_growth) cudaFree(d_growth);
    if (rc != MCB200_OK) return rc;
    for (int e = 0; e < n_est; ++e) {
        if (out && out[e]) std::memcpy(out[e], c->h_stage + in_f + (size_t)e * n, sizeof(float) * n);
        if (out_se && out_se[e]) std::memcpy(out_se[e], c->h_stage + in_f + (size_t)(n_est + e) * n, sizeof(float) * n);
    }
    return MCB200_OK;
}

// ------------------------------------------------------------------------------------------------ trials-minor building blocks
namespace mcb200 {

int host_moments(mcb200_ctx* c, int n, const float* const in[6], float steps, float num_trials_local, int greeks,
                 unsigned flags, const mcb200_bumps* bumps, double* moments_host) {
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (n == 0) return MCB200_OK;
    if (!moments_host) return fail(MCB200_ERR_INVALID, "null moments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const int n_acc = 2 * (greeks ? EST_COUNT_GREEKS : EST_COUNT_PRICE);
    const size_t in_f = 6 * (size_t)n;
    if ((rc = ensure_stage(c, in_f))) return rc;
    if ((rc = ensure(&c->moments, &c->moments_cap, (size_t)n * n_acc))) return rc;
    cudaStream_t st = c->stream;
    for (int k = 0; k < 6; ++k) std::memcpy(c->h_stage + (size_t)k * n, in[k], sizeof(float) * n);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof(float) * in_f, cudaMemcpyHostToDevice, st));
    const float* d = c->d_stage;
    Request rq;
    rq.n = n; rq.steps = steps; rq.num_trials = num_trials_local;
    rq.payoff = greeks ? PAYOFF_GREEKS : ((flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE);
    if (bumps) rq.bumps = to_bumps(bumps);
    rq.opt = OptionArrays{d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, d + 5 * (size_t)n};
    rc = run_sim(c, rq, c->moments, false, nullptr, nullptr, st, nullptr);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CU(cudaMemcpyAsync(moments_host, c->moments, sizeof(double) * (size_t)n * n_acc, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MCB200_OK;
}

int host_finalize(mcb200_ctx* c, int n, int greeks, const float* rate, const float* years, const double* moments_host,
                  double total_paths, double total_num_trials, const mcb200_bumps* bumps, float* const* out,
                  float* const* out_se) {
    const float* in[2] = {rate, years};
    int rc = check_common(c, n, in, 2);
    if (rc) return rc;
    if (n == 0) return MCB200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const int n_est = greeks ? EST_COUNT_GREEKS : EST_COUNT_PRICE;
    const int n_acc = 2 * n_est;
    const size_t in_f = 2 * (size_t)n, out_f = (size_t)n_acc * n;
    if ((rc = ensure_stage(c, in_f + out_f))) return rc;
    if ((rc = ensure(&c->moments, &c->moments_cap, (size_t)n * n_acc))) return rc;
    cudaStream_t st = c->stream;
    std::memcpy(c->h_stage, rate, sizeof(float) * n);
    std::memcpy(c->h_stage + n, years, sizeof(float) * n);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof(float) * in_f, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->moments, moments_host, sizeof(double) * (size_t)n * n_acc, cudaMemcpyHostToDevice, st));
    float* d_out[EST_COUNT_GREEKS]; float* d_se[EST_COUNT_GREEKS];
    for (int e = 0; e < n_est; ++e) {
        d_out[e] = c->d_stage + in_f + (size_t)e * n;
        d_se[e] = c->d_stage + in_f + (size_t)(n_est + e) * n;
    }
    rc = run_finalize(c, n, n_est, c->moments, total_paths, total_num_trials, c->d_stage, c->d_stage + n,
                      bumps ? to_bumps(bumps) : Bumps{1, 1, 1, 1}, d_out, d_se, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CU(cudaMemcpyAsync(c->h_stage + in_f, c->d_stage + in_f, sizeof(float) * out_f, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int e = 0; e < n_est; ++e) {
        if (out && out[e]) std::memcpy(out[e], c->h_stage + in_f + (size_t)e * n, sizeof(float) * n);
        if (out_se && out_se[e]) std::memcpy(out_se[e], c->h_stage + in_f + (size_t)(n_est + e) * n, sizeof(float) * n);
    }
    return MCB200_OK;
}

}  // namespace mcb200

extern "C" int mcb200_price_batch(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                                  const float* rate, const float* years, const float* div, float steps,
                                  float num_trials, unsigned flags, float* call_out, float* put_out, float* call_se,
                                  float* put_se) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return run_host(c, n, in, steps, num_trials, (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE, STREAM_FAST,
                    Bumps{1, 1, 1, 1}, nullptr, out, se, nullptr, false);
}

extern "C" int mcb200_greeks_batch(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                                   const float* rate, const float* years, const float* div, float steps,
                                   float num_trials, const mcb200_bumps* bumps,
                                   float* const out[MCB200_N_GREEK_OUTPUTS],
                                   float* const out_se[MCB200_N_GREEK_OUTPUTS]) {
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    const float* in[6] = {spot, strike, vol, rate, years, div};
    return run_host(c, n, in, steps, num_trials, PAYOFF_GREEKS, STREAM_FAST, to_bumps(bumps), nullptr, out, out_se,
                    nullptr, false);
}

extern "C" int mcb200_price_batch_refstream(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                            const float* vol, const float* rate, const float* years,
                                            const float* div, float steps, float num_trials, unsigned flags,
                                            const uint8_t* chunk_seeds, float* call_out, float* put_out,
                                            float* call_se, float* put_se, float* growth_out) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return run_host(c, n, in, steps, num_trials, (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE, STREAM_REF,
                    Bumps{1, 1, 1, 1}, chunk_seeds, out, se, growth_out, growth_out != nullptr);
}

extern "C" int mcb200_greeks_batch_refstream(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                             const float* vol, const float* rate, const float* years,
                                             const float* div, float steps, float num_trials,
                                             const mcb200_bumps* bumps, const uint8_t* chunk_seeds,
                                             float* const out[MCB200_N_GREEK_OUTPUTS],
                                             float* const out_se[MCB200_N_GREEK_OUTPUTS]) {
    if (!bumps) return fail(MCB200_ERR_INVALID, "null bumps");
    const float* in[6] = {spot, strike, vol, rate, years, div};
    return run_host(c, n, in, steps, num_trials, PAYOFF_GREEKS, STREAM_REF, to_bumps(bumps), chunk_seeds, out, out_se,
                    nullptr, false);
}

extern "C" int mcb200_price_batch_dump(mcb200_ctx* c, int n, const float* spot, const float* strike,
                                       const float* vol, const float* rate, const float* years, const float* div,
                                       float steps, float num_trials, unsigned flags, float* call_out, float* put_out,
                                       float* call_se, float* put_se, float* growth_out) {
    if (!growth_out) return fail(MCB200_ERR_INVALID, "null growth_out");
    const float* in[6] = {spot, strike, vol, rate, years, div};
    float* out[2] = {call_out, put_out};
    float* se[2] = {call_se, put_se};
    return run_host(c, n, in, steps, num_trials, (flags & MCB200_ANTITHETIC) ? PAYOFF_AV : PAYOFF_PRICE, STREAM_FAST,
                    Bumps{1, 1, 1, 1}, nullptr, out, se, growth_out, true);
}

// ------------------------------------------------------------------------------------------------ closed form
static int launch_bs(mcb200_ctx* c, int n, OptionArrays opt, float* const* out, cudaStream_t st) {
    BsArgs a{};
    a.opt = opt; a.n = n;
    for (int e = 0; e < EST_COUNT_GREEKS; ++e) a.out[e] = out ? out[e] : nullptr;
    bs_closed_form_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    return MCB200_OK;
}

extern "C" int mcb200_bs_batch_device(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                                      const float* rate, const float* years, const float* div,
                                      float* const out[MCB200_N_GREEK_OUTPUTS], void* cuda_stream) {
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (n == 0) return MCB200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    return launch_bs(c, n, OptionArrays{spot, strike, vol, rate, years, div}, out,
                     cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

extern "C" int mcb200_bs_batch(mcb200_ctx* c, int n, const float* spot, const float* strike, const float* vol,
                               const float* rate, const float* years, const float* div,
                               float* const out[MCB200_N_GREEK_OUTPUTS]) {
    if (c && c->multi) return mcb200_bs_batch(multi_child(c->multi, 0), n, spot, strike, vol, rate, years, div, out);
    const float* in[6] = {spot, strike, vol, rate, years, div};
    int rc = check_common(c, n, in, 6);
    if (rc) return rc;
    if (n == 0) return MCB200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t in_f = 6 * (size_t)n, out_f = (size_t)MCB200_N_GREEK_OUTPUTS * n;
    if ((rc = ensure_stage(c, in_f + out_f))) return rc;
    for (int k = 0; k < 6; ++k) std::memcpy(c->h_stage + (size_t)k * n, in[k], sizeof(float) * n);
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof(float) * in_f, cudaMemcpyHostToDevice, st));
    const float* d = c->d_stage;
    float* d_out[MCB200_N_GREEK_OUTPUTS];
    for (int e = 0; e < MCB200_N_GREEK_OUTPUTS; ++e) d_out[e] = c->d_stage + in_f + (size_t)e * n;
    rc = launch_bs(c, n, OptionArrays{d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, d + 5 * (size_t)n},
                   d_out, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_stage + in_f, c->d_stage + in_f, sizeof(float) * out_f, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int e = 0; e < MCB200_N_GREEK_OUTPUTS; ++e)
        if (out && out[e]) std::memcpy(out[e], c->h_stage + in_f + (size_t)e * n, sizeof(float) * n);
    return MCB200_OK;
}

extern "C" int mcb200_ctx_export_states(mcb200_ctx* c, int64_t first, int64_t count, uint64_t* states_out) {
    SINGLE_DEVICE_ONLY(c, "mcb200_ctx_export_states");
    if (!c || !states_out) return fail(MCB200_ERR_INVALID, "null argument");
    if (first < 0 || count < 0 || first + count > c->pool_threads)
        return fail(MCB200_ERR_INVALID, "state range [%lld,%lld) outside pool of %lld", (long long)first,
                    (long long)(first + count), (long long)c->pool_threads);
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(states_out, c->pool + 4 * first, (size_t)count * 32, cudaMemcpyDeviceToHost));
    return MCB200_OK;
}

extern "C" int mcb200_ctx_set_timing(mcb200_ctx* c, int enabled) {
    SINGLE_DEVICE_ONLY(c, "mcb200_ctx_set_timing");
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (enabled && !c->ev0) { CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); }
    c->timing = enabled != 0;
    c->timed_once = false;
    return MCB200_OK;
}

extern "C" int mcb200_ctx_last_sim_ms(mcb200_ctx* c, float* ms_out) {
    SINGLE_DEVICE_ONLY(c, "mcb200_ctx_last_sim_ms");
    if (!c || !ms_out) return fail(MCB200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->timing || !c->timed_once) return fail(MCB200_ERR_INVALID, "no timed simulate-kernel launch recorded");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms_out, c->ev0, c->ev1));
    return MCB200_OK;
}

#ifdef MCB_CHECKED
// bits: 1 pool index, 2 option index, 4 slot index, 8 arrival count, 16 final-sum index, 32 exchange-buffer index or
// cross-GPU arrival count, 256 a guard region was overwritten, 512 an arrival counter is not back to zero
extern "C" int mcb200_debug_check(mcb200_ctx* c, unsigned* bits_out) {
    if (!c || !bits_out) return fail(MCB200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    unsigned bits = 0;
    if (c->check) CU(cudaMemcpy(&bits, c->check, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (guards_damaged()) bits |= 256u;
    if (c->arrivals && c->arrivals_cap) {
        std::vector<unsigned> h(c->arrivals_cap);
        CU(cudaMemcpy(h.data(), c->arrivals, sizeof(unsigned) * c->arrivals_cap, cudaMemcpyDeviceToHost));
        for (unsigned v : h) if (v != 0u) { bits |= 512u; break; }
    }
    *bits_out = bits;
    return MCB200_OK;
}
#endif

#ifdef MCB_TRACE
extern "C" int mcb200_debug_set_trace(mcb200_ctx* c, unsigned long long* device_buf) {
    if (!c) return fail(MCB200_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->trace = device_buf;
    return MCB200_OK;
}
#endif

extern "C" int mcb200_microbench(mcb200_ctx* c, mcb200_pipe_rate* out, int cap) {
    if (!c || !out || cap <= 0) return -fail(MCB200_ERR_INVALID, "bad microbench arguments");
    if (c->multi) return mcb200_microbench(multi_child(c->multi, 0), out, cap);
    std::lock_guard<std::mutex> lk(c->mu);
    if (cudaSetDevice(c->device) != cudaSuccess) return -fail(MCB200_ERR_CUDA, "cudaSetDevice failed");
    int n = run_microbench(c->stream, c->sm_count, out, cap);
    if (n < 0) return -fail(MCB200_ERR_CUDA, "microbench failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->launches += 2 * (uint64_t)n;
    return n;
}

// ------------------------------------------------------------------------------------------------ default context + drop-ins
// Shared ownership: a drop-in holds a reference for the duration of its call, so mcb200_default_ctx_reset() on another
// thread retires the old context only after the last call using it has returned.
static std::shared_ptr<mcb200_ctx> g_default;
static std::mutex g_default_mu;

static int default_ctx_ref(std::shared_ptr<mcb200_ctx>* out) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default) {
        const char* dev = std::getenv("MCB200_DEVICE");
        const char* devs = std::getenv("MCB200_DEVICES");        // e.g. "0,1,2,3": a multi-device default context
        const char* seed = std::getenv("MCB200_SEED");
        const char* blk = std::getenv("MCB200_SUBSTREAM_BLOCK");
        const uint64_t sd = seed ? std::strtoull(seed, nullptr, 0) : 0x5EED5EED5EED5EEDull;
        const uint32_t bk = blk ? (uint32_t)std::strtoul(blk, nullptr, 0) : 0u;
        mcb200_ctx* raw = nullptr;
        int rc;
        if (devs && *devs) {
            int ids[64], n = 0;
            for (const char* p = devs; *p && n < 64;) {
                char* end = nullptr;
                long v = std::strtol(p, &end, 10);
                if (end == p) break;
                ids[n++] = (int)v;
                p = (*end == ',') ? end + 1 : end;
            }
            rc = mcb200_ctx_create_multi(&raw, ids, n, sd, bk);
        } else {
            rc = mcb200_ctx_create(&raw, dev ? std::atoi(dev) : 0, sd, bk);
        }
        if (rc != MCB200_OK) return rc;
        g_default = std::shared_ptr<mcb200_ctx>(raw, [](mcb200_ctx* p) { mcb200_ctx_destroy(p); });
    }
    *out = g_default;
    return MCB200_OK;
}

extern "C" int mcb200_default_ctx(mcb200_ctx** out) {
    std::shared_ptr<mcb200_ctx> ref;
    int rc = default_ctx_ref(&ref);
    if (rc != MCB200_OK) return rc;
    if (out) *out = ref.get();             // borrowed: valid until the next mcb200_default_ctx_reset()
    return MCB200_OK;
}

extern "C" int mcb200_default_ctx_reset(int device, uint64_t seed, uint32_t substream_block) {
    mcb200_ctx* raw = nullptr;
    int rc = mcb200_ctx_create(&raw, device, seed, substream_block);
    if (rc != MCB200_OK) return rc;
    std::shared_ptr<mcb200_ctx> fresh(raw, [](mcb200_ctx* p) { mcb200_ctx_destroy(p); });
    std::shared_ptr<mcb200_ctx> old;
    {
        std::lock_guard<std::mutex> lk(g_default_mu);
        old.swap(g_default);
        g_default = fresh;
    }
    return MCB200_OK;                      // `old` is destroyed here, or when the last drop-in using it returns
}

static const float kNaN = std::numeric_limits<float>::quiet_NaN();

// One option through the batched path; `which` picks the estimator to return.
static float dropin_price(float spot, float strike, float vol, float rate, float years, float div, float steps,
                          float num_trials, unsigned flags, int which) {
    std::shared_ptr<mcb200_ctx> ref;
    if (default_ctx_ref(&ref) != MCB200_OK) return kNaN;
    mcb200_ctx* c = ref.get();
    float outv[2] = {kNaN, kNaN};
    int rc = mcb200_price_batch(c, 1, &spot, &strike, &vol, &rate, &years, &div, steps, num_trials, flags, &outv[0],
                                &outv[1], nullptr, nullptr);
    return rc == MCB200_OK ? outv[which] : kNaN;
}

static float dropin_greek(float spot, float strike, float vol, float rate, float years, float div, float steps,
                          float num_trials, mcb200_bumps bumps, int which) {
    std::shared_ptr<mcb200_ctx> ref;
    if (default_ctx_ref(&ref) != MCB200_OK) return kNaN;
    mcb200_ctx* c = ref.get();
    float vals[MCB200_N_GREEK_OUTPUTS];
    float* out[MCB200_N_GREEK_OUTPUTS];
    for (int e = 0; e < MCB200_N_GREEK_OUTPUTS; ++e) { vals[e] = kNaN; out[e] = (e == which) ? &vals[e] : nullptr; }
    int rc = mcb200_greeks_batch(c, 1, &spot, &strike, &vol, &rate, &years, &div, steps, num_trials, &bumps, out, nullptr);
    return rc == MCB200_OK ? vals[which] : kNaN;
}

// unused bumps are 0 (their legs collapse onto the base leg and stay finite); only the requested estimator is read back
extern "C" float mc_simd_call_price(float s, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_price(s, k, v, r, t, q, steps, n, 0u, 0);
}
extern "C" float mc_simd_put_price(float s, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_price(s, k, v, r, t, q, steps, n, 0u, 1);
}
extern "C" float mc_simd_call_price_av(float s, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_price(s, k, v, r, t, q, steps, n, MCB200_ANTITHETIC, 0);
}
extern "C" float mc_simd_put_price_av(float s, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_price(s, k, v, r, t, q, steps, n, MCB200_ANTITHETIC, 1);
}
extern "C" float mc_simd_call_delta(float s, float ds, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{ds, 0.0f, 0.0f, 0.0f}, EST_CALL_DELTA);
}
extern "C" float mc_simd_put_delta(float s, float ds, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{ds, 0.0f, 0.0f, 0.0f}, EST_PUT_DELTA);
}
extern "C" float mc_simd_gamma(float s, float ds, float k, float v, float r, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{ds, 0.0f, 0.0f, 0.0f}, EST_GAMMA);
}
extern "C" float mc_simd_vega(float s, float k, float v, float dv, float r, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{0.0f, dv, 0.0f, 0.0f}, EST_VEGA);
}
extern "C" float mc_simd_call_rho(float s, float k, float v, float r, float dr, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{0.0f, 0.0f, dr, 0.0f}, EST_CALL_RHO);
}
extern "C" float mc_simd_put_rho(float s, float k, float v, float r, float dr, float t, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{0.0f, 0.0f, dr, 0.0f}, EST_PUT_RHO);
}
extern "C" float mc_simd_call_theta(float s, float k, float v, float r, float t, float dt, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{0.0f, 0.0f, 0.0f, dt}, EST_CALL_THETA);
}
extern "C" float mc_simd_put_theta(float s, float k, float v, float r, float t, float dt, float q, float steps, float n) {
    return dropin_greek(s, k, v, r, t, q, steps, n, mcb200_bumps{0.0f, 0.0f, 0.0f, dt}, EST_PUT_THETA);
}
