// mcb200_multi.cu -- the multi-device context: one process, several GPUs behind the host-buffer entry points.
//
// The reference's only parallelism lives inside a call: rayon fans 8-path chunks out over the cores
// (/root/reference/src/mc_simd.rs:49-51) and reduces the partial sums (mc_simd.rs:71-74).  On a B200 node the same two
// steps exist one level up: a batch fans out over the GPUs of the box and the per-GPU results are gathered (options-major)
// or the per-GPU moment sums added (trials-minor).  This file is that level: one per-device context, one host worker
// thread and one stream per listed device; the parent context owns no device memory.  The split policy is
// mcb200_shard_query (mcb200_plan.cpp); the sum of a trials-minor batch runs on the host in device order, in fp64, so
// a given (seed, device count) reproduces its results bit for bit.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/mcb200.h"
#include "mcb200_internal.h"
#include "mcb200_plan.h"

namespace mcb200 {

// One worker = one host thread bound to one device context.  The parent posts a task and waits; tasks of different
// workers run concurrently (H2D copy, launch and D2H copy of every GPU overlap with the others').
struct Worker {
    mcb200_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = true, stop = false;
    int rc = MCB200_OK;
    char err[512] = "";

    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_task || stop; });
            if (stop) return;
            std::function<int()> t = std::move(task);
            has_task = false;
            lk.unlock();
            const int r = t();
            const char* msg = r != MCB200_OK ? mcb200_last_error() : "";     // thread-local to THIS worker
            lk.lock();
            rc = r;
            std::strncpy(err, msg, sizeof(err) - 1);
            done = true;
            cv.notify_all();
        }
    }
    void post(std::function<int()> t) {
        std::lock_guard<std::mutex> lk(mu);
        task = std::move(t); has_task = true; done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

struct MultiRuntime {
    std::vector<Worker*> workers;
    std::vector<double> moments;          // [device][n][n_acc] scratch of the trials-minor split
    std::vector<double> total;            // [n][n_acc]
    int grid_slots = 0, grid_slots_greeks = 0;
    std::mutex mu;                        // one batch at a time per multi-device context
};

void multi_destroy(MultiRuntime* m) {
    if (!m) return;
    for (Worker* w : m->workers) {
        if (w->th.joinable()) {
            { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; w->cv.notify_all(); }
            w->th.join();
        }
        if (w->ctx) mcb200_ctx_destroy(w->ctx);
        delete w;
    }
    delete m;
}

int multi_device_count(MultiRuntime* m) { return (int)m->workers.size(); }
mcb200_ctx* multi_child(MultiRuntime* m, int i) { return m->workers[(size_t)i]->ctx; }

uint64_t multi_launches(MultiRuntime* m) {
    uint64_t total = 0;
    for (Worker* w : m->workers) {
        mcb200_info inf;
        if (mcb200_ctx_info(w->ctx, &inf) == MCB200_OK) total += inf.kernel_launches;
    }
    return total;
}

// Runs fn(device index) on every worker whose `active` flag is set; returns the first failure in device order.
static int run_on_all(MultiRuntime* m, const std::vector<char>& active, const std::function<int(int)>& fn) {
    const int D = (int)m->workers.size();
    for (int d = 0; d < D; ++d)
        if (active[(size_t)d]) m->workers[(size_t)d]->post([fn, d] { return fn(d); });
    int rc = MCB200_OK;
    for (int d = 0; d < D; ++d) {
        if (!active[(size_t)d]) continue;
        const int r = m->workers[(size_t)d]->wait();
        if (r != MCB200_OK && rc == MCB200_OK) rc = set_error(r, "device %d: %s", d, m->workers[(size_t)d]->err);
    }
    return rc;
}

int multi_set_seed(MultiRuntime* m, uint64_t seed, uint32_t first_block) {
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<char> all(m->workers.size(), 1);
    return run_on_all(m, all, [m, seed, first_block](int d) {
        return mcb200_ctx_set_seed(m->workers[(size_t)d]->ctx, seed, first_block + (uint32_t)d);
    });
}

static float* const* offset_rows(float* const* rows, int n_rows, int lo, float** store) {
    if (!rows) return nullptr;
    for (int e = 0; e < n_rows; ++e) store[e] = rows[e] ? rows[e] + lo : nullptr;
    return store;
}

int multi_run_host(MultiRuntime* m, int n, const float* const in[6], float steps, float num_trials, int greeks,
                   unsigned flags, const mcb200_bumps* bumps, float* const* out, float* const* out_se) {
    std::lock_guard<std::mutex> lk(m->mu);
    const int D = (int)m->workers.size();
    const int n_est = greeks ? MCB200_N_GREEK_OUTPUTS : 2;
    const int n_acc = 2 * n_est;
    const int slots = greeks ? m->grid_slots_greeks : m->grid_slots;
    std::vector<mcb200_shard> sh((size_t)D);
    std::vector<char> active((size_t)D, 0);
    for (int d = 0; d < D; ++d) {
        const int rc = make_shard(n, steps, num_trials, D, d, slots, &sh[(size_t)d]);
        if (rc != MCB200_OK) return set_error(rc, "invalid request: n=%d steps=%g num_trials=%g", n, steps, num_trials);
        active[(size_t)d] = (char)sh[(size_t)d].active;
    }
    if (n == 0) return MCB200_OK;
    const mcb200_bumps b = bumps ? *bumps : mcb200_bumps{1.0f, 1.0f, 1.0f, 1.0f};

    // One device's share through the ordinary single-device host entry points, on option block [lo, hi).
    auto block_call = [&, n_est](int d, int lo, int hi, float trials) -> int {
        mcb200_ctx* c = m->workers[(size_t)d]->ctx;
        float* o_store[MCB200_N_GREEK_OUTPUTS]; float* s_store[MCB200_N_GREEK_OUTPUTS];
        float* const* o = offset_rows(out, n_est, lo, o_store);
        float* const* s = offset_rows(out_se, n_est, lo, s_store);
        if (greeks)
            return mcb200_greeks_batch(c, hi - lo, in[0] + lo, in[1] + lo, in[2] + lo, in[3] + lo, in[4] + lo, in[5] + lo,
                                       steps, trials, &b, o, s);
        return mcb200_price_batch(c, hi - lo, in[0] + lo, in[1] + lo, in[2] + lo, in[3] + lo, in[4] + lo, in[5] + lo, steps,
                                  trials, flags, o ? o[0] : nullptr, o ? o[1] : nullptr, s ? s[0] : nullptr,
                                  s ? s[1] : nullptr);
    };

    const int mode = sh[0].mode;
    if (mode == MCB200_SHARD_SINGLE) {       // also the zero-path corner (num_trials < 8): zeros, like one device
        std::vector<char> only0((size_t)D, 0);
        only0[0] = 1;
        return run_on_all(m, only0, [&](int d) { return block_call(d, 0, n, num_trials); });
    }
    if (mode == MCB200_SHARD_OPTIONS)
        return run_on_all(m, active, [&](int d) {
            return block_call(d, sh[(size_t)d].option_begin, sh[(size_t)d].option_end, num_trials);
        });

    // trials-minor: per-device moment sums -> host sum in device order -> one finalisation (device 0's arithmetic)
    const size_t per_dev = (size_t)n * (size_t)n_acc;
    m->moments.assign(per_dev * (size_t)D, 0.0);
    int rc = run_on_all(m, active, [&](int d) {
        return host_moments(m->workers[(size_t)d]->ctx, n, in, steps, sh[(size_t)d].num_trials_local, greeks, flags, &b,
                            m->moments.data() + per_dev * (size_t)d);
    });
    if (rc != MCB200_OK) return rc;
    m->total.assign(per_dev, 0.0);
    for (int d = 0; d < D; ++d) {            // fixed order: device 0, 1, 2, ...
        if (!active[(size_t)d]) continue;
        const double* src = m->moments.data() + per_dev * (size_t)d;
        for (size_t i = 0; i < per_dev; ++i) m->total[i] += src[i];
    }
    std::vector<char> only0((size_t)D, 0);
    only0[0] = 1;
    return run_on_all(m, only0, [&](int) {
        return host_finalize(m->workers[0]->ctx, n, greeks, in[3], in[4], m->total.data(), sh[0].total_paths,
                             (double)num_trials, &b, out, out_se);
    });
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_ctx_create_multi(mcb200_ctx** out, const int* devices, int n_devices, uint64_t seed,
                                       uint32_t first_substream_block) {
    if (!out) return set_error(MCB200_ERR_INVALID, "mcb200_ctx_create_multi: null output pointer");
    *out = nullptr;
    if (!devices || n_devices <= 0 || n_devices > 64)
        return set_error(MCB200_ERR_INVALID, "mcb200_ctx_create_multi: need 1..64 devices");
    MultiRuntime* m = new (std::nothrow) MultiRuntime();
    if (!m) return set_error(MCB200_ERR_NOMEM, "out of host memory");
    for (int d = 0; d < n_devices; ++d) {
        Worker* w = new (std::nothrow) Worker();
        if (!w) { multi_destroy(m); return set_error(MCB200_ERR_NOMEM, "out of host memory"); }
        m->workers.push_back(w);
        const int rc = mcb200_ctx_create(&w->ctx, devices[d], seed, first_substream_block + (uint32_t)d);
        if (rc != MCB200_OK) { multi_destroy(m); return rc; }        // message already recorded on this thread
        w->th = std::thread([w] { w->loop(); });
    }
    mcb200_info inf;
    mcb200_ctx_info(m->workers[0]->ctx, &inf);
    m->grid_slots = inf.grid_slots; m->grid_slots_greeks = inf.grid_slots_greeks;
    for (Worker* w : m->workers) {           // the split policy assumes equal devices
        mcb200_info other;
        mcb200_ctx_info(w->ctx, &other);
        if (other.grid_slots != inf.grid_slots || other.grid_slots_greeks != inf.grid_slots_greeks) {
            multi_destroy(m);
            return set_error(MCB200_ERR_INVALID, "mcb200_ctx_create_multi: the devices differ in SM count");
        }
    }
    mcb200_ctx* parent = ctx_new_parent(m, devices[0], seed, first_substream_block);
    if (!parent) { multi_destroy(m); return set_error(MCB200_ERR_NOMEM, "out of host memory"); }
    *out = parent;
    return MCB200_OK;
}

extern "C" int mcb200_ctx_device_count(mcb200_ctx* c) {
    if (!c) return 0;
    MultiRuntime* m = ctx_multi(c);
    return m ? multi_device_count(m) : 1;
}

extern "C" int mcb200_ctx_child(mcb200_ctx* c, int i, mcb200_ctx** child) {
    if (!c || !child) return set_error(MCB200_ERR_INVALID, "null argument");
    MultiRuntime* m = ctx_multi(c);
    if (!m) {
        if (i != 0) return set_error(MCB200_ERR_INVALID, "a single-device context has one child: itself");
        *child = c;
        return MCB200_OK;
    }
    if (i < 0 || i >= multi_device_count(m)) return set_error(MCB200_ERR_INVALID, "device index %d out of range", i);
    *child = multi_child(m, i);
    return MCB200_OK;
}
