#!/usr/bin/env python3
"""Where does a small batch spend its time?  Per-warp %globaltimer stamps inside simulate_kernel (an -DMCB_TRACE build of
the library under an alternate name, so the production library is untouched) plus host-side CUDA events.

    python tools/kernel_timeline.py [c1 c2 c3 ...]      (GPU box; rebuilds lib/libmcb200_trace.so with nvcc)

Stamps per logical warp (lane 0): 0 kernel entry, 1 generator state + plan ready, 2 end of the LAST run of rounds,
3 end of the last flush (slot, fence, arrival, possibly the option's finalisation), 4 after the state write-back.
Reports, relative to the earliest stamp of the launch: percentiles of each stamp, the event-timed duration of the
same launch and the launch-to-launch period of a back-to-back queue.
"""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MCB200_LIB_NAME"] = "libmcb200_trace.so"
os.environ["MCB200_NVCC_FLAGS"] = (os.environ.get("MCB200_NVCC_FLAGS", "") + " -DMCB_TRACE").strip()

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2", "c3"]
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    pkg.build(force=True)
    from bench import workload, columns, path_steps
    lib = pkg._native.lib()
    lib.mcb200_debug_set_trace.argtypes = [C.c_void_p, C.c_void_p]
    dev = torch.device("cuda", 0)
    ctx = pkg.Context(device=0, seed=7)
    info = ctx.info()
    out = {"sm_count": info["sm_count"], "workloads": {}}
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    lib.mcb200_debug_set_trace(ctx._h, None)
    # ---- floors: what an (almost) empty launch costs between two events on this box, for a kernel with ~100 bytes of
    #      parameters (closed-form kernel, 1 option) and for simulate_kernel itself with next to no work (1 option x 8 trials)
    def event_us(fn, reps=50):
        iso = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            iso.append(1e3 * a.elapsed_time(b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            fn()
        b.record()
        torch.cuda.synchronize()
        period = 1e3 * a.elapsed_time(b) / 200
        # primed: a ~40 us kernel (256 MiB fill) runs while the host enqueues event, launch, event -- the interval is
        # then pure GPU-side time, the way bench.py's timed steps see it
        primed = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            big.fill_(1)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            primed.append(1e3 * a.elapsed_time(b))
        two = []
        for _ in range(reps):                                        # two launches behind the primer: 2nd - 1st = period on the GPU
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            big.fill_(1); big.fill_(1)
            a.record(); fn(); b.record(); fn(); c.record()
            torch.cuda.synchronize()
            two.append(1e3 * b.elapsed_time(c))
        return {"isolated_us_median": float(np.median(iso)), "isolated_us_min": float(np.min(iso)),
                "host_bound_back_to_back_period_us": period, "primed_gpu_side_us_median": float(np.median(primed)),
                "primed_gpu_side_us_min": float(np.min(primed)), "primed_second_launch_us_median": float(np.median(two))}
    big = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    one = [torch.full((1,), v, dtype=torch.float32, device=dev) for v in (100.0, 110.0, 0.25, 0.05, 0.5, 0.02)]
    o10 = torch.zeros(10, 1, device=dev)
    o4 = torch.zeros(4, 1, device=dev)
    lib.mcb200_bs_batch_device.argtypes = None
    N = pkg._native
    def bs_launch():
        N.check(lib.mcb200_bs_batch_device(ctx._h, 1, *[C.c_void_p(c.data_ptr()) for c in one],
                                           N._VoidOut(*[o10[i].data_ptr() for i in range(10)]), C.c_void_p(st.cuda_stream)))
    def tiny_sim():
        ctx.price_batch_device(1, [c.data_ptr() for c in one], 2.0, 8.0, [o4[i].data_ptr() for i in range(4)], stream=st.cuda_stream)
    def small_sim():
        ctx.price_batch_device(1, [c.data_ptr() for c in one], 100.0, 1000.0, [o4[i].data_ptr() for i in range(4)], stream=st.cuda_stream)
    out["floors"] = {"closed_form_kernel_1_option": event_us(bs_launch), "simulate_kernel_1_option_8_trials_2_steps": event_us(tiny_sim),
                     "simulate_kernel_1_option_1000_trials_100_steps": event_us(small_sim)}
    c1w = workload("c1")
    c1cols = [torch.from_numpy(c).to(dev) for c in columns(c1w)]
    c1out = torch.zeros(4, 112, device=dev)
    def c1_sim():
        ctx.price_batch_device(112, [c.data_ptr() for c in c1cols], 100.0, 1000.0, [c1out[i].data_ptr() for i in range(4)], stream=st.cuda_stream)
    out["floors"]["simulate_kernel_c1"] = event_us(c1_sim)
    print("floors", json.dumps(out["floors"]), flush=True)
    for name in names:
        w = workload(name)
        cols = [torch.from_numpy(c).to(dev) for c in columns(w)]
        n = len(w["spots"])
        greeks = w["kind"] == "greeks"
        res = torch.zeros(20 if greeks else 4, n, device=dev)
        plan = ctx.plan(n, w["steps"], w["trials"], greeks=greeks)
        trace = torch.zeros(plan["warps"] * 8, dtype=torch.int64, device=dev)
        ptrs = [c.data_ptr() for c in cols]

        def launch():
            if greeks:
                ctx.greeks_batch_device(n, ptrs, w["steps"], w["trials"], (0.01, 0.01, 0.01, 0.001),
                                        [res[i].data_ptr() for i in range(10)], [res[10 + i].data_ptr() for i in range(10)],
                                        stream=st.cuda_stream)
            else:
                ctx.price_batch_device(n, ptrs, w["steps"], w["trials"], [res[i].data_ptr() for i in range(4)],
                                       antithetic=w["av"], stream=st.cuda_stream)
        lib.mcb200_debug_set_trace(ctx._h, None)
        for _ in range(20):
            launch()
        torch.cuda.synchronize()
        # back-to-back period and isolated event time WITHOUT stamps being written
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200 if path_steps(w) < 1e10 else 5
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        period_us = 1e3 * e0.elapsed_time(e1) / reps
        iso = []
        for _ in range(20 if reps > 5 else 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); launch(); b.record()
            torch.cuda.synchronize()
            iso.append(1e3 * a.elapsed_time(b))
        # one traced launch
        lib.mcb200_debug_set_trace(ctx._h, C.c_void_p(trace.data_ptr()))
        launch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        trace.zero_()
        torch.cuda.synchronize()
        a.record(); launch(); b.record()
        torch.cuda.synchronize()
        traced_us = 1e3 * a.elapsed_time(b)
        lib.mcb200_debug_set_trace(ctx._h, None)
        raw = trace.cpu().numpy().reshape(-1, 8)
        np.save(os.path.join(ROOT, "gpurun_out", "timeline_raw_%s.npy" % name), raw)
        t = raw[:, :5].astype(np.float64)
        t0 = t[:, 0].min()
        rel = (t - t0) * 1e-3                                           # us
        pct = lambda x: [float(np.percentile(x, q)) for q in (0, 50, 90, 99, 100)]
        rec = {"plan": {k: plan[k] for k in ("warps", "grid", "rounds_per_option", "total_rounds", "max_rounds_per_warp")},
               "back_to_back_period_us": period_us, "isolated_event_us_median": float(np.median(iso)),
               "traced_launch_event_us": traced_us, "stamps_us_min_p50_p90_p99_max": {
                   "0_entry": pct(rel[:, 0]), "1_state_loaded": pct(rel[:, 1]), "2_paths_done": pct(rel[:, 2]),
                   "3_flushed": pct(rel[:, 3]), "4_state_stored": pct(rel[:, 4])},
               "per_warp_us_p50": {"load": float(np.median(rel[:, 1] - rel[:, 0])),
                                   "paths": float(np.median(rel[:, 2] - rel[:, 1])),
                                   "last_flush": float(np.median(rel[:, 3] - rel[:, 2])),
                                   "store": float(np.median(rel[:, 4] - rel[:, 3]))},
               "last_flush_us_max": float((rel[:, 3] - rel[:, 2]).max())}
        out["workloads"][name] = rec
        print(name, json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_timeline.json"), "w") as f:
        json.dump(out, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
