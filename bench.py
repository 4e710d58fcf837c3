#!/usr/bin/env python3
"""bench.py -- GBM path-steps/s of the fused Monte Carlo path kernel on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: every option of the workload priced from fresh paths.
Default workload = BASELINE.json configs[1] (C2): mc_simd call_price + put_price, 112 spots x 20000 trials x 100 steps
(benches/benchmark.rs:10-16 grid).  Rank r prices its own copy of the batch on substream block r (weak scaling,
no data-path collective).  One JSON line is printed by rank 0.

  value     device-resident: inputs already in HBM, K steps each timed with CUDA events on the launch stream
            (an L2 flush between steps, outside the events), max over ranks.
  e2e       the same metric through the public host-buffer call (mcb200_price_batch): host arrays -> pinned staging ->
            H2D -> kernels -> D2H -> host arrays inside the timed region, wall clock around K calls.
  roofline  the fused path kernel alone (events around its launch inside the library), against the issue/MUFU bound
            of SURVEY.md section 8d: 8 path-steps per clock per SM at the max SM clock of MEASURED_PEAKS.json.
  cpu_baseline  the oracle's SIMD+threads port of the reference algorithm on the box's host cores (bounded sample).
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PACKAGE = "monte-carlo-options-simd_b200"

GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)            # benches/benchmark.rs:10-16
SEED = 0x5EED5EED5EED5EED
# The reference publishes wall times, not path-steps/s (README.md:107,122: 112 x 1000 x 100 in ~15 ms and 112 x 20000 x 100
# in ~207 ms on an M1).  BASELINE.json.published is empty, so vs_baseline is null; the rate DERIVED from those times is
# reported separately as config.vs_readme_m1_derived.
README_M1_DERIVED = {"c2": 112 * 20000 * 100 / 0.207, "c1": 112 * 1000 * 100 / 0.015}
ROOFLINE_PATH_STEPS_PER_CLK_PER_SM = 8.0                              # SURVEY.md section 8d contract figure


def workload(name):
    spots112 = np.arange(50, 162, dtype=np.float32)
    if name == "c1":
        return dict(name="C1: mc_simd call_price 112 spots x 1000 trials x 100 steps", spots=spots112, strikes=None,
                    steps=100.0, trials=1000.0, kind="price", av=False)
    if name == "c2":
        return dict(name="C2: mc_simd call_price+put_price 112 spots x 20000 trials x 100 steps", spots=spots112,
                    strikes=None, steps=100.0, trials=20000.0, kind="price", av=False)
    if name == "c3":
        return dict(name="C3: mc_simd call_price_av+put_price_av 112 spots x 2000 trials x 1000 steps", spots=spots112,
                    strikes=None, steps=1000.0, trials=2000.0, kind="price", av=True)
    if name == "c4":
        return dict(name="C4: fused CRN Greeks 4096 spots x 1e5 trials x 252 steps",
                    spots=(50.0 + 112.0 * np.arange(4096) / 4096.0).astype(np.float32), strikes=None, steps=252.0,
                    trials=1e5, kind="greeks", av=False)
    if name.startswith("c5"):
        # full C5 is 1e5 pairs x 1e6 trials x 252 steps (2.52e13 path-steps, ~10 s per GPU-step); c5r keeps the
        # 1e5-pair grid with 1e4 trials (SURVEY.md section 8d allows the reduced-trial variant; stated in the name)
        trials = 1e6 if name == "c5" else 1e4
        sp = (50.0 + 112.0 * np.arange(400) / 400.0).astype(np.float32)
        st = (80.0 + 60.0 * np.arange(250) / 250.0).astype(np.float32)
        return dict(name="%s: batched sweep 1e5 spot/strike pairs x %g trials x 252 steps" % (name.upper(), trials),
                    spots=np.repeat(sp, 250), strikes=np.tile(st, 400), steps=252.0, trials=trials, kind="price",
                    av=False)
    raise SystemExit("unknown workload %r" % name)


def columns(w):
    n = len(w["spots"])
    strikes = w["strikes"] if w["strikes"] is not None else np.full(n, GRID["strike"], np.float32)
    return [w["spots"], strikes] + [np.full(n, GRID[k], np.float32) for k in ("vol", "r", "t", "q")]


def path_steps(w):
    return float(len(w["spots"])) * (8 * (int(w["trials"]) // 8)) * (2 * (int(w["steps"]) // 2))


# ---------------------------------------------------------------------------------------------- clocks (NVML)
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every ~20 ms on a thread while armed."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop, self._armed, self._thread = threading.Event(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:                       # NVML missing: report nulls rather than fail the bench
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop.is_set():
            if self._armed.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, nm in names.items():
                        if mask & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(0.02)

    def arm(self, on=True):
        (self._armed.set if on else self._armed.clear)()

    def summary(self):
        self._stop.set()
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the path kernel, from the committed ncu --set full
    capture of this workload (profiles/r1_ncu_<workload>_v3_summary.json); None when no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_%s_v3_summary.json" % workload)) as f:
            m = json.load(f)["metrics"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_sample(w, budget_s):
    """Bounded sample of the workload for the CPU port: a prefix of the option list sized from a calibration call."""
    from oracle import oracle
    cols = columns(w)
    kind = 2 if w["av"] else 0
    m = min(len(cols[0]), 4)
    t_warm = time.perf_counter()                 # the first second of calls pays for thread creation and page-in
    while time.perf_counter() - t_warm < 1.5:
        oracle.baseline_price_loop(*[c[:m] for c in cols], w["steps"], w["trials"], kind)
    per_opt = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        oracle.baseline_price_loop(*[c[:m] for c in cols], w["steps"], w["trials"], kind)
        per_opt = min(per_opt, (time.perf_counter() - t0) / m)
    n = int(max(1, min(len(cols[0]), budget_s / max(per_opt, 1e-9))))
    return [c[:n] for c in cols], kind, n


def run_cpu_step(cols, w, kind):
    from oracle import oracle
    t0 = time.perf_counter()
    _, ps = oracle.baseline_price_loop(*cols, w["steps"], w["trials"], kind)
    if w["kind"] == "price" and not w["av"]:
        # the workload prices call AND put; the reference needs a second simulation for the put (mc_simd.rs:500-521)
        _, ps2 = oracle.baseline_price_loop(*cols, w["steps"], w["trials"], kind + 1)
        ps += ps2                      # every simulated path-step is counted: the CPU arm's rate is not halved
    return time.perf_counter() - t0, ps


def reference_arm(args, w):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference is a Rust
    crate that cannot be built in this image, so this is the oracle port (kind "port"), all host threads."""
    from oracle import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = max(1, args.steps + args.warmup)
    cols, kind, n = cpu_sample(w, budget_s=min(2.0, 150.0 / total))
    for _ in range(args.warmup):
        run_cpu_step(cols, w, kind)
    t, ps = 0.0, 0.0
    for _ in range(args.steps):
        dt, p = run_cpu_step(cols, w, kind)
        t += dt; ps += p
    value = ps / t
    sample = "%d of %d options per step (prefix of the grid), call and put each simulated like the reference" % (
        n, len(w["spots"]))
    line = {"impl": "reference", "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "note": "every simulated path-step is counted; the reference simulates "
                                                      "call and put separately (mc_simd.rs:477-521), the B200 arm "
                                                      "prices both from one pass"},
            "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": oracle.baseline_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- the B200 arm
def b200_arm(args, w):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on STDOUT at the first collective when NCCL_DEBUG is set; stdout must carry the
        # one JSON line only, so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pkg = importlib.import_module(PACKAGE)
    ctx = pkg.Context(device=local_rank, seed=SEED, substream_block=rank)
    info0 = ctx.info()

    cols_h = columns(w)
    n = len(cols_h[0])
    cols_d = [torch.from_numpy(c).to(dev) for c in cols_h]
    in_ptrs = [c.data_ptr() for c in cols_d]
    greeks = w["kind"] == "greeks"
    n_out = 20 if greeks else 4
    out_d = torch.zeros(n_out, n, dtype=torch.float32, device=dev)
    bumps = (0.01, 0.01, 0.01, 0.001)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2
    # all timed work and its events go on ONE explicit stream (a NULL stream would mean the context's private stream)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device():
        if greeks:
            ctx.greeks_batch_device(n, in_ptrs, w["steps"], w["trials"], bumps, [out_d[i].data_ptr() for i in range(10)],
                                    [out_d[10 + i].data_ptr() for i in range(10)], stream=stream)
        else:
            ctx.price_batch_device(n, in_ptrs, w["steps"], w["trials"], [out_d[i].data_ptr() for i in range(4)],
                                   antithetic=w["av"], stream=stream)

    # e2e: one public C-ABI call per step on HOST buffers (mcb200_price_batch / mcb200_greeks_batch); the ctypes
    # marshalling of the pointers is done once, the call itself copies in, launches, copies out and synchronizes
    prepared = (ctx.prepare_greeks_batch(*cols_h, w["steps"], w["trials"], *bumps) if greeks else
                ctx.prepare_price_batch(*cols_h, w["steps"], w["trials"], antithetic=w["av"]))

    def step_host():
        return prepared.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- value: device-resident, per-step CUDA events, L2 flushed between steps
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.info()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if sampler:
        sampler.arm(True)
    for e0, e1 in ev:
        flush.fill_(1)
        e0.record()
        step_device()
        e1.record()
    barrier()
    if sampler:
        sampler.arm(False)
    launches = ctx.info()["kernel_launches"] - launches0
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ps_step = path_steps(w)
    value = world * ps_step * args.steps / (ms_total * 1e-3)

    # ---- roofline: the fused path kernel alone, timed inside the library on its launch stream; >= ~1 s under load
    ctx.set_timing(True)
    sim_ms = []
    t_end = time.perf_counter() + args.roof_seconds
    if sampler:
        sampler.arm(True)
    while len(sim_ms) < args.steps or (time.perf_counter() < t_end and len(sim_ms) < 20000):
        step_device()
        sim_ms.append(ctx.last_sim_ms())
    if sampler:
        sampler.arm(False)
    ctx.set_timing(False)
    sim_avg_ms = float(np.mean(sim_ms))
    peaks = measured_peaks()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    achieved = ps_step / (sim_avg_ms * 1e-3)
    peak = info0["sm_count"] * sm_max_mhz * 1e6 * ROOFLINE_PATH_STEPS_PER_CLK_PER_SM

    # ---- e2e: public host-buffer call, wall clock (each call ends with a stream synchronize)
    for _ in range(args.warmup):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_host()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ps_step * args.steps / float(te.item())
    n_est = 10 if greeks else 2
    h2d, d2h = 6 * n * 4, 2 * n_est * n * 4

    if rank == 0:
        clocks = sampler.summary()
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            cols, kind, n_cpu = cpu_sample(w, budget_s=2.0)
            t_cpu, ps_cpu, reps = 0.0, 0.0, 0
            while t_cpu < args.cpu_seconds:
                dt, p = run_cpu_step(cols, w, kind)
                t_cpu += dt; ps_cpu += p; reps += 1
            cpu = {"value": ps_cpu / t_cpu, "unit": "path-steps/s", "cores": oracle.baseline_threads(), "kind": "port",
                   "sample": "%d of %d options x %d repeats (%.1f s), call and put each simulated like the reference"
                             % (n_cpu, n, reps, t_cpu)}
        frac = achieved / peak
        line = {
            "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "options_per_gpu": n, "trials": w["trials"], "steps_per_path": w["steps"],
                       "outputs": "all Greeks + prices + SE" if greeks else "call + put + SE from one pass",
                       "l2": "flushed between timed steps (256 MiB write, outside the events); kernel reads %d B" % h2d,
                       "parallelism": "options-major, one process per GPU, substream block = rank, no data-path collective",
                       "launch": "one fused kernel per step: paths + payoffs + deterministic reduce + finalize",
                       "vs_readme_m1_derived": (value / README_M1_DERIVED[args.workload]
                                                if args.workload in README_M1_DERIVED else None),
                       "readme_m1_note": "the reference publishes wall times on an M1 (README.md:107,122); "
                                         "path-steps/s derived from them, call_price only"},
            "options_per_sec": world * n * args.steps / (ms_total * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(te.item()) / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "issue/mufu (not hbm, not tensor: SURVEY 8d)", "achieved": achieved, "peak": peak,
                         "unit": "path-steps/s", "frac": frac, "traffic": ncu_traffic(args.workload),
                         "traffic_note": "DRAM bytes per launch (ncu): the per-thread generator states, read once; "
                                         "path data never leaves registers",
                         "inner_step_ceiling_frac": 3.95 * 2.0 / ROOFLINE_PATH_STEPS_PER_CLK_PER_SM,
                         "inner_step_ceiling_note": "the inner step (quad_fast: 3 xoshiro256++ outputs -> 4 Box-Muller "
                                                    "pairs) alone sustains 3.95 pairs/clk/SM on B200 with the ALU pipe "
                                                    "88% busy (profiles/r1_microbench_v2_pipes.json): 99% of this peak",
                         "kernel": "simulate_kernel", "kernel_ms": sim_avg_ms, "launches_timed": len(sim_ms),
                         "peak_basis": "%d SMs x %.0f MHz (MEASURED_PEAKS.json sm_max_mhz%s) x 8 path-steps/clk/SM"
                                       % (info0["sm_count"], sm_max_mhz, "" if peaks else ", fallback"),
                         "frac_at_measured_clock": (achieved / (info0["sm_count"] * clocks["sm_mhz"] * 1e6 * 8.0))
                         if clocks.get("sm_mhz") else None},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--roof-seconds", type=float, default=1.0,
                    help="minimum duration of the kernel-only roofline pass (clock sampling needs ~1 s)")
    args = ap.parse_args()
    w = workload(args.workload)
    if args.impl == "reference":
        reference_arm(args, w)
    else:
        b200_arm(args, w)


if __name__ == "__main__":
    main()
