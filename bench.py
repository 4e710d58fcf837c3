#!/usr/bin/env python3
"""bench.py -- GBM path-steps/s of the fused Monte Carlo path kernels on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5r] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: every option of the workload priced from fresh paths.
Default workload = BASELINE.json configs[4] with the reduced trial count SURVEY.md section 8d allows for a single GPU
(C5R: the 1e5 spot/strike pairs of the sweep x 1e4 trials x 252 steps; the full 1e6-trial sweep is 10.6 s per step and
is timed once as a sub-record).  One JSON line is printed by rank 0.

  N = 1     value / roofline / e2e on the default workload, plus `sub` records (their own roofline each) for C1, C2, C3
            (price kernels on the reference's own 112-spot grid), C4 (the fused-Greeks kernel), the reference-stream
            parity kernel on C2 and one full-size C5 step (`sub.c5_full`, also reported at every N > 1).
  N > 1     STRONG scaling of the same one batch: the library's shard policy (mcb200_shard_query) gives every rank
            its options-major block (or trials-minor range); every timed step ends with the exchange the split needs
            -- results gathered to rank 0 over NCCL (options-major), or the moment sums all-reduced and finalised
            (trials-minor) -- inside the timed region.
  value     inputs already in HBM, K steps each timed with CUDA events on the launch stream (an L2 flush between steps,
            outside the events), max over ranks.
  e2e       N = 1: the public host-buffer C-ABI call (mcb200_price_batch / mcb200_greeks_batch): host arrays -> pinned
            staging -> H2D -> kernel -> D2H -> host arrays, wall clock.  N > 1: pinned host shard -> H2D -> the
            device-pointer C-ABI call -> NCCL gather -> D2H on rank 0, wall clock, max over ranks.
  roofline  the fused path kernel alone (events around its launch inside the library), against the issue/MUFU bound of
            SURVEY.md section 8d (8 path-steps per clock per SM at MEASURED_PEAKS.json's max SM clock), and against the
            ceilings of the pipes the loop's own SASS occupies (read from the loaded library with cuobjdump).
  cpu_baseline  the oracle's SIMD+threads port of the reference algorithm on the box's host cores (bounded sample).
"""
import argparse
import importlib
import json
import os
import re
import shutil
import subprocess
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
BUMPS = (0.01, 0.01, 0.01, 0.001)                                      # the reference tests' bump sizes (mc_simd.rs:840-904)
# The reference publishes wall times, not path-steps/s (README.md:107,122: 112 x 1000 x 100 in ~15 ms and 112 x 20000 x 100
# in ~207 ms on an M1).  BASELINE.json.published is empty, so vs_baseline is null; rates DERIVED from those times are
# reported in the sub-records of the same shapes.
README_M1_DERIVED = {"c2": 112 * 20000 * 100 / 0.207, "c1": 112 * 1000 * 100 / 0.015}
ROOFLINE_PATH_STEPS_PER_CLK_PER_SM = 8.0                              # SURVEY.md section 8d contract figure
# measured pipe rates of this GPU family, thread-instructions per clock per SM (profiles/r1_microbench_v2_pipes.json,
# re-measured in profiles/r2_microbench_tri_vs_quad.json): integer ALU pipe, XU (MUFU) pipe, instruction issue
PIPE_RATE = {"alu": 64.0, "xu": 16.0, "issue": 128.0}


def workload(name):
    spots112 = np.arange(50, 162, dtype=np.float32)
    if name == "c1":
        return dict(key="c1", name="C1: mc_simd call_price 112 spots x 1000 trials x 100 steps", spots=spots112,
                    strikes=None, steps=100.0, trials=1000.0, kind="price", av=False)
    if name == "c2":
        return dict(key="c2", name="C2: mc_simd call_price+put_price 112 spots x 20000 trials x 100 steps", spots=spots112,
                    strikes=None, steps=100.0, trials=20000.0, kind="price", av=False)
    if name == "c3":
        return dict(key="c3", name="C3: mc_simd call_price_av+put_price_av 112 spots x 2000 trials x 1000 steps",
                    spots=spots112, strikes=None, steps=1000.0, trials=2000.0, kind="price", av=True)
    if name == "c4":
        return dict(key="c4", name="C4: fused CRN Greeks 4096 spots x 1e5 trials x 252 steps",
                    spots=(50.0 + 112.0 * np.arange(4096) / 4096.0).astype(np.float32), strikes=None, steps=252.0,
                    trials=1e5, kind="greeks", av=False)
    if name in ("c5", "c5r"):
        # full C5 is 1e5 pairs x 1e6 trials x 252 steps (2.52e13 path-steps, ~10 s per GPU-step); c5r keeps the
        # 1e5-pair grid with 1e4 trials (SURVEY.md section 8d allows the reduced-trial variant; stated in the name)
        trials = 1e6 if name == "c5" else 1e4
        sp = (50.0 + 112.0 * np.arange(400) / 400.0).astype(np.float32)
        st = (80.0 + 60.0 * np.arange(250) / 250.0).astype(np.float32)
        label = "C5: batched sweep" if name == "c5" else "C5R (config 5 grid, REDUCED trial count 1e4 instead of 1e6): batched sweep"
        return dict(key=name, name="%s 1e5 spot/strike pairs x %g trials x 252 steps" % (label, trials),
                    spots=np.repeat(sp, 250), strikes=np.tile(st, 400), steps=252.0, trials=trials, kind="price",
                    av=False)
    raise SystemExit("unknown workload %r" % name)


def columns(w):
    n = len(w["spots"])
    strikes = w["strikes"] if w["strikes"] is not None else np.full(n, GRID["strike"], np.float32)
    return [w["spots"], strikes] + [np.full(n, GRID[k], np.float32) for k in ("vol", "r", "t", "q")]


def path_steps(w, n=None, trials=None):
    n = len(w["spots"]) if n is None else n
    trials = w["trials"] if trials is None else trials
    return float(n) * (8 * (int(trials) // 8)) * (2 * (int(w["steps"]) // 2))


def config_for(w, world):
    """The SAME dict from both arms (the driver compares them key by key)."""
    return {"workload": w["name"], "options": int(len(w["spots"])), "trials": float(w["trials"]),
            "steps_per_path": float(w["steps"]), "grid": "benches/benchmark.rs:10-16 conventions (K=110 or 80..140, "
            "vol 0.25, r 0.05, T 0.5, q 0.02)", "n_gpus": int(world),
            "outputs": "all Greeks + prices" if w["kind"] == "greeks" else "call + put prices",
            "l2": "flushed between timed steps (256 MiB write, outside the events); inputs are 24 B per option"}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the path kernel from the committed ncu --set full
    capture of this workload, with the file it came from; (None, None) when no capture exists.  ncu cannot run inside
    a timed bench, so this is the one figure of the line that is not measured live."""
    for tag in ("r2", "r1"):
        for name in ("%s_ncu_%s_summary.json" % (tag, key), "%s_ncu_%s_v3_summary.json" % (tag, key)):
            path = os.path.join(ROOT, "profiles", name)
            try:
                with open(path) as f:
                    m = json.load(f)["metrics"]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                val = sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                return val, "profiles/" + name
            except Exception:
                continue
    return None, None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


_SASS_CACHE = {}


def sass_loop_mix(lib_path, payoff):
    """Per-pair instruction mix of the hot loop of simulate_kernel<STREAM_FAST, payoff> in the LOADED library (cuobjdump):
    the ceilings of the pipes the kernel really occupies, in path-steps per clock per SM.  None if cuobjdump is absent."""
    if (lib_path, payoff) in _SASS_CACHE:
        return _SASS_CACHE[(lib_path, payoff)]
    out = None
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        if lib_path not in _SASS_CACHE:
            _SASS_CACHE[lib_path] = subprocess.run([exe, "-sass", lib_path], capture_output=True, text=True, timeout=120,
                                                   check=True).stdout
        txt = _SASS_CACHE[lib_path]
        body = next(f for f in re.split(r"\n\s*Function : ", txt)[1:]
                    if ("simulate_kernelILi0ELi%dELb0E" % payoff) in f.split("\n")[0])
        ins = []
        for line in body.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
        best = None
        for addr, op, rest in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) <= addr:
                    loop = [x for x in ins if int(t.group(1), 16) <= x[0] <= addr]
                    mufu = sum(1 for x in loop if x[1].startswith("MUFU"))
                    if mufu >= 3 and (best is None or mufu / len(loop) > best[0] / len(best[1])):
                        best = (mufu, loop)
        mufu, loop = best
        pairs = mufu / 3.0                                           # LG2 + SQRT + SIN per Box-Muller pair
        base = [x[1].split(".")[0] for x in loop]
        alu = sum(b in ("LOP3", "SHF", "IADD3", "LEA", "PRMT", "VIADD", "ISETP", "IADD", "SEL", "MOV", "I2FP") for b in base)
        out = {"pairs_per_iteration": pairs, "instr_per_pair": len(loop) / pairs, "alu_per_pair": alu / pairs,
               "mufu_per_pair": mufu / pairs}
        out["ceilings_path_steps_per_clk_per_sm"] = {
            "alu": 2.0 * PIPE_RATE["alu"] / out["alu_per_pair"], "xu": 2.0 * PIPE_RATE["xu"] / out["mufu_per_pair"],
            "issue": 2.0 * PIPE_RATE["issue"] / out["instr_per_pair"]}
    except Exception as e:                                           # the figure is explanatory, never fatal
        out = {"error": "%s: %s" % (type(e).__name__, e)}
    _SASS_CACHE[(lib_path, payoff)] = out
    return out


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
    crate that cannot be built in this image, so this is the oracle port (kind "port"), ALL host threads (set
    explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers).  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = oracle.baseline_use_all_cores()
    total = max(1, args.steps + args.warmup)
    cols, kind, n = cpu_sample(w, budget_s=min(2.0, 150.0 / total))
    for _ in range(args.warmup):
        run_cpu_step(cols, w, kind)
    t, ps = 0.0, 0.0
    for _ in range(args.steps):
        dt, p = run_cpu_step(cols, w, kind)
        t += dt; ps += p
    value = ps / t
    sample = ("%d of %d options per step (prefix of the grid), call and put each simulated like the reference; "
              "every simulated path-step is counted" % (n, len(w["spots"])))
    line = {"impl": "reference", "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_for(w, args.gpus),
            "note": "the reference simulates call and put separately (mc_simd.rs:477-521); the B200 arm prices both "
                    "from one pass and counts each path-step once",
            "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": cores, "cpu": cpu_model(), "kind": "port",
                             "sample": sample, "isa": "AVX2+FMA 8-lane (x86-64-v3), OpenMP"},
            "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- the B200 arm
class Bench:
    """One process = one GPU.  Measures a workload three ways (device-resident value, kernel-only roofline, e2e)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on STDOUT at the first collective when NCCL_DEBUG is set; stdout must carry
            # the one JSON line only, so file descriptor 1 points at stderr while the communicator comes up
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.pkg = importlib.import_module(PACKAGE)
        self.ctx = self.pkg.Context(device=self.local_rank, seed=SEED, substream_block=self.rank)
        self.info = self.ctx.info()
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)       # > 126 MB L2
        # all timed work and its events go on ONE explicit stream (a NULL stream would mean the context's private stream)
        self.tstream = torch.cuda.Stream(device=self.dev)
        torch.cuda.synchronize()
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0
        self.sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        peaks = measured_peaks()
        self.sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        self.peak_basis = "MEASURED_PEAKS.json sm_max_mhz" if peaks else "fallback 1965 MHz (no MEASURED_PEAKS.json)"
        self.peak = self.info["sm_count"] * self.sm_max_mhz * 1e6 * ROOFLINE_PATH_STEPS_PER_CLK_PER_SM

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ one workload
    def measure(self, w, steps, warmup, roof_seconds, sample_clocks):
        torch, dist, ctx = self.torch, self.dist, self.ctx
        greeks = w["kind"] == "greeks"
        n_total = len(w["spots"])
        n_est = 10 if greeks else 2
        slots = self.info["grid_slots_greeks" if greeks else "grid_slots"]
        shard = self.pkg.shard_query(n_total, w["steps"], w["trials"], self.world, self.rank, slots)
        mode = shard["mode"]
        lo, hi = shard["option_begin"], shard["option_end"]
        active = bool(shard["active"])
        cols_all = columns(w)
        cols_h = [np.ascontiguousarray(c[lo:hi]) for c in cols_all] if active else [c[:0] for c in cols_all]
        n = hi - lo if active else 0
        trials_local = shard["num_trials_local"] if active else 0.0
        cols_d = [torch.from_numpy(c).to(self.dev) for c in cols_h]
        in_ptrs = [c.data_ptr() for c in cols_d]
        out_d = torch.zeros(2 * n_est, max(n, 1), dtype=torch.float32, device=self.dev)
        gather_opts = mode == "options" and self.world > 1
        n_max = -(-n_total // self.world) if gather_opts else n
        send = torch.zeros(2 * n_est, n_max, dtype=torch.float32, device=self.dev) if gather_opts else None
        recv = ([torch.zeros_like(send) for _ in range(self.world)] if gather_opts and self.rank == 0 else None)
        mom = torch.zeros(max(n, 1), 2 * n_est, dtype=torch.float64, device=self.dev) if mode == "trials" else None
        val_ptrs = [out_d[i].data_ptr() for i in range(n_est)]
        se_ptrs = [out_d[n_est + i].data_ptr() for i in range(n_est)]

        def kernel_only():
            """The path kernel of this rank's share (what the roofline pass times)."""
            if not active:
                return
            if mode == "trials":
                if greeks:
                    ctx.greeks_moments_device(n, in_ptrs, w["steps"], trials_local, BUMPS, mom.data_ptr(), stream=self.stream)
                else:
                    ctx.price_moments_device(n, in_ptrs, w["steps"], trials_local, mom.data_ptr(), antithetic=w["av"],
                                             stream=self.stream)
            elif greeks:
                ctx.greeks_batch_device(n, in_ptrs, w["steps"], trials_local, BUMPS, val_ptrs, se_ptrs, stream=self.stream)
            else:
                ctx.price_batch_device(n, in_ptrs, w["steps"], trials_local, val_ptrs + se_ptrs, antithetic=w["av"],
                                       stream=self.stream)

        def exchange():
            """The exchange step the split needs, on the same stream as the kernel."""
            if mode == "trials":
                dist.all_reduce(mom, op=dist.ReduceOp.SUM)              # n x 2*n_est doubles over NCCL
                if greeks:
                    ctx.greeks_finalize_device(n, in_ptrs[3], in_ptrs[4], mom.data_ptr(), shard["total_paths"], w["trials"],
                                               BUMPS, val_ptrs, se_ptrs, stream=self.stream)
                else:
                    ctx.price_finalize_device(n, in_ptrs[3], in_ptrs[4], mom.data_ptr(), shard["total_paths"], w["trials"],
                                              val_ptrs + se_ptrs, stream=self.stream)
            elif gather_opts:
                send[:, :n].copy_(out_d[:, :n])
                dist.gather(send, recv, dst=0)                          # every rank's block to rank 0 over NCCL

        def step_device():
            kernel_only()
            if self.world > 1 and mode != "single":
                exchange()

        # ---- value: device-resident, per-step CUDA events, L2 flushed between steps
        for _ in range(warmup):
            step_device()
        self.barrier()
        launches0 = ctx.info()["kernel_launches"]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if sample_clocks and self.sampler:
            self.sampler.arm(True)
        for e0, e1 in ev:
            self.flush.fill_(1)
            e0.record()
            step_device()
            e1.record()
        self.barrier()
        if sample_clocks and self.sampler:
            self.sampler.arm(False)
        launches = ctx.info()["kernel_launches"] - launches0
        ms_total = self.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev))
        ps_step = path_steps(w)                                           # the WHOLE batch, all ranks together
        value = ps_step * steps / (ms_total * 1e-3)

        # ---- roofline: the fused path kernel alone, timed inside the library on its launch stream; >= ~1 s under load
        sim_avg_ms, n_timed = None, 0
        if active:
            ctx.set_timing(True)
            sim_ms = []
            t_end = time.perf_counter() + roof_seconds
            if sample_clocks and self.sampler:
                self.sampler.arm(True)
            while len(sim_ms) < min(steps, 3) or (time.perf_counter() < t_end and len(sim_ms) < 20000):
                kernel_only()
                sim_ms.append(ctx.last_sim_ms())
            if sample_clocks and self.sampler:
                self.sampler.arm(False)
            ctx.set_timing(False)
            sim_avg_ms, n_timed = float(np.mean(sim_ms)), len(sim_ms)
        self.barrier()
        ps_local = path_steps(w, n=n, trials=trials_local) if active else 0.0

        # ---- e2e
        h2d, d2h = 6 * n_total * 4, 2 * n_est * n_total * 4
        if self.world == 1:
            prepared = (ctx.prepare_greeks_batch(*cols_h, w["steps"], w["trials"], *BUMPS) if greeks else
                        ctx.prepare_price_batch(*cols_h, w["steps"], w["trials"], antithetic=w["av"]))
            step_host = prepared.run
            e2e_path = "one mcb200_%s_batch call per step on host buffers (H2D, kernel, D2H inside)" % ("greeks" if greeks else "price")
        else:
            pin_in = [torch.from_numpy(c).pin_memory() for c in cols_h]
            pin_out = torch.zeros(self.world, 2 * n_est, n_max, dtype=torch.float32).pin_memory() if (gather_opts and self.rank == 0) \
                else torch.zeros(2 * n_est, max(n, 1), dtype=torch.float32).pin_memory()

            def step_host():
                for d, p in zip(cols_d, pin_in):
                    d.copy_(p, non_blocking=True)
                step_device()
                if gather_opts:
                    if self.rank == 0:
                        for r in range(self.world):
                            pin_out[r].copy_(recv[r], non_blocking=True)
                elif self.rank == 0 or mode == "trials":
                    pin_out.copy_(out_d, non_blocking=True)
                self.tstream.synchronize()
            e2e_path = ("per rank: pinned host shard -> H2D -> device-pointer C-ABI call -> NCCL %s -> D2H%s, wall clock, max "
                        "over ranks" % ("all_reduce of the moment sums + finalize" if mode == "trials" else "gather to rank 0",
                                        "" if mode == "trials" else " on rank 0"))
        for _ in range(warmup):
            step_host()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        e2e_s = self.max_over_ranks(time.perf_counter() - t0)
        self.barrier()

        rec = {"workload": w["name"], "value": value, "ms_per_step": ms_total / steps, "steps": steps,
               "options_per_sec": n_total * steps / (ms_total * 1e-3), "gpu_launches": int(launches),
               "e2e": {"value": ps_step * steps / e2e_s, "unit": "path-steps/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / steps, "path": e2e_path},
               "shard": {"mode": mode, "options_on_rank0": n, "trials_on_rank0": trials_local}}
        if self.rank == 0 and active:
            achieved = ps_local / (sim_avg_ms * 1e-3)
            mix = sass_loop_mix(self.pkg.LIB_PATH, 2 if greeks else (1 if w["av"] else 0))
            per_clk = achieved / (self.info["sm_count"] * self.sm_max_mhz * 1e6)
            traffic, traffic_src = ncu_traffic(w["key"])
            roof = {"bound": "issue/mufu (not hbm, not tensor: SURVEY 8d)", "achieved": achieved, "peak": self.peak,
                    "unit": "path-steps/s", "frac": achieved / self.peak, "kernel": "simulate_kernel",
                    "kernel_ms": sim_avg_ms, "launches_timed": n_timed, "path_steps_per_launch": ps_local,
                    "peak_basis": "%d SMs x %.0f MHz (%s) x 8 path-steps/clk/SM (SURVEY 8d contract: 2 MUFU + 16 issue "
                                  "slots per path-step)" % (self.info["sm_count"], self.sm_max_mhz, self.peak_basis),
                    "path_steps_per_clk_per_sm": per_clk, "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_note": "DRAM bytes per launch from an ncu capture (the per-thread generator states; path "
                                    "data never leaves registers); algorithmic I/O is 24 B in + %d B out per option" % (8 * n_est)}
            if mix and "ceilings_path_steps_per_clk_per_sm" in mix:
                c = mix["ceilings_path_steps_per_clk_per_sm"]
                roof.update({"sass_loop": {k: mix[k] for k in ("pairs_per_iteration", "instr_per_pair", "alu_per_pair", "mufu_per_pair")},
                             "alu_bound_frac": per_clk / c["alu"], "xu_bound_frac": per_clk / c["xu"],
                             "issue_bound_frac": per_clk / c["issue"],
                             "pipe_basis": "loop SASS of the loaded library (cuobjdump) against measured pipe rates: ALU 64, "
                                           "MUFU 16, issue 128 thread-instructions/clk/SM (profiles/r2_microbench_tri_vs_quad.json)"})
            elif mix:
                roof["sass_loop"] = mix
            rec["roofline"] = roof
        return rec

    def refstream_record(self, w, reps):
        """Kernel-only rate of the reference-stream (per-path parity) kernel on the same batch: the kernel that matches
        the oracle path by path is not the one timed above, so its own throughput is published beside it."""
        ctx = self.ctx
        cols = columns(w)
        n_chunks = int(w["trials"]) // 8
        rng = np.random.default_rng(1)
        seeds = rng.integers(0, 256, size=len(w["spots"]) * n_chunks * 256, dtype=np.uint8)
        ctx.set_timing(True)
        ms = []
        for _ in range(reps + 1):
            ctx.price_batch_refstream(*cols, w["steps"], w["trials"], seeds)
            ms.append(ctx.last_sim_ms())
        ctx.set_timing(False)
        k_ms = float(np.mean(ms[1:]))
        achieved = path_steps(w) / (k_ms * 1e-3)
        return {"workload": w["name"] + " -- STREAM_REF kernel (reference draw order: 256-byte seed per 8-path chunk, two "
                "64-bit outputs per pair, 53-bit -> f64 -> f32 uniforms, sin + cos)", "kernel_ms": k_ms,
                "roofline": {"achieved": achieved, "peak": self.peak, "frac": achieved / self.peak, "unit": "path-steps/s"},
                "note": "test-only kernel; reads %d MB of chunk seeds per launch" % (seeds.size // 1000000)}


def b200_arm(args, w):
    b = Bench(args)
    main = b.measure(w, args.steps, args.warmup, args.roof_seconds, sample_clocks=True)
    subs = {}
    if b.world == 1 and not args.no_subs:
        for key, k_steps in (("c1", 100), ("c2", 100), ("c3", 100), ("c4", 5)):
            if key == w["key"]:
                continue
            sw = workload(key)
            r = b.measure(sw, k_steps, 3, 0.3, sample_clocks=False)
            if key in README_M1_DERIVED:
                r["vs_readme_m1_derived"] = r["value"] / README_M1_DERIVED[key]
            subs[key] = r
        subs["c2_refstream"] = b.refstream_record(workload("c2"), 5)
    if args.full_c5 and not args.no_subs and w["key"] != "c5":
        # the configuration the metric's "sharded over 2/4/8 B200" wording belongs to, at FULL size, at every N: one
        # warm-up step and one timed step (10 s each on one GPU, 1.25 s on eight), same shard policy and exchange
        subs["c5_full"] = b.measure(workload("c5"), 1, 1, 0.0, sample_clocks=False)
        subs["c5_full"]["note"] = "full-size config 5 (1e5 pairs x 1e6 trials x 252 steps): one warm-up step and one timed step"
    if b.rank == 0:
        clocks = b.sampler.summary()
        cpu = None
        if b.world == 1 and not args.no_cpu:
            from oracle import oracle
            cores = oracle.baseline_use_all_cores()
            cols, kind, n_cpu = cpu_sample(w, budget_s=2.0)
            t_cpu, ps_cpu, reps = 0.0, 0.0, 0
            while t_cpu < args.cpu_seconds:
                dt, p = run_cpu_step(cols, w, kind)
                t_cpu += dt; ps_cpu += p; reps += 1
            cpu = {"value": ps_cpu / t_cpu, "unit": "path-steps/s", "cores": cores, "cpu": cpu_model(), "kind": "port",
                   "isa": "AVX2+FMA 8-lane (x86-64-v3), OpenMP",
                   "sample": "%d of %d options x %d repeats (%.1f s), call and put each simulated like the reference"
                             % (n_cpu, len(w["spots"]), reps, t_cpu)}
        roof = main.get("roofline")
        if roof and clocks.get("sm_mhz"):
            roof["frac_at_measured_clock"] = roof["achieved"] / (b.info["sm_count"] * clocks["sm_mhz"] * 1e6 * 8.0)
        line = {
            "metric": "gbm_path_steps_per_sec", "value": main["value"], "unit": "path-steps/s", "n_gpus": b.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if b.world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(w, b.world),
            "parallelism": {"shard_mode": main["shard"]["mode"], "policy": "mcb200_shard_query (library): one process per "
                            "GPU, substream block = rank; options-major blocks gathered to rank 0 / trials-minor moment "
                            "sums all-reduced, inside the timed region", "options_on_rank0": main["shard"]["options_on_rank0"]},
            "launch": "one fused kernel per step and rank: paths + payoffs + deterministic reduce + finalize",
            "options_per_sec": main["options_per_sec"], "clocks": clocks, "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"], "roofline": roof, "cpu_baseline": cpu, "sub": subs or None,
        }
        print(json.dumps(line), flush=True)
    b.ctx.close()
    if b.world > 1:
        b.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5r")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-subs", action="store_true", help="skip the C1-C4 / refstream / full-C5 sub-records (N = 1)")
    ap.add_argument("--full-c5", type=int, default=1, help="time one full-size C5 step as a sub-record (~50 s / N)")
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
