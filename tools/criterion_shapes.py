#!/usr/bin/env python3
"""The reference's criterion shapes (benches/benchmark.rs:18-103) on the B200 path and on the CPU port.

For each shape the reference times ONE pass of `for spot in 50..162 { mc_simd::call_price(spot, 110, .25, .05, .5, .02,
steps, trials) }` (benches/benchmark.rs:40-42).  This driver times, per pass:
  serial   the same loop through the 12-function drop-in API (112 blocking mc_simd_call_price calls),
  batched  one mcb200_price_batch call for the 112 spots (call + put + standard errors),
  cpu      the oracle's SIMD+threads port of the reference algorithm, same loop (test infrastructure, baseline only).
Prints a markdown table and one JSON line; README.md:95-124 (M1) numbers are quoted beside them.
"""
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("monte-carlo-options-simd_b200")
from oracle import oracle   # noqa: E402  (baseline leg only)

SHAPES = [(80, 100, 3.29e-3), (1000, 100, 15e-3), (10000, 100, 104e-3), (2000, 1000, 205e-3), (20000, 100, 207e-3)]
K, V, R, T, Q = 110.0, 0.25, 0.05, 0.5, 0.02


def timed(fn, min_time=0.4, min_iter=5):
    for _ in range(3):
        fn()
    ts = []
    t_end = time.perf_counter() + min_time
    while len(ts) < min_iter or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), len(ts)


def main():
    spots = np.arange(50, 162, dtype=np.float32)
    cols = [spots] + [np.full(len(spots), v, np.float32) for v in (K, V, R, T, Q)]
    pkg.mc_simd.reset_default_context(device=0, seed=2024)
    rows = []
    with pkg.Context(device=0, seed=7) as ctx:
        for trials, steps, m1 in SHAPES:
            call = pkg.mc_simd.call_price

            def serial():
                for s in range(50, 162):
                    call(float(s), K, V, R, T, Q, float(steps), float(trials))
            prepared = ctx.prepare_price_batch(*cols, float(steps), float(trials))
            t_serial, n1 = timed(serial)
            t_batch, n2 = timed(prepared.run)
            t_cpu, n3 = timed(lambda: oracle.baseline_price_loop(*cols, float(steps), float(trials), 0), min_time=1.0, min_iter=3)
            ps = 112.0 * (8 * (trials // 8)) * (2 * (steps // 2))
            rows.append(dict(trials=trials, steps=steps, serial_ms=1e3 * t_serial, batched_ms=1e3 * t_batch,
                             cpu_ms=1e3 * t_cpu, m1_ms=1e3 * m1, path_steps=ps,
                             batched_path_steps_per_s=ps / t_batch, serial_path_steps_per_s=ps / t_serial,
                             cpu_path_steps_per_s=ps / t_cpu, cpu_threads=oracle.baseline_threads()))
    print("| trials x steps | M1 mc_simd (README) | CPU port (%d thr) | B200 112 serial drop-ins | B200 one batched call | batched path-steps/s |"
          % rows[0]["cpu_threads"])
    print("|---|---|---|---|---|---|")
    for r in rows:
        print("| %d x %d | %.2f ms | %.2f ms | %.3f ms | %.3f ms | %.3g |" % (
            r["trials"], r["steps"], r["m1_ms"], r["cpu_ms"], r["serial_ms"], r["batched_ms"], r["batched_path_steps_per_s"]))
    print("CRITERION_JSON " + json.dumps(rows))


if __name__ == "__main__":
    main()
