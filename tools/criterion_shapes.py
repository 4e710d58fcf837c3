#!/usr/bin/env python3
"""The reference's criterion shapes (benches/benchmark.rs:18-103) on the B200 path and on the CPU ports, with
criterion-style statistics.

For each shape the reference times ONE pass of `for spot in 50..162 { mc_simd::call_price(spot, 110, .25, .05, .5, .02,
steps, trials) }` (benches/benchmark.rs:40-42; the group at :102 registers 80x100, 1000x100, 10000x100 and 20000x100,
the 2000x1000 function at :67-79 exists but is not registered -- all five are run here).  Arms, per pass:
  serial   the same loop through the 12-function drop-in API (112 blocking mc_simd_call_price calls),
  batched  one mcb200_price_batch call for the 112 spots (call + put + standard errors),
  cpu      the oracle's SIMD+threads port of the reference algorithm, same loop ("monte carlo fast" rows of criterion),
  scalar   the oracle's restatement of the scalar pricer src/mc.rs, same loop ("monte carlo" rows), few samples.
Statistics follow criterion's report: >= 100 samples (fewer only where one pass takes seconds), mean with a 95 %
bootstrap confidence interval, median, standard deviation, and Tukey outlier counts (mild: beyond 1.5 IQR, severe:
beyond 3 IQR, low / high).  Writes a markdown table + JSON; --write-baseline replaces the table between the
`<!-- criterion:begin -->` / `<!-- criterion:end -->` markers of BASELINE.md.

    python tools/criterion_shapes.py [--samples 100] [--write-baseline] [--out profiles/r2_criterion_shapes]
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)

# (trials, steps, README.md M1 time of mc_simd::call_price in seconds, README.md M1 time of mc::call_price)
SHAPES = [(80, 100, 3.29e-3, 33.649e-3), (1000, 100, 15e-3, 419e-3), (10000, 100, 104e-3, 4.26), (2000, 1000, 205e-3, 8.22),
          (20000, 100, 207e-3, 8.27)]
K, V, R, T, Q = 110.0, 0.25, 0.05, 0.5, 0.02


def criterion_stats(samples, n_boot=10000, seed=0):
    """mean +- 95 % bootstrap CI, median, std and Tukey outliers of a list of per-pass times (seconds)."""
    x = np.asarray(samples, np.float64)
    rng = np.random.default_rng(seed)
    if len(x) > 1:
        boots = x[rng.integers(0, len(x), size=(n_boot, len(x)))].mean(axis=1)
        lo, hi = np.percentile(boots, [2.5, 97.5])
    else:
        lo = hi = x[0]
    q1, q3 = np.percentile(x, [25, 75])
    iqr = q3 - q1
    out = {"low_severe": int((x < q1 - 3 * iqr).sum()), "low_mild": int(((x >= q1 - 3 * iqr) & (x < q1 - 1.5 * iqr)).sum()),
           "high_mild": int(((x > q3 + 1.5 * iqr) & (x <= q3 + 3 * iqr)).sum()), "high_severe": int((x > q3 + 3 * iqr).sum())}
    return {"n": int(len(x)), "mean": float(x.mean()), "ci95": [float(lo), float(hi)], "median": float(np.median(x)),
            "std": float(x.std(ddof=1)) if len(x) > 1 else 0.0, "outliers": out}


def sample(fn, n, warmup=3, budget_s=None):
    """n timed passes after `warmup`; with a budget, stop early once it is spent (at least one sample)."""
    for _ in range(warmup):
        fn()
    ts = []
    t_end = time.perf_counter() + budget_s if budget_s else None
    while len(ts) < n:
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
        if t_end and time.perf_counter() > t_end:
            break
    return ts


def run(shapes=SHAPES, samples=100, cpu_budget_s=8.0, scalar_budget_s=6.0, seed=2024, device=0):
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    from oracle import oracle   # baseline arms only (test infrastructure)
    threads = oracle.baseline_use_all_cores()
    spots = np.arange(50, 162, dtype=np.float32)
    cols = [spots] + [np.full(len(spots), v, np.float32) for v in (K, V, R, T, Q)]
    pkg.mc_simd.reset_default_context(device=device, seed=seed)
    rows = []
    with pkg.Context(device=device, seed=7) as ctx:
        bs = ctx.bs_batch(*cols)
        exact = bs["call"]
        near = (bs["call_delta"] > 0.25) & (bs["call_delta"] < 0.75)       # far from the money the sample SE of a few
                                                                           # hundred paths is not a usable yardstick
        for trials, steps, m1_fast, m1_scalar in shapes:
            call = pkg.mc_simd.call_price
            last = {}

            def serial():
                last["serial"] = [call(float(s), K, V, R, T, Q, float(steps), float(trials)) for s in range(50, 162)]
            prepared = ctx.prepare_price_batch(*cols, float(steps), float(trials))

            def batched():
                last["batched"] = prepared.run()

            def cpu():
                oracle.baseline_price_loop(*cols, float(steps), float(trials), 0)

            def scalar():
                for s in range(50, 162):
                    oracle.mc_scalar(float(s), K, V, R, T, Q, float(steps), float(trials), 1.0, seed=s)
            ps = 112.0 * (8 * (trials // 8)) * (2 * (steps // 2))
            row = {"trials": trials, "steps": steps, "path_steps": ps, "m1_mc_simd_s": m1_fast, "m1_mc_scalar_s": m1_scalar,
                   "cpu_threads": threads,
                   "serial": criterion_stats(sample(serial, samples)), "batched": criterion_stats(sample(batched, samples)),
                   "cpu": criterion_stats(sample(cpu, samples, budget_s=cpu_budget_s)),
                   "scalar": criterion_stats(sample(scalar, 3, warmup=0, budget_s=scalar_budget_s))}
            # both B200 arms against the closed form (the oracle the reference's own tests use): worst error in units of
            # the batched call's standard error (serial and batched draw independent paths)
            se = np.maximum(last["batched"]["call_se"], 2e-3)[near]
            row["max_z_batched"] = float(np.max(np.abs(last["batched"]["call"] - exact)[near] / se))
            row["max_z_serial"] = float(np.max(np.abs(np.array(last["serial"], np.float32) - exact)[near] / se))
            row["batched_path_steps_per_s"] = ps / row["batched"]["mean"]
            row["serial_path_steps_per_s"] = ps / row["serial"]["mean"]
            row["cpu_path_steps_per_s"] = ps / row["cpu"]["mean"]
            rows.append(row)
    return rows


def fmt_time(s):
    return "%.3f ms" % (1e3 * s) if s < 1.0 else "%.2f s" % s


def markdown(rows):
    def cell(st):
        o = st["outliers"]
        n_out = sum(o.values())
        return "%s [%s, %s]%s" % (fmt_time(st["mean"]), fmt_time(st["ci95"][0]), fmt_time(st["ci95"][1]),
                                  " (n=%d, %d outl.)" % (st["n"], n_out) if st["n"] > 1 else " (n=1)")
    lines = ["| trials x steps | M1 `mc` / `mc_simd` (README) | scalar port, 1 thread | CPU port (%d thr) | B200, 112 serial drop-ins | "
             "B200, one batched call | batched path-steps/s | max z near the money (batched / serial) |" % rows[0]["cpu_threads"],
             "|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append("| %d x %d | %s / %s | %s | %s | %s | %s | %.3g | %.1f / %.1f |" % (
            r["trials"], r["steps"], fmt_time(r["m1_mc_scalar_s"]), fmt_time(r["m1_mc_simd_s"]), cell(r["scalar"]), cell(r["cpu"]),
            cell(r["serial"]), cell(r["batched"]), r["batched_path_steps_per_s"], r["max_z_batched"], r["max_z_serial"]))
    lines.append("")
    lines.append("Time per pass over the reference's 112-spot criterion loop (`benches/benchmark.rs:40-42`): mean [95 % bootstrap "
                 "CI] (samples, Tukey outliers), like criterion's report.  `tools/criterion_shapes.py` under gpurun, 1 x B200.")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "criterion_shapes"))
    ap.add_argument("--write-baseline", action="store_true")
    ap.add_argument("--from-json", default=None, help="format rows measured earlier (no GPU needed) instead of measuring")
    args = ap.parse_args()
    rows = json.load(open(args.from_json)) if args.from_json else run(samples=args.samples)
    md = markdown(rows)
    print(md)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".md", "w") as f:
        f.write(md + "\n")
    with open(args.out + ".json", "w") as f:
        json.dump(rows, f, indent=1)
    if args.write_baseline:
        path = os.path.join(ROOT, "BASELINE.md")
        text = open(path).read()
        a, b = "<!-- criterion:begin -->", "<!-- criterion:end -->"
        if a in text and b in text:
            text = text[:text.index(a) + len(a)] + "\n" + md + "\n" + text[text.index(b):]
            open(path, "w").write(text)
            print("BASELINE.md table rewritten")


if __name__ == "__main__":
    main()
