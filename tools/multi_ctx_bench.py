#!/usr/bin/env python3
"""The multi-device context (mcb200_ctx_create_multi) measured end to end: ONE process, ONE host-buffer C-ABI call per
step, the library splits the batch over D GPUs (mcb200_shard_query policy), D = 1, 2, 4, ... up to the GPUs present.

    python tools/multi_ctx_bench.py [--workloads c5r c4 c2 few] [--reps 10]      (gpurun --gpus 8)

"few" = 5 options x 8e6 trials x 100 steps: the trials-minor split (per-device moment sums, host sum in device order, one
finalisation).  Wall clock around `reps` calls after 3 warm-up calls; host buffers in and out inside the timed region.
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
from bench import BUMPS, SEED, columns, path_steps, workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="*", default=["c5r", "c4", "c2", "few"])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    n_gpu = torch.cuda.device_count()
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    out = {"gpus_present": n_gpu, "results": []}
    for name in args.workloads:
        if name == "few":
            w = dict(key="few", name="5 options x 8e6 trials x 100 steps (trials-minor)", spots=np.linspace(90, 130, 5).astype(np.float32),
                     strikes=None, steps=100.0, trials=8e6, kind="price", av=False)
        else:
            w = workload(name)
        cols = columns(w)
        base = None
        for d in [x for x in (1, 2, 4, 8) if x <= n_gpu]:
            with pkg.Context(devices=list(range(d)), seed=SEED) as ctx:
                slots = ctx.info()["grid_slots_greeks" if w["kind"] == "greeks" else "grid_slots"]
                mode = pkg.shard_query(len(w["spots"]), w["steps"], w["trials"], d, 0, slots)["mode"]
                call = (ctx.prepare_greeks_batch(*cols, w["steps"], w["trials"], *BUMPS) if w["kind"] == "greeks" else
                        ctx.prepare_price_batch(*cols, w["steps"], w["trials"], antithetic=w["av"]))
                reps = args.reps if path_steps(w) < 1e12 else max(3, args.reps // 3)
                for _ in range(3):
                    call.run()
                t0 = time.perf_counter()
                for _ in range(reps):
                    res = call.run()
                dt = (time.perf_counter() - t0) / reps
            rate = path_steps(w) / dt
            base = base or rate
            rec = {"workload": w["name"], "devices": d, "shard_mode": mode, "ms_per_call": 1e3 * dt, "path_steps_per_s": rate,
                   "speedup_vs_1": rate / base, "reps": reps,
                   "finite": bool(all(np.all(np.isfinite(v)) for v in res.values()))}
            out["results"].append(rec)
            print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multi_ctx_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
