#!/usr/bin/env python3
"""Trials-sharded pricing across N GPUs: NCCL all-reduce route vs the fused peer-memory combine.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/xgpu_bench.py

A few options with a very large trial count (the case trials-minor sharding exists for): every rank simulates
trials/N of every option.  Route A = mcb200_price_moments_device + all_reduce(fp64) + mcb200_price_finalize_device;
route B = mcb200_price_xgpu_device (one launch per GPU, the kernels combine through rank 0's exchange buffer).
Per-step wall time between barriers (max over ranks), after warm-up; rank 0 prints one JSON line.
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    sh = pkg.sharding
    n, steps, trials = 8, 100.0, float(os.environ.get("XGPU_TRIALS", 8e6))
    spots = torch.linspace(90.0, 130.0, n, dtype=torch.float32, device=dev)
    cols = [spots] + [torch.full((n,), v, dtype=torch.float32, device=dev) for v in
                      (GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])]
    ptrs = [c.data_ptr() for c in cols]
    local_trials, total_paths = sh.partition_trials(trials, world, rank)
    mom = torch.zeros(n, 4, dtype=torch.float64, device=dev)
    vals = torch.zeros(4, n, dtype=torch.float32, device=dev)
    tstream = torch.cuda.Stream(device=dev)        # NULL would mean the context's private stream: not ordered with NCCL
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    reps, warm = 30, 5

    def timed(step):
        for _ in range(warm):
            step()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with pkg.Context(device=local, seed=5, substream_block=rank) as ctx:
        def nccl_step():
            ctx.price_moments_device(n, ptrs, steps, float(local_trials), mom.data_ptr(), stream=stream)
            sh.combine_moments(mom)
            ctx.price_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), mom.data_ptr(), float(total_paths), trials,
                                      [vals[i].data_ptr() for i in range(4)], stream=stream)
            return vals.cpu()                            # the result is needed on the host side of every rank
        t_nccl = timed(nccl_step)
        a = vals.cpu().numpy().copy()

        handle = [ctx.xgpu_create(n, world) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.xgpu_open(handle[0], n, world)
        dist.barrier()

        fused_out = {}

        def fused_step():
            ctx.price_xgpu_device(n, ptrs, steps, float(local_trials), rank, float(total_paths), trials, stream=stream)
            fused_out.update(ctx.xgpu_results(n))        # device-side wait for all ranks + D2H of the results
        t_fused = timed(fused_step)
        b = fused_out

        def single_step():                               # the same total work on ONE GPU, for the scaling reference
            ctx.price_batch_device(n, ptrs, steps, trials, [vals[i].data_ptr() for i in range(4)], stream=stream)
            torch.cuda.synchronize()
        t_single = timed(single_step) if world > 1 else None
        ctx.xgpu_close()
    if rank == 0:
        ps = n * total_paths * 2 * (int(steps) // 2)
        print("XGPU_JSON " + json.dumps({
            "n_gpus": world, "options": n, "trials": trials, "steps": steps,
            "nccl_route_ms": 1e3 * t_nccl, "fused_route_ms": 1e3 * t_fused, "one_gpu_same_work_ms": t_single and 1e3 * t_single,
            "nccl_path_steps_per_s": ps / t_nccl, "fused_path_steps_per_s": ps / t_fused,
            "call_nccl": a[0].tolist(), "call_fused": b["call"].tolist(), "call_se_fused": b["call_se"].tolist()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
