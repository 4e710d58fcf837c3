"""torchrun worker for tests/test_multigpu.py: one process per GPU, NCCL.  Prints one JSON line from rank 0."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    sh = pkg.sharding
    out = {}
    # an explicit stream: the library reads a NULL stream as "the context's own stream", which is NOT ordered with
    # torch's default stream where the all-reduce would run
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # ---- options-major: each rank prices its contiguous block on its own long_jump substream block, no collective on
    #      the data path; the blocks are gathered only to be checked
    spots = np.arange(50, 162, dtype=np.float32)
    lo, hi = sh.partition_options(len(spots), world, rank)
    with pkg.Context(device=local, seed=77, substream_block=rank) as ctx:
        res = ctx.price_batch(spots[lo:hi], GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"], 100.0, 20000.0)
        again = ctx.price_batch(spots[lo:hi], GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"], 100.0, 20000.0)
        ctx.set_seed(77, rank)
        replay = ctx.price_batch(spots[lo:hi], GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"], 100.0, 20000.0)
    blocks = [None] * world
    dist.all_gather_object(blocks, dict(lo=lo, hi=hi, call=res["call"].tolist(), call_se=res["call_se"].tolist(),
                                        put=res["put"].tolist(), put_se=res["put_se"].tolist(),
                                        replay_equal=bool(np.array_equal(res["call"], replay["call"])),
                                        fresh_differs=bool(not np.array_equal(res["call"], again["call"]))))
    out["options_major"] = blocks

    # ---- trials-minor: every rank simulates a disjoint trial range of the SAME options; the fp64 moment sums are
    #      all-reduced over NCCL and finalised once (results land on every rank)
    tspots = torch.tensor([90.0, 100.0, 110.0, 120.0, 130.0], dtype=torch.float32, device=dev)
    n = tspots.numel()
    cols = [tspots] + [torch.full((n,), v, dtype=torch.float32, device=dev) for v in
                       (GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])]
    ptrs = [c.data_ptr() for c in cols]
    trials = 400000.0
    local_trials, total_paths = sh.partition_trials(trials, world, rank)
    mom = torch.zeros(n, 4, dtype=torch.float64, device=dev)
    vals = torch.zeros(4, n, dtype=torch.float32, device=dev)
    with pkg.Context(device=local, seed=99, substream_block=rank) as ctx:
        ctx.price_moments_device(n, ptrs, 100.0, float(local_trials), mom.data_ptr(), stream=stream)
        sh.combine_moments(mom)                         # ncclAllReduce(fp64, sum) on the same stream
        ctx.price_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), mom.data_ptr(), float(total_paths), trials,
                                  [vals[i].data_ptr() for i in range(4)], stream=stream)
        torch.cuda.synchronize()
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
    out["trials_minor"] = dict(call=vals[0].tolist(), put=vals[1].tolist(), call_se=vals[2].tolist(),
                               put_se=vals[3].tolist(), spots=tspots.tolist(), total_paths=total_paths,
                               same_on_all_ranks=all(bool(torch.equal(g, vals)) for g in gathered))
    # ---- trials-minor, FUSED: no collective and no second kernel -- every rank's path kernel stores its moment sums
    #      into rank 0's exchange buffer (CUDA IPC peer memory over NVLink) and the last arrival finalises there
    with pkg.Context(device=local, seed=99, substream_block=rank) as ctx:
        handle = [ctx.xgpu_create(n, world) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.xgpu_open(handle[0], n, world)
        dist.barrier()
        fused = []
        for rep in range(2):                                  # twice: the arrival counters must reset themselves
            ctx.set_seed(99, rank)
            ctx.price_xgpu_device(n, ptrs, 100.0, float(local_trials), rank, float(total_paths), trials, stream=stream)
            fused.append(ctx.xgpu_results(n))                 # waits on the device for ALL ranks, no barrier needed
        ctx.xgpu_close()
        dist.barrier()
    same = [None] * world
    dist.all_gather_object(same, [fused[0]["call"].tolist(), fused[0]["call_se"].tolist()])
    out["trials_minor_fused"] = dict(call=fused[0]["call"].tolist(), put=fused[0]["put"].tolist(),
                                     call_se=fused[0]["call_se"].tolist(), put_se=fused[0]["put_se"].tolist(),
                                     repeat_equal=bool(all(np.array_equal(fused[0][k], fused[1][k]) for k in fused[0])),
                                     same_on_all_ranks=all(s == same[0] for s in same))
    # ---- the same for the Greeks: fused cross-GPU combine vs moments + all_reduce + finalize on the same seeds
    bumps = (0.01, 0.01, 0.01, 0.001)
    gmom = torch.zeros(n, 20, dtype=torch.float64, device=dev)
    gv, gs = torch.zeros(10, n, device=dev), torch.zeros(10, n, device=dev)
    g_trials = 80000.0
    g_local, g_total = sh.partition_trials(g_trials, world, rank)
    with pkg.Context(device=local, seed=123, substream_block=rank) as ctx:
        ctx.greeks_moments_device(n, ptrs, 100.0, float(g_local), bumps, gmom.data_ptr(), stream=stream)
        sh.combine_moments(gmom)
        ctx.greeks_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), gmom.data_ptr(), float(g_total), g_trials, bumps,
                                   [gv[i].data_ptr() for i in range(10)], [gs[i].data_ptr() for i in range(10)], stream=stream)
        torch.cuda.synchronize()
        handle = [ctx.xgpu_create(n, world, greeks=True) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.xgpu_open(handle[0], n, world, greeks=True)
        dist.barrier()
        ctx.set_seed(123, rank)
        ctx.greeks_xgpu_device(n, ptrs, 100.0, float(g_local), bumps, rank, float(g_total), g_trials, stream=stream)
        gf = ctx.xgpu_results_greeks(n)
        dist.barrier()
        ctx.xgpu_close()
    names = pkg.GREEK_NAMES
    out["greeks_fused_equals_nccl"] = bool(all(np.array_equal(gf[nm], gv[i].cpu().numpy()) and
                                               np.array_equal(gf[nm + "_se"], gs[i].cpu().numpy())
                                               for i, nm in enumerate(names)))
    out["greeks_fused_gamma"] = gf["gamma"].tolist()
    if rank == 0:
        print("MULTIGPU_JSON " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
