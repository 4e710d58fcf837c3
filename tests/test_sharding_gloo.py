"""World-size-2 test of the multi-GPU host logic on CPU (gloo): trials-minor sharding.

Each rank owns a disjoint range of 8-path chunks of the SAME options (sharding.partition_trials), builds the per-option
moment sums {sum, sum^2} x {call, put} of its own paths (here with the CPU oracle standing in for the GPU kernel --
test infrastructure only), the ranks all-reduce the fp64 moments (sharding.combine_moments: NCCL on GPUs, gloo here) and
finalise once (sharding.finalize_price_moments, the host statement of mcb200_price_finalize_device).  The result must
equal the single-process estimate over all chunks."""
import math
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_package

torch = pytest.importorskip("torch")
import torch.distributed as dist            # noqa: E402
import torch.multiprocessing as mp          # noqa: E402

PARAMS = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02, steps=20.0, trials=1000.0)
SPOTS = (95.0, 110.0, 125.0)


def _moments(oracle, spot, seeds, trials_local):
    """[sum call, sum call^2, sum put, sum put^2] over the paths of `seeds` (undiscounted payoffs)."""
    d = oracle.pricing_detail(spot, PARAMS["strike"], PARAMS["vol"], PARAMS["r"], PARAMS["t"], PARAMS["q"],
                              PARAMS["steps"], trials_local, 1.0, seeds, want_paths=True)
    st = np.float64(spot) * d["growth"].astype(np.float64)
    call, put = np.maximum(st - PARAMS["strike"], 0.0), np.maximum(PARAMS["strike"] - st, 0.0)
    return [call.sum(), (call ** 2).sum(), put.sum(), (put ** 2).sum()]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    from oracle import oracle
    sh = pkg.sharding
    # the rank's share comes from the LIBRARY's shard policy (mcb200_shard_query, pure host logic): 3 options on 2 ranks
    # with a batch larger than one (here: tiny, 2-CTA) GPU -> trials-minor
    shard = pkg.shard_query(len(SPOTS), PARAMS["steps"], PARAMS["trials"], world, rank, 2)
    assert shard["mode"] == "trials" and shard["active"]
    local, total_paths = int(shard["num_trials_local"]), int(shard["total_paths"])
    assert (local, total_paths) == sh.partition_trials(PARAMS["trials"], world, rank)
    first_chunk = sum(int(pkg.shard_query(len(SPOTS), PARAMS["steps"], PARAMS["trials"], world, r, 2)["num_trials_local"])
                      for r in range(rank)) // 8
    all_seeds = oracle.chunk_seeds(int(PARAMS["trials"]) // 8, 4321).reshape(-1, 256)
    mine = np.ascontiguousarray(all_seeds[first_chunk:first_chunk + local // 8])
    m = torch.tensor([_moments(oracle, s, mine, float(local)) for s in SPOTS], dtype=torch.float64)
    sh.combine_moments(m)
    n = len(SPOTS)
    res = sh.finalize_price_moments(m.numpy(), np.full(n, PARAMS["r"], np.float32), np.full(n, PARAMS["t"], np.float32),
                                    total_paths, PARAMS["trials"])
    lo, hi = sh.partition_options(n, world, rank)
    if rank == 0:
        np.savez(out_path, call=res["call"], put=res["put"], call_se=res["call_se"], put_se=res["put_se"],
                 moments=m.numpy(), block=np.array([lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_trials_sharding_world2_gloo(tmp_path, oracle, pkg):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    seeds = oracle.chunk_seeds(int(PARAMS["trials"]) // 8, 4321)
    disc = math.exp(-PARAMS["r"] * PARAMS["t"])
    for i, s in enumerate(SPOTS):
        whole = _moments(oracle, s, seeds, PARAMS["trials"])
        assert np.allclose(got["moments"][i], whole, rtol=1e-12)            # same paths, summed in two halves
        ref = oracle.pricing_detail(s, PARAMS["strike"], PARAMS["vol"], PARAMS["r"], PARAMS["t"], PARAMS["q"],
                                    PARAMS["steps"], PARAMS["trials"], 1.0, seeds)
        assert got["call"][i] == pytest.approx(ref["price"], rel=2e-6)      # == the reference algorithm's price
        assert got["call"][i] == pytest.approx(whole[0] / PARAMS["trials"] * disc, rel=1e-6)
        n = 8 * (int(PARAMS["trials"]) // 8)
        var = (whole[1] - whole[0] ** 2 / n) / (n - 1)
        assert got["call_se"][i] == pytest.approx(math.sqrt(var / n) * n / PARAMS["trials"] * disc, rel=1e-5)
    assert list(got["block"]) == [0, 1]                                      # options-major block of rank 0 of 2: [0*3//2, 1*3//2)


def _options_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    from oracle import oracle
    spots = np.linspace(80.0, 140.0, 40).astype(np.float32)
    shard = pkg.shard_query(len(spots), PARAMS["steps"], 800.0, world, rank, 2)      # 40 options, 2 ranks: options-major
    lo, hi = shard["option_begin"], shard["option_end"]
    mine = torch.zeros(len(spots) // world, dtype=torch.float32)
    for j, o in enumerate(range(lo, hi)):                  # seeds follow the OPTION, so the split cannot change a result
        seeds = oracle.chunk_seeds(100, 1000 + o)
        mine[j] = oracle.pricing_detail(float(spots[o]), PARAMS["strike"], PARAMS["vol"], PARAMS["r"], PARAMS["t"],
                                        PARAMS["q"], PARAMS["steps"], 800.0, 1.0, seeds)["price"]
    parts = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)                        # the exchange step of the options-major split (NCCL on GPUs)
    if rank == 0:
        np.savez(out_path, prices=torch.cat(parts).numpy(), mode=np.array([shard["mode"] == "options"]),
                 block=np.array([lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_options_sharding_world2_gloo(tmp_path, oracle, pkg):
    """Options-major split on two ranks: contiguous blocks from mcb200_shard_query, results gathered to rank 0, equal to
    the unsharded loop option by option."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "opt.npz")
    mp.spawn(_options_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got["mode"][0]) and list(got["block"]) == [0, 20]
    spots = np.linspace(80.0, 140.0, 40).astype(np.float32)
    for o, s in enumerate(spots):
        want = oracle.pricing_detail(float(s), PARAMS["strike"], PARAMS["vol"], PARAMS["r"], PARAMS["t"], PARAMS["q"],
                                     PARAMS["steps"], 800.0, 1.0, oracle.chunk_seeds(100, 1000 + o))["price"]
        assert got["prices"][o] == np.float32(want), o
