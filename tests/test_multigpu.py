"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multigpu`).
One process per GPU via torchrun, NCCL: options-major blocks with no data-path collective, and trials-minor sharding
with one all-reduce of the fp64 moment sums."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharding(oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("MULTIGPU_JSON ")][-1]
    out = json.loads(line[len("MULTIGPU_JSON "):])

    blocks = out["options_major"]
    assert [b["lo"] for b in blocks] == [0, 56] and [b["hi"] for b in blocks] == [56, 112]
    spots = np.arange(50, 162, dtype=np.float32)
    zs = []
    for b in blocks:
        assert b["replay_equal"] and b["fresh_differs"]      # per-rank determinism; consecutive calls use fresh paths
        for k, s in enumerate(spots[b["lo"]:b["hi"]]):
            exact = oracle.bs64("call_price", float(s), GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])
            zs.append(abs(b["call"][k] - exact) / max(b["call_se"][k], 2e-3))
    zs = np.array(zs)
    assert (zs > 3).mean() <= 0.03 and zs.max() < 4.5

    tm = out["trials_minor"]
    assert tm["same_on_all_ranks"] and tm["total_paths"] == 400000
    for k, s in enumerate(tm["spots"]):
        for name, key in (("call_price", "call"), ("put_price", "put")):
            exact = oracle.bs64(name, float(s), GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])
            assert abs(tm[key][k] - exact) <= 3.5 * tm[key + "_se"][k] + 1e-4
        assert tm["call_se"][k] < 0.04              # full-sample SE: sigma_payoff (<= ~22) / sqrt(4e5)

    # the fused cross-GPU combine (peer memory + cross-GPU arrival counters, no NCCL on the data path) uses the same
    # seeds and the same trial split: with two ranks the rank-ordered sum equals the all-reduce bit for bit
    fu = out["trials_minor_fused"]
    assert fu["repeat_equal"] and fu["same_on_all_ranks"]
    for key in ("call", "put", "call_se", "put_se"):
        assert fu[key] == tm[key], key
    assert out["greeks_fused_equals_nccl"] and all(g > 0 for g in out["greeks_fused_gamma"])


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_device_context_on_two_real_gpus():
    """mcb200_ctx_create_multi over devices [0, 1] (one process): the options-major blocks equal plain contexts on each
    device bit for bit; a trials-minor batch equals moments(device 0) + moments(device 1) + one finalisation; and the
    fused exchange buffer works between two same-process contexts on DIFFERENT devices (plain peer access)."""
    from conftest import load_package
    pkg = load_package()
    G = (GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])
    spots = np.linspace(70, 150, 64).astype(np.float32)
    with pkg.Context(devices=[0, 1], seed=31) as multi:
        got = multi.price_batch(spots, *G, 100.0, 80000.0)
        few = multi.price_batch(spots[:5], *G, 100.0, 800000.0)      # 5 options on 2 GPUs: trials-minor
        bs = multi.bs_batch(spots, *G)
    for d, (lo, hi) in enumerate(((0, 32), (32, 64))):
        with pkg.Context(device=d, seed=31, substream_block=d) as one:
            want = one.price_batch(spots[lo:hi], *G, 100.0, 80000.0)
        for k in want:
            assert np.array_equal(got[k][lo:hi], want[k]), (d, k)
    z = (got["call"] - bs["call"]) / np.maximum(got["call_se"], 1e-4)
    assert np.abs(z).max() < 4.5
    moms = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        cols = [torch.tensor(spots[:5], device=dev)] + [torch.full((5,), v, dtype=torch.float32, device=dev) for v in G]
        with pkg.Context(device=d, seed=31, substream_block=d) as one:
            with torch.cuda.device(dev):
                # advance the substreams exactly like the multi context did (one options-major batch of 32 options)
                one.price_batch(spots[32 * d:32 * d + 32], *G, 100.0, 80000.0)
                m = torch.zeros(5, 4, dtype=torch.float64, device=dev)
                one.price_moments_device(5, [c.data_ptr() for c in cols], 100.0, 400000.0, m.data_ptr())
                torch.cuda.synchronize(dev)
                moms.append(m.cpu())
    total = (moms[0] + moms[1]).cuda(0)
    dev0 = torch.device("cuda", 0)
    rate = torch.full((5,), GRID["r"], dtype=torch.float32, device=dev0)
    years = torch.full((5,), GRID["t"], dtype=torch.float32, device=dev0)
    vals = torch.zeros(4, 5, device=dev0)
    with pkg.Context(device=0, seed=1) as fin:
        fin.price_finalize_device(5, rate.data_ptr(), years.data_ptr(), total.data_ptr(), 800000.0, 800000.0,
                                  [vals[i].data_ptr() for i in range(4)])
        torch.cuda.synchronize(dev0)
    for i, k in enumerate(("call", "put", "call_se", "put_se")):
        assert np.array_equal(few[k], vals[i].cpu().numpy()), k
    # fused exchange buffer across two real devices of one process
    n = 4
    with pkg.Context(device=0, seed=77, substream_block=0) as c0, pkg.Context(device=1, seed=77, substream_block=1) as c1:
        c0.xgpu_create(n, 2)
        c1.xgpu_attach(c0, n, 2)
        colsd = []
        for d in (0, 1):
            dev = torch.device("cuda", d)
            colsd.append([torch.tensor(spots[:4], device=dev)] + [torch.full((4,), v, dtype=torch.float32, device=dev) for v in G])
        for _ in range(3):
            c0.price_xgpu_device(n, [c.data_ptr() for c in colsd[0]], 100.0, 200000.0, 0, 400000.0, 400000.0)
            c1.price_xgpu_device(n, [c.data_ptr() for c in colsd[1]], 100.0, 200000.0, 1, 400000.0, 400000.0)
            r0, r1 = c0.xgpu_results(n), c1.xgpu_results(n)
            assert all(np.array_equal(r0[k], r1[k]) for k in r0)
            assert np.all(np.abs(r0["call"] - bs["call"][:4]) <= 4.0 * r0["call_se"] + 1e-4)
        c1.xgpu_close(); c0.xgpu_close()
