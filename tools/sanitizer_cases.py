#!/usr/bin/env python3
"""Small batches through the C ABI for compute-sanitizer (tools/run_sanitizer.sh): every kernel family, every arrival /
flush path, sizes that finish in seconds under racecheck.  Results are checked (closed form, repeat equality) so a
sanitizer pass cannot hide a wrong answer."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GRID = (110.0, 0.25, 0.05, 0.5, 0.02)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    spots = np.arange(50, 162, dtype=np.float32)
    with pkg.Context(device=0, seed=11) as ctx:
        if which in ("all", "price"):
            # reduced C2: 112 options share the machine, ~42 warps per option -> slot / fence / arrival / last-warp sum
            a = ctx.price_batch(spots, *GRID, 20.0, 20000.0)
            ctx.set_seed(11)
            b = ctx.price_batch(spots, *GRID, 20.0, 20000.0)
            assert all(np.array_equal(a[k], b[k]) for k in a)
            assert np.all(np.isfinite(a["call"])) and a["call"][60] > 3.0
            # antithetic
            c = ctx.price_batch(spots[:16], *GRID, 20.0, 4000.0, antithetic=True)
            assert np.all(np.isfinite(c["put"]))
        if which in ("all", "crossing"):
            # warps that cross option boundaries, ragged last rounds (8008 = 250 rounds + 8 paths), solo options
            d = ctx.price_batch(np.linspace(80, 140, 41).astype(np.float32), *GRID, 10.0, 8008.0)
            assert np.all(np.isfinite(d["call"])) and np.all(d["call_se"] > 0)
            e = ctx.price_batch(np.linspace(80, 140, 20000).astype(np.float32), *GRID, 4.0, 64.0)   # solo options
            assert np.all(np.isfinite(e["call"]))
        if which in ("all", "greeks"):
            g = ctx.greeks_batch(spots[40:72], *GRID, 20.0, 8000.0)
            assert all(np.all(np.isfinite(g[k])) for k in g)
            z = ctx.greeks_batch_with_z(spots[50:66], *GRID, 20.0, 4000.0)
            assert np.all(np.isfinite(z["gamma_bs"]))
        if which in ("all", "moments"):
            import torch
            dev = torch.device("cuda", 0)
            n = 5
            cols = [torch.full((n,), v, dtype=torch.float32, device=dev) for v in (100.0,) + GRID]
            mom = torch.zeros(n, 4, dtype=torch.float64, device=dev)
            vals = torch.zeros(4, n, device=dev)
            ctx.price_moments_device(n, [c.data_ptr() for c in cols], 20.0, 16000.0, mom.data_ptr())
            ctx.price_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), mom.data_ptr(), 16000.0, 16000.0,
                                      [vals[i].data_ptr() for i in range(4)])
            torch.cuda.synchronize()
            assert torch.isfinite(vals).all()
        if which in ("all", "xgpu") and hasattr(ctx, "xgpu_attach"):
            run_xgpu(pkg)
    print("sanitizer cases ok:", which)


def run_xgpu(pkg):
    """Two contexts of ONE process on device 0 share an exchange buffer: the cross-GPU arrival protocol on one GPU."""
    import torch
    dev = torch.device("cuda", 0)
    n = 6
    cols = [torch.full((n,), v, dtype=torch.float32, device=dev) for v in (100.0,) + GRID]
    ptrs = [c.data_ptr() for c in cols]
    s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    with pkg.Context(device=0, seed=5, substream_block=0) as c0, pkg.Context(device=0, seed=5, substream_block=1) as c1:
        base = c0.xgpu_create(n, 2)
        c1.xgpu_attach(c0, n, 2)
        for _ in range(3):
            c0.price_xgpu_device(n, ptrs, 20.0, 8000.0, 0, 16000.0, 16000.0, stream=s0.cuda_stream)
            c1.price_xgpu_device(n, ptrs, 20.0, 8000.0, 1, 16000.0, 16000.0, stream=s1.cuda_stream)
            r0, r1 = c0.xgpu_results(n), c1.xgpu_results(n)
            assert all(np.array_equal(r0[k], r1[k]) for k in r0)
        c1.xgpu_close(); c0.xgpu_close()


if __name__ == "__main__":
    main()
