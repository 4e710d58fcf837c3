#!/usr/bin/env python3
"""Memory / race evidence without compute-sanitizer (closed on this GPU pool: profiles/r2_sanitizer_closed_on_pool.txt).

Builds lib/libmcb200_checked.so with -DMCB_CHECKED: every long-lived device buffer of a context sits between two
256-byte guard regions, the path kernel checks every pool / option / slot / arrival / exchange-buffer index it forms,
and mcb200_debug_check() reports the OR of the failed checks, damaged guards and arrival counters that did not return
to zero.  Runs the sanitizer case list (tools/sanitizer_cases.py: every kernel family and arrival path, the multi-device
context and the two-rank exchange buffer on one GPU) plus a soak of random batch shapes, checking after every batch, and
compares every batch bit for bit with a replay on a second context (a race in the arrival protocol shows up as a
run-to-run difference: the summation order is fixed by the plan, not by arrival order).

    python tools/checked_build.py            (GPU box)  -> gpurun_out/checked_build.txt
"""
import ctypes as C
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MCB200_LIB_NAME"] = "libmcb200_checked.so"
os.environ["MCB200_NVCC_FLAGS"] = (os.environ.get("MCB200_NVCC_FLAGS", "") + " -DMCB_CHECKED").strip()

import numpy as np  # noqa: E402

G = (110.0, 0.25, 0.05, 0.5, 0.02)
BITS = {1: "pool index", 2: "option index", 4: "slot index", 8: "arrival count", 16: "final-sum index",
        32: "exchange buffer", 256: "guard region overwritten", 512: "arrival counter not reset"}


def main():
    pkg = importlib.import_module("monte-carlo-options-simd_b200")
    pkg.build(force=True)
    lib = pkg._native.lib()
    lib.mcb200_debug_check.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    lines, failures = [], 0

    def check(ctx, what):
        nonlocal failures
        bits = C.c_uint()
        pkg._native.check(lib.mcb200_debug_check(ctx._h, C.byref(bits)))
        bad = [name for b, name in BITS.items() if bits.value & b]
        failures += bool(bad)
        lines.append("%-64s %s" % (what, "FAILED: " + ", ".join(bad) if bad else "clean"))

    import torch
    dev = torch.device("cuda", 0)
    spots = np.arange(50, 162, dtype=np.float32)
    with pkg.Context(device=0, seed=11) as a, pkg.Context(device=0, seed=11) as b:
        def both(what, fn):
            ra, rb = fn(a), fn(b)
            nonlocal failures
            same = all(np.array_equal(ra[k], rb[k]) for k in ra)
            failures += (not same)
            check(a, what + ("" if same else "  [REPLAY DIFFERS]"))
        both("C2-shaped price batch, 112 x 20000 x 20 (42 warps per option)", lambda c: c.price_batch(spots, *G, 20.0, 20000.0))
        both("antithetic, 16 x 4000 x 20", lambda c: c.price_batch(spots[:16], *G, 20.0, 4000.0, antithetic=True))
        both("warps crossing options, ragged rounds: 41 x 8008 x 10", lambda c: c.price_batch(np.linspace(80, 140, 41).astype(np.float32), *G, 10.0, 8008.0))
        both("solo options: 20000 x 64 x 4", lambda c: c.price_batch(np.linspace(80, 140, 20000).astype(np.float32), *G, 4.0, 64.0))
        both("one option over every warp: 1 x 2e6 x 10", lambda c: c.price_batch([100.0], *G, 10.0, 2e6))
        both("Greeks, 32 x 8000 x 20", lambda c: c.greeks_batch(spots[40:72], *G, 20.0, 8000.0))
        both("Greeks solo options, 3000 x 512 x 8", lambda c: c.greeks_batch(np.linspace(80, 140, 3000).astype(np.float32), *G, 8.0, 512.0))
        both("zero paths (5 trials) and one chunk (8 trials)", lambda c: {**c.price_batch(spots[:3], *G, 10.0, 5.0),
                                                                     **{"x" + k: v for k, v in c.price_batch(spots[:3], *G, 10.0, 8.0).items()}})
        rng = np.random.default_rng(3)
        for i in range(40):
            n = int(rng.choice([1, 2, 7, 33, 112, 500, 4097]))
            trials = float(rng.choice([8, 40, 1000, 8008, 50000]))
            steps = float(rng.choice([2, 7, 20, 101]))
            sp = rng.uniform(60, 160, n).astype(np.float32)
            greeks = bool(rng.integers(0, 2))
            both("soak %2d: %s n=%d trials=%g steps=%g" % (i, "greeks" if greeks else "price", n, trials, steps),
                 (lambda c: c.greeks_batch(sp, *G, steps, trials)) if greeks else (lambda c: c.price_batch(sp, *G, steps, trials)))
    # the cross-GPU arrival protocol: two same-process ranks on device 0 (mcb200_xgpu_attach)
    n = 6
    cols = [torch.full((n,), v, dtype=torch.float32, device=dev) for v in (100.0,) + G]
    ptrs = [c.data_ptr() for c in cols]
    s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    with pkg.Context(device=0, seed=5, substream_block=0) as c0, pkg.Context(device=0, seed=5, substream_block=1) as c1:
        c0.xgpu_create(n, 2); c1.xgpu_attach(c0, n, 2)
        first = None
        for rnd in range(4):
            c0.set_seed(5, 0); c1.set_seed(5, 1)
            order = [(c0, 0, s0), (c1, 1, s1)] if rnd % 2 == 0 else [(c1, 1, s1), (c0, 0, s0)]
            for c, rank, st in order:
                c.price_xgpu_device(n, ptrs, 20.0, 8000.0, rank, 16000.0, 16000.0, stream=st.cuda_stream)
            r0, r1 = c0.xgpu_results(n), c1.xgpu_results(n)
            same = all(np.array_equal(r0[k], r1[k]) for k in r0) and (first is None or all(np.array_equal(r0[k], first[k]) for k in r0))
            first = first or r0
            failures += (not same)
            check(c0, "exchange buffer round %d, rank 0%s" % (rnd, "" if same else "  [RESULT DIFFERS]"))
            check(c1, "exchange buffer round %d, rank 1" % rnd)
        c1.xgpu_close(); c0.xgpu_close()
    with pkg.Context(devices=[0, 0], seed=9) as multi:
        multi.price_batch(np.linspace(70, 150, 64).astype(np.float32), *G, 20.0, 80000.0)
        multi.greeks_batch(spots[:5], *G, 20.0, 400000.0)
        for d in range(2):
            check(multi.child(d), "multi-device context, device slot %d (options-major + trials-minor)" % d)
    # self-test of the detector: with a falsified slot capacity the in-kernel slot check must fire
    os.environ["MCB200_CHECK_INJECT"] = "1"
    with pkg.Context(device=0, seed=1) as c:
        c.price_batch(spots, *G, 20.0, 20000.0)
        bits = C.c_uint()
        pkg._native.check(lib.mcb200_debug_check(c._h, C.byref(bits)))
        detected = bool(bits.value & 4)
        failures += (not detected)
        lines.append("%-64s %s" % ("detector self-test (slot capacity falsified to 1)", "fires as expected" if detected else "DID NOT FIRE"))
    del os.environ["MCB200_CHECK_INJECT"]
    lines.append("")
    lines.append("%d batches checked, %d failures" % (len(lines) - 2, failures))
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "checked_build.txt"), "w") as f:
        f.write("# tools/checked_build.py: -DMCB_CHECKED build (guard regions + in-kernel index checks + arrival counters),\n"
                "# every batch replayed on a second context and compared bit for bit\n" + text + "\n")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
