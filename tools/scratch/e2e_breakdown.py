import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("monte-carlo-options-simd_b200")
N = pkg._native
import ctypes as C
import torch
spots = np.arange(50, 162, dtype=np.float32); n = len(spots)
cols = [spots] + [np.full(n, v, np.float32) for v in (110.0, 0.25, 0.05, 0.5, 0.02)]
outs = [np.empty(n, np.float32) for _ in range(4)]
with pkg.Context(device=0, seed=1) as ctx:
    fn = N.lib().mcb200_price_batch
    args = (ctx._h, n, *[N.fptr(a) for a in cols], 100.0, 20000.0, 0, *[N.fptr(o) for o in outs])
    def t(f, k=300):
        for _ in range(20): f()
        t0 = time.perf_counter()
        for _ in range(k): f()
        return (time.perf_counter() - t0) / k * 1e6
    print("python ctx.price_batch      %.1f us" % t(lambda: ctx.price_batch(*cols, 100.0, 20000.0)))
    print("prepared ctypes C-ABI call  %.1f us" % t(lambda: fn(*args)))
    dev = torch.device("cuda:0")
    cd = [torch.from_numpy(c).to(dev) for c in cols]; od = torch.zeros(4, n, device=dev)
    ip = [c.data_ptr() for c in cd]; op = [od[i].data_ptr() for i in range(4)]
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    def devcall():
        ctx.price_batch_device(n, ip, 100.0, 20000.0, op, stream=s.cuda_stream); s.synchronize()
    print("device-resident call + sync %.1f us" % t(devcall))
    for trials in (1000.0, 8.0):
        a2 = (ctx._h, n, *[N.fptr(a) for a in cols], 100.0, trials, 0, *[N.fptr(o) for o in outs])
        print("prepared call, %g trials     %.1f us" % (trials, t(lambda: fn(*a2))))
    one = (ctx._h, 1, *[N.fptr(a) for a in cols], 100.0, 1000.0, 0, *[N.fptr(o) for o in outs])
    print("prepared call, 1 option x 1000 trials  %.1f us" % t(lambda: fn(*one)))
    print("drop-in mc_simd.call_price (1000 trials) %.1f us" % t(lambda: pkg.mc_simd.call_price(100.0, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 1000.0)))
