import importlib, sys, os, math
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("monte-carlo-options-simd_b200")
from oracle import oracle
oracle.build()
GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)
spots = np.linspace(90, 130, 64).astype(np.float32)
ex = np.array([oracle.bs64("gamma", float(s), 110.0, 0.25, 0.05, 0.5, 0.02) for s in spots])
exd = np.array([oracle.bs64("call_delta", float(s), 110.0, 0.25, 0.05, 0.5, 0.02) for s in spots])
for trials in (1e5, 1e6):
    ratios = []
    for seed in range(6):
        with pkg.Context(device=0, seed=seed + 100) as ctx:
            g = ctx.greeks_batch(spots, GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"], 252.0, trials, delta_spot=0.01)
        ratios.append(g["gamma"] / ex)
        z = (g["gamma"] - ex) / np.maximum(g["gamma_se"], 5e-4)
        zd = (g["call_delta"] - exd) / g["call_delta_se"]
        print("fast trials %g seed %d: mean ratio %.4f  mean z %.2f rms z %.2f max|z| %.2f | delta mean z %.2f rms %.2f" % (
            trials, seed, ratios[-1].mean(), z.mean(), np.sqrt((z**2).mean()), np.abs(z).max(), zd.mean(), np.sqrt((zd**2).mean())))
    print("  overall mean ratio %.4f +- %.4f" % (np.mean(ratios), np.std(np.mean(ratios, axis=1)) / math.sqrt(len(ratios))))
# reference-stream mode with fresh random seeds
rng = np.random.default_rng(5)
trials = 1e5
ratios = []
for rep in range(3):
    seeds = rng.integers(0, 256, size=len(spots) * int(trials // 8) * 256, dtype=np.uint8)
    with pkg.Context(device=0, seed=1) as ctx:
        g = ctx.greeks_batch_refstream(spots, GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"], 252.0, trials, seeds, delta_spot=0.01)
    ratios.append(g["gamma"] / ex)
    z = (g["gamma"] - ex) / np.maximum(g["gamma_se"], 5e-4)
    print("ref  trials %g rep %d: mean ratio %.4f mean z %.2f rms z %.2f max|z| %.2f" % (trials, rep, ratios[-1].mean(), z.mean(), np.sqrt((z**2).mean()), np.abs(z).max()))
