#!/usr/bin/env python3
"""Accuracy columns of the BASELINE.md results table: max |z| of the B200 estimates against the bs.rs closed form and
against the CPU oracle (the restatement of the reference algorithm, independent seeds), per config.  Run on the GPU box."""
import importlib
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pkg = importlib.import_module("monte-carlo-options-simd_b200")
from oracle import oracle   # noqa: E402  (checker only)

K, V, R, T, Q = 110.0, 0.25, 0.05, 0.5, 0.02
FLOOR = dict(call=2e-3, put=2e-3, call_delta=2e-4, put_delta=2e-4, gamma=5e-4, vega=2e-4, call_rho=2e-4, put_rho=2e-4,
             call_theta=5e-3, put_theta=5e-3)


def true_price_se(spots, n_paths, num_trials):
    """Exact standard error of the discounted call / put payoff mean over n_paths lognormal draws (closed form second
    moment): the sample SE is useless where only a handful of the paths end in the money (config 1, far OTM)."""
    from scipy.stats import norm
    s0 = np.asarray(spots, np.float64)
    sd = V * math.sqrt(T)
    d2 = (np.log(s0 / K) + (R - Q - 0.5 * V * V) * T) / sd
    d1, d3 = d2 + sd, d2 + 2 * sd
    fwd, disc = s0 * math.exp((R - Q) * T), math.exp(-R * T)
    out = {}
    for key, sgn in (("call", 1.0), ("put", -1.0)):
        m1 = sgn * (fwd * norm.cdf(sgn * d1) - K * norm.cdf(sgn * d2))
        m2 = fwd ** 2 * math.exp(V * V * T) * norm.cdf(sgn * d3) - 2 * K * fwd * norm.cdf(sgn * d1) + K * K * norm.cdf(sgn * d2)
        out[key] = np.sqrt(np.maximum(m2 - m1 ** 2, 0.0) / n_paths) * disc * n_paths / num_trials
    return out


def z_vs_bs(res, bs, keys, true_se=None):
    out = {}
    for k in keys:
        se = np.maximum(res[k + "_se"], FLOOR[k])
        if true_se is not None and k in true_se:
            se = np.maximum(se, true_se[k])
        z = np.abs(res[k].astype(np.float64) - bs[k]) / se
        out[k] = float(z.max())
    return out


def main():
    spots = np.arange(50, 162, dtype=np.float32)
    rows = {}
    with pkg.Context(device=0, seed=20241019) as ctx:
        bs = ctx.bs_batch(spots, K, V, R, T, Q)
        for name, steps, trials, av in (("C1", 100.0, 1000.0, False), ("C2", 100.0, 20000.0, False), ("C3", 1000.0, 2000.0, True)):
            res = ctx.price_batch(spots, K, V, R, T, Q, steps, trials, antithetic=av)
            rows[name] = dict(vs_bs=z_vs_bs(res, bs, ("call", "put"), true_price_se(spots, 8 * (int(trials) // 8), trials)))
            if name == "C2":                                  # vs the oracle on a subset of the grid (CPU, scalar C)
                zs = []
                for i in range(0, 112, 8):
                    seeds = oracle.chunk_seeds(oracle.n_chunks(trials), 1000 + i)
                    ref = oracle.pricing_detail(float(spots[i]), K, V, R, T, Q, steps, trials, 1.0, seeds)
                    zs.append(abs(float(res["call"][i]) - ref["mean"]) / max(math.hypot(float(res["call_se"][i]), ref["se"]), 2e-3))
                rows[name]["vs_oracle_call_subset"] = max(zs)
        s4 = (50.0 + 112.0 * np.arange(4096) / 4096.0).astype(np.float32)
        g = ctx.greeks_batch(s4, K, V, R, T, Q, 252.0, 1e5)
        b4 = ctx.bs_batch(s4, K, V, R, T, Q)
        keys = [k for k in FLOOR if k != "gamma"]
        rows["C4"] = dict(vs_bs=z_vs_bs(g, b4, keys))
        # gamma: use the analytic SE of the rare-event estimator where the sample SE is unreliable (tests explain why)
        true_se = np.sqrt(b4["gamma"].astype(np.float64) * math.exp(-R * T) * 2.0 * K / (3.0 * s4 * 0.01) / 1e5)
        zg = np.abs(g["gamma"].astype(np.float64) - b4["gamma"]) / np.maximum(np.maximum(g["gamma_se"], true_se), FLOOR["gamma"])
        rows["C4"]["vs_bs"]["gamma"] = float(zg.max())
    print("Z_JSON " + json.dumps(rows))


if __name__ == "__main__":
    main()
