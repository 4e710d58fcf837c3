#!/usr/bin/env python3
"""Generate tests/golden/*.json -- golden vectors that pin the CPU oracle.

The reference (/root/reference) is a Rust crate that cannot be built in this image (no rustc/cargo), so the golden
vectors are derived from what its own tests hold for the hot path:
  * bs_closed_form.json: every (function, parameters, tolerance) triple of the reference's tests
    (src/mc_simd.rs:781-907, src/mc.rs:75-121) with the closed-form value the test compares against, computed by an
    INDEPENDENT numpy-float32 emulation of src/bs.rs written here (not by oracle/bs.c), plus an fp64 scipy value.
  * xoshiro256pp_kat.json: known-answer outputs of xoshiro256++ from state {1,2,3,4} and the states after jump() /
    long_jump(), from the pure-Python public-domain algorithm below (not by oracle/xoshiro256pp.h).
Run from the repo root:  python tools/gen_golden.py
"""
import json
import os

import numpy as np
from scipy.stats import norm

f32 = np.float32
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ---------------------------------------------------------------- numpy-f32 emulation of bs.rs
def erf32(x):
    t = f32(np.copysign(f32(1), x)); e = f32(abs(x))
    n, a, r, i, l, d = map(f32, (0.3275911, 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429))
    u = f32(1) / (f32(1) + n * e)
    m = f32(1) - ((((d * u + l) * u + i) * u + r) * u + a) * u * f32(np.exp(f32(-e * e)))
    return f32(t * m)


def cdf32(x):
    return f32(0.5) * (f32(1) + erf32(f32(x / f32(np.sqrt(f32(2))))))


def pdf32(x):
    return f32(np.exp(f32((f32(-1) * x * x) / f32(2)))) / f32(np.sqrt(f32(2) * f32(3.14159)))


def d12(s, k, v, r, t, q):
    d1 = (f32(np.log(f32(s / k))) + (r - q + (v * v) / f32(2)) * t) / (v * f32(np.sqrt(t)))
    return f32(d1), f32(d1 - v * f32(np.sqrt(t)))


def bs32(name, s, k, v, r, t, q):
    s, k, v, r, t, q = map(f32, (s, k, v, r, t, q))
    d1, d2 = d12(s, k, v, r, t, q)
    eq, er, rt = f32(np.exp(f32(-q * t))), f32(np.exp(f32(-r * t))), f32(np.sqrt(t))
    if name == "call_price": return s * eq * cdf32(d1) - k * er * cdf32(d2)
    if name == "put_price": return k * er * (f32(1) - cdf32(d2)) - s * eq * (f32(1) - cdf32(d1))
    if name == "gamma": return (eq * pdf32(d1)) / (s * v * rt)
    if name == "call_delta": return eq * cdf32(d1)
    if name == "put_delta": return eq * (cdf32(d1) - f32(1))
    if name == "vega": return (eq * pdf32(d1) * (s * rt)) / f32(100)
    if name == "call_rho": return k * t * er * cdf32(d2) / f32(100)
    if name == "put_rho": return k * t * er * (cdf32(d2) - f32(1)) / f32(100)
    base = (-eq * s * pdf32(d1) * v) / (f32(2) * rt)
    p1, p2 = r * k * er, q * s * eq
    if name == "call_theta": return base - p1 * cdf32(d2) + p2 * cdf32(d1)
    if name == "put_theta": return base + p1 * cdf32(f32(-d2)) - p2 * cdf32(f32(-d1))
    raise KeyError(name)


def bs64(name, s, k, v, r, t, q):
    d1 = (np.log(s / k) + (r - q + 0.5 * v * v) * t) / (v * np.sqrt(t)); d2 = d1 - v * np.sqrt(t)
    eq, er = np.exp(-q * t), np.exp(-r * t)
    N, n = norm.cdf, norm.pdf
    return {
        "call_price": s * eq * N(d1) - k * er * N(d2), "put_price": k * er * N(-d2) - s * eq * N(-d1),
        "call_delta": eq * N(d1), "put_delta": -eq * N(-d1), "gamma": eq * n(d1) / (s * v * np.sqrt(t)),
        "vega": eq * n(d1) * s * np.sqrt(t) / 100, "call_rho": k * t * er * N(d2) / 100,
        "put_rho": -k * t * er * N(-d2) / 100,
        "call_theta": -eq * s * n(d1) * v / (2 * np.sqrt(t)) - r * k * er * N(d2) + q * s * eq * N(d1),
        "put_theta": -eq * s * n(d1) * v / (2 * np.sqrt(t)) + r * k * er * N(-d2) - q * s * eq * N(-d1),
    }[name]


# (reference test, mc function, bs function, (S,K,vol,r,T,q), steps, trials, bump, tolerance, strict "<" ?)
REF_TESTS = [
    ("mc_simd.rs:781 valid_price1", "call_price", "call_price", (100, 110, .25, .05, .5, .02), 100, 1000, None, 1.25, False),
    ("mc_simd.rs:789 valid_price2", "call_price", "call_price", (90, 110, .2, .05, 1, .02), 100, 10000, None, 1.25, False),
    ("mc_simd.rs:797 valid_price3", "call_price", "call_price", (130, 120, .25, .05, .5, .02), 100, 1000, None, 1.25, False),
    ("mc_simd.rs:805 valid_price_4", "put_price", "put_price", (300, 270, .2, .09, 1, 0), 100, 10000, None, 1.0, False),
    ("mc_simd.rs:813 valid_price_av1", "call_price_av", "call_price", (112, 110, .2, .05, 1, .02), 100, 10000, None, 1.0, False),
    ("mc_simd.rs:821 valid_price_av2", "call_price_av", "call_price", (112, 110, .14, .12, 1, .02), 100, 10000, None, 1.0, False),
    ("mc_simd.rs:829 valid_price_av3", "put_price_av", "put_price", (112, 110, .14, .12, 1, .02), 100, 10000, None, 1.0, False),
    ("mc_simd.rs:837 valid_call_delta", "call_delta", "call_delta", (100, 110, .25, .05, .5, .02), 100, 1000, 0.001, 0.05, True),
    ("mc_simd.rs:845 valid_put_delta", "put_delta", "put_delta", (100, 110, .25, .05, .5, .02), 100, 1000, 0.001, 0.05, True),
    ("mc_simd.rs:853 valid_gamma", "gamma", "gamma", (100, 110, .25, .05, .5, .02), 100, 10000, 0.01, 0.05, True),
    ("mc_simd.rs:861 valid_vega", "vega", "vega", (100, 110, .1, .05, .5, .02), 100, 10000, 0.01, 0.05, True),
    ("mc_simd.rs:869 valid_call_rho", "call_rho", "call_rho", (130, 120, .25, .05, .5, .02), 100, 10000, 0.01, 0.05, True),
    ("mc_simd.rs:877 valid_put_rho", "put_rho", "put_rho", (110, 120, .1, .1, .5, .02), 100, 10000, 0.01, 0.05, True),
    ("mc_simd.rs:885 valid_call_theta", "call_theta", "call_theta", (110, 120, .1, .1, .5, .02), 100, 10000, 0.001, 0.1, True),
    ("mc_simd.rs:893 valid_put_theta", "put_theta", "put_theta", (110, 120, .1, .1, .5, .02), 100, 10000, 0.001, 0.1, True),
    ("mc_simd.rs:901 valid_put_theta2", "put_theta", "put_theta", (130, 120, .1, .1, .5, .02), 100, 10000, 0.001, 0.1, True),
    # the scalar module's tests exercise the same closed form with other parameters (mc.rs:75-121)
    ("mc.rs:75 valid_price1", "call_price", "call_price", (100, 110, .25, .05, .5, 0), 100, 1000, None, 1.25, False),
    ("mc.rs:83 valid_price2", "call_price", "call_price", (105, 110, .25, .05, .5, .02), 100, 10000, None, 1.25, False),
    ("mc.rs:91 valid_price3", "call_price", "call_price", (130, 120, .25, .05, .5, .02), 100, 10000, None, 1.25, False),
    ("mc.rs:99 valid_price_4", "put_price", "put_price", (112, 110, .2, .05, 1, .02), 100, 10000, None, 1.0, False),
    ("mc.rs:107 valid_price_5", "put_price", "put_price", (50, 60, .2, .05, 1, .02), 100, 10000, None, 1.0, False),
    ("mc.rs:115 valid_price_6", "put_price", "put_price", (130, 100, .2, .1, 1, .02), 100, 10000, None, 1.0, False),
]


# ---------------------------------------------------------------- pure-Python xoshiro256++
M = (1 << 64) - 1
rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M


def xo_next(s):
    r = (rotl((s[0] + s[3]) & M, 23) + s[0]) & M
    t = (s[1] << 17) & M
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
    return r


def xo_poly(s, poly):
    acc = [0, 0, 0, 0]
    for w in poly:
        for b in range(64):
            if (w >> b) & 1:
                acc = [a ^ x for a, x in zip(acc, s)]
            xo_next(s)
    s[:] = acc


JUMP = [0x180ec6d33cfd0aba, 0xd5a61266f0c9392c, 0xa9582618e03fc9aa, 0x39abdc4529b1661c]
LONG_JUMP = [0x76e15d3efefdcbbf, 0xc5004e441c522fb3, 0x77710069854ee241, 0x39109bb02acbe635]


def main():
    os.makedirs(OUT, exist_ok=True)
    rows = []
    for ref, mcfn, bsfn, p, steps, trials, bump, tol, strict in REF_TESTS:
        rows.append(dict(reference_test=ref, mc_function=mcfn, bs_function=bsfn, params=list(map(float, p)),
                         steps=steps, trials=trials, bump=bump, tolerance=tol, strict=strict,
                         bs_f32=float(bs32(bsfn, *p)), bs_f64=float(bs64(bsfn, *map(float, p)))))
    with open(os.path.join(OUT, "bs_closed_form.json"), "w") as f:
        json.dump(dict(source="tools/gen_golden.py: numpy-float32 emulation of /root/reference/src/bs.rs + scipy fp64",
                       cases=rows), f, indent=1)
    s = [1, 2, 3, 4]
    outs = [xo_next(s) for _ in range(10)]
    j = [1, 2, 3, 4]; xo_poly(j, JUMP)
    lj = [1, 2, 3, 4]; xo_poly(lj, LONG_JUMP)
    jj = list(j); xo_poly(jj, JUMP)
    with open(os.path.join(OUT, "xoshiro256pp_kat.json"), "w") as f:
        json.dump(dict(source="tools/gen_golden.py: pure-Python xoshiro256++ 1.0 (Blackman & Vigna, public domain)",
                       state=[1, 2, 3, 4], outputs=[str(o) for o in outs],
                       state_after_10=[str(x) for x in s], after_jump=[str(x) for x in j],
                       after_two_jumps=[str(x) for x in jj], after_long_jump=[str(x) for x in lj],
                       jump_poly=[str(x) for x in JUMP], long_jump_poly=[str(x) for x in LONG_JUMP]), f, indent=1)
    for r in rows:
        print("%-34s %-14s f32 %.6f  f64 %.6f" % (r["reference_test"], r["bs_function"], r["bs_f32"], r["bs_f64"]))
    print("KAT", outs[:4])


if __name__ == "__main__":
    main()
