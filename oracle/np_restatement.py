"""TEST INFRASTRUCTURE: a second, independent restatement of the reference's plain pricer in numpy.

Written directly from /root/reference/src/mc_simd.rs:9-21 (speed_update), :25-79 (monte_carlo_pricing) and
src/rand32x8.rs:5-21 (uniform adapter), vectorised over (chunk, lane) instead of looping, sharing no code with
oracle/ref_mc.c.  tests/test_oracle.py compares the two path by path on the same chunk seeds, so a slip in either
restatement shows up as a disagreement.  Same stated conventions as ref_mc.c for the third-party pieces
(PARITY UNPINNED there): seed word k of lane l is the little-endian u64 at byte 64*k + 8*l; uniform = (x >> 11) * 2^-53.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _rotl(x, k):
    return (x << np.uint64(k)) | (x >> np.uint64(64 - k))


def _next(s):
    """xoshiro256++ (Blackman & Vigna, public domain) on arrays of lanes; s is a list of four uint64 arrays."""
    with np.errstate(over="ignore"):
        out = _rotl(s[0] + s[3], 23) + s[0]
        t = s[1] << np.uint64(17)
        s[2] = s[2] ^ s[0]
        s[3] = s[3] ^ s[1]
        s[1] = s[1] ^ s[2]
        s[0] = s[0] ^ s[3]
        s[2] = s[2] ^ t
        s[3] = _rotl(s[3], 45)
    return out


def _uniform_f32(s):
    x = _next(s)
    return ((x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53).astype(np.float32)      # next_f64x8, then `as f32`


def pricing(spot, strike, vol, r, years, q, steps, num_trials, call_mult, seeds):
    """monte_carlo_pricing with injected chunk seeds: returns (price as f32, per-path terminal sums M as f32 [n])."""
    f = np.float32
    spot, strike, vol, r, years, q, steps, num_trials, call_mult = map(f, (spot, strike, vol, r, years, q, steps,
                                                                           num_trials, call_mult))
    dt = years / steps
    nudt = (r - q - f(0.5) * (vol * vol)) * dt
    sidt = vol * np.sqrt(dt)
    two_pi = f(2.0) * f(np.pi)
    drift = steps * nudt
    diff = f(np.sqrt(2.0)) * sidt
    n_chunks = int(num_trials) // 8
    half_steps = int(steps) // 2
    words = np.frombuffer(np.ascontiguousarray(seeds, np.uint8).tobytes(), dtype="<u8").reshape(n_chunks, 4, 8)
    s = [words[:, k, :].copy() for k in range(4)]                   # [chunk, lane]
    m = np.zeros((n_chunks, 8), np.float32)
    for _ in range(half_steps):
        u1 = _uniform_f32(s)
        u2 = _uniform_f32(s)
        ang = two_pi * u2
        # mul_add((-ln u1).sqrt(), sin + cos, acc): fused multiply-add -> do the product-sum in fp64, round once
        rad = np.sqrt(-np.log(u1))
        sc = np.sin(ang) + np.cos(ang)
        m = (rad.astype(np.float64) * sc.astype(np.float64) + m.astype(np.float64)).astype(np.float32)
    expo = (m.astype(np.float64) * np.float64(diff) + np.float64(drift)).astype(np.float32)      # mul_add
    growth = np.exp(expo)
    pay = ((call_mult * spot).astype(np.float64) * growth.astype(np.float64)
           - np.float64(call_mult * strike)).astype(np.float32)                                  # mul_sub
    pay = np.maximum(pay, f(0.0))
    total = pay.sum(axis=0, dtype=np.float32)                       # chunk vectors added lane-wise, then reduce_add
    price = total.sum(dtype=np.float32) / num_trials * np.exp(-r * years)
    return f(price), m.reshape(-1), growth.reshape(-1)
