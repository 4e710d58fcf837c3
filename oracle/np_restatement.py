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


# ---------------------------------------------------------------------------------------------- the other five cores
# and the twelve public functions (mc_simd.rs:82-149, :152-227, :230-309, :312-387, :390-473, :477-779), from the same
# per-path sums: the reference evaluates every leg of a function from ONE stock_price_mult per path.
f = np.float32


def _terminal_sums(steps, num_trials, seeds):
    n_chunks = int(f(num_trials)) // 8
    half_steps = int(f(steps)) // 2
    words = np.frombuffer(np.ascontiguousarray(seeds, np.uint8).tobytes(), dtype="<u8").reshape(n_chunks, 4, 8)
    st = [words[:, k, :].copy() for k in range(4)]
    two_pi = f(2.0) * f(np.pi)
    m = np.zeros((n_chunks, 8), np.float32)
    for _ in range(half_steps):
        u1 = _uniform_f32(st)
        u2 = _uniform_f32(st)
        ang = two_pi * u2
        m = (np.sqrt(-np.log(u1)).astype(np.float64) * (np.sin(ang) + np.cos(ang)).astype(np.float64)
             + m.astype(np.float64)).astype(np.float32)
    return m


def _leg_total(m, spot_signed, strike_signed, drift, diff):
    """reduce_add over lanes of the chunk-summed fast_max(mul_sub(spot, exp(mul_add(m, diff, drift)), strike), 0)."""
    expo = (m.astype(np.float64) * np.float64(diff) + np.float64(drift)).astype(np.float32)
    pay = (np.float64(spot_signed) * np.exp(expo).astype(np.float64) - np.float64(strike_signed)).astype(np.float32)
    return np.maximum(pay, f(0.0)).sum(axis=0, dtype=np.float32).sum(dtype=np.float32)


def _consts(vol, r, q, years, steps):
    dt = years / steps
    nudt = (r - q - f(0.5) * (vol * vol)) * dt
    sidt = vol * np.sqrt(dt)
    return steps * nudt, f(np.sqrt(2.0)) * sidt


def av_price(spot, strike, vol, r, years, q, steps, num_trials, call_mult, seeds):            # mc_simd.rs:82-149
    spot, strike, vol, r, years, q, steps, num_trials, cm = map(f, (spot, strike, vol, r, years, q, steps, num_trials, call_mult))
    m = _terminal_sums(steps, num_trials, seeds)
    drift, diff = _consts(vol, r, q, years, steps)
    pos = _leg_total(m, cm * spot, cm * strike, drift, diff)
    neg = _leg_total(m, cm * spot, cm * strike, drift, -diff)
    return f((f(0.5) * (pos + neg)) / num_trials * np.exp(-r * years))


def spot_prices(spot, dspot, strike, vol, r, years, q, steps, num_trials, call_mult, seeds):   # :152-227
    spot, dspot, strike, vol, r, years, q, steps, num_trials, cm = map(f, (spot, dspot, strike, vol, r, years, q, steps, num_trials, call_mult))
    m = _terminal_sums(steps, num_trials, seeds)
    drift, diff = _consts(vol, r, q, years, steps)
    final = np.exp(-r * years) / num_trials
    tot = [_leg_total(m, cm * s_, strike * cm, drift, diff) for s_ in (spot - dspot, spot, spot + dspot)]
    return tuple(f(t * final) for t in tot)                                                   # (minus, mid, plus)


def vol_prices(spot, strike, vol, dvol, r, years, q, steps, num_trials, seeds):                # :230-309 (call only)
    spot, strike, vol, dvol, r, years, q, steps, num_trials = map(f, (spot, strike, vol, dvol, r, years, q, steps, num_trials))
    m = _terminal_sums(steps, num_trials, seeds)
    out = []
    for v in (vol - dvol, vol + dvol):
        drift, diff = _consts(v, r, q, years, steps)
        out.append(f(_leg_total(m, spot, strike, drift, diff) / num_trials * np.exp(-r * years)))
    return tuple(out)                                                                          # (minus, plus)


def rate_prices(spot, strike, vol, r, dr, years, q, steps, num_trials, call_mult, seeds):      # :312-387
    spot, strike, vol, r, dr, years, q, steps, num_trials, cm = map(f, (spot, strike, vol, r, dr, years, q, steps, num_trials, call_mult))
    m = _terminal_sums(steps, num_trials, seeds)
    out = []
    for rr in (r - dr, r + dr):
        drift, diff = _consts(vol, rr, q, years, steps)
        out.append(f(_leg_total(m, spot * cm, strike * cm, drift, diff) * np.exp(-rr * years) / num_trials))
    return tuple(out)


def time_prices(spot, strike, vol, r, years, dyears, q, steps, num_trials, call_mult, seeds):  # :390-473
    spot, strike, vol, r, years, dyears, q, steps, num_trials, cm = map(f, (spot, strike, vol, r, years, dyears, q, steps, num_trials, call_mult))
    m = _terminal_sums(steps, num_trials, seeds)
    out = []
    for t in (years - dyears, years + dyears):
        drift, diff = _consts(vol, r, q, t, steps)
        out.append(f(_leg_total(m, spot * cm, strike * cm, drift, diff) * np.exp(-r * t) / num_trials))
    return tuple(out)


def public(name, *args, seeds):
    """The twelve public functions of mc_simd (:477-779), positional arguments exactly as in the reference."""
    a = [f(x) for x in args]
    if name in ("call_price", "put_price"):
        return pricing(*a, 1.0 if name == "call_price" else -1.0, seeds)[0]
    if name in ("call_price_av", "put_price_av"):
        return av_price(*a, 1.0 if name == "call_price_av" else -1.0, seeds)
    if name in ("call_delta", "put_delta", "gamma"):
        spot, dspot = a[0], a[1]
        lo, mid, hi = spot_prices(spot, dspot, *a[2:], -1.0 if name == "put_delta" else 1.0, seeds)
        if name == "gamma":
            return f((hi - f(2.0) * mid + lo) / (dspot * dspot))                               # :644
        return f((hi - lo) / (f(2.0) * dspot))                                                 # :592,618
    if name == "vega":
        lo, hi = vol_prices(*a, seeds)
        return f((hi - lo) / (f(200.0) * a[3]))                                                # :670
    if name in ("call_rho", "put_rho"):
        lo, hi = rate_prices(*a, 1.0 if name == "call_rho" else -1.0, seeds)
        return f((hi - lo) / (f(200.0) * a[4]))                                                # :698,725
    if name in ("call_theta", "put_theta"):
        lo, hi = time_prices(*a, 1.0 if name == "call_theta" else -1.0, seeds)
        return f((lo - hi) / (f(2.0) * a[5]))                                                  # :752,778
    raise KeyError(name)
