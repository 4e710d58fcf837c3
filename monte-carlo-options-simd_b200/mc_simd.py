"""mc_simd -- Python mirror of the reference's public module (/root/reference/src/mc_simd.rs:477-779).

Same twelve function names, same positional f32 arguments in the same order, same meaning and error behaviour
(a bare float comes back; NaN on failure, with the reason in last_error()).  Each call goes straight through the C ABI
drop-in of the same name in libmcb200.so and runs on the B200; nothing here computes on the CPU.
"""
import math

from . import _native as N


def _call(name, *args):
    return float(getattr(N.lib(), "mc_simd_" + name)(*[float(a) for a in args]))


def call_price(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials):
    return _call("call_price", spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                 num_trials)


def put_price(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials):
    return _call("put_price", spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                 num_trials)


def call_price_av(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials):
    return _call("call_price_av", spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                 num_trials)


def put_price_av(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials):
    return _call("put_price_av", spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                 num_trials)


def call_delta(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
               num_trials):
    return _call("call_delta", spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield,
                 steps, num_trials)


def put_delta(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
              num_trials):
    return _call("put_delta", spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield,
                 steps, num_trials)


def gamma(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials):
    return _call("gamma", spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield,
                 steps, num_trials)


def vega(spot, strike, volatility, delta_volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
         num_trials):
    return _call("vega", spot, strike, volatility, delta_volatility, risk_free_rate, years_to_expiry, dividend_yield,
                 steps, num_trials)


def call_rho(spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry, dividend_yield, steps,
             num_trials):
    return _call("call_rho", spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry,
                 dividend_yield, steps, num_trials)


def put_rho(spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry, dividend_yield, steps,
            num_trials):
    return _call("put_rho", spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry,
                 dividend_yield, steps, num_trials)


def call_theta(spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry, dividend_yield,
               steps, num_trials):
    return _call("call_theta", spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry,
                 dividend_yield, steps, num_trials)


def put_theta(spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry, dividend_yield,
              steps, num_trials):
    return _call("put_theta", spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry,
                 dividend_yield, steps, num_trials)


def last_error():
    return N.last_error()


def reset_default_context(device=0, seed=0x5EED5EED5EED5EED, substream_block=0):
    """Recreate the process-wide context the twelve functions run on (fixes the seed for reproducible runs)."""
    N.check(N.lib().mcb200_default_ctx_reset(int(device), int(seed) & (2 ** 64 - 1), int(substream_block)))


def isnan(x):
    return math.isnan(x)


__all__ = ["call_price", "put_price", "call_price_av", "put_price_av", "call_delta", "put_delta", "gamma", "vega",
           "call_rho", "put_rho", "call_theta", "put_theta", "last_error", "reset_default_context"]
