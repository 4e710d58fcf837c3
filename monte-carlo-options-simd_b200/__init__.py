"""monte-carlo-options-simd_b200: B200-native (sm_100a) drop-in for the mc_simd hot path of
ashayp22/monte-carlo-options-simd.

    import importlib
    mcb = importlib.import_module("monte-carlo-options-simd_b200")
    mcb.mc_simd.call_price(100.0, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 1000.0)     # the reference's API
    ctx = mcb.Context(device=0, seed=1234)
    ctx.price_batch(spots, 110.0, 0.25, 0.05, 0.5, 0.02, steps=100.0, num_trials=20000.0)   # batched entry point

The compute lives in lib/libmcb200.so (CUDA, built by build.py); this package is the host-side mirror of the
reference's interface and the ctypes plumbing.  No CPU fallback.
"""
from . import _native, mc_simd, sharding  # noqa: F401
from ._native import GREEK_NAMES, LIB_PATH, Mcb200Error  # noqa: F401
from .build import build  # noqa: F401
from .context import Context, plan_query, plan_warp_range, shard_query  # noqa: F401
