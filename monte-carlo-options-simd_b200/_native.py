"""ctypes binding of libmcb200.so (the C ABI in include/mcb200.h).

Fails loudly: if the library has not been built, or no sm_100 CUDA device is present, calls raise -- there is no
CPU fallback anywhere in this package (the CPU oracle under oracle/ is test infrastructure and is never imported here).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", os.environ.get("MCB200_LIB_NAME", "libmcb200.so"))   # alternate: tuning builds

N_GREEK_OUTPUTS = 10
GREEK_NAMES = ("call", "put", "call_delta", "put_delta", "gamma", "vega", "call_rho", "put_rho", "call_theta",
               "put_theta")
ANTITHETIC = 1
IPC_HANDLE_BYTES = 64

_f, _fp, _dp, _vp = C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p


class Mcb200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libmcb200 status %d: %s" % (status, message))
        self.status = status


class Info(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("ctas_per_sm", C.c_int), ("threads_per_cta", C.c_int),
                ("grid_slots", C.c_int), ("grid_slots_greeks", C.c_int), ("pool_threads", C.c_int64), ("kernel_launches", C.c_uint64),
                ("seed", C.c_uint64), ("substream_block", C.c_uint32), ("sm_clock_khz", C.c_int)]


class Plan(C.Structure):
    _fields_ = [("n_paths", C.c_int), ("half_steps", C.c_int), ("rounds_per_option", C.c_int), ("warps", C.c_int),
                ("grid", C.c_int), ("max_rounds_per_warp", C.c_int), ("total_rounds", C.c_int64),
                ("path_steps", C.c_double), ("waves", C.c_int)]


class Shard(C.Structure):
    _fields_ = [("mode", C.c_int), ("n_shards", C.c_int), ("shard", C.c_int), ("active", C.c_int),
                ("option_begin", C.c_int), ("option_end", C.c_int), ("num_trials_local", C.c_float),
                ("total_paths", C.c_double)]


SHARD_MODES = ("single", "options", "trials")


class Bumps(C.Structure):
    _fields_ = [("delta_spot", _f), ("delta_volatility", _f), ("delta_risk_free_rate", _f),
                ("delta_years_to_expiry", _f)]


class PipeRate(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ops_per_clk_per_sm", C.c_double), ("elapsed_ms", C.c_double),
                ("sm_clock_mhz", C.c_double)]


_GreekOut = _fp * N_GREEK_OUTPUTS
_VoidOut = _vp * N_GREEK_OUTPUTS

_lib = None

DROPIN_PRICE = ("call_price", "put_price", "call_price_av", "put_price_av")
DROPIN_GREEK = ("call_delta", "put_delta", "gamma", "vega", "call_rho", "put_rho", "call_theta", "put_theta")


def lib():
    """Load libmcb200.so (building is the job of build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Mcb200Error(-1, "%s not found: run `python monte-carlo-options-simd_b200/build.py` "
                              "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    six = [_fp] * 6
    L.mcb200_ctx_create.restype = C.c_int
    L.mcb200_ctx_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_uint64, C.c_uint32]
    L.mcb200_ctx_destroy.restype = None
    L.mcb200_ctx_destroy.argtypes = [_vp]
    L.mcb200_ctx_set_seed.restype = C.c_int
    L.mcb200_ctx_set_seed.argtypes = [_vp, C.c_uint64, C.c_uint32]
    L.mcb200_last_error.restype = C.c_char_p
    L.mcb200_last_error.argtypes = []
    L.mcb200_ctx_info.restype = C.c_int
    L.mcb200_ctx_info.argtypes = [_vp, C.POINTER(Info)]
    L.mcb200_plan_query.restype = C.c_int
    L.mcb200_plan_query.argtypes = [C.c_int, _f, _f, C.c_int, C.POINTER(Plan)]
    L.mcb200_plan_warp_range.restype = C.c_int
    L.mcb200_plan_warp_range.argtypes = [C.POINTER(Plan), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mcb200_price_batch.restype = C.c_int
    L.mcb200_price_batch.argtypes = [_vp, C.c_int] + six + [_f, _f, C.c_uint] + [_fp] * 4
    L.mcb200_greeks_batch.restype = C.c_int
    L.mcb200_greeks_batch.argtypes = [_vp, C.c_int] + six + [_f, _f, C.POINTER(Bumps), _GreekOut, _GreekOut]
    L.mcb200_price_batch_device.restype = C.c_int
    L.mcb200_price_batch_device.argtypes = [_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.c_uint] + [_vp] * 4 + [_vp]
    L.mcb200_greeks_batch_device.restype = C.c_int
    L.mcb200_greeks_batch_device.argtypes = [_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.POINTER(Bumps), _VoidOut,
                                                                          _VoidOut, _vp]
    L.mcb200_price_moments_device.restype = C.c_int
    L.mcb200_price_moments_device.argtypes = [_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.c_uint, _vp, _vp]
    L.mcb200_price_finalize_device.restype = C.c_int
    L.mcb200_price_finalize_device.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_double, C.c_double] + [_vp] * 4 + [_vp]
    L.mcb200_greeks_moments_device.restype = C.c_int
    L.mcb200_greeks_moments_device.argtypes = [_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.POINTER(Bumps), _vp, _vp]
    L.mcb200_greeks_finalize_device.restype = C.c_int
    L.mcb200_greeks_finalize_device.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_double, C.c_double, C.POINTER(Bumps),
                                                _VoidOut, _VoidOut, _vp]
    L.mcb200_xgpu_create.restype = C.c_int
    L.mcb200_xgpu_create.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.mcb200_xgpu_open.restype = C.c_int
    L.mcb200_xgpu_open.argtypes = [_vp, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int]
    L.mcb200_xgpu_attach.restype = C.c_int
    L.mcb200_xgpu_attach.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int]
    L.mcb200_xgpu_set_timeout.restype = C.c_int
    L.mcb200_xgpu_set_timeout.argtypes = [_vp, C.c_double]
    L.mcb200_ctx_create_multi.restype = C.c_int
    L.mcb200_ctx_create_multi.argtypes = [C.POINTER(_vp), C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_uint32]
    L.mcb200_ctx_device_count.restype = C.c_int
    L.mcb200_ctx_device_count.argtypes = [_vp]
    L.mcb200_ctx_child.restype = C.c_int
    L.mcb200_ctx_child.argtypes = [_vp, C.c_int, C.POINTER(_vp)]
    L.mcb200_shard_query.restype = C.c_int
    L.mcb200_shard_query.argtypes = [C.c_int, _f, _f, C.c_int, C.c_int, C.c_int, C.POINTER(Shard)]
    L.mcb200_xgpu_close.restype = C.c_int
    L.mcb200_xgpu_close.argtypes = [_vp]
    L.mcb200_price_xgpu_device.restype = C.c_int
    L.mcb200_price_xgpu_device.argtypes = ([_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.c_uint, C.c_int, C.c_double,
                                                                           C.c_double, _vp])
    L.mcb200_xgpu_results.restype = C.c_int
    L.mcb200_xgpu_results.argtypes = [_vp, C.c_int] + [_fp] * 4
    L.mcb200_greeks_xgpu_device.restype = C.c_int
    L.mcb200_greeks_xgpu_device.argtypes = ([_vp, C.c_int] + [_vp] * 6 + [_f, _f, C.POINTER(Bumps), C.c_int, C.c_double,
                                                                            C.c_double, _vp])
    L.mcb200_xgpu_results_greeks.restype = C.c_int
    L.mcb200_xgpu_results_greeks.argtypes = [_vp, C.c_int, _GreekOut, _GreekOut]
    L.mcb200_bs_batch.restype = C.c_int
    L.mcb200_bs_batch.argtypes = [_vp, C.c_int] + six + [_GreekOut]
    L.mcb200_bs_batch_device.restype = C.c_int
    L.mcb200_bs_batch_device.argtypes = [_vp, C.c_int] + [_vp] * 6 + [_VoidOut, _vp]
    L.mcb200_price_batch_refstream.restype = C.c_int
    L.mcb200_price_batch_refstream.argtypes = ([_vp, C.c_int] + six + [_f, _f, C.c_uint, C.POINTER(C.c_uint8)]
                                               + [_fp] * 5)
    L.mcb200_greeks_batch_refstream.restype = C.c_int
    L.mcb200_greeks_batch_refstream.argtypes = ([_vp, C.c_int] + six + [_f, _f, C.POINTER(Bumps),
                                                                         C.POINTER(C.c_uint8), _GreekOut, _GreekOut])
    L.mcb200_price_batch_dump.restype = C.c_int
    L.mcb200_price_batch_dump.argtypes = [_vp, C.c_int] + six + [_f, _f, C.c_uint] + [_fp] * 5
    L.mcb200_ctx_export_states.restype = C.c_int
    L.mcb200_ctx_export_states.argtypes = [_vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
    L.mcb200_ctx_set_timing.restype = C.c_int
    L.mcb200_ctx_set_timing.argtypes = [_vp, C.c_int]
    L.mcb200_ctx_last_sim_ms.restype = C.c_int
    L.mcb200_ctx_last_sim_ms.argtypes = [_vp, _fp]
    L.mcb200_microbench.restype = C.c_int
    L.mcb200_microbench.argtypes = [_vp, C.POINTER(PipeRate), C.c_int]
    L.mcb200_default_ctx.restype = C.c_int
    L.mcb200_default_ctx.argtypes = [C.POINTER(_vp)]
    L.mcb200_default_ctx_reset.restype = C.c_int
    L.mcb200_default_ctx_reset.argtypes = [C.c_int, C.c_uint64, C.c_uint32]
    for name in DROPIN_PRICE:
        fn = getattr(L, "mc_simd_" + name); fn.restype = _f; fn.argtypes = [_f] * 8
    for name in DROPIN_GREEK:
        fn = getattr(L, "mc_simd_" + name); fn.restype = _f; fn.argtypes = [_f] * 9
    _lib = L
    return L


def last_error():
    return lib().mcb200_last_error().decode("utf-8", "replace")


def check(status):
    if status != 0:
        raise Mcb200Error(status, last_error())


def as_f32(a, n=None):
    arr = np.ascontiguousarray(a, dtype=np.float32)
    if arr.ndim == 0:
        arr = np.full(n if n is not None else 1, float(arr), np.float32)
    return arr


def fptr(arr):
    return arr.ctypes.data_as(_fp) if arr is not None else None
