"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package never does; it fails loudly if its CUDA library is missing instead of falling back here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_f, _d, _u64p = C.c_float, C.c_double, C.POINTER(C.c_uint64)
_fp, _dp, _u8p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)


def build(force=False):
    """Compile liboracle.so / libcpubaseline.so with the committed Makefile (gcc only)."""
    need = force or not all(os.path.exists(os.path.join(_DIR, n)) for n in ("liboracle.so", "libcpubaseline.so"))
    if not need:
        srcs = [os.path.join(_DIR, n) for n in os.listdir(_DIR) if n.endswith((".c", ".h"))]
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(os.path.getmtime(os.path.join(_DIR, n)) < newest for n in ("liboracle.so", "libcpubaseline.so"))
    if need:
        subprocess.run(["make", "-C", _DIR, "-B", "all"], check=True, capture_output=True)


_lib = None
_base = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_DIR, "liboracle.so"))
        L = _lib
        price_args = [_f] * 8 + [_u8p]
        for name in ("call_price", "put_price", "call_price_av", "put_price_av"):
            fn = getattr(L, "ref_mc_" + name); fn.restype = _f; fn.argtypes = price_args
        for name in ("call_delta", "put_delta", "gamma", "vega", "call_rho", "put_rho", "call_theta", "put_theta"):
            fn = getattr(L, "ref_mc_" + name); fn.restype = _f; fn.argtypes = [_f] * 9 + [_u8p]
        L.ref_mc_pricing.restype = _f
        L.ref_mc_pricing.argtypes = [_f] * 9 + [_u8p, _dp, _dp, _fp, _fp]
        L.ref_mc_set_conventions.restype = None
        L.ref_mc_set_conventions.argtypes = [C.c_int, C.c_int]
        L.ref_mc_chunk_sums.restype = None
        L.ref_mc_chunk_sums.argtypes = [_u8p, C.c_int, _fp]
        L.ref_uniform_lane_stream.restype = None
        L.ref_uniform_lane_stream.argtypes = [_u8p, C.c_int, C.c_int, _fp]
        L.ref_xoshiro_outputs.restype = None
        L.ref_xoshiro_outputs.argtypes = [_u64p, C.c_int, _u64p]
        L.ref_xoshiro_apply_poly.restype = None
        L.ref_xoshiro_apply_poly.argtypes = [_u64p, _u64p]
        L.ref_splitmix64.restype = C.c_uint64
        L.ref_splitmix64.argtypes = [_u64p]
        L.fast_model_stream_paths.restype = None
        L.fast_model_stream_paths.argtypes = [_u64p, C.c_int, C.c_int, _fp]
        L.mc_scalar_price.restype = _f
        L.mc_scalar_price.argtypes = [_f] * 9 + [C.c_uint64, _dp, _dp]
        L.fast_model_tri_terms.restype = None
        L.fast_model_tri_terms.argtypes = [_u64p, C.c_int, _fp, _fp]
        L.fast_model_constants.restype = None
        L.fast_model_constants.argtypes = [_f] * 5 + [_fp, _fp]
        L.fast_model_growth.restype = _f
        L.fast_model_growth.argtypes = [_f, _f, _f]
        for name in BS_NAMES:
            fn = getattr(L, "bsf_" + name); fn.restype = _f; fn.argtypes = [_f] * 6
            fn = getattr(L, "bsd_" + name); fn.restype = _d; fn.argtypes = [_d] * 6
    return _lib


BS_NAMES = ("call_price", "put_price", "call_delta", "put_delta", "gamma", "vega", "call_rho", "put_rho",
            "call_theta", "put_theta")


def _seeds_ptr(seeds):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
    return seeds, seeds.ctypes.data_as(_u8p)


def chunk_seeds(n_chunks, seed):
    """256 random bytes per 8-path chunk, like thread_rng().fill_bytes (mc_simd.rs:52-54) but reproducible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(n_chunks, 256), dtype=np.uint8)


def n_chunks(num_trials):
    return int(num_trials) // 8


def mc(name, *args, seeds):
    """Call ref_mc_<name>(*args, seeds): the reference's public function `name` with injected chunk seeds."""
    keep, ptr = _seeds_ptr(seeds)
    return float(getattr(lib(), "ref_mc_" + name)(*[float(a) for a in args], ptr))


def pricing_detail(spot, strike, vol, r, years, q, steps, num_trials, call_mult, seeds, want_paths=False):
    """ref_mc_pricing with fp64 mean / SE and optionally the per-path growth factors and raw sums."""
    keep, ptr = _seeds_ptr(seeds)
    mean, se = C.c_double(), C.c_double()
    n = n_chunks(num_trials) * 8
    growth = np.empty(n, np.float32) if want_paths else None
    msum = np.empty(n, np.float32) if want_paths else None
    price = lib().ref_mc_pricing(spot, strike, vol, r, years, q, steps, num_trials, call_mult, ptr,
                                 C.byref(mean), C.byref(se),
                                 growth.ctypes.data_as(_fp) if want_paths else None,
                                 msum.ctypes.data_as(_fp) if want_paths else None)
    return dict(price=float(price), mean=mean.value, se=se.value, growth=growth, m=msum)


def bs32(name, s, k, v, r, t, q):
    return float(getattr(lib(), "bsf_" + name)(s, k, v, r, t, q))


def bs64(name, s, k, v, r, t, q):
    return float(getattr(lib(), "bsd_" + name)(s, k, v, r, t, q))


def xoshiro_outputs(state, n):
    st = np.array(state, dtype=np.uint64)
    out = np.empty(n, np.uint64)
    lib().ref_xoshiro_outputs(st.ctypes.data_as(_u64p), n, out.ctypes.data_as(_u64p))
    return out, st


def xoshiro_apply_poly(state, poly):
    st = np.array(state, dtype=np.uint64)
    pl = np.array(poly, dtype=np.uint64)
    lib().ref_xoshiro_apply_poly(st.ctypes.data_as(_u64p), pl.ctypes.data_as(_u64p))
    return st


def set_conventions(seed_layout=0, uniform_mode=0):
    """Switch the two simd_rand conventions the oracle has to assume (ref_mc.c): seed layout 0 = struct-of-arrays,
    1 = array-of-structs; uniform mode 0 = (x >> 11) * 2^-53, 1 = exponent trick.  (0, 0) is the default everywhere."""
    lib().ref_mc_set_conventions(int(seed_layout), int(uniform_mode))


def uniform_lane_stream(seed256, lane, n):
    keep, ptr = _seeds_ptr(seed256)
    out = np.empty(n, np.float32)
    lib().ref_uniform_lane_stream(ptr, lane, n, out.ctypes.data_as(_fp))
    return out


def fast_stream_paths(state, half_steps, n_paths):
    st = np.array(state, dtype=np.uint64)
    out = np.empty(n_paths, np.float32)
    lib().fast_model_stream_paths(st.ctypes.data_as(_u64p), int(half_steps), int(n_paths), out.ctypes.data_as(_fp))
    return out, st


def mc_scalar(spot, strike, vol, r, years, q, steps, num_trials, call_mult=1.0, seed=1):
    """The reference's scalar pricer src/mc.rs (one normal and one exp per step, multiplicative path), seedable:
    dict(price=f32 as mc.rs returns it, mean=fp64 mean, se=fp64 standard error)."""
    mean, se = C.c_double(), C.c_double()
    price = lib().mc_scalar_price(spot, strike, vol, r, years, q, steps, num_trials, call_mult, int(seed),
                                  C.byref(mean), C.byref(se))
    return dict(price=float(price), mean=mean.value, se=se.value)


def fast_tri_terms(state, n):
    """The three Box-Muller terms (pairs A, B, C) of n consecutive triples of the production convention and their
    angles in turns modulo 1: ([n, 3], [n, 3], state after)."""
    st = np.array(state, dtype=np.uint64)
    out = np.empty((n, 3), np.float32)
    turns = np.empty((n, 3), np.float32)
    lib().fast_model_tri_terms(st.ctypes.data_as(_u64p), int(n), out.ctypes.data_as(_fp), turns.ctypes.data_as(_fp))
    return out, turns, st


def fast_constants(vol, r, q, years, steps):
    c2, d2 = C.c_float(), C.c_float()
    lib().fast_model_constants(vol, r, q, years, steps, C.byref(c2), C.byref(d2))
    return c2.value, d2.value


def fast_growth(m, c2, d2):
    return np.array([lib().fast_model_growth(float(x), c2, d2) for x in np.atleast_1d(m)], np.float32)


# ---------------------------------------------------------------- timed CPU baseline (kind "port")
def baseline_lib():
    global _base
    if _base is None:
        build()
        _base = C.CDLL(os.path.join(_DIR, "libcpubaseline.so"))
        _base.cpu_baseline_price_loop.restype = _d
        _base.cpu_baseline_price_loop.argtypes = [C.c_int] + [_fp] * 6 + [_f, _f, C.c_int, C.c_uint64, _fp]
        _base.cpu_baseline_threads.restype = C.c_int
        _base.cpu_baseline_set_threads.restype = None
        _base.cpu_baseline_set_threads.argtypes = [C.c_int]
    return _base


def baseline_price_loop(spot, strike, vol, r, years, q, steps, num_trials, kind, seed=1234):
    """Serial loop of reference-style calls on all host threads; returns (prices, path_steps)."""
    arrs = [np.ascontiguousarray(a, np.float32) for a in (spot, strike, vol, r, years, q)]
    n = len(arrs[0])
    out = np.empty(n, np.float32)
    ps = baseline_lib().cpu_baseline_price_loop(n, *[a.ctypes.data_as(_fp) for a in arrs], steps, num_trials,
                                                kind, seed, out.ctypes.data_as(_fp))
    return out, ps


def baseline_threads():
    return int(baseline_lib().cpu_baseline_threads())


def baseline_use_all_cores():
    """Set the OpenMP team to every core this process may run on (torchrun exports OMP_NUM_THREADS=1); returns it."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    baseline_lib().cpu_baseline_set_threads(n)
    return baseline_threads()
