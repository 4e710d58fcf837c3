"""Context: one device + stream + pool of per-thread xoshiro256++ substreams (wraps mcb200_ctx)."""
import ctypes as C

import numpy as np

from . import _native as N


class Context:
    """Owns an mcb200_ctx.  All batched entry points of the C ABI are methods here.

    substream_block selects the long_jump block of generator substreams: use the global rank of the GPU so that
    GPUs / processes never share draws.
    """

    def __init__(self, device=0, seed=0x5EED5EED5EED5EED, substream_block=0, devices=None, _borrowed=None):
        """devices=[0, 1, ...] creates a multi-device context (mcb200_ctx_create_multi): the host-buffer methods
        (price_batch, greeks_batch, bs_batch, prepare_*) then split every batch over those GPUs; device i draws from
        substream block substream_block + i."""
        self._h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)
        elif devices is not None:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            N.check(N.lib().mcb200_ctx_create_multi(C.byref(self._h), ids, len(devices), int(seed) & (2 ** 64 - 1),
                                                    int(substream_block)))
        else:
            N.check(N.lib().mcb200_ctx_create(C.byref(self._h), int(device), int(seed) & (2 ** 64 - 1),
                                              int(substream_block)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self._owned:
                N.lib().mcb200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def device_count(self):
        return int(N.lib().mcb200_ctx_device_count(self._h))

    def child(self, i):
        """Borrowed per-device context of a multi-device context (valid while the parent lives)."""
        h = C.c_void_p()
        N.check(N.lib().mcb200_ctx_child(self._h, int(i), C.byref(h)))
        return Context(_borrowed=h.value)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ inspection
    def info(self):
        inf = N.Info()
        N.check(N.lib().mcb200_ctx_info(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in N.Info._fields_}

    def set_seed(self, seed, substream_block=0):
        N.check(N.lib().mcb200_ctx_set_seed(self._h, int(seed) & (2 ** 64 - 1), int(substream_block)))

    def plan(self, n_options, steps, num_trials, greeks=False):
        """Launch plan of a batch on this device (the Greeks kernel runs fewer, fatter CTAs per SM)."""
        return plan_query(n_options, steps, num_trials, self.info()["grid_slots_greeks" if greeks else "grid_slots"])

    def export_states(self, first, count):
        out = np.empty((count, 4), np.uint64)
        N.check(N.lib().mcb200_ctx_export_states(self._h, first, count, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def set_timing(self, enabled=True):
        N.check(N.lib().mcb200_ctx_set_timing(self._h, 1 if enabled else 0))

    def last_sim_ms(self):
        ms = C.c_float()
        N.check(N.lib().mcb200_ctx_last_sim_ms(self._h, C.byref(ms)))
        return ms.value

    def microbench(self):
        buf = (N.PipeRate * 128)()
        n = N.lib().mcb200_microbench(self._h, buf, 128)
        if n < 0:
            raise N.Mcb200Error(-n, N.last_error())
        return [dict(name=buf[i].name.decode(), ops_per_clk_per_sm=buf[i].ops_per_clk_per_sm,
                     elapsed_ms=buf[i].elapsed_ms, sm_clock_mhz=buf[i].sm_clock_mhz) for i in range(n)]

    # ------------------------------------------------------------------ host-buffer batches
    @staticmethod
    def _inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield):
        n = max(np.size(a) for a in (spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield))
        arrs = [N.as_f32(a, n) for a in (spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)]
        arrs = [np.full(n, a[0], np.float32) if a.size == 1 and n > 1 else a for a in arrs]
        for a in arrs:
            if a.size != n:
                raise ValueError("option arrays must have the same length")
        return n, arrs

    def price_batch(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                    num_trials, antithetic=False):
        """Call and put prices (+ standard errors) for n options from fresh paths per option."""
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        out = [np.empty(n, np.float32) for _ in range(4)]
        N.check(N.lib().mcb200_price_batch(self._h, n, *[N.fptr(a) for a in arrs], steps, num_trials,
                                           N.ANTITHETIC if antithetic else 0, *[N.fptr(o) for o in out]))
        return dict(call=out[0], put=out[1], call_se=out[2], put_se=out[3])

    def greeks_batch(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                     num_trials, delta_spot=0.01, delta_volatility=0.01, delta_risk_free_rate=0.01,
                     delta_years_to_expiry=0.001):
        """Prices and all Greeks by central differences on common random numbers, one pass."""
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        vals = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        ses = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        bumps = N.Bumps(delta_spot, delta_volatility, delta_risk_free_rate, delta_years_to_expiry)
        N.check(N.lib().mcb200_greeks_batch(self._h, n, *[N.fptr(a) for a in arrs], steps, num_trials, C.byref(bumps),
                                            N._GreekOut(*[N.fptr(v) for v in vals]),
                                            N._GreekOut(*[N.fptr(s) for s in ses])))
        res = {name: vals[i] for i, name in enumerate(N.GREEK_NAMES)}
        res.update({name + "_se": ses[i] for i, name in enumerate(N.GREEK_NAMES)})
        return res

    def bs_batch(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield):
        """Black-Scholes closed form for the batch (the reference's analytic oracle, src/bs.rs), computed on the GPU:
        dict with the ten names of greeks_batch."""
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        vals = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        N.check(N.lib().mcb200_bs_batch(self._h, n, *[N.fptr(a) for a in arrs], N._GreekOut(*[N.fptr(v) for v in vals])))
        return {name: vals[i] for i, name in enumerate(N.GREEK_NAMES)}

    def greeks_batch_with_z(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                            num_trials, **bumps):
        """greeks_batch plus, per estimator, the closed-form value (name + '_bs') and the z-score
        (estimate - closed form) / standard error (name + '_z'; NaN where the sample SE is zero)."""
        res = self.greeks_batch(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                                num_trials, **bumps)
        exact = self.bs_batch(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        for name in N.GREEK_NAMES:
            res[name + "_bs"] = exact[name]
            with np.errstate(divide="ignore", invalid="ignore"):
                res[name + "_z"] = np.where(res[name + "_se"] > 0, (res[name] - exact[name]) / res[name + "_se"], np.nan)
        return res

    def prepare_price_batch(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                            num_trials, antithetic=False):
        """Bind host buffers once for repeated pricing of the same batch shape: returns a PreparedCall whose run()
        is exactly one mcb200_price_batch() C-ABI call (inputs are re-read from the bound arrays on every run, results
        land in .out['call' | 'put' | 'call_se' | 'put_se'])."""
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        out = [np.empty(n, np.float32) for _ in range(4)]
        args = (self._h, n, *[N.fptr(a) for a in arrs], C.c_float(steps), C.c_float(num_trials),
                N.ANTITHETIC if antithetic else 0, *[N.fptr(o) for o in out])
        return PreparedCall(N.lib().mcb200_price_batch, args, arrs,
                            dict(call=out[0], put=out[1], call_se=out[2], put_se=out[3]))

    def prepare_greeks_batch(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                             num_trials, delta_spot=0.01, delta_volatility=0.01, delta_risk_free_rate=0.01,
                             delta_years_to_expiry=0.001):
        """Like prepare_price_batch for mcb200_greeks_batch(); results in .out[name] and .out[name + '_se']."""
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        vals = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        ses = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        bumps = N.Bumps(delta_spot, delta_volatility, delta_risk_free_rate, delta_years_to_expiry)
        args = (self._h, n, *[N.fptr(a) for a in arrs], C.c_float(steps), C.c_float(num_trials), C.byref(bumps),
                N._GreekOut(*[N.fptr(v) for v in vals]), N._GreekOut(*[N.fptr(s) for s in ses]))
        res = {name: vals[i] for i, name in enumerate(N.GREEK_NAMES)}
        res.update({name + "_se": ses[i] for i, name in enumerate(N.GREEK_NAMES)})
        return PreparedCall(N.lib().mcb200_greeks_batch, args, arrs + [bumps], res)

    # ------------------------------------------------------------------ parity / inspection modes
    def price_batch_refstream(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                              num_trials, chunk_seeds, antithetic=False, want_growth=False):
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        seeds = np.ascontiguousarray(chunk_seeds, np.uint8)
        n_chunks = int(num_trials) // 8
        if seeds.size != n * n_chunks * 256:
            raise ValueError("chunk_seeds must hold n * floor(num_trials/8) * 256 bytes")
        out = [np.empty(n, np.float32) for _ in range(4)]
        growth = np.empty((n, n_chunks * 8), np.float32) if want_growth else None
        N.check(N.lib().mcb200_price_batch_refstream(
            self._h, n, *[N.fptr(a) for a in arrs], steps, num_trials, N.ANTITHETIC if antithetic else 0,
            seeds.ctypes.data_as(C.POINTER(C.c_uint8)), *[N.fptr(o) for o in out], N.fptr(growth)))
        return dict(call=out[0], put=out[1], call_se=out[2], put_se=out[3], growth=growth)

    def greeks_batch_refstream(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                               num_trials, chunk_seeds, delta_spot=0.01, delta_volatility=0.01,
                               delta_risk_free_rate=0.01, delta_years_to_expiry=0.001):
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        seeds = np.ascontiguousarray(chunk_seeds, np.uint8)
        if seeds.size != n * (int(num_trials) // 8) * 256:
            raise ValueError("chunk_seeds must hold n * floor(num_trials/8) * 256 bytes")
        vals = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        ses = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        bumps = N.Bumps(delta_spot, delta_volatility, delta_risk_free_rate, delta_years_to_expiry)
        N.check(N.lib().mcb200_greeks_batch_refstream(
            self._h, n, *[N.fptr(a) for a in arrs], steps, num_trials, C.byref(bumps),
            seeds.ctypes.data_as(C.POINTER(C.c_uint8)), N._GreekOut(*[N.fptr(v) for v in vals]),
            N._GreekOut(*[N.fptr(s) for s in ses])))
        res = {name: vals[i] for i, name in enumerate(N.GREEK_NAMES)}
        res.update({name + "_se": ses[i] for i, name in enumerate(N.GREEK_NAMES)})
        return res

    def price_batch_dump(self, spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps,
                         num_trials, antithetic=False):
        n, arrs = self._inputs(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield)
        out = [np.empty(n, np.float32) for _ in range(4)]
        growth = np.empty((n, (int(num_trials) // 8) * 8), np.float32)
        N.check(N.lib().mcb200_price_batch_dump(self._h, n, *[N.fptr(a) for a in arrs], steps, num_trials,
                                                N.ANTITHETIC if antithetic else 0, *[N.fptr(o) for o in out],
                                                N.fptr(growth)))
        return dict(call=out[0], put=out[1], call_se=out[2], put_se=out[3], growth=growth)

    # ------------------------------------------------------------------ device-resident batches (raw pointers)
    def price_batch_device(self, n, in_ptrs, steps, num_trials, out_ptrs, antithetic=False, stream=0):
        """in_ptrs: 6 device pointers (spot, strike, vol, rate, years, div); out_ptrs: call, put, call_se, put_se
        (0 = skip).  Enqueues on `stream` (a cudaStream_t as int; 0 = the context's stream) without synchronizing."""
        N.check(N.lib().mcb200_price_batch_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps, num_trials,
                                                  N.ANTITHETIC if antithetic else 0,
                                                  *[C.c_void_p(p) for p in out_ptrs], C.c_void_p(stream)))

    def greeks_batch_device(self, n, in_ptrs, steps, num_trials, bumps, out_ptrs, se_ptrs, stream=0):
        b = N.Bumps(*bumps)
        N.check(N.lib().mcb200_greeks_batch_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps, num_trials,
                                                   C.byref(b), N._VoidOut(*out_ptrs), N._VoidOut(*se_ptrs),
                                                   C.c_void_p(stream)))

    def price_moments_device(self, n, in_ptrs, steps, num_trials_local, moments_ptr, antithetic=False, stream=0):
        N.check(N.lib().mcb200_price_moments_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps,
                                                    num_trials_local, N.ANTITHETIC if antithetic else 0,
                                                    C.c_void_p(moments_ptr), C.c_void_p(stream)))

    def price_finalize_device(self, n, rate_ptr, years_ptr, moments_ptr, total_paths, total_num_trials, out_ptrs,
                              stream=0):
        N.check(N.lib().mcb200_price_finalize_device(self._h, n, C.c_void_p(rate_ptr), C.c_void_p(years_ptr),
                                                     C.c_void_p(moments_ptr), float(total_paths),
                                                     float(total_num_trials), *[C.c_void_p(p) for p in out_ptrs],
                                                     C.c_void_p(stream)))


    def greeks_moments_device(self, n, in_ptrs, steps, num_trials_local, bumps, moments_ptr, stream=0):
        b = N.Bumps(*bumps)
        N.check(N.lib().mcb200_greeks_moments_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps,
                                                     num_trials_local, C.byref(b), C.c_void_p(moments_ptr),
                                                     C.c_void_p(stream)))

    def greeks_finalize_device(self, n, rate_ptr, years_ptr, moments_ptr, total_paths, total_num_trials, bumps,
                               out_ptrs, se_ptrs, stream=0):
        b = N.Bumps(*bumps)
        N.check(N.lib().mcb200_greeks_finalize_device(self._h, n, C.c_void_p(rate_ptr), C.c_void_p(years_ptr),
                                                      C.c_void_p(moments_ptr), float(total_paths),
                                                      float(total_num_trials), C.byref(b), N._VoidOut(*out_ptrs),
                                                      N._VoidOut(*se_ptrs), C.c_void_p(stream)))


    # ------------------------------------------------------------------ cross-GPU fused combine (no collective)
    def xgpu_create(self, n_max, world, greeks=False):
        """Owner rank: allocate the exchange buffer; returns the 64 CUDA-IPC handle bytes to ship to the other ranks."""
        h = (C.c_uint8 * N.IPC_HANDLE_BYTES)()
        N.check(N.lib().mcb200_xgpu_create(self._h, int(n_max), int(world), 1 if greeks else 0, h))
        return bytes(h)

    def xgpu_open(self, handle, n_max, world, greeks=False):
        h = (C.c_uint8 * N.IPC_HANDLE_BYTES).from_buffer_copy(bytes(handle))
        N.check(N.lib().mcb200_xgpu_open(self._h, h, int(n_max), int(world), 1 if greeks else 0))

    def xgpu_attach(self, owner, n_max, world, greeks=False):
        """Same-process ranks: share `owner`'s exchange buffer through plain peer access (no IPC handle)."""
        N.check(N.lib().mcb200_xgpu_attach(self._h, owner._h, int(n_max), int(world), 1 if greeks else 0))

    def xgpu_set_timeout(self, seconds):
        N.check(N.lib().mcb200_xgpu_set_timeout(self._h, float(seconds)))

    def xgpu_close(self):
        N.check(N.lib().mcb200_xgpu_close(self._h))

    def price_xgpu_device(self, n, in_ptrs, steps, num_trials_local, rank, total_paths, total_num_trials,
                          antithetic=False, stream=0):
        N.check(N.lib().mcb200_price_xgpu_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps, num_trials_local,
                                                 N.ANTITHETIC if antithetic else 0, int(rank), float(total_paths),
                                                 float(total_num_trials), C.c_void_p(stream)))

    def greeks_xgpu_device(self, n, in_ptrs, steps, num_trials_local, bumps, rank, total_paths, total_num_trials,
                           stream=0):
        b = N.Bumps(*bumps)
        N.check(N.lib().mcb200_greeks_xgpu_device(self._h, n, *[C.c_void_p(p) for p in in_ptrs], steps, num_trials_local,
                                                  C.byref(b), int(rank), float(total_paths), float(total_num_trials),
                                                  C.c_void_p(stream)))

    def xgpu_results_greeks(self, n):
        vals = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        ses = [np.empty(n, np.float32) for _ in range(N.N_GREEK_OUTPUTS)]
        N.check(N.lib().mcb200_xgpu_results_greeks(self._h, n, N._GreekOut(*[N.fptr(v) for v in vals]),
                                                   N._GreekOut(*[N.fptr(s) for s in ses])))
        res = {name: vals[i] for i, name in enumerate(N.GREEK_NAMES)}
        res.update({name + "_se": ses[i] for i, name in enumerate(N.GREEK_NAMES)})
        return res

    def xgpu_results(self, n):
        out = [np.empty(n, np.float32) for _ in range(4)]
        N.check(N.lib().mcb200_xgpu_results(self._h, n, *[N.fptr(o) for o in out]))
        return dict(call=out[0], put=out[1], call_se=out[2], put_se=out[3])


class PreparedCall:
    """A C-ABI call with its host buffers bound (the ctypes marshalling is done once, not per call)."""

    def __init__(self, fn, args, keep, out):
        self._fn, self._args, self._keep, self.out = fn, args, keep, out

    def run(self):
        N.check(self._fn(*self._args))
        return self.out


def plan_query(n_options, steps, num_trials, grid_slots):
    """Launch plan (pure host logic; works without a GPU)."""
    p = N.Plan()
    st = N.lib().mcb200_plan_query(int(n_options), float(steps), float(num_trials), int(grid_slots), C.byref(p))
    if st != 0:
        raise N.Mcb200Error(st, "invalid plan request")
    return {k: getattr(p, k) for k, _ in N.Plan._fields_}


def shard_query(n_options, steps, num_trials, n_shards, shard, grid_slots):
    """How a batch is split over n_shards GPUs (mcb200_shard_query; pure host logic): dict with mode in
    ('single', 'options', 'trials'), active, option_begin, option_end, num_trials_local, total_paths."""
    s = N.Shard()
    st = N.lib().mcb200_shard_query(int(n_options), float(steps), float(num_trials), int(n_shards), int(shard),
                                    int(grid_slots), C.byref(s))
    if st != 0:
        raise N.Mcb200Error(st, "invalid shard request")
    d = {k: getattr(s, k) for k, _ in N.Shard._fields_}
    d["mode"] = N.SHARD_MODES[d["mode"]]
    return d


def plan_warp_range(plan, warp):
    """Rounds [first, end) that global warp `warp` simulates under `plan` (a dict from plan_query)."""
    p = N.Plan(**plan)
    a, b = C.c_int64(), C.c_int64()
    st = N.lib().mcb200_plan_warp_range(C.byref(p), int(warp), C.byref(a), C.byref(b))
    if st != 0:
        raise N.Mcb200Error(st, "invalid warp index")
    return a.value, b.value
