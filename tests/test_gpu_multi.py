"""Several GPUs behind the C ABI, proven on ONE GPU: the multi-device context (mcb200_ctx_create_multi) with the same
device listed twice, the cross-GPU arrival protocol with two same-process ranks sharing one exchange buffer
(mcb200_xgpu_attach), and launches of one context arriving on different caller streams.  Everything here is compared
bit for bit with the single-device entry points on the same seeds and substream blocks."""
import numpy as np
import pytest

from conftest import load_package

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GRID = dict(strike=110.0, vol=0.25, r=0.05, t=0.5, q=0.02)
G = (GRID["strike"], GRID["vol"], GRID["r"], GRID["t"], GRID["q"])
SEED = 4242


@pytest.fixture(scope="module")
def pkg():
    return load_package()


def _cols(spots, dev):
    n = len(spots)
    cols = [torch.tensor(np.asarray(spots, np.float32), device=dev)]
    cols += [torch.full((n,), v, dtype=torch.float32, device=dev) for v in G]
    return cols


def test_multi_context_options_major_equals_per_device_contexts(pkg):
    """64 options x 80000 trials on devices [0, 0]: options-major.  Block [0,32) must equal a plain context on substream
    block 0 pricing those 32 options, block [32,64) a plain context on block 1 -- bit for bit, SEs included."""
    spots = np.linspace(70, 150, 64).astype(np.float32)
    with pkg.Context(devices=[0, 0], seed=SEED) as multi:
        assert multi.device_count() == 2
        slots = multi.info()["grid_slots"]
        assert [pkg.shard_query(64, 100.0, 80000.0, 2, d, slots)["mode"] for d in (0, 1)] == ["options", "options"]
        got = multi.price_batch(spots, *G, 100.0, 80000.0)
        again = multi.price_batch(spots, *G, 100.0, 80000.0)          # fresh paths on every call
        multi.set_seed(SEED)
        replay = multi.price_batch(spots, *G, 100.0, 80000.0)
        greeks = multi.greeks_batch(spots, *G, 100.0, 80000.0)
        bs = multi.bs_batch(spots, *G)
    for d, (lo, hi) in enumerate(((0, 32), (32, 64))):
        with pkg.Context(device=0, seed=SEED, substream_block=d) as one:
            want = one.price_batch(spots[lo:hi], *G, 100.0, 80000.0)
            one.price_batch(spots[lo:hi], *G, 100.0, 80000.0)
            one.set_seed(SEED, d)
            one.price_batch(spots[lo:hi], *G, 100.0, 80000.0)
            want_g = one.greeks_batch(spots[lo:hi], *G, 100.0, 80000.0)
        for k in want:
            assert np.array_equal(got[k][lo:hi], want[k]), (d, k)
        for k in want_g:
            assert np.array_equal(greeks[k][lo:hi], want_g[k]), (d, k)
    assert all(np.array_equal(got[k], replay[k]) for k in got) and not np.array_equal(got["call"], again["call"])
    z = (got["call"] - bs["call"]) / np.maximum(got["call_se"], 1e-4)
    assert np.abs(z).max() < 4.5 and (np.abs(z) > 3).mean() < 0.05
    zd = (greeks["call_delta"] - bs["call_delta"]) / np.maximum(greeks["call_delta_se"], 1e-4)
    assert np.abs(zd).max() < 4.5 and (np.abs(zd) > 3).mean() < 0.05


def test_multi_context_trials_minor_equals_moment_route(pkg):
    """5 options x 400000 trials on devices [0, 0]: trials-minor.  Must equal, bit for bit, the explicit route through
    the device-pointer ABI: moments on block 0 (200000 trials) + moments on block 1 (200000 trials), summed in that
    order, one finalisation -- for prices, antithetic prices and all ten Greek estimators."""
    spots = np.array([90.0, 100.0, 110.0, 120.0, 130.0], np.float32)
    dev = torch.device("cuda", 0)
    cols = _cols(spots, dev)
    ptrs = [c.data_ptr() for c in cols]
    n = len(spots)
    bumps = (0.01, 0.01, 0.01, 0.001)
    with pkg.Context(devices=[0, 0], seed=SEED) as multi:
        slots = multi.info()["grid_slots"]
        sh = [pkg.shard_query(n, 100.0, 400000.0, 2, d, slots) for d in (0, 1)]
        assert [s["mode"] for s in sh] == ["trials", "trials"] and [s["num_trials_local"] for s in sh] == [200000.0] * 2
        price = multi.price_batch(spots, *G, 100.0, 400000.0)
        av = multi.price_batch(spots, *G, 100.0, 400000.0, antithetic=True)
        greeks = multi.greeks_batch(spots, *G, 100.0, 400000.0)
        bs = multi.bs_batch(spots, *G)
    moms = {"price": [], "av": [], "greeks": []}
    for d in (0, 1):
        with pkg.Context(device=0, seed=SEED, substream_block=d) as one:
            m = torch.zeros(n, 4, dtype=torch.float64, device=dev)
            one.price_moments_device(n, ptrs, 100.0, 200000.0, m.data_ptr())
            torch.cuda.synchronize(); moms["price"].append(m.clone())
            one.price_moments_device(n, ptrs, 100.0, 200000.0, m.data_ptr(), antithetic=True)
            torch.cuda.synchronize(); moms["av"].append(m.clone())
            mg = torch.zeros(n, 20, dtype=torch.float64, device=dev)
            one.greeks_moments_device(n, ptrs, 100.0, 200000.0, bumps, mg.data_ptr())
            torch.cuda.synchronize(); moms["greeks"].append(mg.clone())
    with pkg.Context(device=0, seed=1) as fin:
        for key, res in (("price", price), ("av", av)):
            total = moms[key][0] + moms[key][1]
            vals = torch.zeros(4, n, device=dev)
            fin.price_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), total.data_ptr(), 400000.0, 400000.0,
                                      [vals[i].data_ptr() for i in range(4)])
            torch.cuda.synchronize()
            for i, k in enumerate(("call", "put", "call_se", "put_se")):
                assert np.array_equal(res[k], vals[i].cpu().numpy()), (key, k)
        total = moms["greeks"][0] + moms["greeks"][1]
        gv, gs = torch.zeros(10, n, device=dev), torch.zeros(10, n, device=dev)
        fin.greeks_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), total.data_ptr(), 400000.0, 400000.0, bumps,
                                   [gv[i].data_ptr() for i in range(10)], [gs[i].data_ptr() for i in range(10)])
        torch.cuda.synchronize()
        for i, name in enumerate(pkg.GREEK_NAMES):
            assert np.array_equal(greeks[name], gv[i].cpu().numpy()), name
            assert np.array_equal(greeks[name + "_se"], gs[i].cpu().numpy()), name
    for k in ("call", "put"):
        assert np.all(np.abs(price[k] - bs[k]) <= 4.0 * price[k + "_se"] + 1e-4)
        assert np.all(av[k + "_se"] < price[k + "_se"])             # the antithetic estimator has the smaller variance


def test_multi_context_small_batch_runs_on_the_first_device(pkg):
    """A batch that does not fill one GPU is not split: identical to a plain context on block 0; device-pointer and
    inspection entry points refuse the parent; ragged trial counts keep the reference's divisor."""
    with pkg.Context(devices=[0, 0, 0], seed=SEED) as multi, pkg.Context(device=0, seed=SEED) as one:
        a = multi.price_batch([100.0], *G, 100.0, 1000.0)
        b = one.price_batch([100.0], *G, 100.0, 1000.0)
        assert all(np.array_equal(a[k], b[k]) for k in a)
        z = multi.price_batch([100.0, 105.0], *G, 100.0, 7.0)           # fewer than 8 trials: nothing simulated
        assert np.all(z["call"] == 0.0) and np.all(z["call_se"] == 0.0)
        with pytest.raises(pkg.Mcb200Error):
            multi.export_states(0, 4)
        with pytest.raises(pkg.Mcb200Error):
            multi.price_batch_device(1, [0] * 6, 100.0, 1000.0, [0] * 4)
        kid = multi.child(2)
        assert kid.info()["substream_block"] == 2 and multi.info()["substream_block"] == 0
        # trials-minor with a trial count that is not a multiple of 8: paths = 8*floor(N/8), mean divides by N
        r = multi.price_batch([100.0, 110.0], *G, 50.0, 1000003.0)
        s = [pkg.shard_query(2, 50.0, 1000003.0, 3, d, multi.info()["grid_slots"]) for d in range(3)]
        assert sum(x["num_trials_local"] for x in s) == 1000000.0 and s[0]["mode"] == "trials"
        bs = multi.bs_batch([100.0, 110.0], *G)
        assert np.all(np.abs(r["call"] - bs["call"] * (1000000.0 / 1000003.0)) <= 4.0 * r["call_se"])


def test_dropins_on_a_multi_device_default_context(pkg, monkeypatch):
    """MCB200_DEVICES=0,0 makes the twelve drop-ins run on a multi-device default context."""
    import subprocess, sys, os
    from conftest import ROOT
    code = ("import importlib; m = importlib.import_module('monte-carlo-options-simd_b200');"
            "v = m.mc_simd.call_price(100.0, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 4000000.0);"
            "g = m.mc_simd.gamma(100.0, 0.01, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 400000.0);"
            "import ctypes as C; h = C.c_void_p(); m._native.lib().mcb200_default_ctx(C.byref(h));"
            "print('DROPIN', v, g, m._native.lib().mcb200_ctx_device_count(h))")
    env = dict(os.environ, MCB200_DEVICES="0,0", PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    _, v, g, count = [l for l in res.stdout.splitlines() if l.startswith("DROPIN")][-1].split()
    assert int(count) == 2 and abs(float(v) - 3.8598) < 0.05 and abs(float(g) - 0.020896) < 0.003


def test_fused_cross_gpu_combine_with_two_ranks_on_one_gpu(pkg):
    """The cross-GPU arrival protocol (peer stores of moment rows, system-scope arrival counters, last arrival
    finalises, epoch word) with two same-process 'ranks' on device 0 sharing one exchange buffer, on two streams.
    Result == moments + sum in rank order + finalize, bit for bit, over four rounds (counters and epoch reset
    themselves); a second launch without fetching the results is refused."""
    dev = torch.device("cuda", 0)
    spots = np.array([85.0, 95.0, 100.0, 105.0, 110.0, 125.0], np.float32)
    cols = _cols(spots, dev)
    ptrs = [c.data_ptr() for c in cols]
    n = len(spots)
    s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    local, total = 48000.0, 96000.0
    with pkg.Context(device=0, seed=SEED, substream_block=0) as c0, pkg.Context(device=0, seed=SEED, substream_block=1) as c1, \
            pkg.Context(device=0, seed=SEED, substream_block=0) as r0, pkg.Context(device=0, seed=SEED, substream_block=1) as r1, \
            pkg.Context(device=0, seed=9) as fin:
        c0.xgpu_create(n, 2)
        c1.xgpu_attach(c0, n, 2)
        c0.xgpu_set_timeout(30.0); c1.xgpu_set_timeout(30.0)
        for rnd in range(4):
            # alternate which rank launches first, so either one can be the last arrival
            order = [(c0, 0, s0), (c1, 1, s1)] if rnd % 2 == 0 else [(c1, 1, s1), (c0, 0, s0)]
            for c, rank, st in order:
                c.price_xgpu_device(n, ptrs, 100.0, local, rank, total, total, stream=st.cuda_stream)
            if rnd == 1:
                with pytest.raises(pkg.Mcb200Error) as err:          # launch and results must alternate
                    c0.price_xgpu_device(n, ptrs, 100.0, local, 0, total, total, stream=s0.cuda_stream)
                assert "pending" in str(err.value)
            got0, got1 = c0.xgpu_results(n), c1.xgpu_results(n)
            assert all(np.array_equal(got0[k], got1[k]) for k in got0)
            # the explicit route on the same substreams (contexts r0, r1 advance in step with c0, c1)
            m0 = torch.zeros(n, 4, dtype=torch.float64, device=dev); m1 = torch.zeros_like(m0)
            r0.price_moments_device(n, ptrs, 100.0, local, m0.data_ptr())
            r1.price_moments_device(n, ptrs, 100.0, local, m1.data_ptr())
            torch.cuda.synchronize()
            tot = m0 + m1
            vals = torch.zeros(4, n, device=dev)
            fin.price_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), tot.data_ptr(), total, total,
                                      [vals[i].data_ptr() for i in range(4)])
            torch.cuda.synchronize()
            for i, k in enumerate(("call", "put", "call_se", "put_se")):
                assert np.array_equal(got0[k], vals[i].cpu().numpy()), (rnd, k)
        with pytest.raises(pkg.Mcb200Error):
            c0.xgpu_results(n)                                          # nothing pending
        c1.xgpu_close(); c0.xgpu_close()
        # Greeks through the same protocol
        c0.xgpu_create(n, 2, greeks=True)
        c1.xgpu_attach(c0, n, 2, greeks=True)
        bumps = (0.01, 0.01, 0.01, 0.001)
        for rnd in range(2):
            c1.greeks_xgpu_device(n, ptrs, 100.0, local, bumps, 1, total, total, stream=s1.cuda_stream)
            c0.greeks_xgpu_device(n, ptrs, 100.0, local, bumps, 0, total, total, stream=s0.cuda_stream)
            g0 = c0.xgpu_results_greeks(n); g1 = c1.xgpu_results_greeks(n)
            assert all(np.array_equal(g0[k], g1[k]) for k in g0)
            m0 = torch.zeros(n, 20, dtype=torch.float64, device=dev); m1 = torch.zeros_like(m0)
            r0.greeks_moments_device(n, ptrs, 100.0, local, bumps, m0.data_ptr())
            r1.greeks_moments_device(n, ptrs, 100.0, local, bumps, m1.data_ptr())
            torch.cuda.synchronize()
            tot = m0 + m1
            gv, gs = torch.zeros(10, n, device=dev), torch.zeros(10, n, device=dev)
            fin.greeks_finalize_device(n, cols[3].data_ptr(), cols[4].data_ptr(), tot.data_ptr(), total, total, bumps,
                                       [gv[i].data_ptr() for i in range(10)], [gs[i].data_ptr() for i in range(10)])
            torch.cuda.synchronize()
            for i, name in enumerate(pkg.GREEK_NAMES):
                assert np.array_equal(g0[name], gv[i].cpu().numpy()) and np.array_equal(g0[name + "_se"], gs[i].cpu().numpy()), name
        c1.xgpu_close(); c0.xgpu_close()


def test_two_caller_streams_on_one_context_serialise(pkg):
    """One context, two caller streams, 50 interleaved batches of different shapes with no host synchronisation in
    between: the library chains the launches with events, so the results equal a single-stream replay bit for bit
    (without the chain the kernels would overlap, share generator states and corrupt each other's slots)."""
    dev = torch.device("cuda", 0)
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    shapes = [(112, 100.0, 20000.0), (7, 50.0, 8008.0), (300, 20.0, 4000.0), (1, 100.0, 250000.0), (41, 10.0, 8008.0)]
    spots = {n: torch.linspace(80, 140, n, device=dev) for n, _, _ in shapes}
    cols = {n: [spots[n]] + [torch.full((n,), v, dtype=torch.float32, device=dev) for v in G] for n, _, _ in shapes}

    def run(streams):
        outs = []
        with pkg.Context(device=0, seed=SEED) as ctx:
            for i in range(50):
                n, steps, trials = shapes[i % len(shapes)]
                out = torch.zeros(4, n, device=dev)
                st = streams[i % len(streams)]
                ctx.price_batch_device(n, [c.data_ptr() for c in cols[n]], steps, trials,
                                       [out[j].data_ptr() for j in range(4)], stream=st.cuda_stream)
                outs.append(out)
            torch.cuda.synchronize()
        return [o.cpu().numpy() for o in outs]

    two = run([sa, sb])
    one = run([sa])
    for i, (x, y) in enumerate(zip(two, one)):
        assert np.array_equal(x, y), i
        assert np.all(np.isfinite(x)) and np.all(x[2] > 0)


def test_dropins_survive_a_concurrent_default_context_reset(pkg):
    """The drop-ins hold a reference to the default context for the duration of a call: resetting it from another thread
    (which used to free the context under the callers' feet) neither crashes nor produces a non-finite price."""
    import threading
    pkg.mc_simd.reset_default_context(device=0, seed=1)
    stop, errors, results = threading.Event(), [], []

    def caller():
        try:
            while not stop.is_set():
                v = pkg.mc_simd.call_price(100.0, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 4000.0)
                g = pkg.mc_simd.gamma(100.0, 0.01, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 4000.0)
                results.append((v, g))
        except Exception as e:          # pragma: no cover
            errors.append(e)
    threads = [threading.Thread(target=caller) for _ in range(4)]
    for t in threads:
        t.start()
    for i in range(25):
        pkg.mc_simd.reset_default_context(device=0, seed=100 + i)
    stop.set()
    for t in threads:
        t.join()
    assert not errors and len(results) > 20
    v = np.array(results)
    assert np.all(np.isfinite(v)) and np.all(np.abs(v[:, 0] - 3.86) < 1.5) and np.all(v[:, 1] >= 0.0)


def test_multi_device_context_from_several_host_threads(pkg):
    """Batches submitted to ONE multi-device context from several host threads are serialised by the library: every
    result is finite and statistically right, and the context is still bit-reproducible afterwards."""
    import threading
    spots = np.linspace(80, 140, 64).astype(np.float32)
    with pkg.Context(devices=[0, 0], seed=SEED) as multi:
        bs = multi.bs_batch(spots, *G)["call"]
        out, errors = [], []

        def worker(k):
            try:
                for _ in range(5):
                    r = multi.price_batch(spots, *G, 50.0, 40000.0) if k % 2 == 0 else multi.price_batch(spots[:5], *G, 50.0, 400000.0)
                    out.append((len(r["call"]), r))
            except Exception as e:      # pragma: no cover
                errors.append(e)
        threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors and len(out) == 20
        for n, r in out:
            z = (r["call"] - bs[:n]) / np.maximum(r["call_se"], 1e-3)
            assert np.all(np.isfinite(r["call"])) and np.abs(z).max() < 5.5
        multi.set_seed(SEED)
        a = multi.price_batch(spots, *G, 50.0, 40000.0)
        multi.set_seed(SEED)
        b = multi.price_batch(spots, *G, 50.0, 40000.0)
        assert all(np.array_equal(a[k], b[k]) for k in a)
