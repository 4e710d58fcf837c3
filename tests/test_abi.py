"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/mcb200.h declares, the launch
planner (pure host logic) behaves, and the product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_package


def _declared_functions():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not header.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(mcb200_[a-z0-9_]+|mc_simd_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(pkg):
    pkg.build()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 30
    for fn in ("mc_simd_call_price", "mc_simd_put_price", "mc_simd_call_price_av", "mc_simd_put_price_av",
               "mc_simd_call_delta", "mc_simd_put_delta", "mc_simd_gamma", "mc_simd_vega", "mc_simd_call_rho",
               "mc_simd_put_rho", "mc_simd_call_theta", "mc_simd_put_theta", "mcb200_price_batch",
               "mcb200_greeks_batch"):
        assert fn in names
    for name in names:
        assert hasattr(lib, name), "libmcb200.so does not export %s" % name


def test_python_mirror_has_reference_api(pkg):
    """Same twelve names and positional parameter order as src/mc_simd.rs:477-779."""
    import inspect
    expect = {
        "call_price": "spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "put_price": "spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "call_price_av": "spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "put_price_av": "spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "call_delta": "spot delta_spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "put_delta": "spot delta_spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "gamma": "spot delta_spot strike volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "vega": "spot strike volatility delta_volatility risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "call_rho": "spot strike volatility risk_free_rate delta_risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "put_rho": "spot strike volatility risk_free_rate delta_risk_free_rate years_to_expiry dividend_yield steps num_trials",
        "call_theta": "spot strike volatility risk_free_rate years_to_expiry delta_years_to_expiry dividend_yield steps num_trials",
        "put_theta": "spot strike volatility risk_free_rate years_to_expiry delta_years_to_expiry dividend_yield steps num_trials",
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(pkg.mc_simd, name)).parameters) == params.split(), name


def test_planner_reference_chunking(pkg):
    """paths = 8*floor(trials/8), pairs = floor(steps)/2 (mc_simd.rs:47,49); the warp ranges tile the rounds exactly
    once, in order, and differ in length by at most one round."""
    for n, steps, trials, slots in [(112, 100.0, 1000.0, 592), (112, 100.0, 20000.0, 592), (112, 1000.0, 2000.0, 592),
                                    (4096, 252.0, 1e5, 592), (100000, 252.0, 1e6, 592), (1, 101.0, 1007.0, 592),
                                    (3, 7.0, 8.0, 16), (1, 252.0, 1e6, 592), (5, 100.0, 3.0, 592), (7, 10.0, 40.0, 2)]:
        p = pkg.plan_query(n, steps, trials, slots)
        assert p["n_paths"] == 8 * (int(trials) // 8)
        assert p["half_steps"] == int(steps) // 2
        assert p["path_steps"] == n * p["n_paths"] * 2 * p["half_steps"]
        if p["n_paths"] == 0:
            assert p["grid"] == 0
            continue
        assert p["rounds_per_option"] == math.ceil(p["n_paths"] / 32)
        assert p["total_rounds"] == n * p["rounds_per_option"]
        waves = max(1, min(6, p["total_rounds"] // (8 * slots * 64)))         # large batches: several waves of CTAs
        assert p["waves"] == waves
        assert p["warps"] == min(8 * slots * waves, p["total_rounds"]) and p["grid"] == math.ceil(p["warps"] / 8)
        assert p["max_rounds_per_warp"] == math.ceil(p["total_rounds"] / p["warps"])
        probe = range(p["warps"]) if p["warps"] <= 4096 else list(range(64)) + list(range(p["warps"] - 64, p["warps"]))
        prev_end, sizes = None, []
        for w in probe:
            a, b = pkg.plan_warp_range(p, w)
            assert a == w * p["total_rounds"] // p["warps"] and b == (w + 1) * p["total_rounds"] // p["warps"]
            if prev_end is not None and w == prev_w + 1:
                assert a == prev_end
            prev_end, prev_w = b, w
            sizes.append(b - a)
        assert pkg.plan_warp_range(p, 0)[0] == 0 and pkg.plan_warp_range(p, p["warps"] - 1)[1] == p["total_rounds"]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == p["max_rounds_per_warp"] or len(sizes) < p["warps"]
        with pytest.raises(pkg.Mcb200Error):
            pkg.plan_warp_range(p, p["warps"])


def test_planner_fills_the_machine(pkg):
    """Any batch shape spreads evenly over all resident warps: the busiest warp is within one round of the mean."""
    slots = 148 * 4
    for n, steps, trials in [(1, 252.0, 1e6), (100000, 252.0, 1e4), (112, 100.0, 20000.0), (112, 1000.0, 2000.0),
                             (4096, 252.0, 1e5)]:
        p = pkg.plan_query(n, steps, trials, slots)
        assert p["grid"] == slots * p["waves"] and (p["waves"] == 1 or p["total_rounds"] // p["warps"] >= 64)
        mean = p["total_rounds"] / p["warps"]
        assert p["max_rounds_per_warp"] < mean + 1
    c2 = pkg.plan_query(112, 100.0, 20000.0, slots)
    assert c2["max_rounds_per_warp"] / (c2["total_rounds"] / c2["warps"]) < 1.02     # 15 vs 14.8 rounds
    small = pkg.plan_query(112, 100.0, 1000.0, slots)                                # fewer rounds than warps
    assert small["warps"] == small["total_rounds"] and small["max_rounds_per_warp"] == 1
    odd = pkg.plan_query(1, 100.0, 10000.0, slots)                                   # 313 rounds: W not a multiple of 8
    assert odd["warps"] == 313 and odd["grid"] == 40


def test_planner_rejects_bad_requests(pkg):
    for n, steps, trials in [(-1, 100.0, 1000.0), (1, 0.0, 1000.0), (1, -2.0, 1000.0), (1, float("nan"), 1000.0),
                             (1, 100.0, float("inf")), (1, 1e10, 1000.0), (1, 100.0, float("nan"))]:
        with pytest.raises(pkg.Mcb200Error):
            pkg.plan_query(n, steps, trials, 1184)
    # legal in the reference, finite results there: no normals for steps in (0,1) (mc_simd.rs:47), an empty chunk range
    # for negative trial counts (mc_simd.rs:49)
    assert pkg.plan_query(1, 0.5, 1000.0, 1184)["half_steps"] == 0
    assert pkg.plan_query(3, 100.0, -8.0, 1184)["n_paths"] == 0 and pkg.plan_query(3, 100.0, -8.0, 1184)["grid"] == 0


def test_shard_policy(pkg):
    """mcb200_shard_query: single device when the batch does not fill one GPU, options-major blocks when they balance to
    within 10 %, trials-minor chunk ranges otherwise; the shards tile the batch exactly once."""
    slots = 592
    for n, steps, trials, world, mode in [(112, 100.0, 1000.0, 8, "single"), (1, 100.0, 1000.0, 4, "single"),
                                          (112, 100.0, 20000.0, 8, "options"), (112, 100.0, 20000.0, 2, "options"),
                                          (20, 100.0, 1e6, 8, "trials"), (100, 100.0, 1e5, 8, "options"), (8, 100.0, 8e6, 8, "options"),
                                          (100000, 252.0, 1e4, 8, "options"), (100000, 252.0, 1e6, 7, "options"),
                                          (5, 100.0, 8e6, 8, "trials"), (4096, 252.0, 1e5, 8, "options"),
                                          (5, 50.0, 1000003.0, 3, "trials"), (100, 10.0, 1e6, 1, "single")]:
        shards = [pkg.shard_query(n, steps, trials, world, r, slots) for r in range(world)]
        assert all(s["mode"] == mode for s in shards), (n, trials, world, shards[0]["mode"])
        paths = 8 * (int(trials) // 8)
        assert all(s["total_paths"] == paths for s in shards)
        if mode == "single":
            assert shards[0]["active"] and (shards[0]["option_begin"], shards[0]["option_end"]) == (0, n)
            assert shards[0]["num_trials_local"] == np.float32(trials) and not any(s["active"] for s in shards[1:])
        elif mode == "options":
            assert shards[0]["option_begin"] == 0 and shards[-1]["option_end"] == n
            assert all(shards[i]["option_end"] == shards[i + 1]["option_begin"] for i in range(world - 1))
            sizes = [s["option_end"] - s["option_begin"] for s in shards]
            assert max(sizes) - min(sizes) <= 1 and all(s["num_trials_local"] == np.float32(trials) for s in shards)
            lo, hi = pkg.sharding.partition_options(n, world, 1 % world)
            assert (lo, hi) == (shards[1 % world]["option_begin"], shards[1 % world]["option_end"])
        else:
            assert all((s["option_begin"], s["option_end"]) == (0, n) for s in shards)
            local = [s["num_trials_local"] for s in shards]
            assert all(t % 8 == 0 for t in local) and sum(local) == paths and max(local) - min(local) <= 8
            assert [pkg.sharding.partition_trials(trials, world, r)[0] for r in range(world)] == local
    with pytest.raises(pkg.Mcb200Error):
        pkg.shard_query(10, 100.0, 1000.0, 4, 4, slots)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_no_cpu_fallback(pkg):
    """Without a GPU the context cannot be created, batches raise and the drop-ins return NaN with a reason."""
    with pytest.raises(pkg.Mcb200Error) as e:
        pkg.Context()
    assert e.value.status == 3
    v = pkg.mc_simd.call_price(100.0, 110.0, 0.25, 0.05, 0.5, 0.02, 100.0, 1000.0)
    assert math.isnan(v) and "no CPU path" in pkg.mc_simd.last_error()


def _build_c_client(pkg, tmp_path):
    """gcc -std=c99 against include/mcb200.h, linked with libmcb200.so only."""
    import subprocess
    pkg.build()
    exe = str(tmp_path / "c_abi_client")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c_client", "c_abi_client.c"), "-L", libdir, "-lmcb200", "-lm",
                    "-Wl,-rpath," + libdir], check=True, capture_output=True)
    return exe


def _build_cpp_client(pkg, tmp_path):
    """g++ -std=c++17 against include/mc_simd.hpp (the reference's module surface, name for name)."""
    import subprocess
    pkg.build()
    exe = str(tmp_path / "cpp_client")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c_client", "cpp_client.cpp"), "-L", libdir, "-lmcb200",
                    "-Wl,-rpath," + libdir], check=True, capture_output=True)
    return exe


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_cpp_mirror_of_mc_simd_without_gpu(pkg, tmp_path):
    import subprocess
    res = subprocess.run([_build_cpp_client(pkg, tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    header = open(os.path.join(ROOT, "include", "mc_simd.hpp")).read()
    for name in ("call_price", "put_price", "call_price_av", "put_price_av", "call_delta", "put_delta", "gamma", "vega",
                 "call_rho", "put_rho", "call_theta", "put_theta"):
        assert re.search(r"inline float %s\(" % name, header), name


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_plain_c_client_links_and_fails_loudly_without_gpu(pkg, tmp_path):
    import subprocess
    res = subprocess.run([_build_c_client(pkg, tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "c client ok (no device)" in res.stdout


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "monte-carlo-options-simd_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "libcpubaseline" not in src, f


def test_jump_table_doubling_chain(oracle):
    """include/xoshiro256_jump_tables.h: entry k applied twice == entry k+1 applied once, starting from the published
    JUMP polynomial (entry 0); so entry k really is jump^(2^k)."""
    text = open(os.path.join(ROOT, "include", "xoshiro256_jump_tables.h")).read()
    def rows_of(name):
        body = re.search(name + r"\[\d*\]\[4\] = \{(.*?)\};", text, flags=re.S).group(1)
        rows = re.findall(r"\{(0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL\}", body)
        return [[int(x, 16) for x in r] for r in rows]
    table, long_table = rows_of("XOSHIRO256_JUMP_POW2"), rows_of("XOSHIRO256_LONG_JUMP_POW2")
    long_jump = [0x76e15d3efefdcbbf, 0xc5004e441c522fb3, 0x77710069854ee241, 0x39109bb02acbe635]
    assert table[0] == [0x180ec6d33cfd0aba, 0xd5a61266f0c9392c, 0xa9582618e03fc9aa, 0x39abdc4529b1661c]
    assert len(table) == 40 and len(long_table) == 32 and long_table[0] == long_jump
    state = [0x9e3779b97f4a7c15, 0xbf58476d1ce4e5b9, 0x94d049bb133111eb, 0x2545f4914f6cdd1d]
    for tab, upto in ((table, 24), (long_table, 31)):
        for k in range(0, upto):
            twice = oracle.xoshiro_apply_poly(oracle.xoshiro_apply_poly(state, tab[k]), tab[k])
            once = oracle.xoshiro_apply_poly(state, tab[k + 1])
            assert list(twice) == list(once), k
    # substream block b = long_jump applied b times = the set bits of b through the power table (what seed_pool does)
    serial = list(state)
    for _ in range(5):
        serial = list(oracle.xoshiro_apply_poly(serial, long_jump))
    fast = list(oracle.xoshiro_apply_poly(oracle.xoshiro_apply_poly(state, long_table[0]), long_table[2]))
    assert serial == fast


def test_sharding_partitions(pkg):
    sh = pkg.sharding
    for n, w in [(112, 1), (112, 2), (113, 8), (5, 8), (100000, 8)]:
        blocks = [sh.partition_options(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1
    for trials, w in [(20000.0, 2), (1000.0, 8), (1e6, 8), (1007.0, 3)]:
        parts = [sh.partition_trials(trials, w, r) for r in range(w)]
        assert all(p[0] % 8 == 0 for p in parts)
        assert sum(p[0] for p in parts) == parts[0][1] == 8 * (int(trials) // 8)
