"""Property tests (hypothesis) of the pure-host logic behind the C ABI: the launch planner and the shard policy must
tile every batch exactly once whatever its shape -- the invariants the kernel's arrival counts and the multi-GPU sums
rely on.  No GPU needed."""
import math

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from conftest import load_package  # noqa: E402

pkg = load_package()


@settings(max_examples=150, deadline=None)
@given(n=st.integers(1, 200000), trials=st.integers(8, 2_000_000), steps=st.integers(1, 2000),
       slots=st.sampled_from([2, 16, 148 * 3, 148 * 4, 132 * 4]))
def test_plan_ranges_tile_the_rounds(n, trials, steps, slots):
    p = pkg.plan_query(n, float(steps), float(trials), slots)
    paths = 8 * (int(np.float32(trials)) // 8)
    assert p["n_paths"] == paths and p["half_steps"] == steps // 2
    assert p["rounds_per_option"] == math.ceil(paths / 32) and p["total_rounds"] == n * p["rounds_per_option"]
    resident = 8 * slots
    assert 1 <= p["waves"] <= 6 and p["warps"] == min(resident * p["waves"], p["total_rounds"])
    assert p["waves"] == 1 or p["total_rounds"] // p["warps"] >= 64          # every warp of a multi-wave plan keeps >= 64 rounds
    assert p["grid"] == math.ceil(p["warps"] / 8) and p["max_rounds_per_warp"] == math.ceil(p["total_rounds"] / p["warps"])
    # ranges: contiguous, non-empty, in order, sizes within one round of each other (probe both ends and the middle)
    W = p["warps"]
    probe = sorted(set([0, 1, W // 2, max(W - 2, 0), W - 1]) & set(range(W)))
    for w in probe:
        a, b = pkg.plan_warp_range(p, w)
        assert a == w * p["total_rounds"] // W and b == (w + 1) * p["total_rounds"] // W and b > a
        assert b - a in (p["total_rounds"] // W, p["total_rounds"] // W + 1)
    assert pkg.plan_warp_range(p, 0)[0] == 0 and pkg.plan_warp_range(p, W - 1)[1] == p["total_rounds"]


@settings(max_examples=150, deadline=None)
@given(n=st.integers(1, 100000), trials=st.integers(0, 3_000_000), steps=st.integers(1, 500), world=st.integers(1, 16))
def test_shards_tile_the_batch(n, trials, steps, world):
    slots = 592
    shards = [pkg.shard_query(n, float(steps), float(trials), world, r, slots) for r in range(world)]
    mode = shards[0]["mode"]
    paths = 8 * (trials // 8)
    assert all(s["mode"] == mode and s["total_paths"] == paths for s in shards)
    if mode == "single":
        assert shards[0]["active"] and (shards[0]["option_begin"], shards[0]["option_end"]) == (0, n)
        assert not any(s["active"] for s in shards[1:])
    elif mode == "options":
        assert n >= world and all(s["active"] for s in shards)
        assert shards[0]["option_begin"] == 0 and shards[-1]["option_end"] == n
        assert all(a["option_end"] == b["option_begin"] for a, b in zip(shards, shards[1:]))
        sizes = [s["option_end"] - s["option_begin"] for s in shards]
        assert max(sizes) * world * 10 <= 11 * n and max(sizes) - min(sizes) <= 1
        assert all(s["num_trials_local"] == np.float32(trials) for s in shards)
    else:
        assert all((s["option_begin"], s["option_end"]) == (0, n) for s in shards)
        local = [int(s["num_trials_local"]) for s in shards]
        assert sum(local) == paths and all(t % 8 == 0 for t in local) and max(local) - min(local) <= 8
        assert all(bool(s["active"]) == (t > 0) for s, t in zip(shards, local))
    # a batch that does not fill one GPU is never split
    whole = pkg.plan_query(n, float(steps), float(trials), slots)
    if world > 1 and paths > 0 and whole["total_rounds"] <= 8 * slots:
        assert mode == "single"
