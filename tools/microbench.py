#!/usr/bin/env python3
"""Measured per-SM pipe rates on the GPU at hand (MUFU / FMA / ALU / issue); JSON on stdout.
Run on the B200 box:  python tools/microbench.py > profiles/microbench_rNN.json"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pkg = importlib.import_module("monte-carlo-options-simd_b200")
with pkg.Context(device=0) as ctx:
    info = ctx.info()
    rates = ctx.microbench()
print(json.dumps({"device": info, "note": "ops = one instruction per thread; pair_fast_loop / xoshiro256pp_next count "
                  "calls (one Box-Muller pair = 2 path-steps; one 64-bit output)", "rates": rates}, indent=1))
