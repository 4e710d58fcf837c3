#!/usr/bin/env python3
"""Static schedule summary of the densest-MUFU loop of a kernel: instruction count, sum of the stall fields of the
control words (cycles one warp needs to issue the loop body alone), yield flags, pipe mix.  CPU-only.

    python tools/sass_sched.py <lib.so> <mangled-name-substring>
"""
import re
import subprocess
import sys
from collections import Counter


def loop_of(lib, name):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    body = next(f for f in re.split(r"\n\s*Function : ", txt)[1:] if name in f.split("\n")[0])
    lines = body.split("\n")
    ins = []
    for i, l in enumerate(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", l)
        if m and i + 1 < len(lines):
            e2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(e2.group(1), 16) if e2 else 0))
    best = None
    for a, t, _ in ins:
        if "BRA" in t.split()[0] or (t.startswith("@") and "BRA" in t):
            mm = re.search(r"0x([0-9a-f]+)", t)
            if mm and int(mm.group(1), 16) <= a:
                lp = [x for x in ins if int(mm.group(1), 16) <= x[0] <= a]
                mu = sum(1 for x in lp if "MUFU" in x[1])
                if mu >= 3 and (best is None or mu / len(lp) > best[0] / len(best[1])):
                    best = (mu, lp)
    return best


def main():
    mufu, lp = loop_of(sys.argv[1], sys.argv[2])
    stalls = [(hi >> 41) & 0xf for _, _, hi in lp]
    ops = Counter(t.split()[-1 if False else 0].split(".")[0] if not t.startswith("@") else t.split()[1].split(".")[0] for _, t, _ in lp)
    pairs = mufu / 3.0
    print("loop %d instr, %d MUFU (%.0f pairs): stall sum %d = %.2f cycles/pair single-warp; stall histogram %s" % (
        len(lp), mufu, pairs, sum(stalls), sum(stalls) / pairs, dict(sorted(Counter(stalls).items()))))
    print("  ops", dict(sorted(ops.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
