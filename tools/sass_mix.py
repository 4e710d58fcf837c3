#!/usr/bin/env python3
"""Instruction mix of the hottest loop of a kernel (the backward branch whose body has the highest MUFU density).

    python tools/sass_mix.py <lib.so|cubin> <mangled-name-substring> [pairs_per_iteration]

Pipes (sm_100a, measured by tools/microbench.py): ALU = LOP3/SHF/IADD3/LEA/PRMT/ISETP/VIADD/SEL..., FMAH = IMAD*,
FP = FFMA/FADD/FMUL, XU = MUFU.  Used to check pipe balance here (no GPU needed) before spending GPU time.
"""
import re
import subprocess
import sys
from collections import Counter


def pipe(op):
    base = op.split(".")[0]
    if base == "MUFU":
        return "XU"
    if base in ("FFMA", "FADD", "FMUL", "FMNMX", "FSEL", "FSETP"):
        return "FP" if base in ("FFMA", "FADD", "FMUL") else "ALU"
    if base in ("IMAD",):
        return "FMAH"
    if base in ("BRA", "BAR", "EXIT", "NOP", "WARPSYNC", "BSSY", "BSYNC"):
        return "CTRL"
    if base.startswith("U") and base not in ("UIADD3X",):
        return "UNIF"
    if base in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "LDCU", "SHFL", "ATOM", "RED"):
        return "MEM"
    return "ALU"


def main():
    lib, name = sys.argv[1], sys.argv[2]
    pairs = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = next((f for f in funcs[1:] if name in f.split("\n")[0]), None)
    if body is None:
        raise SystemExit("function not found")
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    best = None
    for i, (addr, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) <= addr:
                tgt = int(t.group(1), 16)
                loop = [x for x in ins if tgt <= x[0] <= addr]
                mufu = sum(1 for x in loop if x[1].startswith("MUFU"))
                if mufu >= 3 and (best is None or mufu / len(loop) > best[0] / len(best[1])):
                    best = (mufu, loop)
    mufu, loop = best
    ops = Counter(x[1].split(".")[0] + ("." + x[1].split(".")[1] if x[1].startswith(("IMAD.", "MUFU.")) and "." in x[1] else "") for x in loop)
    pipes = Counter(pipe(x[1]) for x in loop)
    print("%s: loop of %d instructions, %d MUFU" % (body.split("\n")[0].strip(), len(loop), mufu))
    print("  pipes:", dict(pipes))
    print("  ops:  ", dict(sorted(ops.items(), key=lambda kv: -kv[1])))
    if pairs is None:
        pairs = mufu / 3.0
    if pairs:
        n = len(loop) / pairs
        alu = pipes["ALU"] / pairs
        fmah = pipes["FMAH"] / pairs
        print("  per pair (%.0f pairs/iter): %.2f instr, ALU %.2f, FMAH %.2f, FP %.2f, XU %.2f" % (
            pairs, n, alu, fmah, pipes["FP"] / pairs, pipes["XU"] / pairs))
        print("  bounds pairs/clk/SM: issue %.2f  alu %.2f  fmah %.2f  xu %.2f" % (
            128 / n, 64 / alu if alu else 0, 64 / fmah if fmah else 0, 16 / (pipes["XU"] / pairs)))


if __name__ == "__main__":
    main()
