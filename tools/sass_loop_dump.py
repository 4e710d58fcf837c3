#!/usr/bin/env python3
"""Dump the hottest loop (highest MUFU density) of a kernel from a library / cubin: one instruction per line, with the
control-code-free text that cuobjdump prints.   python tools/sass_loop_dump.py <lib> <name-substring>"""
import re
import subprocess
import sys


def loop_of(lib, name):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    body = next(f for f in re.split(r"\n\s*Function : ", txt)[1:] if name in f.split("\n")[0])
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, text in ins:
        if text.split()[0].startswith("BRA") or " BRA" in text[:12]:
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) <= addr:
                loop = [x for x in ins if int(t.group(1), 16) <= x[0] <= addr]
                mufu = sum(1 for x in loop if "MUFU" in x[1])
                if mufu >= 3 and (best is None or mufu / len(loop) > best[0] / len(best[1])):
                    best = (mufu, loop)
    return best[1]


if __name__ == "__main__":
    for addr, text in loop_of(sys.argv[1], sys.argv[2]):
        print("%04x  %s" % (addr, text))
