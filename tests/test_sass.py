"""Compile-level guard of the hot loop (no GPU needed): the SASS of the production kernels in libmcb200.so must keep the
instruction mix the design rests on -- a regression here costs throughput one-for-one because
the kernel is bound by the integer ALU pipe.  Uses cuobjdump through tools/sass_mix.py's parser."""
import os
import re
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


def _functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for chunk in re.split(r"\n\s*Function : ", txt)[1:]:
        name = chunk.split("\n")[0].strip()
        ins = []
        for line in chunk.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
        funcs[name] = ins
    return funcs


def _densest_mufu_loop(ins):
    best = None
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) <= addr:
                loop = [x for x in ins if int(t.group(1), 16) <= x[0] <= addr]
                mufu = sum(1 for x in loop if x[1].startswith("MUFU"))
                if mufu >= 3 and (best is None or mufu / len(loop) > best[0] / len(best[1])):
                    best = (mufu, loop)
    return best


@pytest.fixture(scope="module")
def funcs(pkg):
    pkg.build()
    return _functions(pkg.LIB_PATH)


def test_all_kernel_variants_are_in_the_library(funcs):
    names = " ".join(funcs)
    for stream in (0, 1):
        for payoff in (0, 1, 2):
            assert "simulate_kernelILi%dELi%dELb0E" % (stream, payoff) in names
    for k in ("seed_pool_kernel", "finalize_kernel", "bs_closed_form_kernel", "xgpu_wait_kernel"):
        assert k in names


@pytest.mark.parametrize("payoff,pairs_per_iter", [(0, 48), (1, 48), (2, 18)])
def test_tri_loop_instruction_mix(funcs, payoff, pairs_per_iter):
    """tri_fast loop of the production (STREAM_FAST) kernels: 3 MUFU and <= 22.25 instructions per Box-Muller pair, of
    which <= 12.85 on the ALU pipe (LOP3 / SHF / IADD3 / LEA / PRMT + the loop counter); every carry add and every
    low-word shift of the generator on the FMA-heavy pipe (IMAD.X / IMAD); no memory traffic inside the loop."""
    name = next(n for n in funcs if "simulate_kernelILi0ELi%dELb0E" % payoff in n)
    mufu, loop = _densest_mufu_loop(funcs[name])
    assert mufu == 3 * pairs_per_iter                               # LG2 + SQRT + SIN per pair, nothing else on the XU pipe
    ops = [x[1].split(".")[0] for x in loop]
    full = [x[1] for x in loop]
    outputs = 2 * pairs_per_iter // 3                               # generator outputs per iteration
    per_pair = len(loop) / pairs_per_iter
    alu = sum(o in ("LOP3", "SHF", "IADD3", "LEA", "PRMT", "VIADD", "ISETP", "IADD", "SEL", "MOV") for o in ops) / pairs_per_iter
    assert per_pair <= 22.25, per_pair
    assert alu <= 12.85, alu
    assert ops.count("LOP3") == 8 * outputs and ops.count("SHF") == 5 * outputs
    assert ops.count("IADD3") == 2 * outputs                        # low words of the two 64-bit adds ...
    assert sum(f.startswith("IMAD.X") for f in full) == 2 * outputs  # ... their carry halves stay off the ALU pipe
    assert ops.count("PRMT") == 4 * pairs_per_iter // 3 and ops.count("LEA") == pairs_per_iter
    assert not any(o in ("LDL", "STL", "LD", "ST", "LDG", "STG", "LDS", "STS") for o in ops)
    kinds = {x[1] for x in loop if x[1].startswith("MUFU")}
    assert kinds == {"MUFU.LG2", "MUFU.SQRT", "MUFU.SIN"}


def test_price_kernel_does_not_spill(pkg):
    """ptxas resource usage of the production price kernel: within the 64 registers of 4 CTAs/SM, no local memory."""
    res = subprocess.run(["cuobjdump", "-res-usage", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function ", res)
    price = next(b for b in blocks if "simulate_kernelILi0ELi0ELb0E" in b.split("\n")[0])
    m = re.search(r"REG:(\d+).*?STACK:(\d+).*?LOCAL:(\d+)", price, flags=re.S)
    assert m, price[:300]
    assert int(m.group(1)) <= 64 and int(m.group(2)) == 0 and int(m.group(3)) == 0, m.groups()
