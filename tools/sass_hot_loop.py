#!/usr/bin/env python
"""Extract the steady-state loop of a packed extension kernel from the built library and count its instructions by pipe.

    python tools/sass_hot_loop.py [--lib agatha_b200/lib/libagatha_b200.so] [--C 24 --NW 1 --JWS 7] [--out profiles/extend16_c24_hot_loop_r02]

Writes <out>.sass (the loop, two anti-diagonals per iteration) and <out>.json. bench.py uses the counts for the roofline:
the ALU pipe issues 16 lanes per SM sub-partition per clock, so  ceiling = measured ALU lane-op rate / (ALU instructions per
lane-cell). The "new maximum" block (the shared-memory snapshot, executed only on anti-diagonals that raise the running
maximum) is counted separately from the instructions every iteration executes."""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALU = ("VIADDMNMX", "VIMNMX", "PRMT", "LOP3", "SHF", "SEL", "ISETP", "IADD3", "VIADD", "LEA", "PLOP3", "IABS", "POPC", "FLO", "BREV", "I2I", "FSEL", "FMNMX", "MOV ", "IMNMX")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")   # IMAD.MOV / IMAD.U32 / IMAD.SHL / IMAD.IADD included


def pipe_of(op):
    if op.startswith(FMA):
        return "fma"
    if op.startswith(ALU) or op == "MOV":
        return "alu"
    if op.startswith("U") or op.startswith("BRA") or op.startswith("BSSY") or op.startswith("BSYNC") or op.startswith("VOTEU"):
        return "uniform_or_branch"
    if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "SHFL", "CREDUX", "REDUX", "BAR", "VOTE", "S2R", "R2UR")):
        return "lsu_or_other"
    return "other"


def kernel_symbol(C, NW, JWS):
    return "_ZN6agatha15extend16_kernelILi%dELi%dELi%dEEEvNS_9JobArraysENS_12KernelParamsE" % (C, NW, JWS)


def disassemble(lib, sym):
    r = subprocess.run(["cuobjdump", "-sass", "-fun", sym, lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    ins = []
    for line in r.stdout.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    text = re.sub(r"^@!?U?P\d\s+", "", text)
    return text.split()[0]


def analyse(ins, C):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                loops.append((tgt, a))
    P = C // 2
    need = 2 * P * 3                      # per iteration: 2 steps x P registers x (2 VIADDMNMX + 1 VIMNMX3) at least
    best = None
    for lo, hi in loops:
        body = [(a, t) for a, t in ins if lo <= a <= hi]
        n16 = sum(1 for _, t in body if "U16x2" in t)
        has_ldg = any(opcode(t).startswith("LDG") for _, t in body)
        if n16 >= need and not has_ldg and (best is None or len(body) < len(best)):
            best = body
    if best is None:
        return None
    # the "new maximum" blocks: a run of at least P/2 snapshot stores (single STS are the lane-edge / maximum hand-over of the
    # multi-warp shapes and run on every anti-diagonal), from the conditional branch that guards the run to its last store
    ops = [opcode(t) for _, t in best]
    in_block = [False] * len(best)
    i = 0
    while i < len(best):
        if ops[i].startswith("STS"):
            k, last, n_sts = i, i, 0
            while k < len(best) and (k - last) < 6:
                if ops[k].startswith("STS"):
                    last = k; n_sts += 1
                k += 1
            if n_sts >= P // 2:
                j = i
                while j > 0 and not ops[j - 1].startswith("BRA"):
                    j -= 1
                for x in range(j, last + 1):
                    in_block[x] = True
            i = last + 1
        else:
            i += 1
    count = collections.Counter()
    hot = collections.Counter()
    for (a, t), blk in zip(best, in_block):
        p = pipe_of(opcode(t))
        count[p] += 1
        if not blk:
            hot[p] += 1
    mix = collections.Counter(opcode(t) for _, t in best)
    return {"loop_start": hex(best[0][0]), "loop_end": hex(best[-1][0]), "instructions_in_loop": len(best),
            "by_pipe_in_loop": dict(count), "every_iteration": dict(hot), "every_iteration_total": sum(hot.values()),
            "new_maximum_blocks": sum(in_block), "anti_diagonals_per_iteration": 2, "lane_cells_per_iteration": 2 * C,
            "alu_per_lane_cell": hot["alu"] / (2.0 * C), "opcode_mix": dict(mix.most_common())}, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "agatha_b200", "lib", "libagatha_b200.so"))
    ap.add_argument("--C", type=int, default=24)
    ap.add_argument("--NW", type=int, default=1)
    ap.add_argument("--JWS", type=int, default=7)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sym = kernel_symbol(a.C, a.NW, a.JWS)
    ins = disassemble(a.lib, sym)
    if not ins:
        print(json.dumps({"error": "kernel %s not found in %s (or cuobjdump missing)" % (sym, a.lib)}))
        return 1
    res = analyse(ins, a.C)
    if res is None:
        print(json.dumps({"error": "no steady-state loop found"}))
        return 1
    info, body = res
    info["kernel"] = "agatha::extend16_kernel<%d,%d,%d>" % (a.C, a.NW, a.JWS)
    info["kernel_instructions"] = len(ins)
    print(json.dumps(info, indent=1))
    if a.out:
        with open(a.out + ".json", "w") as f:
            json.dump(info, f, indent=1)
        with open(a.out + ".sass", "w") as f:
            f.write("// %s: steady-state loop (two anti-diagonals per iteration), from cuobjdump -sass of the built library\n" % info["kernel"])
            for ad, t in body:
                f.write("/*%04x*/  %s ;\n" % (ad, t))
    return 0


if __name__ == "__main__":
    sys.exit(main())
