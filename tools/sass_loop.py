#!/usr/bin/env python3
"""Static instruction budget of a kernel's hot loop from `cuobjdump -sass`.
    python tools/sass_loop.py libwvb.so 'FixedDecorrILb1EJLin2ELi3ELi2ELi18ELi18EEEELi6ELb1' [--list]
Finds the largest backward branch (the sample loop), prints its instruction count by opcode and by issue pipe
(ALU / FMA-lite (IMAD, IMAD.WIDE, FFMA...) / LSU / control / other), and with --list the body itself."""
import re
import subprocess
import sys
from collections import Counter

FMA = ("IMAD", "FFMA", "FMUL", "FADD", "IMUL", "IDP")  # issued to the FMA pipes on sm_100 (IMAD.MOV / IMAD.SHL included)
ALU = ("IADD", "LOP3", "SHF", "SEL", "ISETP", "LEA", "PRMT", "MOV", "FLO", "POPC", "BREV", "IABS", "IMNMX", "VIMNMX", "PLOP3", "SGXT", "BMSK", "VABSDIFF", "ICMP", "FSEL", "P2R", "R2P", "CS2R", "VIADD", "UIADD", "ULOP", "USHF", "UMOV", "USEL", "UISETP", "UPLOP", "ULEA", "UPRMT", "UFLO", "S2UR", "R2UR", "UIMAD")
LSU = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOM", "RED", "LDC", "ULDC", "LDCU")
CTL = ("BRA", "BSSY", "BSYNC", "WARPSYNC", "EXIT", "NOP", "CALL", "RET", "BAR", "BREAK", "YIELD", "NANOSLEEP", "VOTE", "SHFL", "REDUX", "BMOV", "DEPBAR", "ERRBAR", "MEMBAR", "S2R", "CCTL")


def pipe(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UIMAD",):
        return "uniform"
    for names, p in ((FMA, "fma"), (LSU, "lsu"), (CTL, "control"), (ALU, "alu")):
        if base in names or any(base.startswith(n) for n in names):
            return p
    return "other:" + base


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = [f for f in funcs if pat in f.split("\n", 1)[0]]
    assert body, "no function matches"
    f = body[0]
    print("function:", f.split("\n", 1)[0][:160])
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_\.]+)\s*([^;]*);", f):
        ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
    addr_index = {a: i for i, (a, _, _, _) in enumerate(ins)}
    loops = []
    for i, (a, pred, op, args) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
                loops.append((i - addr_index[int(m.group(1), 16)] + 1, addr_index[int(m.group(1), 16)], i))
    loops.sort(reverse=True)
    n, lo, hi = loops[0]
    print("kernel: %d instructions; hot loop: %d instructions (0x%x .. 0x%x); other loops: %s" % (len(ins), n, ins[lo][0], ins[hi][0], [l[0] for l in loops[1:6]]))
    ops = Counter(op.split(".")[0] if not op.startswith("IMAD") else ".".join(op.split(".")[:2]) for _, _, op, _ in ins[lo:hi + 1])
    pipes = Counter(pipe(op) for _, _, op, _ in ins[lo:hi + 1])
    print("by pipe:", dict(pipes.most_common()))
    print("by opcode:", dict(ops.most_common()))
    if "--list" in sys.argv:
        for a, pred, op, args in ins[lo:hi + 1]:
            print("%05x %-6s %-22s %s" % (a, pred, op, args.strip()))


if __name__ == "__main__":
    main()
