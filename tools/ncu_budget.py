#!/usr/bin/env python3
"""Dynamic instruction budget of a kernel from an ncu report's source page (SASS view):
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_budget.py src.csv <frames> [--top N] [--lines]
Prints warp instructions per `frames` (e.g. samples decoded / 32 lanes) by opcode and pipe, and the hottest source lines."""
import csv
import sys
from collections import Counter

sys.path.insert(0, __import__("os").path.dirname(__file__))
from sass_loop import pipe  # noqa: E402


def main():
    path, frames = sys.argv[1], float(sys.argv[2])
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    iA, iS, iN, iT = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    iSamp = hdr.index("# Samples")
    ops, pipes, total, thr = Counter(), Counter(), 0, 0
    samp = Counter()
    lines = []
    for r in rows[2:]:
        if len(r) <= iT:
            continue
        src = r[iS].strip()
        n = int(r[iN] or 0)
        parts = src.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") else parts[0]
        base = op.split(".")[0] if not op.startswith("IMAD") else ".".join(op.split(".")[:2])
        ops[base] += n
        pipes[pipe(op)] += n
        samp[pipe(op)] += int(r[iSamp] or 0)
        total += n
        thr += int(r[iT] or 0)
        lines.append((n, r[iA], src, int(r[iSamp] or 0)))
    print("warp instructions: %d total, %.1f per frame group; avg active threads %.2f" % (total, total / frames, thr / max(total, 1)))
    print("by pipe (per frame group):", {k: round(v / frames, 1) for k, v in pipes.most_common()})
    print("stall samples by pipe:", dict(samp.most_common()))
    print("by opcode (per frame group):", {k: round(v / frames, 1) for k, v in ops.most_common(40)})
    if "--lines" in sys.argv:
        for n, a, s, sm in lines:
            print("%8.2f %6d  %s" % (n / frames, sm, s))


if __name__ == "__main__":
    main()
