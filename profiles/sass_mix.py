#!/usr/bin/env python
"""Per-kernel dynamic SASS mix and stall reasons from an .ncu-rep captured with --set full --import-source on.

    python profiles/sass_mix.py gpurun_out/prof.ncu-rep k_vertex [top_n]
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def main(rep, kernel, top=14):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
    lines = out.splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    if not starts:
        print("kernel not found"); return
    seg = lines[starts[0] + 1: starts[1] if len(starts) > 1 else None]
    rd = csv.reader(io.StringIO("\n".join(seg)))
    h = next(rd)
    isrc, iex, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    ops, stalls, rows = Counter(), Counter(), []
    total = 0
    for r in rd:
        if len(r) <= iex:
            continue
        ex = int(r[iex] or 0)
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[isrc])
        op = m.group(2) if m else "?"
        ops[op] += ex
        total += ex
        for i, n in stall_cols:
            stalls[n] += int(r[i] or 0)
        rows.append((int(r[isamp] or 0), ex, r[isrc].strip()))
    print("%s: %d warp instructions executed" % (kernel, total))
    print("  opcode mix:", ", ".join("%s %.1f%%" % (o, 100.0 * c / total) for o, c in ops.most_common(top)))
    st = sum(stalls.values()) or 1
    print("  stall reasons:", ", ".join("%s %.1f%%" % (o[6:], 100.0 * c / st) for o, c in stalls.most_common(8)))
    ssum = sum(r[0] for r in rows) or 1
    print("  top stall instructions:")
    for s, ex, src in sorted(rows, reverse=True)[:top]:
        print("    %5.1f%%  x%-8d %s" % (100.0 * s / ssum, ex, src[:90]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)
