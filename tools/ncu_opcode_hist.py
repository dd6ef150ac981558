#!/usr/bin/env python3
"""Histogram of executed warp-instructions and stall samples per SASS opcode from an ncu report.

    python tools/ncu_opcode_hist.py gpurun_out/prof.ncu-rep <kernel regex> [top-n]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True,
                     text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', txt)
for blk in blocks[1:]:
    lines = blk.split("\n")
    print("==", lines[0][:150])
    rd = csv.reader(io.StringIO("\n".join(lines[1:])))
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    inst = defaultdict(int)
    samp = defaultdict(int)
    tot = 0
    tots = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = defaultdict(int)
    for row in rd:
        if len(row) < len(hdr):
            continue
        src = row[ci["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = ".".join(op.split(".")[:2])
        n = int(row[ci["Instructions Executed"]] or 0)
        s = int(row[ci["# Samples"]] or 0)
        inst[op] += n
        samp[op] += s
        tot += n
        tots += s
        for h in stall_cols:
            stalls[h] += int(row[ci[h]] or 0)
    print(f"total warp-instructions {tot:,}  samples {tots:,}")
    for op, n in sorted(inst.items(), key=lambda kv: -kv[1])[:topn]:
        print(f"  {op:24s} {n:14,d} {100.0 * n / max(tot, 1):6.2f}%   samples {100.0 * samp[op] / max(tots, 1):6.2f}%")
    print("  stalls:", ", ".join(f"{h[6:]}={100.0 * v / max(tots, 1):.1f}%" for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
