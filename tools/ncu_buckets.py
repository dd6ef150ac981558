#!/usr/bin/env python3
"""Stall samples of one kernel in an ncu report, bucketed by SASS position (100 instructions per bucket).

    python tools/ncu_buckets.py gpurun_out/prof.ncu-rep <kernel regex> [bucket]
"""
import csv
import io
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', txt)
blk = blocks[1]
print(blk.split("\n")[0][:140])
rd = csv.reader(io.StringIO("\n".join(blk.split("\n")[1:])))
hdr = next(rd)
ci = {h: i for i, h in enumerate(hdr)}
rows = [r for r in rd if len(r) >= len(hdr)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in rows)
print("instructions", len(rows), "samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for b in range(0, len(rows), B):
    seg = rows[b:b + B]
    s = sum(int(r[ci["# Samples"]] or 0) for r in seg)
    ex = sum(int(r[ci["Instructions Executed"]] or 0) for r in seg)
    st = {h: sum(int(r[ci[h]] or 0) for r in seg) for h in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = {}
    for r in seg:
        t = r[ci["Source"]].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    topops = sorted(ops.items(), key=lambda kv: -kv[1])[:3]
    print(f"{b:5d} samples {s:6d} ({100 * s / max(tot, 1):4.1f}%) exec {ex / 1e6:7.1f}M ",
          " ".join(f"{k[6:]}={v}" for k, v in top), " | ", " ".join(f"{k}:{v}" for k, v in topops))
