#!/usr/bin/env python
"""Summarise an ncu report offline: per kernel launch, the SASS instructions with the most stall samples.
usage: tools/ncu_hot.py report.ncu-rep [launch_index] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else None; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for bi, h0 in enumerate(hdrs):
    if which is not None and bi != which: continue
    h = rows[h0]; end = hdrs[bi + 1] if bi + 1 < len(hdrs) else len(rows)
    blk = [r for r in rows[h0 + 1:end] if len(r) == len(h)]
    ix = {n: i for i, n in enumerate(h)}; si = ix["# Samples"]
    tot = sum(int(r[si] or 0) for r in blk) or 1
    stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = {}
    for r in blk:
        for c in stall:
            agg[h[c][6:]] = agg.get(h[c][6:], 0) + int(r[c] or 0)
    print(f"== launch {bi}: {tot} samples; stall mix: " + ", ".join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    for r in sorted(blk, key=lambda r: -int(r[si] or 0))[:topn]:
        st = sorted(((int(r[c] or 0), h[c][6:]) for c in stall), reverse=True)[:2]
        print(f"  {r[0][-5:]} {100*int(r[si] or 0)/tot:5.1f}%  {r[1][:64]:64s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  inst={r[ix['Instructions Executed']]}")
