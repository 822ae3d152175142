#!/usr/bin/env python
"""Stall samples of an ncu capture aggregated by CUDA SOURCE LINE (outermost frame of the inline chain).

usage: tools/ncu_lines.py report.ncu-rep object.o kernel_substring [launch_index] [top_n]

ncu's CSV export of the source page is per SASS instruction; the kernels here are one big function full of inlined
helpers (mbar_wait, umma_*, ...), so the useful key is the line of the .cu file that called the helper.  The map
offset -> line comes from `nvdisasm -gi` of the cubin inside the object file (-lineinfo build), the samples from
`ncu --page source --csv --print-source sass`; both list the function's instructions in the same order.
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict

rep, obj, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 30

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout

# ---- per function: list of (offset, outermost line in the main .cu, innermost file:line)
funcs, cur, name, outer, inner = {}, None, None, None, None
main_cu = os.path.basename(obj).replace(".o", ".cu")
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        name = m.group(1); cur = funcs.setdefault(name, []); outer = inner = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        f, l = os.path.basename(m.group(1)), int(m.group(2))
        if "inlined at" in ln:
            inner = inner or (f, l)
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', ln)
            if m2 and os.path.basename(m2[-1][0]) == main_cu:
                outer = int(m2[-1][1])
        else:
            if f == main_cu:
                outer = l
            inner = inner or (f, l)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur is not None:
        cur.append((int(m.group(1), 16), outer, inner, m.group(2).strip()))
        inner = None

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and rows[i - 1] and rows[i - 1][0] == "Kernel Name" else "?" for i in hdrs]
sel = [k for k, n in enumerate(names) if ksub in n]
k = sel[which] if sel else which
h0 = hdrs[k]; h = rows[h0]; end = hdrs[k + 1] if k + 1 < len(hdrs) else len(rows)
blk = [r for r in rows[h0 + 1:end] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
base = int(blk[0][0], 16)
# pick the disassembled function with the same instruction count whose mangled name contains pieces of the kernel name
cands = [(fn, ins) for fn, ins in funcs.items() if len(ins) == len(blk)]
if not cands:
    sys.exit(f"no function with {len(blk)} instructions in {obj} (have {[len(v) for v in funcs.values()]})")
fn, ins = cands[0]
off2line = {o: (ol, il) for o, ol, il, _ in ins}
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: [0, Counter(), 0])
tot = 0
for r in blk:
    off = int(r[0], 16) - base
    ol, il = off2line.get(off, (None, None))
    s = int(r[ix["# Samples"]] or 0); tot += s
    a = agg[ol]; a[0] += s; a[2] = max(a[2], int(r[ix["Instructions Executed"]] or 0))
    for c in stall_cols:
        a[1][h[c][6:]] += int(r[c] or 0)
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", main_cu)).read().splitlines()
print(f"== {names[k][:90]}: {tot} samples, function {fn[:60]}")
for ol, (s, st, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    code = src[ol - 1].strip()[:86] if ol and ol <= len(src) else "?"
    top = ", ".join(f"{a}:{b}" for a, b in st.most_common(3))
    print(f"{100 * s / max(tot, 1):5.1f}%  L{ol}  x{n:<7d} {code:86s} | {top}")
