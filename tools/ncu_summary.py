#!/usr/bin/env python
"""Offline summary of ncu reports (no GPU needed): one row per captured launch with the metrics DESIGN.md quotes.
usage: tools/ncu_summary.py out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, subprocess, sys
WANT = [("gpu__time_duration.sum", "us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("dram__bytes_read.sum", "rd_MB", 1), ("dram__bytes_write.sum", "wr_MB", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]
out = open(sys.argv[1], "w")
out.write("| report | kernel | " + " | ".join(n for _, n, _ in WANT) + " |\n|---|---|" + "---|" * len(WANT) + "\n")
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        vals = []
        for key, name, _ in WANT:
            i = ix.get(key)
            v = r[i] if i is not None else ""
            u = units[i] if i is not None else ""
            try:
                f = float(v.replace(",", ""))
                if name == "us":
                    f = f / 1e3 if u in ("ns", "nsecond") else (f if u in ("us", "usecond") else f * 1e3 if u in ("ms", "msecond") else f)
                if name.endswith("_MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * f
                vals.append(f"{f:.1f}" if abs(f) < 1e5 else f"{f:.3g}")
            except ValueError:
                vals.append(v)
        kn = r[ix["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:60]
        out.write(f"| {rep.split('/')[-1]} | `{kn}` | " + " | ".join(vals) + " |\n")
out.close()
print(open(sys.argv[1]).read()[:3000])
