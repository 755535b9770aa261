#!/usr/bin/env python
"""Top stall locations of one kernel launch from an ncu report (source page, SASS): usage
   python profiles/stall_table.py report.ncu-rep <kernel regex> [launch index] [top n]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:100])
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
R = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[i_s] or 0) for r in R)
print("total samples", tot)
data = sorted(((int(r[i_s] or 0), k, r) for k, r in enumerate(R)), reverse=True)
for s, k, r in data[:top]:
    best = sorted([(int(r[j] or 0), h) for j, h in stall], reverse=True)[:2]
    print(f"{k:5d} {s:6d} {100 * s / tot:5.1f}% ex={r[i_ex]:>8s} {r[i_src][:66]:66s} {best}")
