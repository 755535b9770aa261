#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum csv) of one pipeline pass: per-kernel time, and TFLOP/s for the
UNet layers (128 boards per launch).  usage: python profiles/launch_table.py gpurun_out/launches.csv [baseline.csv]"""
import csv
import sys

# GFLOP per board of the UNet launches in plan order (stem, inc.3, down1..4, then convT / conv0 / conv3 of up1..4); the four
# max-pools are fused into the producing convs since round 1 and no longer appear as launches
UNET_GF = [0, 4.83, 2.416, 4.83, 2.416, 4.83, 2.416, 4.83, 2.416, 4.83, 1.07, 9.66, 4.83, 1.07, 9.66, 4.83, 1.07, 9.66, 4.83, 1.07,
           9.66, 4.83]


def load(path):
    rows, hdr, out = list(csv.reader(open(path))), None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            out.append((d["Kernel Name"], float(d["Metric Value"].replace(",", "")) / 1e3))
    start = [i for i, l in enumerate(out) if "unet_stem" in l[0]][-1]
    return out[start:start + 46]


cur = load(sys.argv[1])
old = load(sys.argv[2]) if len(sys.argv) > 2 else None
boards = int(sys.argv[3]) if len(sys.argv) > 3 else 128
tot = 0.0
for k, (name, us) in enumerate(cur):
    tf = UNET_GF[k] * boards / us if k < len(UNET_GF) and UNET_GF[k] else 0   # GFLOP / us = PFLOP/s
    prev = f"(was {old[k][1]:8.1f})" if old and k < len(old) else ""
    short = name.replace("cvb::", "").replace("<unnamed>::", "").replace("void ", "")[:46]
    print(f"{k:2d} {short:46s} {us:9.1f} us {prev} {tf * 1e3:7.0f} TF/s" if tf else f"{k:2d} {short:46s} {us:9.1f} us {prev}")
    tot += us
print(f"total {tot:.1f} us = {tot / boards:.2f} us/board")
