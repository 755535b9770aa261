#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (one row per launch, ~2400 metric columns) into the per-launch table kept
under profiles/: duration, DRAM traffic, DRAM / tensor-pipe / issue utilisation, IPC, occupancy, registers.
usage: python profiles/summarize_raw.py gpurun_out/prof_all_raw.csv > profiles/rNN/ncu_full_summary.md"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us", 1e3, "ms"),
    ("dram__bytes_read.sum", "dram_rd_MB", None, None),
    ("dram__bytes_write.sum", "dram_wr_MB", None, None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1, None),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act", 1, None),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%el", 1, None),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_%act", 1, None),
    ("sm__inst_executed.avg.per_cycle_active", "ipc_act", 1, None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1, None),
    ("launch__registers_per_thread", "regs", 1, None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1, None),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "bank_confl", 1, None),
    ("sm__cycles_active.avg", "sm_cycles", 1, None),
    ("sm__inst_executed.sum", "inst", 1, None),
]
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
UNIT_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [c for c in COLS if c[0] in idx]
    print("| # | kernel | grid | " + " | ".join(c[1] for c in names) + " |")
    print("|---|---|---|" + "---|" * len(names))
    for k, r in enumerate(data):
        kn = r[idx["Kernel Name"]].replace("cvb::", "").replace("<unnamed>::", "").replace("void ", "")
        kn = kn.split("(")[0][:40]
        vals = []
        for metric, label, _, _ in names:
            i = idx[metric]
            v, u = num(r[i]), units[i]
            if label == "dur_us":
                v *= UNIT_TO_US.get(u, 1.0)
                vals.append(f"{v:.1f}")
            elif label.endswith("_MB"):
                v *= UNIT_TO_MB.get(u, 1.0)
                vals.append(f"{v:.2f}")
            elif label in ("inst", "bank_confl", "sm_cycles"):
                vals.append(f"{v:.3g}")
            else:
                vals.append(f"{v:.1f}")
        print(f"| {k} | {kn} | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
