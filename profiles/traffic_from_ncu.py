#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch for the kernel classes bench.py reports a roofline for, from an
`ncu --set full --page raw --csv` export of ONE pipeline pass (profiles/gpu_profile_r2.sh).  bench.py cites this file
(`roofline.traffic`, `traffic_source`): a profiler counter cannot be read inside an un-profiled run.
usage: python profiles/traffic_from_ncu.py gpurun_out/prof_all_raw.csv BOARDS > profiles/r02/traffic.json"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, boards):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        return float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)

    names = [r[ix["Kernel Name"]] for r in data]
    start = max(i for i, n in enumerate(names) if "k_unet_stem" in n)          # the last pass in the capture
    last = next(i for i in range(start, len(names)) if "k_head" in names[i])
    cls = {"unet_conv_tc": [], "warp_board": [], "mask_to_quad": [], "resnet_conv_tc": [], "unet_stem": [], "resnet_stem": []}
    seen_warp = False
    for i in range(start, last + 1):
        n, r = names[i], data[i]
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        if "k_warp_board" in n:
            seen_warp = True
            cls["warp_board"].append(b)
        elif "k_mask_to_quad_fast" in n:
            cls["mask_to_quad"].append(b)
        elif "k_unet_stem" in n:
            cls["unet_stem"].append(b)
        elif "k_resnet_stem" in n:
            cls["resnet_stem"].append(b)
        elif "conv" in n and "k_" not in n.split("(")[0].split("::")[-1][:2]:
            cls["resnet_conv_tc" if seen_warp else "unet_conv_tc"].append(b)
    out = {k: {"bytes_per_launch": sum(v) / len(v), "launches": len(v), "boards": boards, "total_bytes": sum(v),
               "source": f"ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one pass over {boards} boards ({path.split('/')[-1]})"}
           for k, v in cls.items() if v}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
