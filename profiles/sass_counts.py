#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS instructions in the shipped library (the tensor-pipe / TMA proof):
    UTCHMMA  tcgen05.mma (kind::f16)        LDTM     tcgen05.ld (TMEM -> registers)
    UTMALDG  cp.async.bulk.tensor load      UTMASTG  cp.async.bulk.tensor store      UTCBAR  tcgen05.commit
    python profiles/sass_counts.py [lib.so] > profiles/r02/sass_counts.txt
The ncu metric used for tensor-pipe utilisation in profiles/*/ncu_full_summary.md is
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active (tensor_%act) and ..._elapsed (tensor_%el)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "chessvision-3lc_b200" / "libchessvision_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "IMMA")
per = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"cvb::\(anonymous namespace\)::|cvb::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = Counter()
        continue
    if cur is None:
        continue
    for n in names:
        if re.search(rf"\b{n}\b", line):
            per[cur][n] += 1
print(f"# {lib}")
print('# arch:', sorted(set(re.findall(r'arch = (sm_[0-9a-z]+)', sass))))
print(f"{'kernel':70s} " + " ".join(f"{n:>8s}" for n in names))
tot = Counter()
for k, c in per.items():
    print(f"{k[:70]:70s} " + " ".join(f"{c[n]:8d}" for n in names))
    tot.update(c)
print(f"{'TOTAL':70s} " + " ".join(f"{tot[n]:8d}" for n in names))
