#!/usr/bin/env bash
# Round 2 ncu evidence for one pass of the image->FEN pipeline over one chunk of 148 boards (= the bench's chunk):
#   launches.csv      gpu__time_duration.sum of every launch (warm-up pass + measured pass)
#   prof_all_raw.csv  `--set full` raw page of every kernel of the SECOND pass
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/prof_launches.log 2>&1
N=$(grep -c 'gpu__time_duration.sum' gpurun_out/launches.csv)
PASS=$((N / 2))
echo "launches per pass: $PASS"
ncu --set full --import-source on --clock-control none -s $PASS -c $PASS -f -o /tmp/prof_all $P > gpurun_out/prof_full.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2> gpurun_out/prof_export.err
ls -la /tmp/prof_all.ncu-rep
SZ=$(stat -c %s /tmp/prof_all.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 0 ] && [ "$SZ" -lt 60000000 ]; then cp /tmp/prof_all.ncu-rep gpurun_out/prof_all.ncu-rep; fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt
python profiles/summarize_raw.py gpurun_out/prof_all_raw.csv > gpurun_out/ncu_full_summary.md; head -60 gpurun_out/ncu_full_summary.md
python profiles/traffic_from_ncu.py gpurun_out/prof_all_raw.csv 148 > gpurun_out/traffic.json; cat gpurun_out/traffic.json | head -40
