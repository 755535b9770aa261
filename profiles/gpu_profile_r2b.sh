#!/usr/bin/env bash
# Final round-2 ncu evidence: pipeline pass over one 148-board chunk (launch list + --set full) and the launch list of the
# UNet training step at batch 32.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/prof_launches.log 2>&1
N=$(grep -c 'gpu__time_duration.sum' gpurun_out/launches.csv)
PASS=$((N / 2))
echo "launches per pass: $PASS"
ncu --set full --clock-control none -s $PASS -c $PASS -f -o /tmp/prof_all $P > gpurun_out/prof_full.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2> gpurun_out/prof_export.err
python profiles/summarize_raw.py gpurun_out/prof_all_raw.csv > gpurun_out/ncu_full_summary.md; head -30 gpurun_out/ncu_full_summary.md | cut -c1-200
python profiles/traffic_from_ncu.py gpurun_out/prof_all_raw.csv 148 > gpurun_out/traffic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/train_launches_b32.csv \
    python bench.py --workload train --train-batch 32 --steps 1 --warmup 3 > gpurun_out/prof_train.log 2>&1
echo "train ncu exit $?"; grep -c gpu__time_duration gpurun_out/train_launches_b32.csv
