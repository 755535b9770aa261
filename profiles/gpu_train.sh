#!/usr/bin/env bash
# UNet training step (configs[4]): bench at two batch sizes + ncu launch list of one step.
mkdir -p gpurun_out
for b in 2 8 32; do
  timeout 300 python bench.py --workload train --train-batch $b --steps 10 --warmup 3 > gpurun_out/train_b$b.json 2>> gpurun_out/train.err; echo "train b=$b exit $?"
  cat gpurun_out/train_b$b.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/train_launches.csv \
    python bench.py --workload train --train-batch 8 --steps 1 --warmup 3 > gpurun_out/train_prof.log 2>&1
tail -3 gpurun_out/train.err
