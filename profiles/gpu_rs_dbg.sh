#!/usr/bin/env bash
# Timing experiments on the row-streaming kernel (results invalid, durations only): CVB_RS_DEBUG bit mask.
mkdir -p gpurun_out
for D in 6 14 22 16; do
  CVB_RS_DEBUG=$D timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv3x3_rs -c 6 --csv --log-file gpurun_out/rs_dbg$D.csv \
      python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > /dev/null 2>&1
  echo "debug $D:"; grep gpu__time_duration gpurun_out/rs_dbg$D.csv | tail -3 | awk -F'","' '{print "   ", $5, $NF}'
done
