#!/usr/bin/env bash
# ncu --set full of the row-streaming conv kernel (mode from $1, default 1) and of the vertical-reuse kernel it
# replaces (CVB_NO_RS=1), on one 128-board pass; raw pages under gpurun_out/.
mkdir -p gpurun_out
MODE=${1:-1}
P="python profiles/prof_step.py --boards 128 --warmup 1 --steps 1"
CVB_RS_MODE=$MODE timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_rs -s 3 -c 3 -f -o /tmp/rs $P > gpurun_out/rs_ncu.log 2>&1
echo "ncu rs exit $?"
ncu -i /tmp/rs.ncu-rep --page raw --csv > gpurun_out/rs_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 12 -f -o /tmp/vr $P > gpurun_out/vr_ncu.log 2>&1
echo "ncu vr exit $?"
ncu -i /tmp/vr.ncu-rep --page raw --csv > gpurun_out/vr_raw.csv 2>/dev/null
cp /tmp/rs.ncu-rep gpurun_out/rs.ncu-rep; cp /tmp/vr.ncu-rep gpurun_out/gen.ncu-rep
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv
