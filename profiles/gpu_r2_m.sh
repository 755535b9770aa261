#!/usr/bin/env bash
# Source-level ncu capture (stall samples per SASS line) of the kernels still furthest below their roofline.
mkdir -p gpurun_out
P="python profiles/prof_step.py --boards 148 --warmup 0 --steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_resnet_stem_tc|conv_convt_kernel|conv3x3_rs_kernel' -c 8 -f -o /tmp/m $P > gpurun_out/m_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/m_ncu.log
cp /tmp/m.ncu-rep gpurun_out/m.ncu-rep; ls -la gpurun_out/m.ncu-rep
