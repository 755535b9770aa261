#!/usr/bin/env bash
# One gpurun call: GPU parity tests, bench (both stem variants), ncu launch list and one full capture of the convs.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
CVB_STEM_FP32=1 python bench.py --no-cpu-baseline > gpurun_out/bench_stem_fp32.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv|stem' -s 43 -c 43 -o gpurun_out/prof_conv \
    python profiles/prof_step.py --boards 32 --warmup 1 --steps 1 > gpurun_out/prof_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mask_to_quad|warp_board' -s 2 -c 2 -o gpurun_out/prof_geom \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_geom.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
