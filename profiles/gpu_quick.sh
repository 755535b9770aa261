#!/usr/bin/env bash
# Quick iteration: GPU parity tests, bench (no CPU baseline), ncu launch list of one pipeline pass.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_launches.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
