#!/usr/bin/env bash
# One gpurun call producing everything profiles/r01 holds: GPU parity suite, bench lines (pipeline, reference arm, training
# step, JPEG front-end), ncu launch lists and one full-set capture of every launch of one 128-board pass.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; echo "reference arm exit $?"
timeout 600 python bench.py --workload train --train-batch 8 > gpurun_out/train_b8.json 2> gpurun_out/train_b8.err; echo "train exit $?"
timeout 600 python bench.py --workload train --train-batch 32 > gpurun_out/train_b32.json 2>> gpurun_out/train_b8.err; echo "train b32 exit $?"
timeout 600 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/decode_n1.json 2> gpurun_out/decode_n1.err; echo "decode exit $?"
P="python profiles/prof_step.py --boards 128 --warmup 1 --steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/prof_launches.log 2>&1
N=$(grep -c 'gpu__time_duration.sum' gpurun_out/launches.csv); PASS=$((N / 2)); echo "launches per pass: $PASS"
timeout 900 ncu --set full --clock-control none -s $PASS -c $PASS -f -o /tmp/prof_all $P > gpurun_out/prof_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2> gpurun_out/prof_export.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_jpeg -c 8 --csv --log-file gpurun_out/decode_launches.csv \
    python bench.py --workload decode --boards 128 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu decode exit $?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt
cat gpurun_out/bench_n1.json | cut -c1-300
