#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_train.log | cut -c1-300
bash profiles/gpu_scale_r2.sh 1
