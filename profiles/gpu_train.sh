#!/usr/bin/env bash
# GPU bring-up / parity of the UNet training step: pytest (tests/test_gpu_train.py) with its printed error figures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -s > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -40 gpurun_out/pytest_train.log
