#!/usr/bin/env bash
# One gpurun call: inference parity (known-good), then training-step bring-up under its own timeout, then bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_train.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -60 gpurun_out/pytest_train.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
