#!/usr/bin/env bash
# conv parity in both staging modes + launch list + short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x > gpurun_out/conv_tests.log 2>&1; echo "conv tests exit $?"; tail -3 gpurun_out/conv_tests.log
CVB_RS_MODE=0 timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k test_conv2d > gpurun_out/conv_tests0.log 2>&1; echo "conv tests mode 0 exit $?"; tail -2 gpurun_out/conv_tests0.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_nets.py -q -x > gpurun_out/pipe_tests.log 2>&1; echo "pipeline tests exit $?"; tail -2 gpurun_out/pipe_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, "found", d.get("found_rate"), d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_launches.log 2>&1
echo "ncu exit $?"
