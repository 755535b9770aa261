#!/usr/bin/env bash
# Row-streaming conv kernel (conv3x3_rs_kernel): parity in both staging modes, A/B bench against the vertical-reuse
# kernel, launch list.  Every step under its own timeout; results under gpurun_out/.
mkdir -p gpurun_out
for MODE in 0 1; do
  CVB_RS_MODE=$MODE timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "test_conv2d" > gpurun_out/rs_conv_mode$MODE.log 2>&1
  echo "mode $MODE conv tests exit $?"; tail -3 gpurun_out/rs_conv_mode$MODE.log
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest (default mode) exit $?"; tail -5 gpurun_out/pytest_gpu.log
CVB_NO_RS=1 timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_nors.json 2> gpurun_out/bench_nors.err; echo "bench no-rs exit $?"
for MODE in 0 1; do
  CVB_RS_MODE=$MODE timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_rs$MODE.json 2> gpurun_out/bench_rs$MODE.err; echo "bench mode $MODE exit $?"
done
python - <<'PY'
import json
for n in ("nors", "rs0", "rs1"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "unet_conv ms", round(d["stage_ms_per_step"]["unet_conv_tc"], 2), "found", d.get("found_rate"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rs.csv \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_launches.log 2>&1
echo "ncu exit $?"
