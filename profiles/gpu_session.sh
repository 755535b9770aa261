#!/usr/bin/env bash
# One gpurun call: the GPU parity suite, a short bench, and the launch list of one 128-board pass.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, "found", d.get("found_rate"), d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_launches.log 2>&1
echo "ncu exit $?"
