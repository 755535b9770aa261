#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for i in 1 2; do
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --api-steps 1 > gpurun_out/bench_pipeline.json 2> gpurun_out/bench_pipeline.err; echo "bench exit $?"; tail -3 gpurun_out/bench_pipeline.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_pipeline.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), d["e2e_api"]["fen_equal_e2e_arm"],
          "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
done
