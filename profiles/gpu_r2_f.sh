#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
run() {
  env $2 timeout 300 python bench.py --no-cpu-baseline --steps 6 --warmup 3 --api-steps 1 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 exit $?"
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
          {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["clocks"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run fused CVB_X=1
run unfused CVB_NO_CONVT_FUSE=1
run fused2 CVB_X=1
