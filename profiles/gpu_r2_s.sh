#!/usr/bin/env bash
# A/B: weights of the Cin = 128 vertical-reuse pair layer streamed instead of resident; fused conv+convT pair with 4 + 12 stages.
mkdir -p gpurun_out
CVB_CONVT_DEEP_A=1 CVB_VR_PAIR_STREAM=1 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -m gpu -q -x > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s.log | cut -c1-200
run() {
  env $2 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 --api-steps 1 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 exit $?"
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
run s_default CVB_X=1
run s_vrstream CVB_VR_PAIR_STREAM=1
run s_deepa CVB_CONVT_DEEP_A=1
run s_both "CVB_VR_PAIR_STREAM=1 CVB_CONVT_DEEP_A=1"
run s_default2 CVB_X=1
