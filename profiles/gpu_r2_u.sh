#!/usr/bin/env bash
# Two epilogue groups for the N = 128 vertical-reuse layers (Cin <= 128), Cin = 64 as a pair: parity + A/B + launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_u.log | cut -c1-200
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
run u_default CVB_X=1
run u_noepg2 CVB_EPG2=0
run u_default2 CVB_X=1
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_u.csv $P > gpurun_out/prof_launches_u.log 2>&1
python profiles/launch_table.py gpurun_out/launches_u.csv profiles/r02/ncu_launches_148boards.csv 148 2>&1 | head -50 | cut -c1-120
